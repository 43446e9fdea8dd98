"""Slab-decomposed run of the hot path on P GPUs (one process per GPU, launched by torchrun) checked on rank 0
against the single-lattice CPU checker.  Driven by tests/test_multigpu.py and scripts/gpu_multi.sh:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_worker.py

Covers: distributed FFT (2-D local, exchange by NCCL all-to-all and by transposes pushed over peer memory, 1-D local) both ways, ghost-plane exchange (updateHalo), deposit
fold (projection_comm), slab migration of particles (moveParticles), parallel.sum / max, and whole cycles of the time loop
-- decomposition independence: P slabs must reproduce the undecomposed result to round-off.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "gevolution-1.2_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402  (first: its bundled NCCL must be the one mapped into the process)
import torch.distributed as dist  # noqa: E402

import common  # noqa: E402
import gevb  # noqa: E402

TOL = 1e-10


def gather(obj):
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def slab(a, ctx, axis=1):
    """local z-slab of a global real array [c][z][y][x]"""
    return np.ascontiguousarray(np.take(a, range(ctx.z0, ctx.z0 + ctx.nzl), axis=axis))


RESULTS = []          # (name, error, tolerance) of every check made on this rank (bench.py reports the worst one)
QUIET = False         # bench.py: no line per check


def check(name, err, tol, failures):
    ok = err <= tol
    RESULTS.append((name, float(err), float(tol)))
    if dist.get_rank() == 0 and not (QUIET and ok):
        print(f"  {'ok ' if ok else 'BAD'} {name}: {err:.3e} (tol {tol:.1e})", flush=True, file=sys.stderr if QUIET else sys.stdout)
    if not ok:
        failures.append((name, float(err)))


def case_fft(ctx, N, failures):
    rng = np.random.default_rng(11)
    f = common.gaussian_field(rng, N, 3, 1.0)
    F, K = gevb.Field(ctx, gevb.REAL, 3, data=slab(f, ctx)), gevb.Field(ctx, gevb.CPLX, 3)
    plan = gevb.PlanFFT(F, K)
    plan.execute(gevb.FFT_FORWARD)
    parts = gather(K.download())                                     # [c][ky_local][kx][kz][2] per rank
    k_all = np.concatenate(parts, axis=1)
    ref = np.fft.rfftn(f, axes=(1, 2, 3))                            # [c][kz][ky][kx]
    ref = ref.transpose(0, 2, 3, 1)                                  # -> [c][ky][kx][kz]
    check(f"N={N} fft forward (slab layout)", common.rel_linf(common.to_cplx(k_all), ref), 1e-12, failures)
    plan.execute(gevb.FFT_BACKWARD)
    back = np.concatenate(gather(F.download()), axis=1)
    check(f"N={N} fft backward (unnormalised)", common.rel_linf(back / N ** 3, f), 1e-12, failures)
    again = np.concatenate(gather(K.download()), axis=1)
    check(f"N={N} fft backward preserves its Fourier input", np.abs(again - k_all).max(), 0.0, failures)
    # divergence-free projection through the distributed layout (k decode on slabs)
    gevb.projectFTvector(K, K, 1.0, 0.0)
    plan.execute(gevb.FFT_BACKWARD)
    B = np.concatenate(gather(F.download()), axis=1)
    div = sum(B[i] - np.roll(B[i], 1, axis=2 - i) for i in range(3))
    check(f"N={N} projectFTvector divergence", np.abs(div).max() / np.abs(B).max(), 1e-11, failures)
    plan.close(); F.close(); K.close()


def case_particles(ctx, chk, N, failures):
    """deposit + fold, halo, migration with many slab crossings"""
    rng = np.random.default_rng(5)
    a = 0.03
    ids, pos, vel = common.quasi_uniform_particles(rng, N, sigma=0.6, a=a, qscale=0.5)
    ids2, pos2, vel2 = common.clustered_particles(rng, N, N ** 3, a=a, qscale=0.5)
    ids, pos, vel = np.concatenate([ids, ids2 + len(ids)]), np.concatenate([pos, pos2]), np.concatenate([vel, vel2])
    phi, chi, Bi = common.metric_fields(rng, N, a)
    phi, chi, Bi = phi * 1e3, chi * 1e4, Bi * 1e4
    mass = 0.31 / len(ids)
    p = gevb.Particles(ctx, mass).add(ids, pos, vel)
    counts = gather(p.count())
    check(f"N={N} particle conservation at add", abs(sum(counts) - len(ids)), 0, failures)
    fphi = gevb.Field(ctx, gevb.REAL, 1, data=slab(phi, ctx)).updateHalo()
    fchi = gevb.Field(ctx, gevb.REAL, 1, data=slab(chi, ctx)).updateHalo()
    fB = gevb.Field(ctx, gevb.REAL, 3, data=slab(Bi, ctx)).updateHalo()
    T00, T0i, Tij = gevb.Field(ctx, gevb.REAL, 1), gevb.Field(ctx, gevb.REAL, 3), gevb.Field(ctx, gevb.REAL, 6)
    p.projection_T00_project(T00, a, fphi, 1.0); T00.projection_comm()
    p.projection_T0i_project(T0i, fphi, 1.0); T0i.projection_comm()
    p.projection_Tij_project(Tij, a, fphi, 1.0); Tij.projection_comm()
    total = T00.sum()
    g00, g0i, gij = (np.concatenate(gather(f.download()), axis=1) for f in (T00, T0i, Tij))
    if dist.get_rank() == 0:
        r00 = chk.projection_T00(N, pos, vel, mass, a, phi, 1.0)
        check(f"N={N} T00 deposit + fold", common.rel_linf(g00, r00), TOL, failures)
        check(f"N={N} T0i deposit + fold", common.rel_linf(g0i, chk.projection_T0i(N, pos, vel, mass, phi, 1.0)), TOL, failures)
        check(f"N={N} Tij deposit + fold", common.rel_linf(gij, chk.projection_Tij(N, pos, vel, mass, a, phi, 1.0)), TOL, failures)
        check(f"N={N} field sum + parallel.sum", abs(total - r00.sum()) / abs(r00.sum()), 1e-12, failures)
    # kick (needs the ghost planes of phi, chi, B) then a drift long enough to cross slabs
    params = np.array([a, a * a * N])
    vmax = p.updateVel(gevb.UPDATE_Q, 0.0095, [fphi, fchi, fB], 3, params)
    vmax_all = ctx.parallel_max([vmax])[0]
    dtau = 0.25 * ctx.nzl / N                                # |v| < 1: nobody moves farther than a quarter slab
    p.moveParticles(gevb.UPDATE_Q, dtau, [fphi, fchi, fB], 3, params)
    parts = gather(p.download())
    cnts = gather(p.cell_counts())
    gid = np.concatenate([q[0] for q in parts]); gpos = np.concatenate([q[1] for q in parts]); gvel = np.concatenate([q[2] for q in parts])
    crossed = sum(abs(len(q[0]) - c) for q, c in zip(parts, counts))
    if dist.get_rank() == 0:
        rvel, rmax = chk.updateVel(N, pos, vel, 0, 0.0095, phi, chi, Bi, 3, params)
        rpos = chk.moveParticles(N, pos, rvel, 0, dtau, phi, chi, Bi, 3, params)
        o = np.argsort(gid)
        check(f"N={N} particle conservation after migration", float(not np.array_equal(gid[o], ids)), 0, failures)
        check(f"N={N} kick velocities", common.rel_linf(gvel[o], rvel), 1e-12, failures)
        check(f"N={N} kick max (parallel.max)", abs(vmax_all - rmax) / rmax, 1e-12, failures)
        check(f"N={N} drift positions", np.abs(gpos[o] - rpos).max(), 1e-14, failures)
        ok, flipped = common.counts_bit_exact(N, gpos[o], rpos, np.concatenate(cnts), chk.cell_index)
        check(f"N={N} per-cell counts after migration (bit-exact; {flipped} particles on a cell face)", float(not ok), 0, failures)
        print(f"  (slab population changes summed over ranks: {crossed})", flush=True)
    for f in (fphi, fchi, fB, T00, T0i, Tij):
        f.close()
    p.close()


def case_time_loop(ctx, chk, N, vector_flag, fused, failures, nsteps=3):
    rng = np.random.default_rng(100 + vector_flag)
    cosmo, ds = common.shipped_cosmology(), common.shipped_settings()
    a0 = 1.0 / (1.0 + ds[3])
    ids, pos, vel = common.quasi_uniform_particles(rng, N, sigma=0.2, a=a0, qscale=3e-3)
    phi, chi, Bi = common.metric_fields(rng, N, a0)
    mass = (cosmo[0] + cosmo[1]) / len(ids)
    gs = gevb.Sim(ctx, 1, vector_flag, ds, cosmo)
    gs.set_fused(fused)
    gs.set_particles(0, ids, pos, vel, mass)               # every rank offers all particles, the library keeps its slab
    gs.set_field("phi", slab(phi, ctx)); gs.set_field("chi", slab(chi, ctx)); gs.set_field("Bi", slab(Bi, ctx))
    plan = gevb.PlanFFT(gs.field("Bi"), gs.field("BiFT"))
    plan.execute(gevb.FFT_FORWARD)                          # BiFT is persistent state of the parabolic solver
    rs = None
    if dist.get_rank() == 0:
        rs = chk.sim(N, 1, vector_flag, ds, cosmo)
        rs.set_particles(0, ids, pos, vel, mass)
        rs.set_field("phi", phi); rs.set_field("chi", chi); rs.set_field("Bi", Bi)
        rs.set_field("BiFT", chk.fft_forward(Bi))
    for step in range(nsteps):
        gs.step()
        fields = {name: np.concatenate(gather(gs.get_field(name)), axis=1) for name in ("phi", "chi", "Bi", "source", "Sij")}
        parts = gather(gs.pcls(0).download())
        st = gs.state()
        if dist.get_rank() == 0:
            rs.step()
            tag = f"N={N} vec={vector_flag} fused={fused} step {step + 1}"
            for name, g in fields.items():
                check(f"{tag} {name}", common.rel_linf(g, rs.get_field(name)), TOL * (1 + step), failures)
            gid = np.concatenate([q[0] for q in parts]); gpos = np.concatenate([q[1] for q in parts]); gvel = np.concatenate([q[2] for q in parts])
            rid, rpos, rvel = rs.get_particles(0)
            o, ro = np.argsort(gid), np.argsort(rid)
            check(f"{tag} particle ids", float(not np.array_equal(gid[o], rid[ro])), 0, failures)
            check(f"{tag} positions", np.abs(gpos[o] - rpos[ro]).max(), 1e-13, failures)
            check(f"{tag} momenta", common.rel_linf(gvel[o], rvel[ro]), TOL, failures)
            check(f"{tag} cells (bit-exact binning)", float(np.count_nonzero(np.floor(gpos[o] * N) != np.floor(rpos[ro] * N))), 0, failures)
            r = rs.state()
            check(f"{tag} T00hom", abs(r["T00hom"] - st["T00hom"]) / abs(r["T00hom"]), TOL, failures)
            check(f"{tag} maxvel", abs(r["maxvel"][0] - st["maxvel"][0]) / r["maxvel"][0], TOL, failures)
    plan.close(); gs.close()
    if rs is not None:
        rs.close()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    chk = None
    if rank == 0:
        import oracle
        oracle.build()
        chk = oracle.load_ref() or oracle.load_ora()
        print(f"multi-GPU parity on {world} ranks; checker: {chk.description}", flush=True)
    failures = []
    sizes = [n for n in (32, 24, 64) if n % world == 0 and n // world >= 2]
    for N in sizes:
        box = [gevb.nccl_unique_id() if rank == 0 else None]          # one NCCL communicator (and id) per context
        dist.broadcast_object_list(box, src=0)
        ctx = gevb.Context(N, device=local, rank=rank, nranks=world, nccl_id=box[0])
        for xch in (0, 1):                                            # NCCL all-to-all + local transpose, then the peer-memory push (default)
            gevb.tuning("fft_exchange", xch)
            if rank == 0:
                print(f" fft_exchange = {xch}", flush=True)
            case_fft(ctx, N, failures)
        case_particles(ctx, chk, N, failures)
        if N == sizes[0]:
            for vector_flag, fused in ((0, 1), (1, 0)):
                case_time_loop(ctx, chk, N, vector_flag, fused, failures)
        ctx.close()
    allf = gather(failures)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        bad = [f for fl in allf for f in fl]
        print("MULTIGPU OK" if not bad else f"MULTIGPU FAILED: {bad}", flush=True)
        sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
