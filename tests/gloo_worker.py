"""Host-side logic of the slab decomposition on CPU ranks (torch.distributed, gloo backend, launched by torchrun from
tests/test_gloo_slab.py).  The device kernels cannot run here, so this checks what does not need them:

  * gevb_slab_geometry (the arithmetic gevb_ctx_create uses) tiles z and ky exactly once over the ranks;
  * bench.py's per-rank synthetic data (particles of the own slab, field slabs) is a partition of the global problem;
  * the exchange index arithmetic of the distributed FFT (fft.cu: local 2-D transforms in the layout [c][ky][zl][kx], block
    ky / nkyl goes to rank d, which files it at [c][kyl][kx][kz = src * nzl + zl], then 1-D transforms along kz) restated
    with numpy FFTs and a real all-to-all reproduces the transform of the undecomposed lattice, forward and backward;
  * the neighbour a slab-crossing particle is sent to (geodesic.cu: periodic distance decides) and the offsets at which the
    ranks write their share of a snapshot (host/output.cpp) restated on the ranks' own counts.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gevolution-1.2_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import gevb  # noqa: E402


def gather(obj):
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def alltoall(send):
    """send[d] -> rank d; returns recv[s] from rank s (equal-sized complex blocks)"""
    s = torch.from_numpy(np.ascontiguousarray(np.stack(send)).view(np.float64).copy())
    r = torch.empty_like(s)
    dist.all_to_all_single(r, s)
    return r.numpy().view(np.complex128).reshape(np.stack(send).shape)


def main():
    dist.init_process_group("gloo")
    rank, P = dist.get_rank(), dist.get_world_size()
    bad = []
    N, nc = 16, 2
    nh = N // 2 + 1
    z0, nzl, ky0, nkyl = gevb.slab_geometry(N, rank, P)
    geo = gather((z0, nzl, ky0, nkyl))
    if sorted(z for g in geo for z in range(g[0], g[0] + g[1])) != list(range(N)) or sorted(k for g in geo for k in range(g[2], g[2] + g[3])) != list(range(N)):
        bad.append("slabs do not tile the lattice")

    # ---- bench.py's synthetic data is a partition
    ids, pos, vel = bench.local_particles(N, z0, nzl, 0.01, 42)
    cz = np.floor(pos[:, 2] * N).astype(int)
    if not np.all((cz >= z0) & (cz < z0 + nzl)):
        bad.append("bench particles outside the own slab")
    all_ids = np.concatenate(gather(ids))
    if not np.array_equal(np.sort(all_ids), np.arange(N ** 3)):
        bad.append("bench particle ids are not a partition")
    f_loc = bench.analytic_field(N, z0, nzl, 1.0, 5, ncomp=nc)
    f_all = np.concatenate(gather(f_loc), axis=1)
    if f_all.shape != (nc, N, N, N) or not np.all(np.isfinite(f_all)):                  # (each slab is normalised to the requested rms by itself)
        bad.append("bench field slabs do not assemble to the lattice")

    # ---- distributed FFT: layouts and exchange arithmetic of fft.cu, restated
    rng = np.random.default_rng(7)
    f = rng.standard_normal((nc, N, N, N))                                              # the same global field on every rank
    loc = f[:, z0:z0 + nzl]                                                             # [c][zl][y][x]
    A = np.fft.rfft2(loc, axes=(2, 3)).transpose(0, 2, 1, 3)                            # 2-D per plane -> [c][ky][zl][kx]
    recv = alltoall([A[:, d * nkyl:(d + 1) * nkyl] for d in range(P)])                  # [src][c][kyl][zl][kx]
    X = np.empty((nc, nkyl, nh, N), dtype=np.complex128)                                # [c][kyl][kx][kz]
    for s in range(P):
        X[:, :, :, s * nzl:(s + 1) * nzl] = recv[s].transpose(0, 1, 3, 2)
    K = np.fft.fft(X, axis=3)                                                           # 1-D along kz
    ref = np.fft.rfftn(f, axes=(1, 2, 3)).transpose(0, 2, 3, 1)[:, ky0:ky0 + nkyl]      # [c][kz][ky][kx] -> [c][ky][kx][kz], own rows
    if np.abs(K - ref).max() > 1e-10 * np.abs(ref).max():
        bad.append("forward slab FFT layout")
    # backward: 1-D inverse along z, block z / nzl goes to rank d as [c][ky][zl][kx], 2-D c2r per plane, unnormalised
    Ainv = np.fft.ifft(K, axis=3) * N                                                   # [c][kyl][kx][z]
    recv = alltoall([Ainv[:, :, :, d * nzl:(d + 1) * nzl].transpose(0, 1, 3, 2) for d in range(P)])   # [src][c][kyl][zl][kx]
    Xb = np.empty((nc, N, nzl, nh), dtype=np.complex128)                                # [c][ky][zl][kx]
    for s in range(P):
        Xb[:, s * nkyl:(s + 1) * nkyl] = recv[s]
    back = np.fft.irfft2(Xb.transpose(0, 2, 1, 3), s=(N, N), axes=(2, 3)) * N * N       # [c][zl][y][x]
    if np.abs(back / N ** 3 - loc).max() > 1e-12:
        bad.append("backward slab FFT layout")

    # ---- migration target of a slab-crossing particle: periodic distance decides (geodesic.cu); the neighbour chosen must be
    #      the rank that owns the particle's new cell (moves are limited to the adjacent slab, main.cpp:281-286)
    for zl in (-2, -1, nzl, nzl + 1):                                                   # local plane index outside [0, nzl)
        cz = (z0 + zl) % N                                                              # global cell the particle is filed under
        d = (cz - z0 + N) % N
        dest = (rank + 1) % P if d < N // 2 else (rank - 1) % P
        if dest != cz // nzl:
            bad.append(f"migration target for zl={zl}")
    # ---- snapshot offsets: exclusive prefix of the ranks' counts (host/output.cpp)
    counts = gather(int(len(ids) // (rank + 2)))
    before, total = sum(counts[:rank]), sum(counts)
    t = torch.tensor([float(counts[r] if r == rank else 0) for r in range(P)], dtype=torch.float64)
    dist.all_reduce(t)                                                                  # what gevb_parallel_sum does on the device
    if [int(v) for v in t.tolist()] != counts or before + counts[rank] > total:
        bad.append("snapshot offsets")

    allbad = gather(bad)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        flat = [b for bl in allbad for b in bl]
        print("GLOO SLAB OK" if not flat else f"GLOO SLAB FAILED: {flat}", flush=True)
        sys.exit(1 if flat else 0)


if __name__ == "__main__":
    main()
