"""GPU parity tests (-m gpu): every hot-path entry point of libgevb.so, called through
the C ABI, against the CPU checker (compiled reference when oracle/_ref is present,
else the C restatement) and against the committed golden vectors.

Tolerances (BASELINE.json north_star): integer / index work bit-exact; FP64 fields
relative L-infinity <= 1e-10 (atomic summation order, FMA contraction).
"""
import os

import numpy as np
import pytest

import common
import golden_cases

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "hotpath_N8.npz")
FIELD_TOL = 1e-10
PCL_TOL = 1e-12


def _check(out, ref_get, names):
    bad = []
    for k in names:
        a, b = np.asarray(out[k]), np.asarray(ref_get(k))
        if k in golden_cases.EXACT:
            if not np.array_equal(a.astype(np.int64), b.astype(np.int64)):
                bad.append((k, "integer mismatch"))
            continue
        tol = PCL_TOL if k.startswith(("kick", "drift")) else FIELD_TOL
        if k in ("pk_kscatter", "pk_pscatter"):
            tol = 1e-5      # variances by cancellation of large sums: summation-order sensitive (north_star: spectra to 1e-5)
        err = common.rel_linf(a, b)
        if not err <= tol:
            bad.append((k, float(err)))
    return bad


def test_golden_vectors(gevb, ctx):
    gold = np.load(GOLDEN)
    c = ctx(8)
    out = golden_cases.run_gpu(gevb, c, golden_cases.inputs(N=8))
    assert _check(out, lambda k: gold[k], gold.files) == []
    # fused variants equal the separate calls to round-off
    assert common.rel_linf(out["T00_fused"], gold["T00"]) <= FIELD_TOL
    assert common.rel_linf(out["Tij_fused"], gold["Tij"]) <= FIELD_TOL


@pytest.mark.parametrize("N,seed", [(16, 5), (24, 6), (32, 7)])
def test_against_checker(gevb, ctx, checker, N, seed):
    inp = golden_cases.inputs(N=N, seed=seed)
    ref_out = golden_cases.run_cpu(checker, inp)
    out = golden_cases.run_gpu(gevb, ctx(N), inp)
    assert _check(out, lambda k: ref_out[k], ref_out.keys()) == []
    assert common.rel_linf(out["ftscalar_fusedB"], out["ftscalar"]) <= 1e-14 and common.rel_linf(out["evolve_fusedchi"], out["evolve"]) <= 1e-14
    # fused kick+drift == updateVel followed by moveParticles
    vel, vmax = checker.updateVel(N, inp["pos"], inp["vel"], 0, inp["dtau_kick"], inp["phi"], inp["chi"], inp["Bi"], 3, inp["params"])
    pos = checker.moveParticles(N, inp["pos"], vel, 0, inp["dtau"], inp["phi"], inp["chi"], inp["Bi"], 3, inp["params"])
    assert common.rel_linf(out["fused_vel"], vel) <= PCL_TOL
    assert np.abs(out["fused_pos"] - pos).max() <= 1e-14
    assert abs(out["fused_max"][0] - vmax) <= 1e-12 * vmax
    # bit-exact re-binning after the drift
    ok, flipped = common.counts_bit_exact(N, out["drift_nf3"], ref_out["drift_nf3"], out["drift_nf3_counts"], checker.cell_index)
    assert ok, f"per-cell counts differ from the reference's ({flipped} particles on a cell face)"


def test_empty_and_single_particle(gevb, ctx, checker):
    N = 8
    c = ctx(N)
    p = gevb.Particles(c, 1.0)
    T = gevb.Field(c, gevb.REAL, 1)
    p.projection_T00_project(T, 0.1, None, 1.0)
    assert p.count() == 0 and np.all(T.download() == 0.0) and p.cell_counts().sum() == 0
    pos = np.array([[np.nextafter(1.0, 0), np.nextafter(1.0, 0), np.nextafter(1.0, 0)]])
    vel = np.array([[0.01, -0.02, 0.03]])
    p.add(np.array([7]), pos, vel)
    T.projection_init(); p.projection_T00_project(T, 0.1, None, 1.0); T.projection_comm()
    assert common.rel_linf(T.download(), checker.projection_T00(N, pos, vel, 1.0, 0.1, None)) <= FIELD_TOL
    assert p.cell_counts()[-1] == 1


def test_error_behaviour(gevb, ctx):
    """wrong field shapes / unknown callbacks are reported, never executed"""
    c = ctx(8)
    p = gevb.Particles(c, 1.0)
    f3 = gevb.Field(c, gevb.REAL, 3)
    with pytest.raises(gevb.GevbError):
        p.projection_T00_project(f3, 0.1, None, 1.0)
    with pytest.raises(gevb.GevbError):
        p.updateVel(5, 0.1, [f3], 1, [1.0, 1.0])
    k1 = gevb.Field(c, gevb.CPLX, 1)
    with pytest.raises(gevb.GevbError):
        gevb.projectFTscalar(k1, k1)


def _cell_index_numpy(N, pos):
    c = np.minimum(np.floor(pos * N).astype(np.int64), N - 1)
    lin = (c[:, 2] * N + c[:, 1]) * N + c[:, 0]
    return lin, np.bincount(lin, minlength=N ** 3).astype(np.uint32)


def _make_sims(gevb, c, checker, N, seed, vector_flag=0, gr=1, baryons=False):
    rng = np.random.default_rng(seed)
    cosmo, ds = common.shipped_cosmology(), common.shipped_settings()
    a0 = 1.0 / (1.0 + ds[3])
    ids, pos, vel = common.quasi_uniform_particles(rng, N, sigma=0.2, a=a0, qscale=3e-3)
    phi, chi, Bi = common.metric_fields(rng, N, a0)
    rs, gs = checker.sim(N, gr, vector_flag, ds, cosmo), gevb.Sim(c, gr, vector_flag, ds, cosmo)
    mass = (cosmo[0] + cosmo[1]) / len(ids)
    if baryons:
        half = len(ids) // 2
        for s in (rs, gs):
            s.set_particles(0, ids[:half], pos[:half], vel[:half], mass)
            s.set_particles(1, ids[half:], pos[half:], vel[half:], mass)
    else:
        for s in (rs, gs):
            s.set_particles(0, ids, pos, vel, mass)
    for s in (rs, gs):
        s.set_field("phi", phi)
        s.set_field("chi", chi)
        s.set_field("Bi", Bi)
    BiFT = checker.fft_forward(Bi)
    rs.set_field("BiFT", BiFT)
    gs.set_field("BiFT", BiFT)
    return rs, gs


def _compare_sims(rs, gs, N, nspecies=1):
    errs = {}
    for name in ("phi", "chi", "Bi", "source", "Sij"):
        errs[name] = common.rel_linf(gs.get_field(name), rs.get_field(name))
    for name in ("scalarFT", "BiFT", "SijFT"):
        errs[name] = common.rel_linf(gs.get_field(name), rs.get_field(name))
    for sp in range(nspecies):
        rid, rpos, rvel = rs.get_particles(sp)
        gid, gpos, gvel = gs.pcls(sp).download()
        ro, go = np.argsort(rid), np.argsort(gid)
        assert np.array_equal(rid[ro], gid[go])
        errs[f"pos{sp}"] = np.abs(gpos[go] - rpos[ro]).max()
        errs[f"vel{sp}"] = common.rel_linf(gvel[go], rvel[ro])
        # bit-exact binning: same particles per cell
        rc = np.floor(rpos[ro] * N).astype(np.int64)
        gc = np.floor(gpos[go] * N).astype(np.int64)
        errs[f"cells{sp}"] = int(np.count_nonzero(rc != gc))
        # ... and the container's own per-cell counts are the histogram of floor(pos/dx) (always compared)
        ok, _ = common.counts_bit_exact(N, gpos[go], rpos[ro], gs.pcls(sp).cell_counts(), _cell_index_numpy)
        errs[f"cells_counts{sp}"] = 0 if ok else 1
    r, g = rs.state(), gs.state()
    for k in ("a", "dtau", "dtau_old", "T00hom"):
        errs["state_" + k] = abs(r[k] - g[k]) / abs(r[k]) if r[k] != 0 else abs(g[k])
    errs["state_tau"] = abs(r["tau"] - g["tau"]) / abs(r["tau"])
    errs["maxvel"] = abs(r["maxvel"][0] - g["maxvel"][0]) / max(abs(r["maxvel"][0]), 1e-300)
    return errs


@pytest.mark.parametrize("fused", [1, 0])
@pytest.mark.parametrize("vector_flag", [0, 1])
def test_time_loop_one_and_more_steps(gevb, ctx, ref, fused, vector_flag):
    """north_star: FP64 fields after one step rel L-inf <= 1e-10, binning bit-exact (same ICs both sides)"""
    N = 16
    rs, gs = _make_sims(gevb, ctx(N), ref, N, seed=100 + vector_flag, vector_flag=vector_flag)
    gs.set_fused(fused)
    for step in range(3):
        rs.step(); gs.step()
        e = _compare_sims(rs, gs, N)
        tol = FIELD_TOL * (1 + step)
        skip = ("state_tau",) + (("scalarFT",) if fused else ())     # fused mode: scalarFT is scratch (its backward FFT may clobber it)
        bad = {k: v for k, v in e.items() if k not in skip and ((k.startswith("cells") and v != 0) or (not k.startswith("cells") and not v <= tol))}
        assert bad == {}, (step, e)
        assert e["state_tau"] < 1e-6    # tau offset comes from the reference's 1e-7 quadrature (bookkeeping only)
    rs.close(); gs.close()


def test_time_loop_N128_many_super_bricks(gevb, ctx, ref):
    """One and two cycles at N = 128 against the compiled reference: 2 x 4 x 4 super-bricks of 64 x 32 x 32 cells, so the
    super-brick index arithmetic in x (sx_shift), k_decode at scale and the multi-brick flush paths are compared with the
    oracle, not just with invariants."""
    N = 128
    rs, gs = _make_sims(gevb, ctx(N), ref, N, seed=128)
    for step in range(2):
        rs.step(); gs.step()
        e = _compare_sims(rs, gs, N)
        skip = ("state_tau", "scalarFT")
        bad = {k: v for k, v in e.items() if k not in skip and ((k.startswith("cells") and v != 0) or (not k.startswith("cells") and not v <= FIELD_TOL * (1 + step)))}
        assert bad == {}, (step, e)
    rs.close(); gs.close()


def test_time_loop_newton_and_baryons(gevb, ctx, ref):
    N = 16
    for gr, bar in ((0, False), (1, True)):
        rs, gs = _make_sims(gevb, ctx(N), ref, N, seed=300 + gr, gr=gr, baryons=bar)
        for _ in range(2):
            rs.step(); gs.step()
        e = _compare_sims(rs, gs, N, nspecies=2 if bar else 1)
        skip = ("state_tau", "scalarFT") + (("chi", "Bi", "Sij", "BiFT", "SijFT") if gr == 0 else ())
        bad = {k: v for k, v in e.items() if k not in skip and ((k.startswith("cells") and v != 0) or (not k.startswith("cells") and not v <= 2 * FIELD_TOL))}
        assert bad == {}, e
        rs.close(); gs.close()


@pytest.mark.parametrize("fused", [1, 0])
def test_time_loop_ncdm_subcycling(gevb, ctx, ref, fused):
    """BASELINE config 4 in small: cdm + baryons + 2 massive-neutrino species whose updates are sub-cycled
    (main.cpp:696-765); neutrino T00 switches on after the first cycle (z_switch_deltancdm), before that the
    homogeneous bg_ncdm stand-in is added (main.cpp:392-397)"""
    N = 16
    rng = np.random.default_rng(77)
    cosmo, ds = common.shipped_cosmology(), common.shipped_settings()
    cosmo, m, T, Om = common.ncdm_model(cosmo)
    a0 = 1.0 / (1.0 + ds[3])
    rs, gs = ref.sim(N, 1, 0, ds, cosmo), gevb.Sim(ctx(N), 1, 0, ds, cosmo)
    gs.set_fused(fused)
    z1 = 1.0 / rs.state()["a"] - 1.0
    for s in (rs, gs):
        # species 0 switches its T00 on only after the first cycle; species 1 from the start
        s.set_ncdm(m, T, Om, [z1 - 0.5, 1e4], [1e4, 1e4], 1e4, 0.05)
        s.set_ncdm_maxvel([0.07, 0.05])
    ids, pos, vel = common.quasi_uniform_particles(rng, N, sigma=0.2, a=a0, qscale=3e-3)
    half = len(ids) // 2
    parts = [(ids[:half], pos[:half], vel[:half], cosmo[0] / half), (ids[half:], pos[half:], vel[half:], cosmo[1] / (len(ids) - half))]
    for k in range(2):
        i2, p2, v2 = common.thermal_particles(rng, N, N // 2, a0, 0.06 * a0 * (1 + k))
        parts.append((i2, p2, v2, Om[k] / len(i2)))
    phi, chi, Bi = common.metric_fields(rng, N, a0)
    for s in (rs, gs):
        for sp, (i_, p_, v_, mass) in enumerate(parts):
            s.set_particles(sp, i_, p_, v_, mass)
        s.set_field("phi", phi); s.set_field("chi", chi); s.set_field("Bi", Bi)
    BiFT = ref.fft_forward(Bi)
    rs.set_field("BiFT", BiFT); gs.set_field("BiFT", BiFT)
    assert abs(rs.state()["dtau"] - gs.state()["dtau"]) <= 1e-13 * rs.state()["dtau"]
    saw_subcycling = False
    for step in range(3):
        rs.step(); gs.step()
        e = _compare_sims(rs, gs, N, nspecies=4)
        (rv, rn), (gv, gn) = rs.ncdm_state(), gs.ncdm_state()
        assert np.array_equal(rn, gn), (rn, gn)
        assert np.abs(rv - gv).max() <= 1e-12
        saw_subcycling |= bool(rn.max() > 1)
        skip = ("state_tau",) + (("scalarFT",) if fused else ())
        tol = FIELD_TOL * (1 + step)
        bad = {k: v for k, v in e.items() if k not in skip and ((k.startswith("cells") and v != 0) or (not k.startswith("cells") and not v <= tol))}
        assert bad == {}, (step, e)
    assert saw_subcycling
    rs.close(); gs.close()


def test_run_to_z0_power_spectra_and_snapshot_statistics(gevb, ctx, ref):
    """north_star: phi, chi and B power spectra and the particle snapshot statistics at z = 0 within 1e-5 of the
    reference CPU run from identical initial conditions (N = 32, z = 100 -> 0: 117 cycles, structure forms --
    at the end 70 % of the cells are empty and the densest holds ~50 particles)"""
    N, numbins = 32, 16
    rs, gs = _make_sims(gevb, ctx(N), ref, N, seed=5)
    ids0, pos0, _ = rs.get_particles(0)
    pos0 = pos0[np.argsort(ids0)]
    ncycles = 0
    while rs.state()["a"] < 1.0 and ncycles < 400:
        rs.step(); gs.step(); ncycles += 1
    r, g = rs.state(), gs.state()
    assert ncycles > 50 and g["cycle"] == r["cycle"] and abs(g["a"] - r["a"]) <= 1e-12 * r["a"]
    assert abs(g["T00hom"] - r["T00hom"]) <= 1e-9 * r["T00hom"]
    # spectra the way writeSpectra takes them (main.cpp:639-679, tools.hpp:53-212): forward FFT of the field, binned |.|^2
    c = gs.ctx
    worst = {}
    for name, ncomp in (("phi", 1), ("chi", 1), ("Bi", 3)):
        fr = rs.get_field(name)
        kb, pw, ks, ps, occ = ref.extractPowerSpectrum(ref.fft_forward(fr), numbins)
        F, K = gevb.Field(c, gevb.REAL, ncomp, data=gs.get_field(name)), gevb.Field(c, gevb.CPLX, ncomp)
        plan = gevb.PlanFFT(F, K)
        plan.execute(gevb.FFT_FORWARD)
        gkb, gpw, gks, gps, gocc = gevb.extractPowerSpectrum(K, numbins)
        assert np.array_equal(occ, gocc)
        m = occ > 0
        worst[name] = float(np.max(np.abs(gpw[m] - pw[m]) / np.abs(pw[m])))
        assert np.max(np.abs(gkb[m] - kb[m]) / kb[m]) <= 1e-12
        plan.close(); F.close(); K.close()
    assert all(v <= 1e-5 for v in worst.values()), worst
    # snapshot statistics: displacement and momentum moments, extremes, and the occupancy histogram
    rid, rpos, rvel = rs.get_particles(0)
    gid, gpos, gvel = gs.pcls(0).download()
    ro, go = np.argsort(rid), np.argsort(gid)
    assert np.array_equal(rid[ro], gid[go])
    rpos, rvel, gpos, gvel = rpos[ro], rvel[ro], gpos[go], gvel[go]
    def stats(pos, vel):
        d = pos - pos0
        d -= np.round(d)
        cells = np.minimum(np.floor(pos * N).astype(np.int64), N - 1)
        occ = np.bincount((cells[:, 2] * N + cells[:, 1]) * N + cells[:, 0], minlength=N ** 3)
        return np.array([np.sqrt((d ** 2).mean()), np.abs(d).max(), np.sqrt((vel ** 2).mean()), np.abs(vel).max(),
                         (vel ** 2).sum(axis=1).max(), float((occ == 0).mean()), float(occ.max()), float((occ.astype(np.float64) ** 2).mean())])
    sr, sg = stats(rpos, rvel), stats(gpos, gvel)
    assert np.max(np.abs(sg - sr) / np.abs(sr)) <= 1e-5, (sr, sg)
    assert np.abs(gpos - rpos).max() <= 1e-7 and common.rel_linf(gvel, rvel) <= 1e-5, (np.abs(gpos - rpos).max(), common.rel_linf(gvel, rvel))
    print("z=0 parity: cycles", ncycles, "spectra rel err", worst, "max |dx|", np.abs(gpos - rpos).max())
    rs.close(); gs.close()


# ---- size-independent properties at the benchmark size --------------------
@pytest.mark.parametrize("N", [128, 256])          # 256^3 grid / 256^3 particles is BASELINE config 2
def test_full_size_properties(gevb, ctx, N):
    rng = np.random.default_rng(9)
    c = ctx(N)
    a = 0.02
    ids, pos, vel = common.quasi_uniform_particles(rng, N, sigma=0.3, a=a)
    p = gevb.Particles(c, 0.3 / len(ids)).add(ids, pos, vel)
    # particle conservation + bit-exact counts against numpy binning
    counts = p.cell_counts()
    cell = np.minimum(np.floor(pos / (1.0 / N)).astype(np.int64), N - 1)
    key = (cell[:, 2] * N + cell[:, 1]) * N + cell[:, 0]
    assert np.array_equal(counts, np.bincount(key, minlength=N ** 3).astype(np.uint32))
    # mass conservation with phi = NULL: mean T00 = Omega
    T00 = gevb.Field(c, gevb.REAL, 1)
    p.projection_T00_project(T00, a, None, 1.0); T00.projection_comm()
    assert abs(T00.sum() / N ** 3 - 0.3) < 1e-12
    # FFT round trip (unnormalised both ways) and preserved Fourier input
    f = common.gaussian_field(rng, N, 3, 1.0)
    F3, K3 = gevb.Field(c, gevb.REAL, 3, data=f), gevb.Field(c, gevb.CPLX, 3)
    plan = gevb.PlanFFT(F3, K3)
    plan.execute(gevb.FFT_FORWARD); plan.execute(gevb.FFT_BACKWARD)
    assert common.rel_linf(F3.download() / N ** 3, f) < 1e-13
    # projectFTvector output is divergence-free (backward differences), tools.hpp:375
    gevb.projectFTvector(K3, K3, 1.0, 0.0)
    plan.execute(gevb.FFT_BACKWARD)
    B = F3.download()
    div = sum(B[i] - np.roll(B[i], 1, axis=2 - i) for i in range(3))
    assert np.abs(div).max() < 1e-11 * np.abs(B).max()
    # drift with zero momenta is the identity and keeps the sort (idempotence)
    before = p.download()
    p0 = gevb.Particles(c, 1.0).add(ids, pos, np.zeros_like(vel))
    p0.moveParticles(gevb.UPDATE_Q_NEWTON, 0.1, None, 0, [a, 1.0])
    i0, x0, v0 = p0.download()
    o0, ob = np.argsort(i0), np.argsort(before[0])
    assert np.array_equal(i0[o0], before[0][ob]) and np.array_equal(x0[o0], before[1][ob])
    k0 = np.minimum(np.floor(x0 * N).astype(np.int64), N - 1)
    key = gevb.storage_key(N, N, k0[:, 0], k0[:, 1], k0[:, 2])      # storage order: bricks, then cells inside the brick
    assert np.all(np.diff(key) >= 0), "brick-major cell order lost"
    assert np.array_equal(p0.cell_counts(), counts)


# ---- the own x-pass of the forward transform (N = 512 only) with the source preparation fused into its load --------------------
def test_own_xpass_512(gevb, ctx):
    """csrc/xpass.cu against numpy (the transform itself, a few planes and lines of the 512^3 spectrum) and against the
    cuFFT path with the separate prepareFTsource kernels (both oracle-checked at small N): same spectrum to 1e-12 of its
    largest coefficient, same T00hom sum."""
    N = 512
    rng = np.random.default_rng(21)
    c = ctx(N)
    f = rng.standard_normal((N, N, N))
    R, K = gevb.Field(c, gevb.REAL, 1, data=f[None]), gevb.Field(c, gevb.CPLX, 1)
    plan = gevb.PlanFFT(R, K)
    gevb.tuning("fft_xpass", 1)
    plan.execute(gevb.FFT_FORWARD)
    cplx = lambda a: a[..., 0] + 1j * a[..., 1] if a.shape[-1] == 2 and not np.iscomplexobj(a) else a
    own = cplx(K.download()[0])                                # [kz][ky][kx]
    gevb.tuning("fft_xpass", 0)
    plan.execute(gevb.FFT_FORWARD)
    lib = cplx(K.download()[0])
    scale = np.abs(lib).max()
    assert np.abs(own - lib).max() < 1e-12 * scale
    # numpy on the part it can do quickly: the full transform of three x-lines' worth of modes through separable passes
    fx = np.fft.rfft(f[:, :, :], axis=2)[:, :, [0, 1, 37, 255, 256]]      # x-pass (all rows), five kx columns
    ref = np.fft.fft(np.fft.fft(fx, axis=1), axis=0)
    assert np.abs(own[:, :, [0, 1, 37, 255, 256]] - ref).max() < 1e-12 * scale
    gevb.tuning("fft_xpass", 0)
    plan.close(); R.close(); K.close()
    del own, lib, fx, ref

    # fused scalar and tensor preparation against the separate kernels + cuFFT
    phi = gevb.Field(c, gevb.REAL, 1, data=1e-5 * rng.standard_normal((1, N, N, N))); phi.updateHalo()
    chi = gevb.Field(c, gevb.REAL, 1, data=1e-7 * rng.standard_normal((1, N, N, N))); chi.updateHalo()
    src0 = 1.0 + 0.1 * rng.standard_normal((1, N, N, N))
    coeffs = (0.3, 1.7e-3, 2.2e-3, 4.1e-3)
    out = {}
    for knob in (1, 0):
        gevb.tuning("fft_xpass", knob)
        S, SF = gevb.Field(c, gevb.REAL, 1, data=src0), gevb.Field(c, gevb.CPLX, 1)
        pl = gevb.PlanFFT(S, SF)
        total = gevb.prepareFTsource_scalar_fft(phi, chi, pl, *coeffs, want_sum=True)
        out[knob] = (cplx(SF.download()[0]), total)
        pl.close(); S.close(); SF.close()
    scale = np.abs(out[0][0]).max()
    assert np.abs(out[1][0] - out[0][0]).max() < 1e-12 * scale
    assert abs(out[1][1] - out[0][1]) < 1e-12 * abs(out[0][1]) and abs(out[0][1] - src0.sum()) < 1e-10 * abs(src0.sum())
    del out, src0
    T0 = 1e-3 * rng.standard_normal((6, N, N, N))
    res = {}
    for knob in (1, 0):
        gevb.tuning("fft_xpass", knob)
        T, TF = gevb.Field(c, gevb.REAL, 6, data=T0), gevb.Field(c, gevb.CPLX, 6)
        pl = gevb.PlanFFT(T, TF)
        gevb.prepareFTsource_tensor_fft(phi, pl, 0.37)
        res[knob] = cplx(TF.download())
        pl.close(); T.close(); TF.close()
    gevb.tuning("fft_xpass", 0)
    for k in range(6):
        scale = np.abs(res[0][k]).max()
        assert np.abs(res[1][k] - res[0][k]).max() < 1e-12 * scale, f"component {k}"


# ---- outputs the parity metrics are defined on (f1 spectra files, f3 Gadget-2 snapshot) --------------------
def _read_gadget2(path):
    """minimal Gadget-2 reader (format 1, int64 IDs): header dict, pos[n][3] float32, vel[n][3] float32, ids[n]"""
    import struct
    raw = open(path, "rb").read()
    assert struct.unpack_from("<I", raw, 0)[0] == 256 and struct.unpack_from("<I", raw, 260)[0] == 256
    npart = struct.unpack_from("<6I", raw, 4)
    mass = struct.unpack_from("<6d", raw, 4 + 24)
    time, redshift = struct.unpack_from("<2d", raw, 4 + 24 + 48)
    box = struct.unpack_from("<d", raw, 4 + 24 + 48 + 16 + 8 + 24 + 8)[0]
    n = npart[1]
    o = 264
    assert struct.unpack_from("<I", raw, o)[0] == 12 * n
    pos = np.frombuffer(raw, dtype="<f4", count=3 * n, offset=o + 4).reshape(n, 3)
    o += 4 + 12 * n
    assert struct.unpack_from("<2I", raw, o) == (12 * n, 12 * n)
    vel = np.frombuffer(raw, dtype="<f4", count=3 * n, offset=o + 8).reshape(n, 3)
    o += 8 + 12 * n
    assert struct.unpack_from("<2I", raw, o) == (12 * n, 8 * n)
    ids = np.frombuffer(raw, dtype="<i8", count=n, offset=o + 8)
    o += 8 + 8 * n
    assert struct.unpack_from("<I", raw, o)[0] == 8 * n and len(raw) == o + 4
    return dict(npart=npart, mass=mass, time=time, redshift=redshift, BoxSize=box), pos, vel, ids


def test_spectra_files_and_gadget2_snapshot(gevb, ctx, ref, tmp_path):
    N, numbins = 16, 8
    rs, gs = _make_sims(gevb, ctx(N), ref, N, seed=41)
    for _ in range(2):
        rs.step(); gs.step()
    st = gs.state()
    a = st["a"]
    # ---- f1: spectra files, phi / chi / hij / B (output.hpp:1945-1981,2151-2155; file format tools.hpp:268-346)
    prefix = str(tmp_path / "pk")
    gs.write_spectra(prefix, 3, numbins, 1 | 2 | 8 | 128)
    cosmo = common.shipped_cosmology()
    fourpiG, box = st["fourpiG"], 320.0
    phi, chi = rs.get_field("phi"), rs.get_field("chi")
    # the reference side of the hij branch: Tij of the particles, source with the snapshot coefficient, TT projection
    rid, rpos, rvel = rs.get_particles(0)
    mass = (cosmo[0] + cosmo[1]) / len(rid)
    Tij = ref.projection_Tij(N, rpos, rvel, mass, a, phi[0])
    hijFT = ref.projectFTtensor(ref.fft_forward(ref.prepareFTsource_tensor(phi[0], Tij, 2. * fourpiG / N / N / a)))
    n3 = float(N) ** 3
    expect = {
        "phi": (ref.extractPowerSpectrum(ref.fft_forward(phi), numbins, deconvolve=False), n3 * n3 * 2 * np.pi ** 2),
        "chi": (ref.extractPowerSpectrum(ref.fft_forward(chi), numbins, deconvolve=False), n3 * n3 * 2 * np.pi ** 2),
        "hij": (ref.extractPowerSpectrum(hijFT, numbins, symmetric=True, deconvolve=False), 2 * np.pi ** 2),
        "B": (ref.extractPowerSpectrum(rs.get_field("BiFT"), numbins, deconvolve=False), a ** 4 * N * N * 2 * np.pi ** 2),
    }
    for tag, ((kb, pw, ks, ps, occ), rescalep) in expect.items():
        lines = open(f"{prefix}003_{tag}.dat").read().splitlines()
        assert lines[0] == f"# power spectrum of {tag}" and lines[1] == "# redshift z=%f" % (1. / a - 1.) and lines[2].startswith("# k ")
        rows = np.array([[float(v) for v in ln.split()] for ln in lines[3:]])
        m = occ > 0
        assert len(rows) == m.sum() and np.array_equal(rows[:, 4].astype(int), occ[m])
        assert np.allclose(rows[:, 0], kb[m] / box, rtol=2e-6) and np.allclose(rows[:, 1], pw[m] / rescalep, rtol=2e-6), tag   # %e keeps 7 digits
        assert np.allclose(rows[:, 3], ps[m] / rescalep / np.sqrt(occ[m]), rtol=1e-4, atol=1e-6 * np.abs(rows[:, 1]).max()), tag
    # EXACT_OUTPUT_REDSHIFTS: a second write past the target redshift interpolates with the stored file (tools.hpp:277-315)
    kb, pw, ks, ps, occ = expect["phi"][0]
    fn = str(tmp_path / "interp.dat")
    gevb.writePowerSpectrum(kb, pw, ks, ps, occ, box, 2.0, fn, "test", 1. / (1. + 10.0), 9.5)       # z = 10 > target: stored as is
    gevb.writePowerSpectrum(kb, 3.0 * pw, ks, ps, occ, box, 2.0, fn, "test", 1. / (1. + 9.0), 9.5)  # z = 9 < target 9.5: half-way
    lines = open(fn).read().splitlines()
    assert lines[1] == "# redshift z=%f" % 9.5
    rows = np.array([[float(v) for v in ln.split()] for ln in lines[3:]])
    assert np.allclose(rows[:, 1], 2.0 * pw[occ > 0] / 2.0, rtol=3e-6)
    # ---- f3: Gadget-2 snapshot with the half-step corrections (Particles_gevolution.hpp:153-199)
    fn = str(tmp_path / "snap_cdm")
    dtau_pos, dtau_vel, tracer = 0.013, 0.021, 3
    gs.save_gadget2(0, fn, tracer, dtau_pos, dtau_vel)
    hdr, gpos, gvel, gids = _read_gadget2(fn)
    sel = rid % tracer == 0
    assert hdr["npart"][1] == sel.sum() and abs(hdr["time"] - a) < 1e-15 and hdr["BoxSize"] == box / 0.001
    assert abs(hdr["mass"][1] - tracer * 2.77459457e11 * mass * box ** 3 / 1e10) <= 1e-12 * hdr["mass"][1]
    o, ro = np.argsort(gids), np.argsort(rid[sel])
    assert np.array_equal(gids[o], rid[sel][ro])
    p, v = rpos[sel][ro], rvel[sel][ro]
    f = phi[0]                                                           # [z][y][x]
    s = p * N
    c = np.minimum(np.floor(s).astype(int), N - 1)
    r = s - np.floor(s)
    def at(dx, dy, dz):
        return f[(c[:, 2] + dz) % N, (c[:, 1] + dy) % N, (c[:, 0] + dx) % N]
    w = [(1 - r[:, i], r[:, i]) for i in range(3)]
    phip = sum(at(i, j, k) * w[0][i] * w[1][j] * w[2][k] for i in (0, 1) for j in (0, 1) for k in (0, 1))
    grad = np.stack([sum((at(1, j, k) - at(0, j, k)) * w[1][j] * w[2][k] for j in (0, 1) for k in (0, 1)),
                     sum((at(i, 1, k) - at(i, 0, k)) * w[0][i] * w[2][k] for i in (0, 1) for k in (0, 1)),
                     sum((at(i, j, 1) - at(i, j, 0)) * w[0][i] * w[1][j] for i in (0, 1) for j in (0, 1))], axis=1)
    v2 = (v ** 2).sum(axis=1)
    e2 = v2 + a * a
    e = np.sqrt(e2)
    ssum = v2 + e2
    corr = 1. + (4. - ssum / e2) * phip
    xpos = np.modf(1. + p + dtau_pos * v * (corr / e)[:, None])[0] * hdr["BoxSize"]
    xvel = (v - dtau_vel * (ssum / e)[:, None] * grad * N) / np.sqrt(a) / 3.335640952e-6 / a
    assert np.allclose(gpos[o], xpos.astype(np.float32), rtol=3e-7, atol=hdr["BoxSize"] * 1e-7)
    assert np.allclose(gvel[o], xvel.astype(np.float32), rtol=3e-6, atol=np.abs(xvel).max() * 1e-6)
    rs.close(); gs.close()


def test_hibernate_and_restart(gevb, ctx, ref, tmp_path):
    """a run continued from a hibernation point is compared with the REFERENCE running on from the same state
    (hibernation.hpp:512-611, ic_read.hpp:290-330).  The files hold the reference's state set: particles, phi, chi and the
    real-space vector potential divided by a^2 N; the restart multiplies it by a^2 / N^2 and rebuilds BiFT by a forward FFT."""
    N = 16
    rs, gs = _make_sims(gevb, ctx(N), ref, N, seed=61, baryons=True)
    for _ in range(2):
        rs.step(); gs.step()
    base = str(tmp_path / "hib")
    a_hib = gs.state()["a"]
    gs.hibernate(base)
    # what the reference writes: B / (a^2 N) (hibernation.hpp:533-538), phi, chi
    B_file = gevb.read_raw_field(base + "_B.bin")
    assert common.rel_linf(B_file, rs.get_field("Bi") / (a_hib * a_hib * N)) <= 2 * FIELD_TOL
    assert common.rel_linf(gevb.read_raw_field(base + "_phi.bin"), rs.get_field("phi")) <= 2 * FIELD_TOL
    assert common.rel_linf(gevb.read_raw_field(base + "_chi.bin"), rs.get_field("chi")) <= 2 * FIELD_TOL
    # the reference runs on uninterrupted; the device run is thrown away and continued from the files in a new simulation
    gs.close()
    cosmo, ds = common.shipped_cosmology(), common.shipped_settings()
    g2 = gevb.Sim(ctx(N), 1, 0, ds, cosmo)
    g2.restore(base)
    assert g2.state()["cycle"] == 2
    # BiFT rebuilt by the forward transform of the stored field (ic_read.hpp:305-319) is the reference's persistent BiFT
    assert common.rel_linf(g2.get_field("BiFT"), rs.get_field("BiFT")) <= 2 * FIELD_TOL
    for _ in range(2):
        rs.step(); g2.step()
    e = _compare_sims(rs, g2, N, nspecies=2)
    skip = ("state_tau", "scalarFT")
    bad = {k: v for k, v in e.items() if k not in skip and ((k.startswith("cells") and v != 0) or (not k.startswith("cells") and not v <= 5 * FIELD_TOL))}
    assert bad == {}, e
    assert g2.state()["cycle"] == rs.state()["cycle"] == 4
    # a file written for another lattice is refused; so is a truncated particle file (no rank is left inside a collective)
    g3 = gevb.Sim(ctx(8), 1, 0, ds, cosmo)
    with pytest.raises(gevb.GevbError):
        g3.restore(base)
    with open(base + ".0.gevb", "r+b") as f:
        f.truncate(600)
    g4 = gevb.Sim(ctx(N), 1, 0, ds, cosmo)
    with pytest.raises(gevb.GevbError):
        g4.restore(base)
    rs.close(); g2.close(); g3.close(); g4.close()


def test_field_snapshot_dumps(gevb, ctx, ref, tmp_path):
    """writeSnapshots' field dumps (output.hpp:98-300) against the reference's fields: phi, chi, T00 as projected
    for the output, B divided by a^2 N with the stored field restored afterwards (output.hpp:212-236), hij = TT projection of
    SijFT transformed back (output.hpp:259-265)"""
    N = 16
    rs, gs = _make_sims(gevb, ctx(N), ref, N, seed=62)
    for _ in range(2):
        rs.step(); gs.step()
    a = gs.state()["a"]
    Bi_before = gs.get_field("Bi")
    prefix = str(tmp_path / "snap000")
    gs.write_field_snapshot(prefix, 1 | 2 | 8 | 16 | 128)
    rd = gevb.read_raw_field
    assert common.rel_linf(rd(prefix + "_phi.bin"), rs.get_field("phi")) <= 2 * FIELD_TOL
    assert common.rel_linf(rd(prefix + "_chi.bin"), rs.get_field("chi")) <= 2 * FIELD_TOL
    assert common.rel_linf(rd(prefix + "_B.bin"), rs.get_field("Bi") / (a * a * N)) <= 2 * FIELD_TOL
    # the rescale is undone by the backward transform of BiFT (output.hpp:232-236)
    assert common.rel_linf(gs.get_field("Bi"), Bi_before) <= 1e-12
    # T00 as output.hpp:155-182 projects it: the current particles with the current phi
    ids, pos, vel = rs.get_particles(0)
    cosmo = common.shipped_cosmology()
    mass = (cosmo[0] + cosmo[1]) / len(ids)
    T00 = ref.projection_T00(N, pos, vel, mass, a, rs.get_field("phi")[0], 1.0)
    assert common.rel_linf(rd(prefix + "_T00.bin"), T00) <= 2 * FIELD_TOL
    # hij: TT projection of the reference's SijFT, transformed back (unnormalised c2r)
    hijFT = ref.projectFTtensor(rs.get_field("SijFT"))
    hij = np.stack([ref.fft_backward(hijFT[c:c + 1])[0] for c in range(6)])
    assert common.rel_linf(rd(prefix + "_hij.bin"), hij) <= 5 * FIELD_TOL
    rs.close(); gs.close()


def _cic_gradient(field, pos, N):
    """CIC gradient of a scalar lattice field [z][y][x] at the particle positions, in the reference's operation order
    (ic_basic.hpp:70-85 / 129-144), already divided by the lattice resolution"""
    s = pos * N
    c = np.minimum(np.floor(s).astype(int), N - 1)
    r = s - np.floor(s)
    def at(dx, dy, dz):
        return field[(c[:, 2] + dz) % N, (c[:, 1] + dy) % N, (c[:, 0] + dx) % N]
    g = np.empty_like(pos)
    g[:, 0] = (1 - r[:, 1]) * (1 - r[:, 2]) * (at(1, 0, 0) - at(0, 0, 0)); g[:, 1] = (1 - r[:, 0]) * (1 - r[:, 2]) * (at(0, 1, 0) - at(0, 0, 0)); g[:, 2] = (1 - r[:, 0]) * (1 - r[:, 1]) * (at(0, 0, 1) - at(0, 0, 0))
    g[:, 0] += r[:, 1] * (1 - r[:, 2]) * (at(1, 1, 0) - at(0, 1, 0)); g[:, 1] += r[:, 0] * (1 - r[:, 2]) * (at(1, 1, 0) - at(1, 0, 0)); g[:, 2] += r[:, 0] * (1 - r[:, 1]) * (at(1, 0, 1) - at(1, 0, 0))
    g[:, 0] += (1 - r[:, 1]) * r[:, 2] * (at(1, 0, 1) - at(0, 0, 1)); g[:, 1] += (1 - r[:, 0]) * r[:, 2] * (at(0, 1, 1) - at(0, 0, 1)); g[:, 2] += (1 - r[:, 0]) * r[:, 1] * (at(0, 1, 1) - at(0, 1, 0))
    g[:, 0] += r[:, 1] * r[:, 2] * (at(1, 1, 1) - at(0, 1, 1)); g[:, 1] += r[:, 0] * r[:, 2] * (at(1, 1, 1) - at(1, 0, 1)); g[:, 2] += r[:, 0] * r[:, 1] * (at(1, 1, 1) - at(1, 1, 0))
    return g * N


@pytest.mark.parametrize("nfields", [1, 2])
def test_ic_callbacks(gevb, ctx, nfields):
    """the two IC-generator callbacks of the drop-in boundary: displace_pcls_ic_basic under moveParticles with its
    MAX reduction output, initialize_q_ic_basic under updateVel; with two fields every eighth ID uses the second
    (baryon treatment = hybrid, ic_basic.hpp:65-68,124-127)"""
    N = 16
    rng = np.random.default_rng(90 + nfields)
    c = ctx(N)
    ids, pos, vel = common.quasi_uniform_particles(rng, N, sigma=0.25, a=0.01)
    xi = common.gaussian_field(rng, N, 2, 3e-3 / N)                 # displacements of a fraction of a cell
    F = [gevb.Field(c, gevb.REAL, 1, data=xi[k:k + 1]) for k in range(2)]
    for f in F:
        f.updateHalo()
    which = ((ids % 8 == 0) & (nfields > 1)).astype(int)
    g = np.where(which[:, None] == 1, _cic_gradient(xi[1], pos, N), _cic_gradient(xi[0], pos, N))
    # velocities first (positions unchanged), then the displacement
    p = gevb.Particles(c, 1.0).add(ids, pos, vel)
    coeff_q = 0.7
    vmax = p.updateVel(gevb.INITIALIZE_Q_IC_BASIC, coeff_q, F[:nfields], nfields, [1.0, 1.0])
    gid, gpos, gvel = p.download()
    o = np.argsort(gid)
    assert common.rel_linf(gvel[o], -g * coeff_q) <= PCL_TOL and np.array_equal(gpos[o], pos)
    assert abs(vmax - np.sqrt(((g * coeff_q) ** 2).sum(axis=1).max())) <= 1e-12 * vmax
    dmax = p.moveParticles_max(gevb.DISPLACE_PCLS_IC_BASIC, 1.0, F[:nfields], nfields)
    gid, gpos, gvel = p.download()
    o = np.argsort(gid)
    want = pos + g
    want -= np.floor(want)
    assert np.abs(gpos[o] - want).max() <= 1e-15 and abs(dmax - np.sqrt((g ** 2).sum(axis=1).max())) <= 1e-12 * dmax
    # re-filed under the new cells, bit-exact
    cell = np.minimum(np.floor(gpos[o] * N).astype(np.int64), N - 1)
    assert np.array_equal(p.cell_counts(), np.bincount((cell[:, 2] * N + cell[:, 1]) * N + cell[:, 0], minlength=N ** 3).astype(np.uint32))
    # a callback under the wrong driver is refused
    with pytest.raises(gevb.GevbError):
        p.moveParticles(gevb.INITIALIZE_Q_IC_BASIC, 1.0, F[:1], 1, [1.0, 1.0])
    with pytest.raises(gevb.GevbError):
        p.updateVel(gevb.DISPLACE_PCLS_IC_BASIC, 1.0, F[:1], 1, [1.0, 1.0])
    p.close()
    for f in F:
        f.close()


def test_extreme_occupancy(gevb, ctx, checker):
    """collisions as the domain has them: thousands of particles in one cell (segmented reduction over whole warps, cells
    that span many warps and batches), next to empty bricks and a cell at the periodic corner"""
    N = 16
    rng = np.random.default_rng(123)
    a = 0.05
    blob = (np.array([5.0, 3.0, 9.0]) + rng.random((3000, 3))) / N                  # one cell
    corner = (np.array([15.0, 15.0, 15.0]) + rng.random((700, 3))) / N              # the cell whose CIC cloud wraps in all directions
    few = rng.random((50, 3))
    pos = np.ascontiguousarray(np.concatenate([blob, corner, few]))
    vel = rng.standard_normal(pos.shape) * 0.02 * a
    ids = np.arange(len(pos), dtype=np.int64)
    phi, chi, Bi = common.metric_fields(rng, N, a)
    c = ctx(N)
    P = gevb.Field(c, gevb.REAL, 1, data=phi); P.updateHalo()
    X = gevb.Field(c, gevb.REAL, 1, data=chi); X.updateHalo()
    B = gevb.Field(c, gevb.REAL, 3, data=Bi); B.updateHalo()
    p = gevb.Particles(c, 1e-4).add(ids, pos, vel)
    counts = p.cell_counts()
    assert counts.max() >= 3000 and counts.sum() == len(ids)
    t00, T = gevb.Field(c, gevb.REAL, 1), gevb.Field(c, gevb.REAL, 6)
    p.projection_T00_Tij_project(t00, T, a, P, 1.0); t00.projection_comm(); T.projection_comm()
    assert common.rel_linf(t00.download(), checker.projection_T00(N, pos, vel, 1e-4, a, phi[0])) <= FIELD_TOL
    assert common.rel_linf(T.download(), checker.projection_Tij(N, pos, vel, 1e-4, a, phi[0])) <= FIELD_TOL
    t0i = gevb.Field(c, gevb.REAL, 3)
    p.projection_T0i_project(t0i, P, 1.0); t0i.projection_comm()
    assert common.rel_linf(t0i.download(), checker.projection_T0i(N, pos, vel, 1e-4, phi[0])) <= FIELD_TOL
    params = [a, a * a * N]
    vmax = p.kick_drift(gevb.UPDATE_Q, 0.03, 3, params, 0.05, 3, params, [P, X, B])
    rv, rmax = checker.updateVel(N, pos, vel, 0, 0.03, phi, chi, Bi, 3, params)
    rp = checker.moveParticles(N, pos, rv, 0, 0.05, phi, chi, Bi, 3, params)
    gid, gpos, gvel = p.download()
    o = np.argsort(gid)
    assert common.rel_linf(gvel[o], rv) <= PCL_TOL and np.abs(gpos[o] - rp).max() <= 1e-14 and abs(vmax - rmax) <= 1e-12 * rmax
    cell = np.minimum(np.floor(gpos[o] * N).astype(np.int64), N - 1)
    assert np.array_equal(p.cell_counts(), np.bincount((cell[:, 2] * N + cell[:, 1]) * N + cell[:, 0], minlength=N ** 3).astype(np.uint32))
    for f in (P, X, B, t00, T, t0i):
        f.close()
    p.close()


# ---- BASELINE config 1: the reference's shipped settings.ini from its own seed ------------------------------------------
def _gpu_sim_like(gevb, c, cosmo, ds, flags, mass, ids, pos, vel, phi, chi, Bi, BiFT, state):
    gs = gevb.Sim(c, int(flags[1]), int(flags[2]), ds, cosmo)
    gs.set_particles(0, ids, pos, vel, float(mass[0]))
    gs.set_field("phi", phi); gs.set_field("chi", chi); gs.set_field("Bi", Bi); gs.set_field("BiFT", BiFT)
    gs.set_state(state[0], state[1], state[2], state[3], int(state[4]), (state[5], state[6]))
    return gs


def test_shipped_settings_fixture(gevb, ctx):
    """initial conditions made by the reference's own parser + generateIC_basic from the shipped settings.ini (Ngrid 16,
    tiling 4; tests/golden/make_golden.py) and the reference's state four cycles later: the device path reproduces it"""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "shipped_settings_N16.npz"))
    N = int(g["flags"][0])
    gs = _gpu_sim_like(gevb, ctx(N), g["cosmo"], g["dsettings"], g["flags"], g["mass"], g["ic_ids"], g["ic_pos"], g["ic_vel"],
                       g["ic_phi"], g["ic_chi"], g["ic_Bi"], g["ic_BiFT"], g["ic_state"])
    ncycles = int(g["ncycles"][0])
    for _ in range(ncycles):
        gs.step()
    tol = FIELD_TOL * ncycles
    for name in ("phi", "chi", "Bi"):
        assert common.rel_linf(gs.get_field(name), g["end_" + name]) <= tol, name
    gid, gpos, gvel = gs.pcls(0).download()
    o = np.argsort(gid)
    assert np.array_equal(gid[o], g["end_ids"]) and np.abs(gpos[o] - g["end_pos"]).max() <= 1e-13 and common.rel_linf(gvel[o], g["end_vel"]) <= tol
    assert np.array_equal(np.floor(gpos[o] * N), np.floor(g["end_pos"] * N))                       # bit-exact binning
    st, e = gs.state(), g["end_state"]
    assert st["cycle"] == int(e[4]) and abs(st["a"] - e[0]) <= 1e-14 * e[0] and abs(st["dtau"] - e[2]) <= 1e-13 * e[2]
    assert abs(st["maxvel"][0] - e[5]) <= 1e-10 * e[5] and abs(st["T00hom"] - e[6]) <= 1e-10 * e[6]
    gs.close()


def test_config1_shipped_settings_and_reference_snapshot(gevb, ctx, ref, tmp_path):
    """BASELINE config 1 as shipped -- settings.ini unchanged: Ngrid 64, 64^3 particles, GR, parabolic B, seed 42 -- with the
    initial conditions generated by the reference itself, six cycles on both sides; then the Gadget-2 snapshot of the device
    path against the file the reference's own saveGadget2 writes"""
    rs = ref.sim_from_settings()
    N = rs.N
    assert N == 64
    ids, pos, vel = rs.get_particles(0)
    assert len(ids) == 64 ** 3
    st = rs.state()
    state = [st["a"], st["tau"], st["dtau"], st["dtau_old"], st["cycle"], st["maxvel"][0], st["maxvel"][1]]
    gs = _gpu_sim_like(gevb, ctx(N), rs.cosmo, rs.dsettings, [N, rs.gr_flag, rs.vector_flag, rs.baryon_flag], rs.mass, ids, pos, vel,
                       rs.get_field("phi"), rs.get_field("chi"), rs.get_field("Bi"), rs.get_field("BiFT"), state)
    for step in range(6):
        rs.step(); gs.step()
        e = _compare_sims(rs, gs, N)
        tol = FIELD_TOL * (1 + step)
        bad = {k: v for k, v in e.items() if k not in ("state_tau", "scalarFT") and ((k.startswith("cells") and v != 0) or (not k.startswith("cells") and not v <= tol))}
        assert bad == {}, (step, e)
    # snapshot: same header, same blocks (particle order in the file is free; float32 values within one unit in the last place)
    fr, fg = str(tmp_path / "ref_snap"), str(tmp_path / "gpu_snap")
    tracer, dtau_pos, dtau_vel = 4, 0.011, 0.017
    rs.save_gadget2(0, fr, tracer, dtau_pos, dtau_vel)
    gs.save_gadget2(0, fg, tracer, dtau_pos, dtau_vel)
    rawr, rawg = open(fr, "rb").read(), open(fg, "rb").read()
    assert len(rawr) == len(rawg)
    hr, pr, vr, ir = _read_gadget2(fr)
    hg, pg, vg, ig = _read_gadget2(fg)
    assert hr["npart"] == hg["npart"] and hr["BoxSize"] == hg["BoxSize"]
    for k in ("time", "redshift"):
        assert abs(hr[k] - hg[k]) <= 1e-12 * abs(hr[k])
    assert np.allclose(hr["mass"], hg["mass"], rtol=1e-14, atol=0)
    import struct
    assert struct.unpack_from("<2i", rawr, 4 + 24 + 48 + 16 + 8 + 24)[1] == struct.unpack_from("<2i", rawg, 4 + 24 + 48 + 16 + 8 + 24)[1] == 1     # num_files
    assert struct.unpack_from("<3d", rawr, 4 + 24 + 48 + 16 + 8 + 24 + 8 + 8) == struct.unpack_from("<3d", rawg, 4 + 24 + 48 + 16 + 8 + 24 + 8 + 8)   # Omega0, OmegaLambda, h
    orr, og = np.argsort(ir), np.argsort(ig)
    assert np.array_equal(ir[orr], ig[og])
    assert np.all(np.abs(pg[og] - pr[orr]) <= np.spacing(np.abs(pr[orr]).astype(np.float32)) + 1e-30)
    assert np.all(np.abs(vg[og] - vr[orr]) <= 4 * np.spacing(np.abs(vr[orr]).astype(np.float32)) + 1e-30)
    rs.close(); gs.close()


@pytest.mark.parametrize("overrides,ngrid,tiling", [("", 64, 16), ("", 16, 4), ("gravity theory = Newton", 16, 4), ("baryon treatment = sample", 16, 4),
                                                    ("baryon treatment = hybrid", 16, 4), ("baryon treatment = ignore\ncorrect displacement = no\nk-domain = cube", 16, 4),
                                                    ("", 24, 6)],
                         ids=["config1_as_shipped", "ngrid16", "newton", "baryons_sampled", "baryons_hybrid", "ignore_cube_uncorrected", "ngrid24"])
def test_product_starts_from_settings_ini(gevb, ctx, ref, tmp_path, overrides, ngrid, tiling):
    """SURVEY 8f-2 / VERDICT r1 missing 2-3: the product reads the reference's settings.ini and generates its own initial
    conditions (host/settings.cpp, host/ic_basic.cpp: template, CIC kernel, Threefry realisation, transfer-function splines on
    the host; FFTs, displacement / momentum callbacks, phi, chi, B on the device) -- no oracle in the product's loop.  Compared
    with the reference's own parser + generateIC_basic from the same file and seed: identical IDs and per-cell counts,
    positions / momenta / fields to round-off; then both sides run three cycles."""
    d = tmp_path
    ref.dump_shipped_files(d)                                          # the three files the shipped run reads (test infrastructure provides the inputs)
    ov = f"Ngrid = {ngrid}\ntiling factor = {tiling}\ntemplate file = {d}/sc1_crystal.dat\nTk file = {d}/class_tk.dat\n" + overrides
    st = gevb.settings_read(d / "settings.ini", ov)
    assert st.ngrid == ngrid
    gs = gevb.sim_from_settings(ctx(ngrid), st)
    rs = ref.sim_from_settings(ngrid=ngrid, tiling=tiling, overrides=overrides)
    N = ngrid
    nspecies = 2 if "sample" in overrides else 1
    e = _compare_sims(rs, gs, N, nspecies=nspecies)
    # the realisation is bit-identical; what differs is the device's FFT / summation order: particle displacements are O(1e-3) box
    # units from gradients of fields that agree to 1e-10 relative
    skip = ("state_tau", "scalarFT", "source", "Sij", "SijFT", "state_T00hom", "state_dtau_old")
    bad = {k: v for k, v in e.items() if k not in skip and ((k.startswith("cells") and v != 0) or (not k.startswith("cells") and not v <= 5 * FIELD_TOL))}
    assert bad == {}, e
    assert abs(rs.state()["maxvel"][0] - gs.state()["maxvel"][0]) <= 1e-9 * rs.state()["maxvel"][0]
    for step in range(3):
        rs.step(); gs.step()
    e = _compare_sims(rs, gs, N, nspecies=nspecies)
    bad = {k: v for k, v in e.items() if k not in ("state_tau", "scalarFT") and ((k.startswith("cells") and v != 0) or (not k.startswith("cells") and not v <= 20 * FIELD_TOL))}
    assert bad == {}, e
    rs.close(); gs.close()


@pytest.mark.parametrize("overrides,ngrid", [("gravity theory = Newton", 16), ("vector method = elliptic", 16), ("baryon treatment = sample\ntiling factor = 4, 4", 16),
                                             ("", 24)], ids=["newton", "elliptic", "baryons_sampled", "ngrid24_not_power_of_two"])
def test_shipped_settings_variants_from_reference_ics(gevb, ctx, ref, overrides, ngrid):
    """the shipped settings.ini with one line changed (Newtonian gravity; elliptic vector method, which adds the T0i deposit and
    projectFTvector; baryons as their own species; a lattice size that is not a power of two, where pos/dx is a true
    division and bricks are partial), initial conditions by the reference's generator, three cycles"""
    rs = ref.sim_from_settings(ngrid, ngrid // 4, overrides=overrides)
    N, nsp = rs.N, 1 + rs.baryon_flag
    st = rs.state()
    gs = gevb.Sim(ctx(N), rs.gr_flag, rs.vector_flag, rs.dsettings, rs.cosmo)
    for sp in range(nsp):
        ids, pos, vel = rs.get_particles(sp)
        gs.set_particles(sp, ids, pos, vel, float(rs.mass[sp]))
    for name in ("phi", "chi", "Bi", "BiFT"):
        gs.set_field(name, rs.get_field(name))
    gs.set_state(st["a"], st["tau"], st["dtau"], st["dtau_old"], st["cycle"], st["maxvel"])
    for step in range(3):
        rs.step(); gs.step()
        e = _compare_sims(rs, gs, N, nspecies=nsp)
        skip = ("state_tau", "scalarFT") + (("chi", "Bi", "Sij", "BiFT", "SijFT", "state_T00hom") if rs.gr_flag == 0 else ())
        tol = FIELD_TOL * (1 + step)
        bad = {k: v for k, v in e.items() if k not in skip and ((k.startswith("cells") and v != 0) or (not k.startswith("cells") and not v <= tol))}
        assert bad == {}, (step, e)
    rs.close(); gs.close()


def test_shipped_run_to_z0_output_files(gevb, ctx, ref, tmp_path):
    """the whole shipped run in small (settings.ini at Ngrid 16, reference-made initial conditions) from z = 100 to z = 0 through
    gevb_sim_run and through the reference's loop, with the file's own output lists: every power-spectrum file (phi, chi, hij, B
    at z = 50, 30, 10, 3, 1, 0) and every Gadget-2 snapshot (z = 30, 10, 3, 0) is compared -- the north-star 1e-5 criterion on
    the spectra and the snapshots at z = 0, on the files a user would read"""
    rs = ref.sim_from_settings(16, 4)
    N = rs.N
    ids, pos, vel = rs.get_particles(0)
    st = rs.state()
    gs = _gpu_sim_like(gevb, ctx(N), rs.cosmo, rs.dsettings, [N, rs.gr_flag, rs.vector_flag, rs.baryon_flag], rs.mass, ids, pos, vel,
                       rs.get_field("phi"), rs.get_field("chi"), rs.get_field("Bi"), rs.get_field("BiFT"),
                       [st["a"], st["tau"], st["dtau"], st["dtau_old"], st["cycle"], st["maxvel"][0], st["maxvel"][1]])
    z_pk, z_snap, mask, numbins, tracer = [50., 30., 10., 3., 1., 0.], [30., 10., 3., 0.], 1 | 2 | 8 | 128, 16, 2
    rdir, gdir = tmp_path / "ref", tmp_path / "gpu"
    rdir.mkdir(); gdir.mkdir()
    rc = rs.run(z_pk, mask, numbins, str(rdir / "pk"), z_snap, tracer, str(rdir / "snap"))
    gc = gs.run(z_pk, mask, numbins, str(gdir / "pk"), z_snap, tracer, str(gdir / "snap"))
    assert rc == gc and rc[1:] == (6, 4), (rc, gc)
    worst = 0.0
    for k in range(len(z_pk)):
        for tag in ("phi", "chi", "hij", "B"):
            lr = open(str(rdir / f"pk{k:03d}_{tag}.dat")).read().splitlines()
            lg = open(str(gdir / f"pk{k:03d}_{tag}.dat")).read().splitlines()
            assert lr[:3] == lg[:3] and len(lr) == len(lg), (k, tag)
            a = np.array([[float(v) for v in ln.split()] for ln in lr[3:]])
            b = np.array([[float(v) for v in ln.split()] for ln in lg[3:]])
            assert np.array_equal(a[:, 4], b[:, 4]) and np.allclose(a[:, 0], b[:, 0], rtol=2e-6)
            err = np.max(np.abs(b[:, 1] - a[:, 1]) / np.abs(a[:, 1]))
            worst = max(worst, err)
            assert err <= 1e-5, (k, tag, err)
            assert np.allclose(a[:, 3], b[:, 3], rtol=1e-3, atol=1e-5 * np.abs(a[:, 1]).max())
    for k, z in enumerate(z_snap):
        hr, pr, vr, ir = _read_gadget2(str(rdir / f"snap{k:03d}_cdm"))
        hg, pg, vg, ig = _read_gadget2(str(gdir / f"snap{k:03d}_cdm"))
        assert hr["npart"] == hg["npart"] and hr["redshift"] == hg["redshift"] == z and hr["time"] == hg["time"] and hr["BoxSize"] == hg["BoxSize"]
        orr, og = np.argsort(ir), np.argsort(ig)
        assert np.array_equal(ir[orr], ig[og])
        d = np.abs(pg[og] - pr[orr])
        d = np.minimum(d, hr["BoxSize"] - d)                                     # a particle may sit on either side of the periodic boundary
        assert d.max() <= 1e-5 * hr["BoxSize"] * (1 + k), (k, d.max())
        assert np.abs(vg[og] - vr[orr]).max() <= 1e-4 * np.abs(vr).max(), k
        # snapshot statistics (north star: 1e-5): rms velocity, extent of the displacements from the lattice
        sr, sg = np.sqrt((vr.astype(np.float64) ** 2).mean()), np.sqrt((vg.astype(np.float64) ** 2).mean())
        assert abs(sg - sr) <= 1e-5 * sr
    print("shipped run: cycles", rc[0], "worst spectrum rel err", worst)
    rs.close(); gs.close()


def test_settings_file_to_output_files_product_alone(gevb, ctx, ref, tmp_path):
    """BASELINE config 1 end to end with NO oracle in the product's loop: the product reads settings.ini (Ngrid 16 here), makes
    its own initial conditions from the file's seed, runs the main loop to z = 0 and writes the file's spectra and Gadget-2
    snapshots (gevb_settings_read -> gevb_sim_create_from_settings -> gevb_sim_run_settings, what scripts/run_settings.py
    does); the reference does the same from the same file.  Spectra within 1e-5, same bins and counts; snapshot headers equal,
    same particle IDs, positions within 1e-5 of the box."""
    d = tmp_path
    ref.dump_shipped_files(d)
    gdir, rdir = d / "gpu_out", d / "ref_out"
    gdir.mkdir(); rdir.mkdir()
    ov = (f"Ngrid = 16\ntiling factor = 4\ntemplate file = {d}/sc1_crystal.dat\nTk file = {d}/class_tk.dat\noutput path = {gdir}/\n"
          "Pk bins = 16\ntracer factor = 2\nsnapshot outputs = Gadget2\nPk file base = pk\nsnapshot file base = snap")
    st = gevb.settings_read(d / "settings.ini", ov)
    gs = gevb.sim_from_settings(ctx(16), st)
    gc = gs.run_settings(st)
    rs = ref.sim_from_settings(16, 4)
    z_pk, z_snap = list(st.z_pk)[:st.num_pk], list(st.z_snapshot)[:st.num_snapshot]
    rc = rs.run(z_pk, st.pk_mask, st.numbins, str(rdir / "pk"), z_snap, 2, str(rdir / "snap"))
    assert rc == gc and rc[1:] == (6, 4), (rc, gc)
    for k in range(len(z_pk)):
        for tag in ("phi", "chi", "hij", "B"):
            lr = open(str(rdir / f"pk{k:03d}_{tag}.dat")).read().splitlines()
            lg = open(str(gdir / f"pk{k:03d}_{tag}.dat")).read().splitlines()
            assert lr[:3] == lg[:3] and len(lr) == len(lg), (k, tag)
            a = np.array([[float(v) for v in ln.split()] for ln in lr[3:]])
            b = np.array([[float(v) for v in ln.split()] for ln in lg[3:]])
            assert np.array_equal(a[:, 4], b[:, 4])
            assert np.max(np.abs(b[:, 1] - a[:, 1]) / np.abs(a[:, 1])) <= 1e-5, (k, tag)
    for k, z in enumerate(z_snap):
        hr, pr, vr, ir = _read_gadget2(str(rdir / f"snap{k:03d}_cdm"))
        hg, pg, vg, ig = _read_gadget2(str(gdir / f"snap{k:03d}_cdm"))
        assert hr["npart"] == hg["npart"] and hr["redshift"] == hg["redshift"] == z and hr["BoxSize"] == hg["BoxSize"]
        orr, og = np.argsort(ir), np.argsort(ig)
        assert np.array_equal(ir[orr], ig[og])
        dd = np.abs(pg[og] - pr[orr])
        dd = np.minimum(dd, hr["BoxSize"] - dd)
        assert dd.max() <= 1e-5 * hr["BoxSize"] * (1 + k), (k, dd.max())
    rs.close(); gs.close()
