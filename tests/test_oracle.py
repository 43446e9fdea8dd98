"""CPU tests of the checkers themselves (run with -m "not gpu").

The reference ships no tests or golden vectors, so the oracle
is pinned three ways:
  1. tests/golden/hotpath_N8.npz -- outputs of the reference's own gevolution.hpp
     (compiled against the LATfield2 shim, oracle/_ref) on seeded inputs;
  2. the compiled reference itself where oracle/_ref is present (this container);
  3. the analytic invariants the reference's own diagnostics check
     (tools.hpp:364-431, main.cpp:465-468).
"""
import os

import numpy as np
import pytest

import common
import golden_cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "hotpath_N8.npz")
TOL = 1e-13   # CPU restatement vs compiled reference: same arithmetic, different summation layout


def _compare(out, gold, tol):
    bad = []
    for k in gold.files:
        a, b = np.asarray(out[k]), gold[k]
        if k in golden_cases.EXACT:
            if not np.array_equal(a.astype(np.int64), b.astype(np.int64)):
                bad.append((k, "integer mismatch"))
        else:
            err = common.rel_linf(a, b)
            if not err <= tol:
                bad.append((k, err))
    return bad


def test_restatement_matches_golden(ora):
    gold = np.load(GOLDEN)
    out = golden_cases.run_cpu(ora, golden_cases.inputs(N=8))
    assert set(gold.files) == set(out.keys())
    assert _compare(out, gold, TOL) == []


def test_compiled_reference_reproduces_golden(ref):
    gold = np.load(GOLDEN)
    out = golden_cases.run_cpu(ref, golden_cases.inputs(N=8))
    assert _compare(out, gold, 0.0) == []


@pytest.mark.parametrize("N", [6, 12, 16])
def test_restatement_matches_reference_other_sizes(ora, ref, N):
    """non-power-of-two and larger lattices, fresh seed"""
    inp = golden_cases.inputs(N=N, seed=77 + N)
    a, b = golden_cases.run_cpu(ora, inp), golden_cases.run_cpu(ref, inp)

    class G:
        files = list(b.keys())
        __getitem__ = lambda self, k: b[k]
    assert _compare(a, G(), TOL) == []


def test_fft_matches_numpy(ora):
    rng = np.random.default_rng(3)
    N = 16
    f = rng.standard_normal((2, N, N, N))
    F = ora.fft_forward(f)
    Fn = np.fft.rfftn(f, axes=(1, 2, 3))
    assert common.rel_linf(common.to_cplx(F), Fn) < 1e-14
    assert common.rel_linf(ora.fft_backward(F), f * N ** 3) < 1e-14      # unnormalised both ways


# ---- invariants (verified against the reference's diagnostics) ----
def Dp(f, i): return np.roll(f, -1, axis=2 - i) - f
def Dm(f, i): return f - np.roll(f, 1, axis=2 - i)


def test_invariant_poisson(ora):
    rng = np.random.default_rng(1)
    N = 16
    src = rng.standard_normal((1, N, N, N)); src -= src.mean()
    pot = ora.fft_backward(ora.solveModifiedPoissonFT(ora.fft_forward(src), 1.0, 0.0))[0]
    lap = sum(Dp(Dm(pot, i), i) for i in range(3)) * N * N
    assert np.abs(lap - src[0]).max() < 1e-12


def test_invariant_vector_projection_divergence_free(ora):
    rng = np.random.default_rng(2)
    N = 16
    B = ora.fft_backward(ora.projectFTvector(ora.fft_forward(rng.standard_normal((3, N, N, N))), 1.0))
    mdiv, mcurl = ora.computeVectorDiagnostics(B)
    assert mdiv < 1e-11 * np.abs(B).max() * N and mcurl > 1.0


def test_invariant_scalar_projection_recovers_chi(ora):
    rng = np.random.default_rng(4)
    N = 16
    chi = rng.standard_normal((N, N, N))
    trace = rng.standard_normal((N, N, N))
    idx = {(0, 0): 0, (0, 1): 1, (0, 2): 2, (1, 1): 3, (1, 2): 4, (2, 2): 5}
    S = np.zeros((6, N, N, N))
    for i in range(3):
        S[idx[(i, i)]] = Dp(Dm(chi, i), i) + trace
    for (i, j) in ((0, 1), (0, 2), (1, 2)):
        S[idx[(i, j)]] = Dp(Dp(chi, i), j)
    back = ora.fft_backward(ora.projectFTscalar(ora.fft_forward(S)))[0]
    assert np.abs(back - (chi - chi.mean())).max() < 1e-12


def test_invariant_tensor_projection_transverse_traceless(ora):
    rng = np.random.default_rng(5)
    N = 16
    h = ora.fft_backward(ora.projectFTtensor(ora.fft_forward(rng.standard_normal((6, N, N, N)))))
    mdiv, mtrace, mnorm = ora.computeTensorDiagnostics(h)
    assert mdiv < 1e-11 * mnorm * N and mtrace < 1e-12 * mnorm


def test_mass_conservation(ora):
    """with phi = NULL the deposited -a^3 T^0_0 averages to Omega (main.cpp:465-468)"""
    rng = np.random.default_rng(6)
    N = 12
    ids, pos, vel = common.clustered_particles(rng, N, 5000)
    T00 = ora.projection_T00(N, pos, vel, 0.3 / len(pos), 0.1, None)
    assert abs(T00.mean() - 0.3) < 1e-14
    rho = ora.scalarProjectionCIC(N, pos, 0.3 / len(pos))
    assert abs(rho.mean() - 0.3) < 1e-14


def test_uniform_lattice_symmetry(ora):
    """particles at cell centres, q = 0, phi = 0: T00 uniform, Tij = T0i = 0"""
    N = 8
    rng = np.random.default_rng(0)
    ids, pos, vel = common.quasi_uniform_particles(rng, N, sigma=0.0)
    vel[:] = 0.0
    a = 0.05
    T00 = ora.projection_T00(N, pos, vel, 0.3 / len(pos), a, np.zeros((1, N, N, N)))
    assert np.abs(T00 - 0.3).max() < 1e-14
    assert np.abs(ora.projection_Tij(N, pos, vel, 0.3 / len(pos), a, None)).max() == 0.0
    assert np.abs(ora.projection_T0i(N, pos, vel, 0.3 / len(pos), None)).max() == 0.0


def test_binning_edges(ora):
    """cell = floor(pos/dx): positions on cell boundaries and next to the box edge"""
    N = 8
    dx = 1.0 / N
    pos = np.array([[0.0, 0.0, 0.0], [dx, 2 * dx, 3 * dx], [np.nextafter(dx, 0), 0.5, 0.5],
                    [np.nextafter(1.0, 0), np.nextafter(1.0, 0), np.nextafter(1.0, 0)], [0.999, 0.0, 0.4375]])
    cell, counts = ora.cell_index(N, pos)
    expect = [0, (3 * N + 2) * N + 1, (4 * N + 4) * N + 0, (7 * N + 7) * N + 7, (3 * N + 0) * N + 7]
    assert cell.tolist() == expect and counts.sum() == len(pos)


def test_drift_wrap_convention(ora):
    """periodic wrap: result in [0,1); -eps maps to 0 (documented edge semantics)"""
    N = 8
    pos = np.array([[0.01, 0.99, 0.5], [1e-18, 0.5, 0.5]])
    vel = np.array([[-0.05, 0.05, 0.0], [-1e-17, 0.0, 0.0]])
    out = ora.moveParticles(N, pos, vel, 1, 1.0, None, None, None, 0, [1.0, 1.0])
    assert np.all(out >= 0.0) and np.all(out < 1.0)
    assert abs(out[0, 0] - 0.96) < 1e-15 and abs(out[0, 1] - 0.04) < 1e-15
    assert out[1, 0] == 0.0


def test_empty_particle_set(ora):
    N = 6
    pos = np.zeros((0, 3)); vel = np.zeros((0, 3))
    assert np.all(ora.projection_T00(N, pos, vel, 1.0, 0.1, None) == 0.0)
    assert np.all(ora.projection_Tij(N, pos, vel, 1.0, 0.1, None) == 0.0)


def test_background_matches_reference(ora, ref):
    cosmo = common.shipped_cosmology()
    fourpiG = 1.5 * 320.0 ** 2 / 2997.92458 ** 2
    for a in (0.0099, 0.1, 1.0):
        assert ora.Hconf(a, fourpiG, cosmo) == ref.Hconf(a, fourpiG, cosmo)
        assert ora.rungekutta4bg(a, fourpiG, cosmo, 0.01) == ref.rungekutta4bg(a, fourpiG, cosmo, 0.01)
    # Omega_m + Omega_Lambda + Omega_rad = 1  =>  Hconf(1) = sqrt(2 fourpiG / 3)
    assert abs(ora.Hconf(1.0, fourpiG, cosmo) - np.sqrt(2 * fourpiG / 3)) < 1e-15


def test_reference_time_loop_runs(ref):
    """two cycles of the reference's main loop on the shim: T00hom = Omega_m to O(phi), a advances"""
    rng = np.random.default_rng(11)
    N = 8
    cosmo = common.shipped_cosmology()
    sim = ref.sim(N, 1, 0, common.shipped_settings(), cosmo)
    ids, pos, vel = common.quasi_uniform_particles(rng, N, sigma=0.1, a=1 / 101.0)
    sim.set_particles(0, ids, pos, vel, (cosmo[0] + cosmo[1]) / len(ids))
    phi, chi, Bi = common.metric_fields(rng, N, 1 / 101.0)
    sim.set_field("phi", phi)
    s0 = sim.state()
    sim.step(); sim.step()
    s = sim.state()
    assert s["cycle"] == 2 and s["a"] > s0["a"] and s["dtau_old"] > 0
    assert abs(s["T00hom"] - (cosmo[0] + cosmo[1])) < 1e-3
    ids2, pos2, vel2 = sim.get_particles()
    assert sorted(ids2.tolist()) == ids.tolist() and np.all((pos2 >= 0) & (pos2 < 1))
    sim.close()


# ---- the reference's shipped configuration from its own seed (parser + generateIC_basic compiled into oracle/_ref) ----
SHIPPED = os.path.join(os.path.dirname(__file__), "golden", "shipped_settings_N16.npz")


def test_shipped_settings_ic_generation(ref):
    """settings.ini of the reference (GR, parabolic, blend, seed 42) with Ngrid = 16, tiling 4: what the reference's own
    parser and IC generator produce is deterministic, depends on the seed, conserves mass, and equals the committed fixture"""
    s = ref.sim_from_settings(16, 4)
    assert (s.N, s.gr_flag, s.vector_flag, s.baryon_flag) == (16, 1, 0, 0)
    cosmo = common.shipped_cosmology()
    assert np.allclose(s.cosmo, cosmo, rtol=1e-12) and np.allclose(s.dsettings, common.shipped_settings(), rtol=0, atol=0)
    ids, pos, vel = s.get_particles(0)
    assert len(ids) == 4 ** 3 * 4 ** 3 and sorted(ids.tolist()) == list(range(len(ids)))     # sc1 template: 4^3 points, tiled 4^3 times
    assert abs(s.mass[0] - (cosmo[0] + cosmo[1]) / len(ids)) <= 1e-15 * s.mass[0]             # baryon treatment = blend
    assert np.all((pos >= 0) & (pos < 1)) and 1e-5 < np.abs(s.get_field("phi")).max() < 1e-3
    g = np.load(SHIPPED)
    o = np.argsort(ids)
    assert np.array_equal(ids[o], g["ic_ids"][np.argsort(g["ic_ids"])])
    assert np.abs(pos[o] - g["ic_pos"][np.argsort(g["ic_ids"])]).max() <= 1e-14 and common.rel_linf(vel[o], g["ic_vel"][np.argsort(g["ic_ids"])]) <= 1e-12
    assert common.rel_linf(s.get_field("phi"), g["ic_phi"]) <= 1e-12
    st = s.state()
    assert abs(st["a"] - 1 / 101.0) < 1e-15 and abs(st["dtau"] - g["ic_state"][2]) <= 1e-14 and st["maxvel"][0] > 0
    # mass conservation through the first cycle, and the state after the fixture's cycles
    for _ in range(int(g["ncycles"][0])):
        s.step()
    assert abs(s.state()["T00hom"] / (cosmo[0] + cosmo[1]) - 1.0) < 1e-3
    i2, p2, v2 = s.get_particles(0)
    o2 = np.argsort(i2)
    assert np.abs(p2[o2] - g["end_pos"]).max() <= 1e-13 and common.rel_linf(s.get_field("phi"), g["end_phi"]) <= 1e-11
    # another seed gives another realisation
    s3 = ref.sim_from_settings(16, 4, seed=7)
    assert np.abs(s3.get_field("phi") - g["ic_phi"]).max() > 1e-6
    s.close(); s3.close()


def test_reference_gadget2_writer_runs(ref, tmp_path):
    """the reference's own saveGadget2 over the MPI-IO stand-in: file size and block markers of the Gadget-2 layout"""
    import struct
    s = ref.sim_from_settings(16, 4)
    fn = str(tmp_path / "snap")
    s.save_gadget2(0, fn, 2, 0.01, 0.02)
    raw = open(fn, "rb").read()
    n = 4096 // 2
    assert len(raw) == 264 + (8 + 12 * n) * 2 + 8 + 8 * n
    assert struct.unpack_from("<I", raw, 0)[0] == 256 and struct.unpack_from("<6I", raw, 4)[1] == n
    assert struct.unpack_from("<I", raw, 264)[0] == 12 * n and struct.unpack_from("<I", raw, len(raw) - 4)[0] == 8 * n
    s.close()


def test_restatement_on_reference_made_ics(ora, ref):
    """the plain-C restatement against the compiled reference on physical initial conditions (the reference generator's
    particles and metric fields for the shipped settings.ini), not only on the seeded synthetic inputs"""
    s = ref.sim_from_settings(16, 4)
    N, a = s.N, s.state()["a"]
    ids, pos, vel = s.get_particles(0)
    phi, chi, Bi = s.get_field("phi"), s.get_field("chi"), s.get_field("Bi")
    mass = float(s.mass[0])
    for name, args in (("projection_T00", (N, pos, vel, mass, a, phi[0])), ("projection_Tij", (N, pos, vel, mass, a, phi[0])), ("projection_T0i", (N, pos, vel, mass, phi[0]))):
        assert common.rel_linf(getattr(ora, name)(*args), getattr(ref, name)(*args)) <= 1e-12, name
    params = [a, a * a * N]
    vo, mo = ora.updateVel(N, pos, vel, 0, 0.05, phi, chi, Bi, 3, params)
    vr, mr = ref.updateVel(N, pos, vel, 0, 0.05, phi, chi, Bi, 3, params)
    assert common.rel_linf(vo, vr) <= 1e-13 and abs(mo - mr) <= 1e-13 * mr
    po = ora.moveParticles(N, pos, vr, 0, 0.07, phi, chi, Bi, 3, params)
    pr = ref.moveParticles(N, pos, vr, 0, 0.07, phi, chi, Bi, 3, params)
    assert np.abs(po - pr).max() <= 1e-15
    T = ref.projection_Tij(N, pos, vel, mass, a, phi[0])
    assert common.rel_linf(ora.prepareFTsource_tensor(phi[0], T, 0.3), ref.prepareFTsource_tensor(phi[0], T, 0.3)) <= 1e-13
    SF = ref.fft_forward(ref.prepareFTsource_tensor(phi[0], T, 0.3))
    assert common.rel_linf(ora.projectFTscalar(SF), ref.projectFTscalar(SF)) <= 1e-12
    BF = s.get_field("BiFT")
    assert common.rel_linf(ora.evolveFTvector(SF, BF, 0.02), ref.evolveFTvector(SF, BF, 0.02)) <= 1e-12
    s.close()


def test_reference_run_with_outputs(ref, tmp_path):
    """the shipped settings.ini at Ngrid 16 to z = 0 with its own output lists (Pk at z = 50, 30, 10, 3, 1, 0; snapshots at
    z = 30, 10, 3, 0): every file appears, stamped with the exact target redshift (EXACT_OUTPUT_REDSHIFTS)"""
    s = ref.sim_from_settings(16, 4)
    z_pk, z_snap = [50., 30., 10., 3., 1., 0.], [30., 10., 3., 0.]
    cycles, npk, nsnap = s.run(z_pk, 1 | 2 | 8 | 128, 16, str(tmp_path / "pk"), z_snap, 2, str(tmp_path / "snap"))
    assert (npk, nsnap) == (6, 4) and 40 < cycles < 400 and s.state()["a"] >= 1.0
    for k, z in enumerate(z_pk):
        for tag in ("phi", "chi", "hij", "B"):
            lines = open(str(tmp_path / f"pk{k:03d}_{tag}.dat")).read().splitlines()
            assert lines[0] == f"# power spectrum of {tag}" and lines[1] == "# redshift z=%f" % z and len(lines) > 8
    import struct
    for k, z in enumerate(z_snap):
        raw = open(str(tmp_path / f"snap{k:03d}_cdm"), "rb").read()
        time, redshift = struct.unpack_from("<2d", raw, 4 + 24 + 48)
        assert redshift == z and abs(time - 1. / (1. + z)) < 1e-15 and struct.unpack_from("<6I", raw, 4)[1] == 2048
    s.close()
