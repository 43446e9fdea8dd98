"""Seeded hot-path cases whose reference outputs are committed under tests/golden/.

`inputs(N, seed)` builds the inputs deterministically; `run_cpu(checker, inp)`
evaluates every hot-path function on a CPU checker (oracle/oracle.py);
`run_gpu(gevb, ctx, inp)` evaluates the same cases through the C ABI.
tests/golden/make_golden.py stores run_cpu(<compiled reference>) once.
"""
import numpy as np

import common


def inputs(N=8, seed=20221):
    rng = np.random.default_rng(seed)
    a = 0.03
    inp = dict(N=N, a=a, mass=0.31 / (2 * N ** 3), dtau=0.011, dtau_kick=0.0095)
    phi, chi, Bi = common.metric_fields(rng, N, a)
    # larger amplitudes than cosmological so that every non-linear term is visible above round-off
    inp["phi"] = phi * 1e3
    inp["chi"] = chi * 1e4
    inp["Bi"] = Bi * 1e4
    ids, pos, vel = common.quasi_uniform_particles(rng, N, sigma=0.6, a=a, qscale=0.5)
    ids2, pos2, vel2 = common.clustered_particles(rng, N, N ** 3, a=a, qscale=0.5)
    inp["ids"] = np.concatenate([ids, ids2 + len(ids)])
    inp["pos"] = np.ascontiguousarray(np.concatenate([pos, pos2]))
    inp["vel"] = np.ascontiguousarray(np.concatenate([vel, vel2]))
    inp["source"] = 1.0 + 0.3 * common.gaussian_field(rng, N, 1, 1.0)
    inp["Tij"] = common.gaussian_field(rng, N, 6, 1.0, slope=-1.0)
    inp["Si"] = common.gaussian_field(rng, N, 3, 1.0, slope=-1.0)
    inp["params"] = np.array([a, a * a * N])
    return inp


def run_cpu(o, inp):
    N, a, mass = inp["N"], inp["a"], inp["mass"]
    phi, chi, Bi, pos, vel = inp["phi"], inp["chi"], inp["Bi"], inp["pos"], inp["vel"]
    out = {}
    out["fft_fwd"] = o.fft_forward(inp["Tij"])
    out["fft_bwd"] = o.fft_backward(out["fft_fwd"])
    out["prep_scalar"] = o.prepareFTsource_scalar(phi, chi, inp["source"], 0.31, 0.7, 1.3, 0.2)
    out["prep_tensor"] = o.prepareFTsource_tensor(phi, inp["Tij"], 1.7)
    srcFT = o.fft_forward(inp["source"])
    SijFT = out["fft_fwd"]
    SiFT = o.fft_forward(inp["Si"])
    out["poisson_mod"] = o.solveModifiedPoissonFT(srcFT, 2.5, 0.8)
    out["poisson"] = o.solveModifiedPoissonFT(srcFT, 2.5, 0.0)
    out["ftscalar"] = o.projectFTscalar(SijFT)
    out["ftscalar_add"] = o.projectFTscalar(SijFT, srcFT)
    out["evolve"] = o.evolveFTvector(SijFT, SiFT, 0.37)
    out["ftvector"] = o.projectFTvector(SiFT, 1.3, 0.0)
    out["ftvector_mod"] = o.projectFTvector(SiFT, 1.3, 0.6)
    out["fttensor"] = o.projectFTtensor(SijFT)
    for tag, ph in (("", phi), ("_nophi", None)):
        out["T00" + tag] = o.projection_T00(N, pos, vel, mass, a, ph, 1.0)
        out["T0i" + tag] = o.projection_T0i(N, pos, vel, mass, ph, 1.0)
        out["Tij" + tag] = o.projection_Tij(N, pos, vel, mass, a, ph, 1.0)
    out["cic"] = o.scalarProjectionCIC(N, pos, mass)
    cell, counts = o.cell_index(N, pos)
    out["cell"], out["counts"] = cell, counts
    for nf in (1, 2, 3):
        v, m = o.updateVel(N, pos, vel, 0, inp["dtau_kick"], phi, chi, Bi, nf, inp["params"])
        out[f"kick_nf{nf}"], out[f"kick_nf{nf}_max"] = v, np.array([m])
    v, m = o.updateVel(N, pos, vel, 1, inp["dtau_kick"], phi, chi, Bi, 1, inp["params"])
    out["kick_newton"], out["kick_newton_max"] = v, np.array([m])
    for nf in (0, 1, 2, 3):
        out[f"drift_nf{nf}"] = o.moveParticles(N, pos, vel, 0, inp["dtau"], phi, chi, Bi, nf, inp["params"])
    out["drift_newton"] = o.moveParticles(N, pos, vel, 1, inp["dtau"] * a, None, None, None, 0, inp["params"])
    pk = o.extractPowerSpectrum(srcFT, 16)
    for k, name in enumerate(("kbin", "power", "kscatter", "pscatter", "occupation")):
        out["pk_" + name] = np.asarray(pk[k], dtype=np.float64)
    return out


# outputs that are integer / index work: bit-exact contract
EXACT = ("cell", "counts", "pk_occupation")


def run_gpu(g, ctx, inp):
    """Same cases through libgevb.so (single rank). Particle outputs are mapped back to input order by ID."""
    N, a, mass = inp["N"], inp["a"], inp["mass"]
    pos, vel, ids = inp["pos"], inp["vel"], inp["ids"]
    out = {}
    F = lambda data, nc=1, kind=None: g.Field(ctx, g.REAL if kind is None else kind, nc, data=data)
    phi, chi, Bi = F(inp["phi"]).updateHalo(), F(inp["chi"]).updateHalo(), F(inp["Bi"], 3).updateHalo()
    # FFT
    T = g.Field(ctx, g.REAL, 6, symmetric=True, data=inp["Tij"])
    TFT = g.Field(ctx, g.CPLX, 6, symmetric=True)
    plan6 = g.PlanFFT(T, TFT)
    plan6.execute(g.FFT_FORWARD)
    out["fft_fwd"] = TFT.download()
    T.projection_init()
    plan6.execute(g.FFT_BACKWARD)
    out["fft_bwd"] = T.download()
    assert np.array_equal(TFT.download(), out["fft_fwd"]), "backward FFT must preserve its Fourier input"
    # real-space source preparation (aliased as in main.cpp:472,539)
    src = F(inp["source"])
    g.prepareFTsource_scalar(phi, chi, src, 0.31, src, 0.7, 1.3, 0.2)
    out["prep_scalar"] = src.download()
    T.upload(inp["Tij"])
    g.prepareFTsource_tensor(phi, T, T, 1.7)
    out["prep_tensor"] = T.download()
    # Fourier kernels
    src.upload(inp["source"])
    sFT = g.Field(ctx, g.CPLX, 1)
    plan1 = g.PlanFFT(src, sFT)
    plan1.execute(g.FFT_FORWARD)
    srcFT = sFT.download()
    S3 = F(inp["Si"], 3)
    S3FT = g.Field(ctx, g.CPLX, 3)
    plan3 = g.PlanFFT(S3, S3FT)
    plan3.execute(g.FFT_FORWARD)
    SiFT = S3FT.download()
    pot = g.Field(ctx, g.CPLX, 1)
    g.solveModifiedPoissonFT(sFT, pot, 2.5, 0.8); out["poisson_mod"] = pot.download()
    g.solveModifiedPoissonFT(sFT, pot, 2.5, 0.0); out["poisson"] = pot.download()
    g.projectFTscalar(TFT, pot, 0); out["ftscalar"] = pot.download()
    pot.upload(srcFT); g.projectFTscalar(TFT, pot, 1); out["ftscalar_add"] = pot.download()
    B3 = g.Field(ctx, g.CPLX, 3, data=SiFT)
    g.evolveFTvector(TFT, B3, 0.37); out["evolve"] = B3.download()
    # fused projectFTscalar + evolveFTvector: same two results from one pass over SijFT
    B3.upload(SiFT); c1 = g.Field(ctx, g.CPLX, 1)
    g.projectFTscalar_evolveFTvector(TFT, c1, B3, 0.37); out["ftscalar_fusedB"], out["evolve_fusedchi"] = c1.download(), B3.download()
    B3.upload(SiFT); g.projectFTvector(B3, B3, 1.3, 0.0); out["ftvector"] = B3.download()
    B3.upload(SiFT); g.projectFTvector(B3, B3, 1.3, 0.6); out["ftvector_mod"] = B3.download()
    g.projectFTtensor(TFT, TFT); out["fttensor"] = TFT.download()
    # projections
    p = g.Particles(ctx, mass).add(ids, pos, vel)
    t00, t0i = g.Field(ctx, g.REAL, 1), g.Field(ctx, g.REAL, 3)
    for tag, ph in (("", phi), ("_nophi", None)):
        t00.projection_init(); p.projection_T00_project(t00, a, ph, 1.0); t00.projection_comm(); out["T00" + tag] = t00.download()
        t0i.projection_init(); p.projection_T0i_project(t0i, ph, 1.0); t0i.projection_comm(); out["T0i" + tag] = t0i.download()
        T.projection_init(); p.projection_Tij_project(T, a, ph, 1.0); T.projection_comm(); out["Tij" + tag] = T.download()
    t00.projection_init(); T.projection_init()
    p.projection_T00_Tij_project(t00, T, a, phi, 1.0); t00.projection_comm(); T.projection_comm()
    out["T00_fused"], out["Tij_fused"] = t00.download(), T.download()
    t00.projection_init(); p.scalarProjectionCIC_project(t00); t00.projection_comm(); out["cic"] = t00.download()
    out["counts"] = p.cell_counts()
    did, dpos, dvel = p.download()
    order = np.argsort(did)
    assert np.array_equal(did[order], ids)
    dx = 1.0 / N
    c = np.minimum(np.floor(dpos / dx).astype(np.int64), N - 1)
    keys = (c[:, 2] * N + c[:, 1]) * N + c[:, 0]
    bkey = g.storage_key(N, N, c[:, 0], c[:, 1], c[:, 2])      # storage order: bricks, then cells inside the brick
    assert np.all(np.diff(bkey) >= 0), "device particle order is not brick-major cell-sorted"
    out["cell"] = keys[order].astype(np.int32)
    assert np.array_equal(dpos[order], pos) and np.array_equal(dvel[order], vel)
    p.close()
    fields = [phi, chi, Bi]

    def fresh():
        return g.Particles(ctx, mass).add(ids, pos, vel)

    def by_id(pp):
        i, x, v = pp.download()
        o = np.argsort(i)
        return x[o], v[o]

    for nf in (1, 2, 3):
        pp = fresh(); m = pp.updateVel(g.UPDATE_Q, inp["dtau_kick"], fields, nf, inp["params"])
        out[f"kick_nf{nf}"], out[f"kick_nf{nf}_max"] = by_id(pp)[1], np.array([m]); pp.close()
    pp = fresh(); m = pp.updateVel(g.UPDATE_Q_NEWTON, inp["dtau_kick"], fields, 1, inp["params"])
    out["kick_newton"], out["kick_newton_max"] = by_id(pp)[1], np.array([m]); pp.close()
    for nf in (0, 1, 2, 3):
        pp = fresh(); pp.moveParticles(g.UPDATE_Q, inp["dtau"], fields, nf, inp["params"])
        out[f"drift_nf{nf}"] = by_id(pp)[0]; out[f"drift_nf{nf}_counts"] = pp.cell_counts(); pp.close()
    pp = fresh(); pp.moveParticles(g.UPDATE_Q_NEWTON, inp["dtau"] * a, None, 0, inp["params"])
    out["drift_newton"] = by_id(pp)[0]; pp.close()
    # fused kick + drift == kick then drift
    pp = fresh()
    m = pp.kick_drift(g.UPDATE_Q, inp["dtau_kick"], 3, inp["params"], inp["dtau"], 3, inp["params"], fields)
    out["fused_pos"], out["fused_vel"] = by_id(pp); out["fused_max"] = np.array([m]); pp.close()
    pk = g.extractPowerSpectrum(sFT, 16)
    for k, name in enumerate(("kbin", "power", "kscatter", "pscatter", "occupation")):
        out["pk_" + name] = np.asarray(pk[k], dtype=np.float64)
    return out
