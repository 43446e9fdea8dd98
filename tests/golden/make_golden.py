"""Regenerate tests/golden/hotpath_N8.npz and tests/golden/shipped_settings_N16.npz from the compiled reference (oracle/_ref).

Run in the build container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py
The reference ships no golden vectors of its own; these are
outputs of the reference's own gevolution.hpp / tools.hpp compiled against the
single-rank LATfield2 shim, on the seeded inputs of tests/golden_cases.py.

shipped_settings_N16.npz: the reference's shipped settings.ini (GR, parabolic B, baryon treatment = blend, seed 42) with
Ngrid = 16 and tiling factor = 4, run by the reference's own parser and IC generator (generateIC_basic) and then NCYCLES
cycles of its main loop: initial conditions (particles, phi, chi, Bi, BiFT, loop scalars) and the state after the cycles.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import golden_cases  # noqa: E402
import oracle  # noqa: E402

if __name__ == "__main__":
    oracle.build()
    ref = oracle.load_ref()
    assert ref is not None, "oracle/_ref/libgevref.so missing: the reference tree is needed to make golden vectors"
    inp = golden_cases.inputs(N=8)
    out = golden_cases.run_cpu(ref, inp)
    path = os.path.join(HERE, "hotpath_N8.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays;", ref.description)

    NCYCLES = 4
    s = ref.sim_from_settings(16, 4)
    ids, pos, vel = s.get_particles(0)
    st = s.state()
    ic = dict(cosmo=s.cosmo, dsettings=s.dsettings, flags=np.array([s.N, s.gr_flag, s.vector_flag, s.baryon_flag]), mass=s.mass,
              ic_ids=ids, ic_pos=pos, ic_vel=vel, ic_phi=s.get_field("phi"), ic_chi=s.get_field("chi"), ic_Bi=s.get_field("Bi"), ic_BiFT=s.get_field("BiFT"),
              ic_state=np.array([st["a"], st["tau"], st["dtau"], st["dtau_old"], st["cycle"], st["maxvel"][0], st["maxvel"][1]]), ncycles=np.array([NCYCLES]))
    for _ in range(NCYCLES):
        s.step()
    ids, pos, vel = s.get_particles(0)
    o = np.argsort(ids)
    st = s.state()
    ic.update(end_ids=ids[o], end_pos=pos[o], end_vel=vel[o], end_phi=s.get_field("phi"), end_chi=s.get_field("chi"), end_Bi=s.get_field("Bi"),
              end_state=np.array([st["a"], st["tau"], st["dtau"], st["dtau_old"], st["cycle"], st["maxvel"][0], st["T00hom"]]))
    path = os.path.join(HERE, "shipped_settings_N16.npz")
    np.savez_compressed(path, **ic)
    print("wrote", path, os.path.getsize(path), "bytes")
