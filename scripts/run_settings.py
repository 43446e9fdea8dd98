#!/usr/bin/env python
"""Run a gevolution settings.ini with the product alone (no oracle anywhere): settings reader -> basic IC generator ->
main loop with the file's power-spectrum and Gadget-2 outputs.  This is BASELINE config 1 when given the reference's
shipped settings.ini, class_tk.dat and sc1_crystal.dat (paths in the file are relative to the working directory, as
for the reference's binary).

    python scripts/run_settings.py settings.ini ["key = value" overrides ...]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gevolution-1.2_b200"))
import gevb  # noqa: E402


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    st = gevb.settings_read(sys.argv[1], "\n".join(sys.argv[2:]))
    out = st.output_path.decode()
    if out and not os.path.isdir(out):
        os.makedirs(out, exist_ok=True)
    ctx = gevb.Context(st.ngrid, device=0)
    t0 = time.perf_counter()
    sim = gevb.sim_from_settings(ctx, st)
    t1 = time.perf_counter()
    n = sim.pcls(0).count()
    print(f"initial conditions: Ngrid {st.ngrid}, {n} cdm particles, seed {st.seed}, z_in {st.z_in} ({t1 - t0:.2f} s)")
    cycles, npk, nsnap = sim.run_settings(st)
    ctx.sync()
    t2 = time.perf_counter()
    s = sim.state()
    print(f"{cycles} cycles to z = {1.0 / s['a'] - 1.0:.4f}: {npk} spectra sets, {nsnap} snapshots in {t2 - t1:.2f} s "
          f"({n * cycles / (t2 - t1):.3e} particle-steps/s including outputs)")
    sim.close(); ctx.close()


if __name__ == "__main__":
    main()
