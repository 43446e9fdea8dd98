"""Regenerate tests/golden/hotpath_N8.npz from the compiled reference (oracle/_ref).

Run in the build container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py
The reference ships no golden vectors of its own; these are
outputs of the reference's own gevolution.hpp / tools.hpp compiled against the
single-rank LATfield2 shim, on the seeded inputs of tests/golden_cases.py.
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
