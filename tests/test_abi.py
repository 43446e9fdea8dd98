"""C-ABI checks that need no GPU: the library loads, exports every symbol that
include/gevb.h declares, and refuses to work without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gevb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gevb_[a-zA-Z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(gevb):
    L = gevb.lib()
    declared = _declared_symbols()
    assert len(declared) >= 55
    missing = [s for s in declared if not hasattr(L, s)]
    assert missing == []
    assert sorted(gevb.SYMBOLS) == declared, "gevb/__init__.py SYMBOLS out of sync with include/gevb.h"


def test_version_and_error_strings(gevb):
    L = gevb.lib()
    assert b"gevb" in L.gevb_version()
    assert isinstance(L.gevb_last_error(), bytes)


def test_argument_validation_needs_no_device(gevb):
    L = gevb.lib()
    h = ctypes.c_void_p()
    assert L.gevb_ctx_create(ctypes.byref(h), 7, 0, 0, 1, None) != 0        # odd Ngrid
    assert b"Ngrid" in L.gevb_last_error()
    assert L.gevb_ctx_create(ctypes.byref(h), 16, 0, 0, 3, None) != 0       # not divisible
    assert L.gevb_ctx_create(ctypes.byref(h), 16, 0, 1, 2, None) != 0       # nccl id missing


def test_documented_knobs_exist(gevb):
    """every tuning knob include/gevb.h documents is known to the library (no device needed), anything else is refused"""
    import re
    hdr = open(os.path.join(ROOT, "include", "gevb.h")).read()
    block = hdr[hdr.index("kernel-variant knobs"):hdr.index("int gevb_tuning")]
    knobs = re.findall(r'^ \*   "([a-z0-9_]+)"', block, flags=re.M)
    assert len(knobs) >= 10 and "deposit_variant" in knobs and "geodesic_tma" in knobs
    lib = gevb.lib()
    for k in knobs:
        assert lib.gevb_tuning(k.encode(), 0) == 0, k
    assert lib.gevb_tuning(b"no_such_knob", 1) != 0 and b"unknown knob" in lib.gevb_last_error()
    # back to the defaults for whatever runs next in this process
    for k, v in (("geodesic_variant", 5), ("fft_exchange", 1), ("fft_overlap", 2), ("fft_decomposed", 1), ("deposit_variant", 18), ("fft_l2_planes", 0),
                 ("rebin_variant", 2), ("peer_comm", 1), ("geodesic_tma", 1), ("tma_l2_promotion", 0), ("fft_xpass", 0)):
        assert lib.gevb_tuning(k.encode(), v) == 0


def test_no_cpu_fallback(gevb):
    """without a CUDA device context creation must fail loudly, not fall back"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gevb.GevbError) as e:
        gevb.Context(16)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_reference_oracle():
    """the product tree must not import, link or mention the oracle"""
    prod = os.path.join(ROOT, "gevolution-1.2_b200")
    for dirpath, _, files in os.walk(prod):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".py", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"


def test_host_background_matches_reference(gevb, ref):
    """product host scalars (gevolution-1.2_b200/host/background.hpp) against the reference's background.hpp,
    with and without ncdm species; the horizon only to the reference's own quadrature request (1e-7)"""
    import numpy as np
    import common
    cosmo = common.shipped_cosmology()
    fourpiG = 1.5 * 320.0 ** 2 / 2997.92458 ** 2
    for a in (0.0099, 0.05, 0.3, 1.0):
        g = gevb.background_eval(cosmo, a, fourpiG, 0.01)
        assert abs(g["Hconf"] - ref.Hconf(a, fourpiG, cosmo)) <= 1e-14 * g["Hconf"]
        assert abs(g["a_next"] - ref.rungekutta4bg(a, fourpiG, cosmo, 0.01)) <= 1e-14 * a
        assert abs(g["particleHorizon"] - ref.particleHorizon(a, fourpiG, cosmo)) <= 1e-6 * g["particleHorizon"]
    c2, m, T, Om = common.ncdm_model(cosmo)
    for a in (0.0099, 0.05, 0.3, 1.0):
        g = gevb.background_eval(c2, a, fourpiG, 0.01, m, T, Om)
        r = ref.bg_ncdm(a, c2, m, T, Om)
        assert r > 0 and abs(g["bg_ncdm"] - r) <= 1e-12 * r
    # non-relativistic limit: bg_ncdm -> Omega_ncdm (background.hpp:78 with w >> 1)
    g = gevb.background_eval(c2, 1.0, fourpiG, 0.0, m, T, Om)
    assert abs(g["bg_ncdm"] - Om.sum()) < 2e-4 * Om.sum()


def test_power_spectrum_file_format_needs_no_device(gevb, tmp_path):
    """host logic of the spectra writer (tools.hpp:268-346): header, one line per occupied bin, rescaling, and the
    EXACT_OUTPUT_REDSHIFTS interpolation between the stored spectrum and the current one"""
    import numpy as np
    k = np.array([1.0, 2.0, 3.0, 4.0]); p = np.array([10.0, 20.0, 0.0, 40.0])
    ks = np.array([0.1, 0.2, 0.0, 0.4]); ps = np.array([1.0, 2.0, 0.0, 4.0])
    occ = np.array([3, 5, 0, 7], dtype=np.int32)
    fn = str(tmp_path / "pk.dat")
    gevb.writePowerSpectrum(k, p, ks, ps, occ, 2.0, 5.0, fn, "power spectrum of phi", 1.0 / 11.0, 9.5)   # z = 10, above the target
    lines = open(fn).read().splitlines()
    assert lines[:3] == ["# power spectrum of phi", "# redshift z=10.000000", "# k              Pk             sigma(k)       sigma(Pk)      count"]
    rows = np.array([[float(v) for v in ln.split()] for ln in lines[3:]])
    assert rows.shape == (3, 5)                                             # the empty bin is skipped
    assert np.allclose(rows[:, 0], k[occ > 0] / 2.0) and np.allclose(rows[:, 1], p[occ > 0] / 5.0)
    assert np.allclose(rows[:, 2], ks[occ > 0] / 2.0) and np.allclose(rows[:, 3], ps[occ > 0] / 5.0 / np.sqrt(occ[occ > 0]), rtol=1e-6)
    assert np.array_equal(rows[:, 4].astype(int), occ[occ > 0])
    # next cycle, z = 9 < 9.5: weight = (10 - 9.5) / (1 + 10 - 10) = 0.5 -> mean of the stored and the new spectrum, labelled z = 9.5
    gevb.writePowerSpectrum(k, 3.0 * p, ks, ps, occ, 2.0, 5.0, fn, "power spectrum of phi", 1.0 / 10.0, 9.5)
    lines = open(fn).read().splitlines()
    assert lines[1] == "# redshift z=9.500000"
    rows = np.array([[float(v) for v in ln.split()] for ln in lines[3:]])
    assert np.allclose(rows[:, 1], 2.0 * p[occ > 0] / 5.0, rtol=1e-6)
    # a target that is not reached yet (or disabled): plain overwrite
    gevb.writePowerSpectrum(k, p, ks, ps, occ, 2.0, 5.0, fn, "power spectrum of phi", 1.0 / 10.0, -1.0)
    rows = np.array([[float(v) for v in ln.split()] for ln in open(fn).read().splitlines()[3:]])
    assert np.allclose(rows[:, 1], p[occ > 0] / 5.0)


def test_reference_shaped_loop_compiles_against_the_drop_in_header():
    """a main.cpp-shaped loop (same names, argument order and types as the reference's statements) compiles
    against include/gevolution_b200.hpp -- syntax and overload resolution only, nothing is executed"""
    import subprocess
    src = os.path.join(ROOT, "tests", "dropin", "main_loop_shape.cpp")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"), src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_power_spectrum_file_equals_the_reference_writer(gevb, ref, tmp_path):
    """byte for byte: gevb_writePowerSpectrum against the reference's own writePowerSpectrum (tools.hpp:268-346), first
    write, interpolated second write past the target redshift, and a later plain overwrite"""
    import numpy as np
    rng = np.random.default_rng(3)
    nb = 24
    k = np.sort(rng.random(nb)) * 50; p = rng.random(nb) * 1e-9; ks = rng.random(nb) * 0.1; ps = rng.random(nb) * 1e-10
    occ = rng.integers(0, 40, nb).astype(np.int32); occ[[3, 11]] = 0
    fa, fb = str(tmp_path / "a.dat"), str(tmp_path / "b.dat")
    calls = [(p, 1.0 / 11.0, 9.5), (2.5 * p, 1.0 / 10.2, 9.5), (0.7 * p, 1.0 / 4.0, 3.0), (0.9 * p, 1.0 / 3.9, 3.0), (p, 0.5, -1.0)]
    for pw, a, zt in calls:
        gevb.writePowerSpectrum(k, pw, ks, ps, occ, 320.0, 7.3e4, fa, "power spectrum of phi", a, zt)
        ref.writePowerSpectrum(k, pw, ks, ps, occ, 320.0, 7.3e4, fb, "power spectrum of phi", a, zt)
        assert open(fa, "rb").read() == open(fb, "rb").read(), (a, zt)
