"""Host-side pieces of the product's basic IC generator and its settings.ini reader against the compiled reference
(no device needed): same template, same 27-site convolution kernel, same Gaussian realisation from the same seed --
bit for bit, since they are integer / single-precision-intermediate work -- and the same parsed settings."""
import numpy as np
import pytest

import common


@pytest.fixture(scope="module")
def shipped(ref, tmp_path_factory):
    d = tmp_path_factory.mktemp("shipped")
    return d, ref.dump_shipped_files(d)


def _fold(k27, N):
    full = np.zeros((N, N, N))
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                full[dz % N, dy % N, dx % N] += k27[dz + 1, dy + 1, dx + 1]
    return full


def test_template_loader(gevb_host, shipped):
    d, tmpl = shipped
    mine = gevb_host.ic_load_template(d / "sc1_crystal.dat")
    assert mine.dtype == np.float32 and np.array_equal(mine, tmpl)
    with pytest.raises(gevb_host.GevbError):
        gevb_host.ic_load_template(d / "settings.ini")          # not a Gadget-2 file


@pytest.mark.parametrize("N,tile", [(16, 4), (24, 6), (64, 16)])
def test_cic_kernel_shipped_template(gevb_host, ref, shipped, N, tile):
    """generateCICKernel (ic_basic.hpp:737-1052) for the shipped sc1 template, and the standard kernel"""
    _, tmpl = shipped
    assert np.array_equal(_fold(gevb_host.ic_cic_kernel(N, tmpl, tile), N), ref.generateCICKernel(N, tmpl, tile))
    assert np.array_equal(_fold(gevb_host.ic_cic_kernel(N), N), ref.generateCICKernel(N))


def test_cic_kernel_random_template(gevb_host, ref):
    """a random template exercises every octant and the reference's mixed float / double operand types"""
    rng = np.random.default_rng(3)
    for N, tile, n in ((16, 4, 50), (32, 2, 200), (16, 1, 30)):
        t = rng.random((n, 3)).astype(np.float32)
        assert np.array_equal(_fold(gevb_host.ic_cic_kernel(N, t, tile), N), ref.generateCICKernel(N, t, tile))


@pytest.mark.parametrize("N", [8, 16, 32])
def test_gaussian_realisation(gevb_host, ref, N):
    """generateDisplacementField (ic_basic.hpp:1090-1379): Threefry stream positions per row, zeroed modes (k-sphere / cube),
    deconvolution on / off, the modified-kernel coefficient -- identical bits from the same seed"""
    rng = np.random.default_rng(N)
    pot = rng.standard_normal((N, N, N // 2 + 1, 2)) + 2.0
    x = np.geomspace(0.5, 400, 60)
    y = np.exp(-x / 50) * x ** -1.5
    for seed, ksphere, deconv, coeff in ((42, 1, 1, 0.0), (42, 0, 1, 0.3), (7, 1, 0, 0.0)):
        a = ref.generateDisplacementField(pot, coeff, x, y, seed, ksphere, deconv)
        b = gevb_host.ic_displacement_field(pot, coeff, x, y, seed, ksphere, deconv)
        assert np.array_equal(a, b), (seed, ksphere, deconv, coeff)
    # another seed gives another field
    assert not np.array_equal(gevb_host.ic_displacement_field(pot, 0.0, x, y, 1, 1, 1), gevb_host.ic_displacement_field(pot, 0.0, x, y, 2, 1, 1))


def test_settings_reader(gevb_host, shipped):
    """the shipped settings.ini through the product's reader: the values the reference's parser derives (parser.hpp:759-1800)"""
    d, _ = shipped
    st = gevb_host.settings_read(d / "settings.ini")
    assert (st.ngrid, st.gr_flag, st.vector_flag, st.baryon_flag, st.seed, st.ksphere, st.correct_displacement) == (64, 1, 0, 2, 42, 1, 1)
    assert st.tiling[0] == 16 and st.numbins == 1024 and st.movelimit == 64.0
    assert np.allclose(np.array(st.cosmo), common.shipped_cosmology(), rtol=1e-15, atol=0)
    assert np.allclose([st.boxsize, st.Cf, st.steplimit, st.z_in, st.z_relax], common.shipped_settings(), rtol=0, atol=0)
    assert list(st.z_pk)[:st.num_pk] == [50.0, 30.0, 10.0, 3.0, 1.0, 0.0] and list(st.z_snapshot)[:st.num_snapshot] == [30.0, 10.0, 3.0, 0.0]
    assert st.pk_mask == (1 | 8 | 2 | 128) and st.snapshot_mask == (1 | 8 | 512)           # phi, B, chi, hij ; phi, B, Gadget2
    assert st.tk_file == b"class_tk.dat" and st.template_file[0].value == b"sc1_crystal.dat"
    assert (st.A_s, st.n_s, st.k_pivot) == (2.215e-9, 0.9619, 0.05)
    # overrides replace the file's lines; unsupported configurations are refused, not half-read
    st2 = gevb_host.settings_read(d / "settings.ini", "Ngrid = 16\ntiling factor = 4\ngravity theory = Newton\nvector method = elliptic\nbaryon treatment = sample\nseed = 7")
    assert (st2.ngrid, st2.tiling[0], st2.tiling[1], st2.gr_flag, st2.vector_flag, st2.baryon_flag, st2.seed, st2.movelimit) == (16, 4, 4, 0, 1, 1, 7, 16.0)
    for bad in ("mPk file = pk.dat", "IC generator = read from disk", "m_ncdm = 0.1, 0.2"):
        with pytest.raises(gevb_host.GevbError):
            gevb_host.settings_read(d / "settings.ini", bad)
    with pytest.raises(gevb_host.GevbError):
        gevb_host.settings_read(d / "no_such_file.ini")


# ---- the product's settings reader against the reference's own parser on variants of the shipped file -----------------------------
_BASE = """
template file = sc1_crystal.dat
Tk file = class_tk.dat
IC generator = basic
boxsize = 320.0
Ngrid = 64
tiling factor = 16
initial redshift = 100.0
Courant factor = 48.0
time step limit = 0.04
seed = 42
"""

_VARIANTS = {
    "minimal: parser defaults everywhere": _BASE,
    "comments, blank lines, odd spacing": "# a comment line\n\n   boxsize=200.5   # trailing comment\nNgrid   =   32\n" + _BASE.replace("boxsize = 320.0", "").replace("Ngrid = 64", "")
                                          + "\ngravity theory = Newton   # N-body gauge\n",
    "first match wins": _BASE + "seed = 7\nNgrid = 128\n",
    "physical densities and radiation": _BASE + "h = 0.7\nomega_b = 0.0224\nomega_cdm = 0.12\nT_cmb = 2.7255\nN_ur = 2.0328\n",
    "fractional densities": _BASE + "h = 0.6774\nOmega_b = 0.0486\nOmega_cdm = 0.2589\nOmega_g = 5.4e-5\nOmega_ur = 3.7e-5\n",
    "N_eff alias": _BASE + "N_eff = 3.5\n",
    "dark-energy fluid": _BASE + "Omega_fld = 0.65\nw0_fld = -0.9\nwa_fld = 0.1\n",
    "elliptic, sample, outputs": _BASE + "vector method = elliptic\nbaryon treatment = sample\nsnapshot redshifts = 0, 3, 30, 10\nsnapshot outputs = phi, B, Gadget2, chi, hij, T00\n"
                                 "Pk redshifts = 1, 50, 0\nPk outputs = phi, chi, hij, B, T00, delta\nPk bins = 512\ntracer factor = 4\n",
    "hybrid baryons, relaxation, k-domain": _BASE + "baryon treatment = hybrid\nrelaxation redshift = 50\nk-domain = cube\ncorrect displacement = no\nmove limit = 8\n"
                                            "A_s = 2.1e-9\nn_s = 0.965\nk_pivot = 0.002\n",
    "two tiling factors and templates": _BASE.replace("tiling factor = 16", "tiling factor = 16, 8").replace("template file = sc1_crystal.dat", "template file = sc1_crystal.dat, sc1_crystal.dat")
                                        + "baryon treatment = sample\n",
}


@pytest.mark.parametrize("name", list(_VARIANTS))
def test_settings_reader_matches_the_reference_parser(gevb_host, ref, tmp_path, name):
    """gevb_settings_read (host/settings.cpp) against loadParameterFile + parseMetadata of the compiled reference (parser.hpp:122,759)
    on the same text: every value a basic run consults, bit for bit (the derived cosmology included)."""
    text = _VARIANTS[name]
    f = tmp_path / "settings.ini"
    f.write_text(text)
    want = ref.parse_settings(text)
    st = gevb_host.settings_read(f)
    got = {"ngrid": st.ngrid, "gr_flag": st.gr_flag, "vector_flag": st.vector_flag, "baryon_flag": st.baryon_flag, "seed": st.seed, "ksphere": st.ksphere,
           "correct_displacement": st.correct_displacement, "tiling0": st.tiling[0], "tracer0": st.tracer_factor[0], "numbins": st.numbins,
           "pk_mask": st.pk_mask, "snapshot_mask": st.snapshot_mask, "num_pk": st.num_pk, "num_snapshot": st.num_snapshot}
    if st.baryon_flag in (1, 3):                                     # the second species exists: its tiling and tracer factors count
        got["tiling1"], got["tracer1"] = st.tiling[1], st.tracer_factor[1]
    for k, v in got.items():
        assert v == want[k], (name, k, v, want[k])
    for k in "boxsize Cf steplimit movelimit z_in z_relax A_s n_s k_pivot".split():
        assert getattr(st, k) == want[k], (name, k, getattr(st, k), want[k])
    assert np.array_equal(np.array(st.cosmo), want["cosmo"]), (name, np.array(st.cosmo) - want["cosmo"])
    assert np.array_equal(np.array(st.z_pk)[:st.num_pk], want["z_pk"]) and np.array_equal(np.array(st.z_snapshot)[:st.num_snapshot], want["z_snapshot"]), name
