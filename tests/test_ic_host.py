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
