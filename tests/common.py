"""Seeded synthetic inputs shared by the CPU and GPU tests."""
import numpy as np

# shipped settings.ini:28-32 through parser.hpp:1658-1748
H = 0.67556
C_PLANCK_LAW = 4.48147e-7


def shipped_cosmology():
    Omega_b = 0.022032 / H / H
    Omega_cdm = 0.12038 / H / H
    Omega_g = 2.7255
    Omega_g = Omega_g * Omega_g / H
    Omega_g = Omega_g * Omega_g * C_PLANCK_LAW
    Omega_ur = 3.046 * (7. / 8.) * (4. / 11.) ** (4. / 3.) * Omega_g
    Omega_rad = Omega_g + Omega_ur
    Omega_m = Omega_cdm + Omega_b
    Omega_fld = 0.0
    Omega_Lambda = 1. - Omega_m - Omega_rad - Omega_fld
    #       Omega_cdm, Omega_b, Omega_m, Omega_Lambda, Omega_fld, w0_fld, wa_fld, Omega_g, Omega_ur, Omega_rad, h
    return np.array([Omega_cdm, Omega_b, Omega_m, Omega_Lambda, Omega_fld, -1.0, 0.0, Omega_g, Omega_ur, Omega_rad, H])


def shipped_settings(z_in=100.0):
    # boxsize, Cf, steplimit, z_in, z_relax  (settings.ini:37-44; relaxation redshift defaults to z_in, parser.hpp:1246)
    return np.array([320.0, 48.0, 0.04, z_in, z_in])


def gaussian_field(rng, N, ncomp=1, rms=1.0, slope=-3.0):
    """Real Gaussian random field with power ~ k^slope, zero mean, given rms; shape (ncomp,N,N,N)."""
    k = np.fft.fftfreq(N) * N
    kx = np.fft.rfftfreq(N) * N
    k2 = k[:, None, None] ** 2 + k[None, :, None] ** 2 + kx[None, None, :] ** 2
    k2[0, 0, 0] = 1.0
    amp = k2 ** (slope / 4.0)
    amp[0, 0, 0] = 0.0
    out = np.empty((ncomp, N, N, N))
    for c in range(ncomp):
        w = rng.standard_normal((N, N, N // 2 + 1)) + 1j * rng.standard_normal((N, N, N // 2 + 1))
        f = np.fft.irfftn(w * amp, s=(N, N, N))
        out[c] = f * (rms / f.std())
    return out


def metric_fields(rng, N, a=0.02):
    """phi, chi, Bi of cosmological magnitude: phi ~1e-5, chi ~1e-7, B (stored as a^2 N B) ~1e-8 a^2 N."""
    phi = gaussian_field(rng, N, 1, 1e-5)
    chi = gaussian_field(rng, N, 1, 1e-7)
    Bi = gaussian_field(rng, N, 3, 1e-8 * a * a * N)
    return phi, chi, Bi


def quasi_uniform_particles(rng, N, per_dim=None, sigma=0.05, a=0.02, qscale=1e-3):
    """One particle per cell at the centre + Gaussian displacement (sigma cells); q ~ N(0, (qscale a)^2)."""
    n = per_dim or N
    g = (np.arange(n) + 0.5) / n
    z, y, x = np.meshgrid(g, g, g, indexing="ij")
    pos = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    pos += rng.standard_normal(pos.shape) * (sigma / N)
    pos -= np.floor(pos)
    pos[pos >= 1.0] = 0.0
    vel = rng.standard_normal(pos.shape) * (qscale * a)
    ids = np.arange(len(pos), dtype=np.int64)
    return ids, np.ascontiguousarray(pos), np.ascontiguousarray(vel)


def clustered_particles(rng, N, npart, a=0.5, qscale=1e-2, nblobs=24):
    """Clustered (low-z like) distribution: Gaussian blobs of different widths + a uniform floor."""
    centres = rng.random((nblobs, 3))
    widths = 10 ** rng.uniform(-2.2, -1.0, nblobs)
    which = rng.integers(0, nblobs, npart)
    pos = centres[which] + rng.standard_normal((npart, 3)) * widths[which, None]
    floor = rng.random(npart) < 0.2
    pos[floor] = rng.random((floor.sum(), 3))
    pos -= np.floor(pos)
    pos[pos >= 1.0] = 0.0
    vel = rng.standard_normal(pos.shape) * (qscale * a)
    ids = np.arange(npart, dtype=np.int64)
    return ids, np.ascontiguousarray(pos), np.ascontiguousarray(vel)


def rel_linf(a, b):
    """relative L-infinity error of a against b (the north-star field metric)."""
    scale = np.abs(b).max()
    return np.abs(a - b).max() / (scale if scale > 0 else 1.0)


def to_cplx(F):
    return F[..., 0] + 1j * F[..., 1]


def ncdm_model(cosmo, masses=(0.1, 0.2)):
    """Two massive-neutrino species (BASELINE config 4): m [eV], T_ncdm [T_cmb] and Omega_ncdm = m / (93.14 eV h^2);
    returns (cosmo with Omega_m, Omega_Lambda adjusted as parser.hpp:1700-1748 does, m, T, Omega)."""
    m = np.array(masses, dtype=np.float64)
    T = np.full(len(m), 0.71611)
    Om = m / 93.14 / H / H
    c = np.array(cosmo, dtype=np.float64)
    c[2] = c[0] + c[1] + Om.sum()
    c[3] = 1.0 - c[2] - c[9] - c[4]
    return c, m, T, Om


def thermal_particles(rng, N, per_dim, a, qrms):
    """hot species: coarser lattice (per_dim^3) with large Gaussian displacements and momenta q ~ N(0, qrms^2)"""
    g = (np.arange(per_dim) + 0.5) / per_dim
    z, y, x = np.meshgrid(g, g, g, indexing="ij")
    pos = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1) + rng.standard_normal((per_dim ** 3, 3)) * (0.4 / N)
    pos -= np.floor(pos)
    pos[pos >= 1.0] = 0.0
    vel = rng.standard_normal(pos.shape) * qrms
    return np.arange(len(pos), dtype=np.int64), np.ascontiguousarray(pos), np.ascontiguousarray(vel)


def counts_bit_exact(N, dev_pos, ref_pos, dev_counts, cell_index):
    """Unconditional form of the bit-exact binning contract (BASELINE.json): the device's per-cell counts must equal the
    histogram of floor(pos/dx).  A particle may be filed one cell away from the reference's only if the two positions sit on
    either side of a cell face within rounding (|pos N - face| < 1e-9); such particles are re-filed as the device has them and
    the counts must then be identical.  Returns (ok, flipped) -- never skips the integer comparison."""
    import numpy as np
    dev_cells, ref_cells = np.floor(dev_pos * N), np.floor(ref_pos * N)
    flipped = np.any(dev_cells != ref_cells, axis=1)
    if flipped.any():
        s = ref_pos[flipped] * N
        near_face = np.abs(s - np.rint(s)).min(axis=1) < 1e-9
        if not near_face.all():
            return False, int(flipped.sum())
    expect_pos = np.where(flipped[:, None], dev_pos, ref_pos)
    _, expect = cell_index(N, expect_pos)
    return bool(np.array_equal(np.asarray(dev_counts).ravel(), np.asarray(expect).ravel())), int(flipped.sum())
