"""oracle/oracle.py -- TEST INFRASTRUCTURE (ctypes loader for the CPU checkers).

Two libraries share one flat-array C interface (layouts in oracle/ref_driver.cpp):

* ``ref``  -> oracle/_ref/libgevref.so : the reference's own gevolution.hpp /
  tools.hpp / background.hpp compiled against the single-rank LATfield2 shim
  (built from /root/reference in the build container; travels prebuilt to the
  GPU box).
* ``ora``  -> oracle/libgev_oracle.so  : the independent plain-C restatement
  (oracle/gev_oracle.c), each function citing the reference file:line it follows.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs may
import this module.  The product (gevolution-1.2_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libgevref.so")
ORA_SO = os.path.join(HERE, "libgev_oracle.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_vp = C.c_void_p
_d = C.c_double
_i = C.c_int
_l = C.c_long

# name -> (restype, argtypes); pointers that may be NULL are declared c_void_p
_SIGS = {
    "fft_forward": (None, [_i, _i, _dp, _dp]),
    "fft_backward": (None, [_i, _i, _dp, _dp]),
    "prepareFTsource_scalar": (None, [_i, _dp, _dp, _dp, _d, _dp, _d, _d, _d]),
    "prepareFTsource_tensor": (None, [_i, _dp, _dp, _dp, _d]),
    "solveModifiedPoissonFT": (None, [_i, _dp, _dp, _d, _d]),
    "projectFTscalar": (None, [_i, _dp, _dp, _i]),
    "evolveFTvector": (None, [_i, _dp, _dp, _d]),
    "projectFTvector": (None, [_i, _dp, _dp, _d, _d]),
    "projectFTtensor": (None, [_i, _dp, _dp]),
    "projection_T00": (None, [_i, _l, _dp, _dp, _d, _d, _vp, _d, _dp]),
    "projection_T0i": (None, [_i, _l, _dp, _dp, _d, _vp, _d, _dp]),
    "projection_Tij": (None, [_i, _l, _dp, _dp, _d, _d, _vp, _d, _dp]),
    "scalarProjectionCIC": (None, [_i, _l, _dp, _d, _dp]),
    "updateVel": (_d, [_i, _l, _dp, _dp, _i, _d, _vp, _vp, _vp, _i, _dp]),
    "moveParticles": (None, [_i, _l, _dp, _dp, _i, _d, _vp, _vp, _vp, _i, _dp]),
    "cell_index": (None, [_i, _l, _dp, _i32p, _u32p]),
    "extractPowerSpectrum": (None, [_i, _i, _i, _dp, _dp, _dp, _dp, _dp, _i32p, _i, _i, _i]),
    "computeVectorDiagnostics": (None, [_i, _dp, _dp, _dp]),
    "computeTensorDiagnostics": (None, [_i, _dp, _dp, _dp, _dp]),
    "Hconf": (_d, [_d, _d, _dp]),
    "rungekutta4bg": (_d, [_d, _d, _dp, _d]),
    "particleHorizon": (_d, [_d, _d, _dp]),
    "sim_create": (_vp, [_i, _i, _i, _dp, _dp]),
    "sim_destroy": (None, [_vp]),
    "sim_set_particles": (None, [_vp, _i, _l, _i64p, _dp, _dp, _d]),
    "sim_set_field": (None, [_vp, _i, _dp]),
    "sim_get_field": (None, [_vp, _i, _dp]),
    "sim_num_particles": (_l, [_vp, _i]),
    "sim_get_particles": (None, [_vp, _i, _i64p, _dp, _dp]),
    "sim_get_state": (None, [_vp, _dp]),
    "sim_set_state": (None, [_vp, _dp]),
    "sim_get_timers": (None, [_vp, _dp]),
    "sim_step": (None, [_vp]),
    "writePowerSpectrum": (None, [_dp, _dp, _dp, _dp, _i32p, _i, _d, _d, C.c_char_p, C.c_char_p, _d, _d]),
    "sim_create_from_settings": (_vp, [_i, _i, _i, C.c_char_p]),
    "sim_get_config": (None, [_vp, _dp, _dp, _i32p, _dp]),
    "sim_save_gadget2": (_i, [_vp, _i, C.c_char_p, _i, _d, _d]),
    "sim_run": (_i, [_vp, _dp, _i, _i, _i, C.c_char_p, _dp, _i, _i, C.c_char_p, _i, _i32p]),
    "sim_set_ncdm": (None, [_vp, _i, _dp, _dp, _dp, _dp, _dp, _d, _d]),
    "sim_set_ncdm_maxvel": (None, [_vp, _dp]),
    "sim_get_ncdm_state": (None, [_vp, _dp, _i32p]),
    "bg_ncdm": (_d, [_d, _dp, _i, _dp, _dp, _dp]),
    "generateCICKernel": (None, [_i, _l, _vp, _i, _dp]),
    "generateDisplacementField": (None, [_i, _dp, _d, _i, _dp, _dp, C.c_uint, _i, _i]),
    "dump_shipped_files": (_l, [C.c_char_p, np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS"), _l]),
    "parse_settings": (C.c_int, [C.c_char_p, np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS"), np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")]),
}

FIELD_IDS = {"phi": 0, "chi": 1, "Bi": 2, "source": 3, "Sij": 4, "scalarFT": 10, "BiFT": 11, "SijFT": 12}
FIELD_COMPS = {"phi": 1, "chi": 1, "Bi": 3, "source": 1, "Sij": 6, "scalarFT": 1, "BiFT": 3, "SijFT": 6}


def build(force=False):
    """Compile the checkers (make). `_ref` is rebuilt only where /root/reference exists."""
    args = ["make", "-C", HERE, "-s"] + (["-B"] if force else [])
    subprocess.run(args, check=True)


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Thin, typed view of one checker library; ``prefix`` is 'ref' or 'ora'."""

    def __init__(self, path, prefix):
        self.path, self.prefix = path, prefix
        self.lib = C.CDLL(path)
        self.fn = {}
        for name, (res, args) in _SIGS.items():
            f = getattr(self.lib, f"{prefix}_{name}", None)
            if f is None:
                continue
            f.restype, f.argtypes = res, args
            self.fn[name] = f
        d = getattr(self.lib, f"{prefix}_describe")
        d.restype = C.c_char_p
        self.description = d().decode()

    # ---- FFT -------------------------------------------------------------
    def fft_forward(self, real):
        real = np.ascontiguousarray(real, dtype=np.float64)
        nc, N = real.shape[0], real.shape[-1]
        out = np.zeros((nc, N, N, N // 2 + 1, 2))
        self.fn["fft_forward"](N, nc, real, out)
        return out

    def fft_backward(self, cplx):
        cplx = np.ascontiguousarray(cplx, dtype=np.float64)
        nc, N = cplx.shape[0], cplx.shape[1]
        out = np.zeros((nc, N, N, N))
        self.fn["fft_backward"](N, nc, cplx, out)
        return out

    # ---- real-space source preparation --------------------------------------
    def prepareFTsource_scalar(self, phi, chi, source, bgmodel, coeff, coeff2, coeff3):
        N = phi.shape[-1]
        out = np.zeros_like(source)
        self.fn["prepareFTsource_scalar"](N, phi, chi, source, bgmodel, out, coeff, coeff2, coeff3)
        return out

    def prepareFTsource_tensor(self, phi, Tij, coeff):
        N = phi.shape[-1]
        out = np.zeros_like(Tij)
        self.fn["prepareFTsource_tensor"](N, phi, Tij, out, coeff)
        return out

    # ---- Fourier-space kernels -----------------------------------------------
    def solveModifiedPoissonFT(self, src, coeff, modif=0.0):
        out = np.zeros_like(src)
        self.fn["solveModifiedPoissonFT"](src.shape[1], src, out, coeff, modif)
        return out

    def projectFTscalar(self, SijFT, chiFT=None):
        N = SijFT.shape[1]
        add = 0 if chiFT is None else 1
        out = np.zeros((1,) + SijFT.shape[1:]) if chiFT is None else chiFT.copy()
        self.fn["projectFTscalar"](N, SijFT, out, add)
        return out

    def evolveFTvector(self, SijFT, BiFT, a2dtau):
        out = BiFT.copy()
        self.fn["evolveFTvector"](SijFT.shape[1], SijFT, out, a2dtau)
        return out

    def projectFTvector(self, SiFT, coeff=1.0, modif=0.0):
        out = np.zeros_like(SiFT)
        self.fn["projectFTvector"](SiFT.shape[1], SiFT, out, coeff, modif)
        return out

    def projectFTtensor(self, SijFT):
        out = np.zeros_like(SijFT)
        self.fn["projectFTtensor"](SijFT.shape[1], SijFT, out)
        return out

    # ---- projections -----------------------------------------------------------
    def projection_T00(self, N, pos, vel, mass, a, phi=None, coeff=1.0):
        out = np.zeros((1, N, N, N))
        self.fn["projection_T00"](N, len(pos), pos, vel, mass, a, _opt(phi), coeff, out)
        return out

    def projection_T0i(self, N, pos, vel, mass, phi=None, coeff=1.0):
        out = np.zeros((3, N, N, N))
        self.fn["projection_T0i"](N, len(pos), pos, vel, mass, _opt(phi), coeff, out)
        return out

    def projection_Tij(self, N, pos, vel, mass, a, phi=None, coeff=1.0):
        out = np.zeros((6, N, N, N))
        self.fn["projection_Tij"](N, len(pos), pos, vel, mass, a, _opt(phi), coeff, out)
        return out

    def scalarProjectionCIC(self, N, pos, mass):
        out = np.zeros((1, N, N, N))
        self.fn["scalarProjectionCIC"](N, len(pos), pos, mass, out)
        return out

    # ---- kick / drift ------------------------------------------------------------
    def updateVel(self, N, pos, vel, kind, dtau, phi, chi, Bi, nfields, params):
        vel = vel.copy()
        params = np.ascontiguousarray(params, dtype=np.float64)
        r = self.fn["updateVel"](N, len(pos), pos, vel, kind, dtau, _opt(phi), _opt(chi), _opt(Bi), nfields, params)
        return vel, r

    def moveParticles(self, N, pos, vel, kind, dtau, phi, chi, Bi, nfields, params):
        pos = pos.copy()
        params = np.ascontiguousarray(params, dtype=np.float64)
        self.fn["moveParticles"](N, len(pos), pos, vel, kind, dtau, _opt(phi), _opt(chi), _opt(Bi), nfields, params)
        return pos

    def cell_index(self, N, pos):
        cell = np.zeros(len(pos), dtype=np.int32)
        counts = np.zeros(N * N * N, dtype=np.uint32)
        self.fn["cell_index"](N, len(pos), pos, cell, counts)
        return cell, counts

    # ---- analysis ------------------------------------------------------------------
    def extractPowerSpectrum(self, fldFT, numbins, symmetric=False, deconvolve=True, ktype=1):
        N = fldFT.shape[1]
        kbin, power, ksc, psc = (np.zeros(numbins) for _ in range(4))
        occ = np.zeros(numbins, dtype=np.int32)
        self.fn["extractPowerSpectrum"](N, fldFT.shape[0], int(symmetric), fldFT, kbin, power, ksc, psc, occ, numbins, int(deconvolve), ktype)
        return kbin, power, ksc, psc, occ

    def writePowerSpectrum(self, kbin, power, kscatter, pscatter, occupation, rescalek, rescalep, filename, description, a, z_target=-1.0):
        """the reference's own file writer (tools.hpp:268-346)"""
        arr = [np.ascontiguousarray(v, dtype=np.float64) for v in (kbin, power, kscatter, pscatter)]
        occ = np.ascontiguousarray(occupation, dtype=np.int32)
        self.fn["writePowerSpectrum"](*arr, occ, len(occ), rescalek, rescalep, filename.encode(), description.encode(), a, z_target)

    def computeVectorDiagnostics(self, Bi):
        a, b = np.zeros(1), np.zeros(1)
        self.fn["computeVectorDiagnostics"](Bi.shape[-1], Bi, a, b)
        return a[0], b[0]

    def computeTensorDiagnostics(self, hij):
        a, b, c = np.zeros(1), np.zeros(1), np.zeros(1)
        self.fn["computeTensorDiagnostics"](hij.shape[-1], hij, a, b, c)
        return a[0], b[0], c[0]

    # ---- background ------------------------------------------------------------------
    def Hconf(self, a, fourpiG, cosmo):
        return self.fn["Hconf"](a, fourpiG, np.ascontiguousarray(cosmo, dtype=np.float64))

    def rungekutta4bg(self, a, fourpiG, cosmo, dtau):
        return self.fn["rungekutta4bg"](a, fourpiG, np.ascontiguousarray(cosmo, dtype=np.float64), dtau)

    def bg_ncdm(self, a, cosmo, m_ncdm, T_ncdm, Omega_ncdm):
        arr = [np.ascontiguousarray(v, dtype=np.float64) for v in (m_ncdm, T_ncdm, Omega_ncdm)]
        return self.fn["bg_ncdm"](a, np.ascontiguousarray(cosmo, dtype=np.float64), len(arr[0]), *arr)

    def particleHorizon(self, a, fourpiG, cosmo):
        return self.fn["particleHorizon"](a, fourpiG, np.ascontiguousarray(cosmo, dtype=np.float64))

    # ---- pieces of the IC generator (compiled reference only) ---------------------------
    def generateCICKernel(self, N, pcldata=None, numtile=1):
        """generateCICKernel (ic_basic.hpp:737): the kernel field [z][y][x]; pcldata (n, 3) float32 template or None (standard kernel)"""
        out = np.zeros((N, N, N))
        if pcldata is None:
            self.fn["generateCICKernel"](N, 0, None, 1, out)
        else:
            p = np.ascontiguousarray(pcldata, dtype=np.float32)
            self.fn["generateCICKernel"](N, len(p), p.ctypes.data_as(C.c_void_p), numtile, out)
        return out

    def generateDisplacementField(self, potFT, coeff, spline_x, spline_y, seed, ksphere=0, deconvolve_f=1):
        """generateDisplacementField (ic_basic.hpp:1090) applied to a copy of potFT [kz][ky][kx][2]"""
        out = np.ascontiguousarray(potFT, dtype=np.float64).copy()
        x, y = np.ascontiguousarray(spline_x, dtype=np.float64), np.ascontiguousarray(spline_y, dtype=np.float64)
        self.fn["generateDisplacementField"](out.shape[0], out, coeff, len(x), x, y, seed, ksphere, deconvolve_f)
        return out

    def dump_shipped_files(self, directory):
        """settings.ini, class_tk.dat, sc1_crystal.dat of the reference written to `directory`; returns the template's
        positions as loadHomogeneousTemplate (ic_basic.hpp:191) delivers them"""
        buf = np.zeros((4096, 3), dtype=np.float32)
        n = self.fn["dump_shipped_files"](str(directory).encode(), buf, len(buf))
        return buf[:n].copy()

    def parse_settings(self, text):
        """the reference's parser (parser.hpp:122,759) on a settings text (compiled reference only): dict of what it derives"""
        ints, dbl = np.zeros(16, dtype=np.int32), np.zeros(84)
        n = self.fn["parse_settings"](text.encode(), ints, dbl)
        if n <= 0:
            raise RuntimeError("reference parser read no parameters")
        names = ("ngrid gr_flag vector_flag baryon_flag seed ksphere correct_displacement tiling0 tiling1 tracer0 tracer1 numbins "
                 "pk_mask snapshot_mask num_pk num_snapshot").split()
        out = {k: int(v) for k, v in zip(names, ints)}
        for k, v in zip("boxsize Cf steplimit movelimit z_in z_relax A_s n_s k_pivot".split(), dbl[:9]):
            out[k] = float(v)
        out["cosmo"] = dbl[9:20].copy()
        out["z_pk"] = dbl[20:20 + out["num_pk"]].copy()
        out["z_snapshot"] = dbl[52:52 + out["num_snapshot"]].copy()
        return out

    # ---- stateful simulation -----------------------------------------------------------
    def sim(self, N, gr_flag, vector_flag, dsettings, cosmo):
        return Sim(self, N, gr_flag, vector_flag, dsettings, cosmo)

    def sim_from_settings(self, ngrid=0, tiling=0, seed=-1, overrides=""):
        """the reference's shipped settings.ini run from its own seed: its parser and generateIC_basic (compiled reference
        only); ngrid / tiling factor / seed override the file's values when given, `overrides` holds further
        "key = value" lines that replace the file's lines of the same key (e.g. "gravity theory = Newton")"""
        h = self.fn["sim_create_from_settings"](ngrid, tiling, seed, overrides.encode())
        if not h:
            raise RuntimeError("reference IC generation failed")
        s = Sim.__new__(Sim)
        s.o, s.h = self, h
        cosmo, ds, flags, mass = np.zeros(11), np.zeros(5), np.zeros(4, dtype=np.int32), np.zeros(2)
        self.fn["sim_get_config"](h, cosmo, ds, flags, mass)
        s.N, s.cosmo, s.dsettings, s.gr_flag, s.vector_flag, s.baryon_flag, s.mass = int(flags[0]), cosmo, ds, int(flags[1]), int(flags[2]), int(flags[3]), mass
        return s


class Sim:
    """main.cpp:217-246 state + one-cycle stepping, on the CPU checker."""

    def __init__(self, ora, N, gr_flag, vector_flag, dsettings, cosmo):
        self.o, self.N = ora, N
        self.h = ora.fn["sim_create"](N, gr_flag, vector_flag,
                                      np.ascontiguousarray(dsettings, dtype=np.float64),
                                      np.ascontiguousarray(cosmo, dtype=np.float64))

    def close(self):
        if self.h:
            self.o.fn["sim_destroy"](self.h)
            self.h = None

    def set_particles(self, species, ids, pos, vel, mass):
        self.o.fn["sim_set_particles"](self.h, species, len(ids), np.ascontiguousarray(ids, dtype=np.int64),
                                       np.ascontiguousarray(pos), np.ascontiguousarray(vel), mass)

    def set_field(self, name, data):
        self.o.fn["sim_set_field"](self.h, FIELD_IDS[name], np.ascontiguousarray(data, dtype=np.float64))

    def get_field(self, name):
        N, nc = self.N, FIELD_COMPS[name]
        out = np.zeros((nc, N, N, N)) if FIELD_IDS[name] < 10 else np.zeros((nc, N, N, N // 2 + 1, 2))
        self.o.fn["sim_get_field"](self.h, FIELD_IDS[name], out)
        return out

    def get_particles(self, species=0):
        n = self.o.fn["sim_num_particles"](self.h, species)
        ids = np.zeros(n, dtype=np.int64)
        pos, vel = np.zeros((n, 3)), np.zeros((n, 3))
        self.o.fn["sim_get_particles"](self.h, species, ids, pos, vel)
        return ids, pos, vel

    def state(self):
        s = np.zeros(9)
        self.o.fn["sim_get_state"](self.h, s)
        return dict(a=s[0], tau=s[1], dtau=s[2], dtau_old=s[3], cycle=int(s[4]), maxvel=(s[5], s[6]), T00hom=s[7], fourpiG=s[8])

    def set_state(self, a, tau, dtau, dtau_old, cycle, maxvel=(0.0, 0.0)):
        self.o.fn["sim_set_state"](self.h, np.array([a, tau, dtau, dtau_old, cycle, maxvel[0], maxvel[1]], dtype=np.float64))

    def set_ncdm(self, m_ncdm, T_ncdm, Omega_ncdm, z_switch_deltancdm, z_switch_Bncdm, z_switch_linearchi, movelimit):
        arr = [np.ascontiguousarray(v, dtype=np.float64) for v in (m_ncdm, T_ncdm, Omega_ncdm, z_switch_deltancdm, z_switch_Bncdm)]
        self.o.fn["sim_set_ncdm"](self.h, len(arr[0]), *arr, z_switch_linearchi, movelimit)

    def set_ncdm_maxvel(self, maxvel):
        v = np.zeros(4); v[:len(maxvel)] = maxvel
        self.o.fn["sim_set_ncdm_maxvel"](self.h, v)

    def ncdm_state(self):
        v, n = np.zeros(4), np.zeros(4, dtype=np.int32)
        self.o.fn["sim_get_ncdm_state"](self.h, v, n)
        return v, n

    def run(self, z_pk, pk_mask, numbins, pk_prefix, z_snapshot, tracer_factor, snap_prefix, max_cycles=100000):
        """the reference's main loop with its spectrum / Gadget-2 outputs at the requested redshifts (main.cpp:605-693)"""
        zp, zs = np.ascontiguousarray(z_pk, dtype=np.float64), np.ascontiguousarray(z_snapshot, dtype=np.float64)
        out = np.zeros(3, dtype=np.int32)
        self.o.fn["sim_run"](self.h, zp, len(zp), pk_mask, numbins, pk_prefix.encode(), zs, len(zs), tracer_factor, snap_prefix.encode(), max_cycles, out)
        return tuple(int(v) for v in out)

    def save_gadget2(self, species, filename, tracer_factor=1, dtau_pos=0.0, dtau_vel=0.0):
        """the reference's own saveGadget2 (Particles_gevolution.hpp:30-251)"""
        return self.o.fn["sim_save_gadget2"](self.h, species, filename.encode(), tracer_factor, dtau_pos, dtau_vel)

    def timers(self):
        t = np.zeros(6)
        self.o.fn["sim_get_timers"](self.h, t)
        return dict(projection=t[0], gravity_solver=t[1], fft=t[2], update_q=t[3], move_particles=t[4], cycle=t[5])

    def step(self):
        self.o.fn["sim_step"](self.h)


def load_ref():
    """The compiled reference (None if oracle/_ref was never built here)."""
    return Oracle(REF_SO, "ref") if os.path.exists(REF_SO) else None


def load_ora():
    if not os.path.exists(ORA_SO):
        build()
    return Oracle(ORA_SO, "ora")


def best():
    """Preferred checker: compiled reference when present, else the C restatement."""
    return load_ref() or load_ora()
