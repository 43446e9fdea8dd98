"""gevb -- ctypes view of libgevb.so, the B200-native gevolution hot path.

Thin host-side mirror of include/gevb.h used by tests/, bench.py and
__graft_entry__.py.  All compute happens in the hand-written sm_100a kernels
behind the C ABI; there is NO CPU or PyTorch fallback: a missing library or a
missing GPU raises.

Names follow the reference's vocabulary (Field, PlanFFT, Particles,
projection_T00_project, updateVel, moveParticles ...; reference gevolution.hpp
and main.cpp:372-879).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GEVB_LIB") or os.path.join(os.path.dirname(_HERE), "libgevb.so")   # GEVB_LIB: an alternative build of the library

REAL, CPLX = 0, 1
FFT_FORWARD, FFT_BACKWARD = 1, -1
UPDATE_Q, UPDATE_Q_NEWTON, DISPLACE_PCLS_IC_BASIC, INITIALIZE_Q_IC_BASIC = 0, 1, 2, 3

FIELD_IDS = {"phi": 0, "chi": 1, "Bi": 2, "source": 3, "Sij": 4, "scalarFT": 10, "BiFT": 11, "SijFT": 12}
FIELD_COMPS = {"phi": 1, "chi": 1, "Bi": 3, "source": 1, "Sij": 6, "scalarFT": 1, "BiFT": 3, "SijFT": 6}

# every symbol include/gevb.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "gevb_last_error", "gevb_version", "gevb_tuning", "gevb_nccl_unique_id", "gevb_ctx_create", "gevb_slab_geometry", "gevb_ctx_destroy", "gevb_ctx_sync",
    "gevb_ctx_geometry", "gevb_ctx_stream", "gevb_ctx_launch_count",
    "gevb_ctx_timing", "gevb_ctx_timing_read", "gevb_timing_num_classes", "gevb_timing_class_name", "gevb_parallel_sum", "gevb_parallel_max",
    "gevb_field_create", "gevb_field_destroy", "gevb_field_upload", "gevb_field_download", "gevb_field_components",
    "gevb_field_device_ptr", "gevb_projection_init", "gevb_field_updateHalo", "gevb_projection_comm", "gevb_field_sum",
    "gevb_field_add_constant", "gevb_field_scale", "gevb_field_ctx", "gevb_field_save_raw", "gevb_field_load_raw", "gevb_plan_create", "gevb_plan_destroy", "gevb_plan_execute", "gevb_plan_set_preserve_input", "gevb_pcls_create",
    "gevb_pcls_destroy", "gevb_pcls_reset", "gevb_pcls_add", "gevb_pcls_count", "gevb_pcls_download", "gevb_pcls_cell_counts", "gevb_pcls_mass", "gevb_brick_dims",
    "gevb_projection_T00_project", "gevb_projection_T0i_project", "gevb_projection_Tij_project",
    "gevb_scalarProjectionCIC_project", "gevb_projection_T00_Tij_project", "gevb_prepareFTsource_scalar",
    "gevb_prepareFTsource_scalar_sum", "gevb_prepareFTsource_tensor", "gevb_prepareFTsource_scalar_fft", "gevb_prepareFTsource_tensor_fft", "gevb_solveModifiedPoissonFT", "gevb_projectFTscalar", "gevb_evolveFTvector",
    "gevb_projectFTscalar_evolveFTvector", "gevb_projectFTvector", "gevb_projectFTtensor", "gevb_updateVel", "gevb_moveParticles", "gevb_moveParticles_max", "gevb_kick_drift",
    "gevb_extractPowerSpectrum", "gevb_writePowerSpectrum", "gevb_pcls_saveGadget2", "gevb_pcls_gadget2_arrays", "gevb_pcls_ctx", "gevb_ctx_ranks",
    "gevb_sim_write_spectra", "gevb_sim_save_gadget2", "gevb_sim_write_field_snapshot", "gevb_sim_hibernate", "gevb_sim_restore", "gevb_sim_run", "gevb_background_eval", "gevb_sim_create", "gevb_sim_destroy", "gevb_sim_set_ncdm", "gevb_sim_set_ncdm_maxvel", "gevb_sim_get_ncdm_state", "gevb_sim_set_particles", "gevb_sim_set_field",
    "gevb_sim_get_field", "gevb_sim_field", "gevb_sim_pcls", "gevb_sim_get_state", "gevb_sim_set_state",
    "gevb_sim_set_fused", "gevb_sim_step",
    "gevb_settings_read", "gevb_sim_create_from_settings", "gevb_sim_run_settings", "gevb_ic_cic_kernel", "gevb_ic_displacement_field",
    "gevb_ic_load_template", "gevb_field_set_sites",
]

_lib = None


class GevbError(RuntimeError):
    pass


def lib():
    """Load libgevb.so (in-tree build); raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GevbError(f"{LIB_PATH} not found: build it with `make -C {os.path.dirname(LIB_PATH)}` "
                            "(or __graft_entry__.build()); there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L):
    vp, d, i, i64 = C.c_void_p, C.c_double, C.c_int, C.c_int64
    pd = C.POINTER(C.c_double)
    L.gevb_last_error.restype = C.c_char_p
    L.gevb_version.restype = C.c_char_p
    L.gevb_ctx_stream.restype = vp
    L.gevb_ctx_stream.argtypes = [vp]
    L.gevb_ctx_launch_count.restype = i64
    L.gevb_ctx_launch_count.argtypes = [vp]
    L.gevb_field_device_ptr.restype = vp
    L.gevb_field_device_ptr.argtypes = [vp]
    L.gevb_sim_field.restype = vp
    L.gevb_sim_field.argtypes = [vp, i]
    L.gevb_sim_pcls.restype = vp
    L.gevb_sim_pcls.argtypes = [vp, i]
    L.gevb_pcls_mass.restype = d
    L.gevb_pcls_mass.argtypes = [vp]
    L.gevb_brick_dims.restype = None
    L.gevb_brick_dims.argtypes = [C.POINTER(i)] * 2
    L.gevb_timing_class_name.restype = C.c_char_p
    L.gevb_timing_class_name.argtypes = [i]
    L.gevb_timing_num_classes.restype = i
    L.gevb_timing_num_classes.argtypes = []
    sig = {
        "gevb_nccl_unique_id": [vp], "gevb_tuning": [C.c_char_p, i],
        "gevb_ctx_create": [C.POINTER(vp), i, i, i, i, vp],
        "gevb_ctx_destroy": [vp], "gevb_ctx_sync": [vp],
        "gevb_slab_geometry": [i, i, i] + [C.POINTER(i)] * 4,
        "gevb_ctx_geometry": [vp] + [C.POINTER(i)] * 5,
        "gevb_ctx_timing": [vp, i], "gevb_ctx_timing_read": [vp, pd, C.POINTER(C.c_int64)],
        "gevb_parallel_sum": [vp, pd, i], "gevb_parallel_max": [vp, pd, i],
        "gevb_field_create": [vp, C.POINTER(vp), i, i, i],
        "gevb_field_destroy": [vp], "gevb_field_upload": [vp, vp], "gevb_field_download": [vp, vp],
        "gevb_field_components": [vp], "gevb_projection_init": [vp], "gevb_field_updateHalo": [vp],
        "gevb_projection_comm": [vp], "gevb_field_sum": [vp, i, pd], "gevb_field_add_constant": [vp, i, d],
        "gevb_plan_create": [C.POINTER(vp), vp, vp], "gevb_plan_destroy": [vp], "gevb_plan_execute": [vp, i], "gevb_plan_set_preserve_input": [vp, i],
        "gevb_pcls_create": [vp, C.POINTER(vp), d], "gevb_pcls_destroy": [vp], "gevb_pcls_reset": [vp, d],
        "gevb_pcls_add": [vp, i64, vp, vp, vp], "gevb_pcls_count": [vp, C.POINTER(i64)],
        "gevb_pcls_download": [vp, vp, vp, vp], "gevb_pcls_cell_counts": [vp, vp],
        "gevb_projection_T00_project": [vp, vp, d, vp, d], "gevb_projection_T0i_project": [vp, vp, vp, d],
        "gevb_projection_Tij_project": [vp, vp, d, vp, d], "gevb_scalarProjectionCIC_project": [vp, vp],
        "gevb_projection_T00_Tij_project": [vp, vp, vp, d, vp, d],
        "gevb_prepareFTsource_scalar": [vp, vp, vp, d, vp, d, d, d], "gevb_prepareFTsource_scalar_sum": [vp, vp, vp, d, vp, d, d, d, C.POINTER(C.c_double)], "gevb_prepareFTsource_tensor": [vp, vp, vp, d],
        "gevb_prepareFTsource_scalar_fft": [vp, vp, vp, d, d, d, d, C.POINTER(C.c_double)], "gevb_prepareFTsource_tensor_fft": [vp, vp, d],
        "gevb_solveModifiedPoissonFT": [vp, vp, d, d], "gevb_projectFTscalar": [vp, vp, i],
        "gevb_evolveFTvector": [vp, vp, d], "gevb_projectFTscalar_evolveFTvector": [vp, vp, vp, d], "gevb_projectFTvector": [vp, vp, d, d], "gevb_projectFTtensor": [vp, vp],
        "gevb_updateVel": [vp, i, d, C.POINTER(vp), i, pd, pd],
        "gevb_moveParticles": [vp, i, d, C.POINTER(vp), i, pd],
        "gevb_moveParticles_max": [vp, i, d, C.POINTER(vp), i, pd, pd],
        "gevb_kick_drift": [vp, i, d, i, pd, d, i, pd, C.POINTER(vp), pd],
        "gevb_extractPowerSpectrum": [vp, vp, vp, vp, vp, vp, i, i, i],
        "gevb_sim_create": [C.POINTER(vp), vp, i, i, pd, pd], "gevb_sim_destroy": [vp], "gevb_background_eval": [pd, i, pd, pd, pd, d, d, d, pd], "gevb_sim_set_ncdm": [vp, i, pd, pd, pd, pd, pd, d, d], "gevb_sim_set_ncdm_maxvel": [vp, pd], "gevb_sim_get_ncdm_state": [vp, pd, C.POINTER(i)],
        "gevb_sim_set_particles": [vp, i, i64, vp, vp, vp, d], "gevb_sim_set_field": [vp, i, vp],
        "gevb_sim_get_field": [vp, i, vp], "gevb_sim_get_state": [vp, pd], "gevb_sim_set_state": [vp, pd],
        "gevb_sim_set_fused": [vp, i], "gevb_sim_step": [vp],
        "gevb_writePowerSpectrum": [vp, vp, vp, vp, vp, i, d, d, C.c_char_p, C.c_char_p, d, d],
        "gevb_pcls_saveGadget2": [vp, C.c_char_p, vp, i, d, d, vp],
        "gevb_pcls_gadget2_arrays": [vp, d, d, i, d, d, vp, vp, vp, vp, C.POINTER(i64)],
        "gevb_ctx_ranks": [vp, C.POINTER(i), C.POINTER(i)],
        "gevb_sim_write_spectra": [vp, C.c_char_p, i, i, i, d],
        "gevb_sim_save_gadget2": [vp, i, C.c_char_p, i, d, d],
        "gevb_sim_hibernate": [vp, C.c_char_p], "gevb_sim_restore": [vp, C.c_char_p], "gevb_sim_write_field_snapshot": [vp, C.c_char_p, i],
        "gevb_field_scale": [vp, d], "gevb_field_save_raw": [vp, C.c_char_p], "gevb_field_load_raw": [vp, C.c_char_p],
        "gevb_sim_run": [vp, pd, i, i, i, C.c_char_p, pd, i, i, C.c_char_p, i, C.POINTER(i)],
        "gevb_settings_read": [C.c_char_p, C.c_char_p, vp], "gevb_sim_create_from_settings": [C.POINTER(vp), vp, vp],
        "gevb_sim_run_settings": [vp, vp, i, C.POINTER(i)],
        "gevb_ic_cic_kernel": [i, i64, vp, i, pd], "gevb_ic_displacement_field": [i, pd, d, i, pd, pd, C.c_uint, i, i],
        "gevb_ic_load_template": [C.c_char_p, C.POINTER(i64), C.POINTER(C.POINTER(C.c_float))],
        "gevb_field_set_sites": [vp, i, i, vp, vp],
    }
    for name, args in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = C.c_int, args


def writePowerSpectrum(kbin, power, kscatter, pscatter, occupation, rescalek, rescalep, filename, description, a, z_target=-1.0):
    occ = np.ascontiguousarray(occupation, dtype=np.int32)
    arrs = [np.ascontiguousarray(v, dtype=np.float64) for v in (kbin, power, kscatter, pscatter)]
    _ck(lib().gevb_writePowerSpectrum(*[_ptr(v) for v in arrs], _ptr(occ), len(occ), rescalek, rescalep, filename.encode(), description.encode(), a, z_target), "gevb_writePowerSpectrum")


def slab_geometry(ngrid, rank, nranks):
    """(z0, nz_local, ky0, nky_local) of `rank` -- host arithmetic only, no device"""
    v = [C.c_int() for _ in range(4)]
    _ck(lib().gevb_slab_geometry(ngrid, rank, nranks, *[C.byref(x) for x in v]), "gevb_slab_geometry")
    return tuple(x.value for x in v)


def tuning(knob, value):
    """kernel-variant knob for ablation runs (gevb_tuning)"""
    _ck(lib().gevb_tuning(knob.encode(), int(value)), "gevb_tuning")


def _ck(status, what):
    if status != 0:
        raise GevbError(f"{what}: {lib().gevb_last_error().decode()}")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _darr(v):
    a = np.ascontiguousarray(v, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def brick_dims():
    b, sp = (C.c_int * 3)(), (C.c_int * 3)()
    lib().gevb_brick_dims(b, sp)
    return tuple(b), tuple(sp)


def storage_key(N, nzl, cx, cy, czl):
    """Sort key of the device particle order (include/gevb.h, gevb_brick_dims): super-brick, brick, cell -- all z-major."""
    (bx, by, bz), (sx, sy, sz) = brick_dims()
    nsx, nsy = -(-N // (bx * sx)), -(-N // (by * sy))
    cx, cy, czl = (np.asarray(v, dtype=np.int64) for v in (cx, cy, czl))
    kx, ky, kz = cx // bx, cy // by, czl // bz                        # brick coordinates
    sup = ((kz // sz) * nsy + ky // sy) * nsx + kx // sx
    loc = ((kz % sz) * sy + ky % sy) * sx + kx % sx
    cell = ((czl % bz) * by + cy % by) * bx + cx % bx
    return (sup * (sx * sy * sz) + loc) * (bx * by * bz) + cell


def nccl_unique_id():
    buf = (C.c_char * 128)()
    _ck(lib().gevb_nccl_unique_id(C.cast(buf, C.c_void_p)), "gevb_nccl_unique_id")
    return bytes(buf)


class Context:
    """Lattice geometry + device + communicator of one rank (Lattice + parallel)."""

    def __init__(self, ngrid, device=0, rank=0, nranks=1, nccl_id=None):
        self.h = C.c_void_p()
        idbuf = None if nccl_id is None else C.create_string_buffer(nccl_id, 128)
        _ck(lib().gevb_ctx_create(C.byref(self.h), ngrid, device, rank, nranks, C.cast(idbuf, C.c_void_p) if idbuf else None), "gevb_ctx_create")
        g = [C.c_int() for _ in range(5)]
        lib().gevb_ctx_geometry(self.h, *[C.byref(x) for x in g])
        self.N, self.z0, self.nzl, self.ky0, self.nkyl = (x.value for x in g)
        self.rank, self.nranks, self.device = rank, nranks, device
        self.nh = self.N // 2 + 1

    def close(self):
        if self.h:
            lib().gevb_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def sync(self):
        _ck(lib().gevb_ctx_sync(self.h), "gevb_ctx_sync")

    @property
    def stream(self):
        return lib().gevb_ctx_stream(self.h)

    @property
    def launches(self):
        return lib().gevb_ctx_launch_count(self.h)

    def timing(self, enable):
        _ck(lib().gevb_ctx_timing(self.h, int(enable)), "gevb_ctx_timing")

    def timing_read(self):
        """{entry point: (milliseconds, calls)} accumulated since the last read"""
        n = lib().gevb_timing_num_classes()
        ms, cnt = np.zeros(n), np.zeros(n, dtype=np.int64)
        _ck(lib().gevb_ctx_timing_read(self.h, ms.ctypes.data_as(C.POINTER(C.c_double)), cnt.ctypes.data_as(C.POINTER(C.c_int64))), "gevb_ctx_timing_read")
        return {lib().gevb_timing_class_name(k).decode(): (float(ms[k]), int(cnt[k])) for k in range(n) if cnt[k] > 0}

    def parallel_sum(self, v):
        a, p = _darr(v)
        a = a.copy()
        _ck(lib().gevb_parallel_sum(self.h, a.ctypes.data_as(C.POINTER(C.c_double)), a.size), "parallel.sum")
        return a

    def parallel_max(self, v):
        a, p = _darr(v)
        a = a.copy()
        _ck(lib().gevb_parallel_max(self.h, a.ctypes.data_as(C.POINTER(C.c_double)), a.size), "parallel.max")
        return a

    # shapes of the host arrays exchanged by upload/download
    def real_shape(self, ncomp):
        return (ncomp, self.nzl, self.N, self.N)

    def cplx_shape(self, ncomp):
        if self.nranks == 1:
            return (ncomp, self.N, self.N, self.nh, 2)
        return (ncomp, self.nkyl, self.nh, self.N, 2)     # [c][ky][kx][kz] on slabs


class Field:
    def __init__(self, ctx, kind=REAL, ncomp=1, symmetric=False, data=None, handle=None):
        self.ctx, self.kind, self.ncomp = ctx, kind, ncomp
        self.owned = handle is None
        if handle is None:
            self.h = C.c_void_p()
            _ck(lib().gevb_field_create(ctx.h, C.byref(self.h), kind, ncomp, int(symmetric)), "gevb_field_create")
        else:
            self.h = C.c_void_p(handle)
        if data is not None:
            self.upload(data)

    @property
    def shape(self):
        return self.ctx.real_shape(self.ncomp) if self.kind == REAL else self.ctx.cplx_shape(self.ncomp)

    def close(self):
        if self.owned and self.h:
            lib().gevb_field_destroy(self.h)
        self.h = C.c_void_p()

    def upload(self, data):
        a = np.ascontiguousarray(data, dtype=np.float64)
        assert a.shape == self.shape, f"host array {a.shape} != field {self.shape}"
        _ck(lib().gevb_field_upload(self.h, _ptr(a)), "gevb_field_upload")
        return self

    def download(self):
        out = np.empty(self.shape, dtype=np.float64)
        _ck(lib().gevb_field_download(self.h, _ptr(out)), "gevb_field_download")
        return out

    def projection_init(self):
        _ck(lib().gevb_projection_init(self.h), "projection_init")

    def updateHalo(self):
        _ck(lib().gevb_field_updateHalo(self.h), "updateHalo")
        return self

    def projection_comm(self):
        _ck(lib().gevb_projection_comm(self.h), "projection_comm")

    def sum(self, comp=0):
        out = C.c_double()
        _ck(lib().gevb_field_sum(self.h, comp, C.byref(out)), "gevb_field_sum")
        return out.value

    def add_constant(self, value, comp=0):
        _ck(lib().gevb_field_add_constant(self.h, comp, value), "gevb_field_add_constant")


class PlanFFT:
    def __init__(self, real_field, cplx_field):
        self.h = C.c_void_p()
        self.real, self.cplx = real_field, cplx_field
        _ck(lib().gevb_plan_create(C.byref(self.h), real_field.h, cplx_field.h), "gevb_plan_create")

    def execute(self, direction):
        _ck(lib().gevb_plan_execute(self.h, direction), "PlanFFT.execute")

    def preserve_input(self, keep):
        _ck(lib().gevb_plan_set_preserve_input(self.h, int(bool(keep))), "PlanFFT.preserve_input")

    def close(self):
        if self.h:
            lib().gevb_plan_destroy(self.h)
            self.h = C.c_void_p()


def _handles(fields):
    arr = (C.c_void_p * 3)()
    for k, f in enumerate(fields or []):
        arr[k] = f.h if f is not None else None
    return arr


class Particles:
    """Cell-sorted FP64 SoA particle container (Particles_gevolution)."""

    def __init__(self, ctx, mass, handle=None):
        self.ctx = ctx
        self.owned = handle is None
        if handle is None:
            self.h = C.c_void_p()
            _ck(lib().gevb_pcls_create(ctx.h, C.byref(self.h), mass), "gevb_pcls_create")
        else:
            self.h = C.c_void_p(handle)

    def close(self):
        if self.owned and self.h:
            lib().gevb_pcls_destroy(self.h)
        self.h = C.c_void_p()

    def add(self, ids, pos, vel):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        vel = np.ascontiguousarray(vel, dtype=np.float64)
        assert pos.shape == (len(ids), 3) and vel.shape == (len(ids), 3)
        _ck(lib().gevb_pcls_add(self.h, len(ids), _ptr(ids), _ptr(pos), _ptr(vel)), "gevb_pcls_add")
        return self

    def count(self):
        n = C.c_int64()
        _ck(lib().gevb_pcls_count(self.h, C.byref(n)), "gevb_pcls_count")
        return n.value

    def download(self):
        n = self.count()
        ids = np.empty(n, dtype=np.int64)
        pos, vel = np.empty((n, 3)), np.empty((n, 3))
        _ck(lib().gevb_pcls_download(self.h, _ptr(ids), _ptr(pos), _ptr(vel)), "gevb_pcls_download")
        return ids, pos, vel

    def cell_counts(self):
        out = np.empty(self.ctx.nzl * self.ctx.N * self.ctx.N, dtype=np.uint32)
        _ck(lib().gevb_pcls_cell_counts(self.h, _ptr(out)), "gevb_pcls_cell_counts")
        return out

    # ---- projections ------------------------------------------------------------
    def projection_T00_project(self, T00, a=1.0, phi=None, coeff=1.0):
        _ck(lib().gevb_projection_T00_project(self.h, T00.h, a, phi.h if phi else None, coeff), "projection_T00_project")

    def projection_T0i_project(self, T0i, phi=None, coeff=1.0):
        _ck(lib().gevb_projection_T0i_project(self.h, T0i.h, phi.h if phi else None, coeff), "projection_T0i_project")

    def projection_Tij_project(self, Tij, a=1.0, phi=None, coeff=1.0):
        _ck(lib().gevb_projection_Tij_project(self.h, Tij.h, a, phi.h if phi else None, coeff), "projection_Tij_project")

    def projection_T00_Tij_project(self, T00, Tij, a, phi, coeff=1.0):
        _ck(lib().gevb_projection_T00_Tij_project(self.h, T00.h, Tij.h, a, phi.h, coeff), "projection_T00_Tij_project")

    def scalarProjectionCIC_project(self, rho):
        _ck(lib().gevb_scalarProjectionCIC_project(self.h, rho.h), "scalarProjectionCIC_project")

    # ---- geodesic updates ----------------------------------------------------------
    def updateVel(self, fn, dtau, fields, nfields, params):
        pa, pp = _darr(params)
        out = C.c_double()
        _ck(lib().gevb_updateVel(self.h, fn, dtau, _handles(fields), nfields, pp, C.byref(out)), "updateVel")
        return out.value

    def moveParticles(self, fn, dtau, fields, nfields, params):
        pa, pp = _darr(params)
        _ck(lib().gevb_moveParticles(self.h, fn, dtau, _handles(fields), nfields, pp), "moveParticles")

    def moveParticles_max(self, fn, dtau, fields, nfields, params=(1.0, 1.0)):
        """moveParticles with the callback's reduction output (largest displacement of displace_pcls_ic_basic)"""
        pa, pp = _darr(params)
        out = C.c_double()
        _ck(lib().gevb_moveParticles_max(self.h, fn, dtau, _handles(fields), nfields, pp, C.byref(out)), "moveParticles")
        return out.value

    def kick_drift(self, fn, dtau_kick, nf_kick, params_kick, dtau_drift, nf_drift, params_drift, fields):
        ka, kp = _darr(params_kick)
        da, dp = _darr(params_drift)
        out = C.c_double()
        _ck(lib().gevb_kick_drift(self.h, fn, dtau_kick, nf_kick, kp, dtau_drift, nf_drift, dp, _handles(fields), C.byref(out)), "kick_drift")
        return out.value


# ---- free functions under the reference's names -----------------------------------
def prepareFTsource_scalar(phi, chi, source, bgmodel, result, coeff, coeff2, coeff3):
    _ck(lib().gevb_prepareFTsource_scalar(phi.h, chi.h, source.h, bgmodel, result.h, coeff, coeff2, coeff3), "prepareFTsource")


def prepareFTsource_tensor(phi, Tij, Sij, coeff):
    _ck(lib().gevb_prepareFTsource_tensor(phi.h, Tij.h, Sij.h, coeff), "prepareFTsource")


def prepareFTsource_scalar_fft(phi, chi, plan_source, bgmodel, coeff, coeff2, coeff3, want_sum=False):
    """prepareFTsource + plan.execute(FFT_FORWARD) in one call; returns the sum of the incoming source if asked for"""
    out = C.c_double(0.0)
    _ck(lib().gevb_prepareFTsource_scalar_fft(phi.h, chi.h, plan_source.h, bgmodel, coeff, coeff2, coeff3, C.byref(out) if want_sum else None), "prepareFTsource")
    return out.value if want_sum else None


def prepareFTsource_tensor_fft(phi, plan_Sij, coeff):
    _ck(lib().gevb_prepareFTsource_tensor_fft(phi.h, plan_Sij.h, coeff), "prepareFTsource")


def solveModifiedPoissonFT(sourceFT, potFT, coeff, modif=0.0):
    _ck(lib().gevb_solveModifiedPoissonFT(sourceFT.h, potFT.h, coeff, modif), "solveModifiedPoissonFT")


def projectFTscalar(SijFT, chiFT, add=0):
    _ck(lib().gevb_projectFTscalar(SijFT.h, chiFT.h, add), "projectFTscalar")


def evolveFTvector(SijFT, BiFT, a2dtau):
    _ck(lib().gevb_evolveFTvector(SijFT.h, BiFT.h, a2dtau), "evolveFTvector")


def projectFTscalar_evolveFTvector(SijFT, chiFT, BiFT, a2dtau):
    _ck(lib().gevb_projectFTscalar_evolveFTvector(SijFT.h, chiFT.h, BiFT.h, a2dtau), "projectFTscalar_evolveFTvector")


def projectFTvector(SiFT, BiFT, coeff=1.0, modif=0.0):
    _ck(lib().gevb_projectFTvector(SiFT.h, BiFT.h, coeff, modif), "projectFTvector")


def projectFTtensor(SijFT, hijFT):
    _ck(lib().gevb_projectFTtensor(SijFT.h, hijFT.h), "projectFTtensor")


def extractPowerSpectrum(fldFT, numbins, deconvolve=True, ktype=1):
    kbin, power, ksc, psc = (np.zeros(numbins) for _ in range(4))
    occ = np.zeros(numbins, dtype=np.int32)
    _ck(lib().gevb_extractPowerSpectrum(fldFT.h, _ptr(kbin), _ptr(power), _ptr(ksc), _ptr(psc), _ptr(occ), numbins, int(deconvolve), ktype), "extractPowerSpectrum")
    return kbin, power, ksc, psc, occ


def background_eval(cosmo, a, fourpiG, dtau=0.0, m_ncdm=(), T_ncdm=(), Omega_ncdm=()):
    """host-side background (no device): dict(Hconf, bg_ncdm, particleHorizon, a_next)"""
    keep = [_darr(v) for v in (cosmo, m_ncdm, T_ncdm, Omega_ncdm)]
    out = np.zeros(4)
    _ck(lib().gevb_background_eval(keep[0][1], len(m_ncdm), keep[1][1], keep[2][1], keep[3][1], a, fourpiG, dtau, out.ctypes.data_as(C.POINTER(C.c_double))), "gevb_background_eval")
    return dict(Hconf=out[0], bg_ncdm=out[1], particleHorizon=out[2], a_next=out[3])


class Sim:
    """State of main.cpp:217-246 on the device + one-cycle stepping (host loop is C++)."""

    def __init__(self, ctx, gr_flag, vector_flag, dsettings, cosmo):
        self.ctx = ctx
        self.h = C.c_void_p()
        ds, dsp = _darr(dsettings)
        co, cop = _darr(cosmo)
        _ck(lib().gevb_sim_create(C.byref(self.h), ctx.h, gr_flag, vector_flag, dsp, cop), "gevb_sim_create")

    def close(self):
        if self.h:
            lib().gevb_sim_destroy(self.h)
            self.h = C.c_void_p()

    def set_particles(self, species, ids, pos, vel, mass):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        vel = np.ascontiguousarray(vel, dtype=np.float64)
        _ck(lib().gevb_sim_set_particles(self.h, species, len(ids), _ptr(ids), _ptr(pos), _ptr(vel), mass), "gevb_sim_set_particles")

    def field(self, name):
        kind = REAL if FIELD_IDS[name] < 10 else CPLX
        return Field(self.ctx, kind, FIELD_COMPS[name], handle=lib().gevb_sim_field(self.h, FIELD_IDS[name]))

    def pcls(self, species=0):
        return Particles(self.ctx, 0.0, handle=lib().gevb_sim_pcls(self.h, species))

    def set_field(self, name, data):
        a = np.ascontiguousarray(data, dtype=np.float64)
        assert a.shape == self.field(name).shape, (a.shape, self.field(name).shape)
        _ck(lib().gevb_sim_set_field(self.h, FIELD_IDS[name], _ptr(a)), "gevb_sim_set_field")

    def get_field(self, name):
        return self.field(name).download()

    def state(self):
        s = np.zeros(9)
        _ck(lib().gevb_sim_get_state(self.h, s.ctypes.data_as(C.POINTER(C.c_double))), "gevb_sim_get_state")
        return dict(a=s[0], tau=s[1], dtau=s[2], dtau_old=s[3], cycle=int(s[4]), maxvel=(s[5], s[6]), T00hom=s[7], fourpiG=s[8])

    def set_state(self, a, tau, dtau, dtau_old, cycle, maxvel=(0.0, 0.0)):
        s = np.array([a, tau, dtau, dtau_old, cycle, maxvel[0], maxvel[1]], dtype=np.float64)
        _ck(lib().gevb_sim_set_state(self.h, s.ctypes.data_as(C.POINTER(C.c_double))), "gevb_sim_set_state")

    def set_ncdm(self, m_ncdm, T_ncdm, Omega_ncdm, z_switch_deltancdm, z_switch_Bncdm, z_switch_linearchi, movelimit):
        keep = [_darr(v) for v in (m_ncdm, T_ncdm, Omega_ncdm, z_switch_deltancdm, z_switch_Bncdm)]
        _ck(lib().gevb_sim_set_ncdm(self.h, len(m_ncdm), *[k[1] for k in keep], z_switch_linearchi, movelimit), "gevb_sim_set_ncdm")

    def set_ncdm_maxvel(self, maxvel):
        v = np.zeros(4); v[:len(maxvel)] = maxvel
        _ck(lib().gevb_sim_set_ncdm_maxvel(self.h, v.ctypes.data_as(C.POINTER(C.c_double))), "gevb_sim_set_ncdm_maxvel")

    def ncdm_state(self):
        v, n = np.zeros(4), np.zeros(4, dtype=np.int32)
        _ck(lib().gevb_sim_get_ncdm_state(self.h, v.ctypes.data_as(C.POINTER(C.c_double)), n.ctypes.data_as(C.POINTER(C.c_int))), "gevb_sim_get_ncdm_state")
        return v, n

    def write_spectra(self, prefix, pkcount, numbins, mask, z_target=-1.0):
        _ck(lib().gevb_sim_write_spectra(self.h, prefix.encode(), pkcount, numbins, mask, z_target), "gevb_sim_write_spectra")

    def save_gadget2(self, species, filename, tracer_factor=1, dtau_pos=0.0, dtau_vel=0.0):
        _ck(lib().gevb_sim_save_gadget2(self.h, species, filename.encode(), tracer_factor, dtau_pos, dtau_vel), "gevb_sim_save_gadget2")

    def run(self, z_pk, pk_mask, numbins, pk_prefix, z_snapshot, tracer_factor, snap_prefix, max_cycles=100000):
        """the main loop with its outputs (gevb_sim_run); returns (cycles, spectra sets, snapshots)"""
        zp, zpp = _darr(z_pk)
        zs, zsp = _darr(z_snapshot)
        out = (C.c_int * 3)()
        _ck(lib().gevb_sim_run(self.h, zpp, len(zp), pk_mask, numbins, pk_prefix.encode(), zsp, len(zs), tracer_factor, snap_prefix.encode(), max_cycles, out), "gevb_sim_run")
        return tuple(out)

    def hibernate(self, filebase):
        _ck(lib().gevb_sim_hibernate(self.h, filebase.encode()), "gevb_sim_hibernate")

    def restore(self, filebase):
        _ck(lib().gevb_sim_restore(self.h, filebase.encode()), "gevb_sim_restore")

    def run_settings(self, settings, max_cycles=100000):
        """the main loop with the outputs of a settings file (gevb_sim_run_settings); returns (cycles, spectra sets, snapshots)"""
        out = (C.c_int * 3)()
        _ck(lib().gevb_sim_run_settings(self.h, C.byref(settings), max_cycles, out), "gevb_sim_run_settings")
        return tuple(out)

    def write_field_snapshot(self, prefix, mask):
        """writeSnapshots' field dumps (output.hpp:98-300): <prefix>_<T00|B|phi|chi|hij>.bin"""
        _ck(lib().gevb_sim_write_field_snapshot(self.h, prefix.encode(), int(mask)), "gevb_sim_write_field_snapshot")

    def set_fused(self, fused):
        lib().gevb_sim_set_fused(self.h, int(bool(fused)))

    def step(self):
        _ck(lib().gevb_sim_step(self.h), "gevb_sim_step")


def read_raw_field(filename):
    """the flat binary field file of gevb_field_save_raw: float64 [comp][z][y][x] behind a 32-byte header"""
    with open(filename, "rb") as f:
        hdr = f.read(32)
        if hdr[:8] != b"GEVBFLD1":
            raise GevbError(f"{filename}: not a gevb field file")
        n, ncomp = np.frombuffer(hdr[8:16], dtype=np.int32)
        return np.fromfile(f, dtype=np.float64, count=int(ncomp) * int(n) ** 3).reshape(int(ncomp), int(n), int(n), int(n))


# ---- settings.ini + basic IC generator (host side of the product; no device needed for the pieces) -------------------
MAX_OUTPUTS, PATH_MAX = 32, 512


class Settings(C.Structure):
    """gevb_settings (include/gevb.h): the subset of the reference's settings.ini the hot path needs"""
    _fields_ = [("ngrid", C.c_int), ("gr_flag", C.c_int), ("vector_flag", C.c_int), ("baryon_flag", C.c_int), ("seed", C.c_int), ("ksphere", C.c_int),
                ("correct_displacement", C.c_int), ("tiling", C.c_int * 2), ("tracer_factor", C.c_int * 2), ("numbins", C.c_int), ("pk_mask", C.c_int),
                ("snapshot_mask", C.c_int), ("num_pk", C.c_int), ("num_snapshot", C.c_int),
                ("boxsize", C.c_double), ("Cf", C.c_double), ("steplimit", C.c_double), ("movelimit", C.c_double), ("z_in", C.c_double), ("z_relax", C.c_double),
                ("A_s", C.c_double), ("n_s", C.c_double), ("k_pivot", C.c_double), ("cosmo", C.c_double * 11),
                ("z_pk", C.c_double * MAX_OUTPUTS), ("z_snapshot", C.c_double * MAX_OUTPUTS),
                ("template_file", (C.c_char * PATH_MAX) * 2), ("tk_file", C.c_char * PATH_MAX), ("output_path", C.c_char * PATH_MAX),
                ("basename_generic", C.c_char * 128), ("basename_pk", C.c_char * 128), ("basename_snapshot", C.c_char * 128)]


def settings_read(filename, overrides=""):
    st = Settings()
    _ck(lib().gevb_settings_read(str(filename).encode(), overrides.encode(), C.byref(st)), "gevb_settings_read")
    return st


def ic_cic_kernel(N, pcldata=None, numtile=1):
    """generateCICKernel on its 27 sites: array [dz+1][dy+1][dx+1]"""
    out = np.zeros(27)
    if pcldata is None:
        _ck(lib().gevb_ic_cic_kernel(N, 0, None, 1, out.ctypes.data_as(C.POINTER(C.c_double))), "gevb_ic_cic_kernel")
    else:
        p = np.ascontiguousarray(pcldata, dtype=np.float32)
        _ck(lib().gevb_ic_cic_kernel(N, len(p), p.ctypes.data_as(C.c_void_p), numtile, out.ctypes.data_as(C.POINTER(C.c_double))), "gevb_ic_cic_kernel")
    return out.reshape(3, 3, 3)


def ic_displacement_field(potFT, coeff, spline_x, spline_y, seed, ksphere=0, deconvolve_f=1):
    out = np.ascontiguousarray(potFT, dtype=np.float64).copy()
    x, y = np.ascontiguousarray(spline_x, dtype=np.float64), np.ascontiguousarray(spline_y, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    _ck(lib().gevb_ic_displacement_field(out.shape[0], out.ctypes.data_as(dp), coeff, len(x), x.ctypes.data_as(dp), y.ctypes.data_as(dp), seed, ksphere, deconvolve_f),
        "gevb_ic_displacement_field")
    return out


def ic_load_template(filename):
    n, ptr = C.c_int64(0), C.POINTER(C.c_float)()
    _ck(lib().gevb_ic_load_template(str(filename).encode(), C.byref(n), C.byref(ptr)), "gevb_ic_load_template")
    out = np.ctypeslib.as_array(ptr, shape=(n.value, 3)).copy()
    C.CDLL(None).free(ptr)
    return out


def sim_from_settings(ctx, settings):
    """main.cpp:184-340 with IC generator = basic: a Sim with particles and metric fields from the settings' own seed"""
    h = C.c_void_p()
    _ck(lib().gevb_sim_create_from_settings(C.byref(h), ctx.h, C.byref(settings)), "gevb_sim_create_from_settings")
    s = Sim.__new__(Sim)
    s.ctx, s.h = ctx, h
    s.settings = settings
    return s
