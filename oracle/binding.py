"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY — imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by sparkl_b200/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from sparkl_b200 import abi
from sparkl_b200.engine import Engine, declare

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False, name: str = "liboracle.so") -> str:
    so = os.path.join(_HERE, name)
    src = os.path.join(_HERE, "mpm_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, name], stdout=subprocess.DEVNULL)
    return so


_LIB_FAST = None


def lib_fast():
    """liboracle_fast.so: the same source built for speed (ORC_FAST: f32 Jacobi SVD, f32 node sums, -O3) — the CPU ARM of bench.py
    only; every parity check uses lib()"""
    global _LIB_FAST
    if _LIB_FAST is None:
        _LIB_FAST = _declare_all(C.CDLL(build(name="liboracle_fast.so")))
    return _LIB_FAST


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _declare_all(C.CDLL(build()))
    return _LIB


def _declare_all(_LIB):
    if True:
        declare(_LIB, "orc_")
        _LIB.orc_create.argtypes = [C.c_float]
        _LIB.orc_create.restype = C.c_void_p
        _LIB.orc_destroy.argtypes = [C.c_void_p]
        _LIB.orc_step.argtypes = [C.c_void_p, C.POINTER(abi.SolverParameters), C.POINTER(C.c_float * 3),
                                  C.POINTER(abi.StepTimings), C.POINTER(C.c_float)]
        _LIB.orc_step.restype = C.c_int
        vp = C.c_void_p
        _LIB.orc_halo_begin.argtypes = [vp]
        _LIB.orc_halo_active_keys.argtypes = [vp, vp, C.c_uint32]
        _LIB.orc_halo_tag.argtypes = [vp, vp, C.c_uint32]
        _LIB.orc_halo_count.argtypes = [vp, C.POINTER(C.c_uint32)]
        _LIB.orc_halo_pack.argtypes = [vp, vp, vp, C.c_uint32]
        _LIB.orc_halo_merge.argtypes = [vp, vp, vp, C.c_uint32]
        for nme in ("orc_halo_begin", "orc_halo_active_keys", "orc_halo_tag", "orc_halo_count", "orc_halo_pack", "orc_halo_merge"):
            getattr(_LIB, nme).restype = C.c_int
        _LIB.orc_set_threads.argtypes = [C.c_int]
        _LIB.orc_set_threads.restype = C.c_int
        _LIB.orc_block_key_of_point.argtypes = [C.c_void_p, C.c_float]
        _LIB.orc_block_key_of_point.restype = C.c_uint64
        _LIB.orc_dt_from_sec.argtypes = [C.c_float]
        _LIB.orc_dt_from_sec.restype = C.c_uint64
        _LIB.orc_dt_into_sec.argtypes = [C.c_uint64]
        _LIB.orc_dt_into_sec.restype = C.c_float
    return _LIB


class OracleEngine(Engine):
    def __init__(self, cell_width: float, fast: bool = False):
        l = lib_fast() if fast else lib()
        super().__init__(l, "orc_", C.c_void_p(l.orc_create(cell_width)))
        self.cell_width = cell_width
        self.device = -1

    def __del__(self):
        try:
            self._lib.orc_destroy(self._h)
        except Exception:
            pass

    def step(self, params: abi.SolverParameters, gravity):
        g = (C.c_float * 3)(*gravity)
        t = abi.StepTimings()
        sim = C.c_float()
        rc = self._lib.orc_step(self._h, C.byref(params), C.byref(g), C.byref(t), C.byref(sim))
        if rc:
            raise abi.SparklError(rc, "orc_step")
        return sim.value, t


def set_threads(n: int, fast: bool = False) -> int:
    return (lib_fast() if fast else lib()).orc_set_threads(n)


def svd3(F: np.ndarray):
    """F: 3x3 row-major numpy -> (U, s, Vt) as computed by the oracle's nalgebra stand-in."""
    l = lib()
    f = np.ascontiguousarray(F.T, dtype=np.float32)  # column-major image
    u, s, vt = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(9, np.float32)
    l.orc_svd3(f.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), vt.ctypes.data_as(C.c_void_p))
    return u.reshape(3, 3).T, s, vt.reshape(3, 3).T


def block_key_of_point(p, cell_width: float) -> int:
    a = np.asarray(p, np.float32)
    return lib().orc_block_key_of_point(a.ctypes.data_as(C.c_void_p), cell_width)


def kirchhoff_stress(model: abi.ParticleModel, F: np.ndarray, phase=1.0, hardening=1.0, jp=1.0, mass=1.0, volume0=1.0, vel_grad=None):
    l = lib()
    m = np.frombuffer(model.to_gpu_model(), np.uint8).copy()
    d = np.zeros(1, abi.DATA); d["elastic_hardening"] = hardening; d["plastic_hardening"] = 1.0
    v = np.zeros(1, abi.VOLUME)
    v["mass"], v["volume0"], v["plastic_deformation_gradient_det"] = mass, volume0, jp
    v["deformation_gradient"] = np.asarray(F, np.float32).T.reshape(9)
    g = np.zeros(9, np.float32) if vel_grad is None else np.asarray(vel_grad, np.float32).T.reshape(9).copy()
    out = np.zeros(9, np.float32)
    l.orc_kirchhoff_stress(m.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), C.c_float(phase),
                           g.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out.reshape(3, 3).T


def plastic_update(model: abi.ParticleModel, F: np.ndarray, data: np.ndarray, jp=1.0, phase=1.0):
    l = lib()
    m = np.frombuffer(model.to_gpu_model(), np.uint8).copy()
    v = np.zeros(1, abi.VOLUME)
    v["mass"], v["volume0"], v["plastic_deformation_gradient_det"] = 1.0, 1.0, jp
    v["deformation_gradient"] = np.asarray(F, np.float32).T.reshape(9)
    d = data.copy()
    l.orc_plastic_update(m.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), C.c_float(phase))
    return v["deformation_gradient"][0].reshape(3, 3).T.copy(), float(v["plastic_deformation_gradient_det"][0]), d


def cuboid_project(collider: abi.CuboidCollider, pt):
    l = lib()
    c = np.frombuffer(collider.to_gpu_collider(), np.uint8).copy()
    p = np.asarray(pt, np.float32).copy()
    out = np.zeros(3, np.float32)
    inside = C.c_int()
    l.orc_cuboid_project(c.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.byref(inside))
    return out, bool(inside.value)


def collider_project(collider, pt, max_dist):
    """GpuColliderShape::project_point_with_max_dist as restated by the oracle: (point, is_inside) or None"""
    l = lib()
    l.orc_collider_project.restype = C.c_int
    c = np.frombuffer(collider.to_gpu_collider(), np.uint8).copy()
    p = np.asarray(pt, np.float32).copy()
    out = np.zeros(3, np.float32)
    inside = C.c_int()
    found = l.orc_collider_project(c.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), C.c_float(max_dist), out.ctypes.data_as(C.c_void_p), C.byref(inside))
    return (out, bool(inside.value)) if found else None


# ------------------------------------------------------------------------------------------------ dim2 oracle
_LIB2D = None


def build2d(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle2d.so")
    src = os.path.join(_HERE, "mpm_oracle2d.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle2d.so"], stdout=subprocess.DEVNULL)
    return so


def lib2d():
    """oracle/liboracle2d.so: the dim2 restatement (mpm_oracle2d.cpp), same `orc_*` entry points with the dim2 struct layouts"""
    global _LIB2D
    if _LIB2D is None:
        l = C.CDLL(build2d())
        declare(l, "orc_")
        l.orc_create.argtypes = [C.c_float]
        l.orc_create.restype = C.c_void_p
        l.orc_destroy.argtypes = [C.c_void_p]
        l.orc_step.argtypes = [C.c_void_p, C.POINTER(abi.SolverParameters), C.POINTER(C.c_float * 3), C.POINTER(abi.StepTimings), C.POINTER(C.c_float)]
        l.orc_step.restype = C.c_int
        l.orc_block_key_of_point.argtypes = [C.c_void_p, C.c_float]
        l.orc_block_key_of_point.restype = C.c_uint64
        l.orc_collider_project.restype = C.c_int
        _LIB2D = l
    return _LIB2D


class OracleEngine2D(Engine):
    def __init__(self, cell_width: float):
        from sparkl_b200 import abi2d
        l = lib2d()
        super().__init__(l, "orc_", C.c_void_p(l.orc_create(cell_width)), types=abi2d)
        self.cell_width = cell_width
        self.device = -1

    def __del__(self):
        try:
            self._lib.orc_destroy(self._h)
        except Exception:
            pass

    def step(self, params: abi.SolverParameters, gravity):
        g = (C.c_float * 3)(*gravity)
        t = abi.StepTimings()
        sim = C.c_float()
        rc = self._lib.orc_step(self._h, C.byref(params), C.byref(g), C.byref(t), C.byref(sim))
        if rc:
            raise abi.SparklError(rc, "orc_step")
        return sim.value, t


def svd2(F: np.ndarray):
    """F: 2x2 row-major numpy -> (U, s, Vt) of the dim2 oracle's closed-form nalgebra stand-in"""
    l = lib2d()
    f = np.ascontiguousarray(F.T, dtype=np.float32)
    u, s, vt = np.zeros(4, np.float32), np.zeros(2, np.float32), np.zeros(4, np.float32)
    l.orc_svd2(f.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), vt.ctypes.data_as(C.c_void_p))
    return u.reshape(2, 2).T, s, vt.reshape(2, 2).T


def block_key_of_point2d(p, cell_width: float) -> int:
    a = np.asarray(p, np.float32)
    return lib2d().orc_block_key_of_point(a.ctypes.data_as(C.c_void_p), cell_width)


def kirchhoff_stress2d(model, F: np.ndarray, phase=1.0, hardening=1.0, jp=1.0, mass=1.0, volume0=1.0, vel_grad=None):
    from sparkl_b200 import abi2d
    l = lib2d()
    m = np.frombuffer(model.to_gpu_model(), np.uint8).copy()
    d = np.zeros(1, abi2d.DATA); d["elastic_hardening"] = hardening; d["plastic_hardening"] = 1.0
    v = np.zeros(1, abi2d.VOLUME)
    v["mass"], v["volume0"], v["plastic_deformation_gradient_det"] = mass, volume0, jp
    v["deformation_gradient"] = np.asarray(F, np.float32).T.reshape(4)
    g = np.zeros(4, np.float32) if vel_grad is None else np.asarray(vel_grad, np.float32).T.reshape(4).copy()
    out = np.zeros(4, np.float32)
    l.orc_kirchhoff_stress(m.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), C.c_float(phase),
                           g.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out.reshape(2, 2).T


def plastic_update2d(model, F: np.ndarray, data: np.ndarray, jp=1.0, phase=1.0):
    from sparkl_b200 import abi2d
    l = lib2d()
    m = np.frombuffer(model.to_gpu_model(), np.uint8).copy()
    v = np.zeros(1, abi2d.VOLUME)
    v["mass"], v["volume0"], v["plastic_deformation_gradient_det"] = 1.0, 1.0, jp
    v["deformation_gradient"] = np.asarray(F, np.float32).T.reshape(4)
    d = data.copy()
    l.orc_plastic_update(m.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), C.c_float(phase))
    return v["deformation_gradient"][0].reshape(2, 2).T.copy(), float(v["plastic_deformation_gradient_det"][0]), d


def collider_project2d(collider, pt, max_dist):
    l = lib2d()
    c = np.frombuffer(collider.to_gpu_collider(), np.uint8).copy()
    p = np.asarray(pt, np.float32).copy()
    out = np.zeros(2, np.float32)
    inside = C.c_int()
    found = l.orc_collider_project(c.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), C.c_float(max_dist), out.ctypes.data_as(C.c_void_p), C.byref(inside))
    return (out, bool(inside.value)) if found else None


# ------------------------------------------------------------------------------------------------ oracle-A
_REF = None
REF_CUBIN = os.path.join(_HERE, "_ref", "sparkl3d-kernels.cubin")
REF_LIB = os.path.join(_HERE, "_ref", "libref_harness.so")


def ref_available() -> bool:
    """the reference's own GPU kernels (assembled from its shipped PTX by `make -C oracle ref`) + a CUDA driver"""
    if not (os.path.exists(REF_CUBIN) and os.path.exists(REF_LIB)):
        return False
    try:
        return bool(ref_lib().ref_available())
    except OSError:
        return False


def ref_lib():
    global _REF
    if _REF is None:
        _REF = C.CDLL(REF_LIB)
        declare(_REF, "ref_")
        vp = C.c_void_p
        _REF.ref_available.restype = C.c_int
        _REF.ref_create.argtypes = [C.c_int, C.c_float, C.c_char_p, C.POINTER(vp)]
        _REF.ref_create.restype = C.c_int
        _REF.ref_destroy.argtypes = [vp]
        _REF.ref_step.argtypes = [vp, C.POINTER(abi.SolverParameters), C.POINTER(C.c_float * 3), C.POINTER(abi.StepTimings), C.POINTER(C.c_float)]
        _REF.ref_step.restype = C.c_int
    return _REF


class RefEngine(Engine):
    """The reference's unmodified GPU kernels driven by oracle/ref_harness.cpp (oracle-A)."""

    def __init__(self, cell_width: float, device: int = 0):
        l = ref_lib()
        h = C.c_void_p()
        rc = l.ref_create(device, cell_width, REF_CUBIN.encode(), C.byref(h))
        if rc:
            raise abi.SparklError(rc, "ref_create")
        super().__init__(l, "ref_", h)
        self.cell_width, self.device = cell_width, device

    def __del__(self):
        try:
            self._lib.ref_destroy(self._h)
        except Exception:
            pass

    def step(self, params: abi.SolverParameters, gravity):
        g = (C.c_float * 3)(*gravity)
        t = abi.StepTimings()
        sim = C.c_float()
        rc = self._lib.ref_step(self._h, C.byref(params), C.byref(g), C.byref(t), C.byref(sim))
        if rc:
            raise abi.SparklError(rc, "ref_step")
        return sim.value, t


# ------------------------------------------------------------------------------------------------ oracle-A, dim2
_REF2D = None
REF2D_CUBIN = os.path.join(_HERE, "_ref", "sparkl2d-kernels.cubin")
REF2D_LIB = os.path.join(_HERE, "_ref", "libref_harness2d.so")


def ref2d_available() -> bool:
    """the reference's own dim2 GPU kernels (resources/sparkl2d-kernels.ptx assembled by `make -C oracle ref`) + a CUDA driver"""
    if not (os.path.exists(REF2D_CUBIN) and os.path.exists(REF2D_LIB)):
        return False
    try:
        return bool(ref2d_lib().ref_available())
    except OSError:
        return False


def ref2d_lib():
    global _REF2D
    if _REF2D is None:
        _REF2D = C.CDLL(REF2D_LIB)
        declare(_REF2D, "ref_")
        vp = C.c_void_p
        _REF2D.ref_available.restype = C.c_int
        _REF2D.ref_create.argtypes = [C.c_int, C.c_float, C.c_char_p, C.POINTER(vp)]
        _REF2D.ref_create.restype = C.c_int
        _REF2D.ref_destroy.argtypes = [vp]
        _REF2D.ref_step.argtypes = [vp, C.POINTER(abi.SolverParameters), C.POINTER(C.c_float * 3), C.POINTER(abi.StepTimings), C.POINTER(C.c_float)]
        _REF2D.ref_step.restype = C.c_int
    return _REF2D


class RefEngine2D(Engine):
    """The reference's unmodified dim2 GPU kernels driven by the -DSPK_DIM=2 build of oracle/ref_harness.cpp.  Collider-free
    scenes only (colliders_upload with a non-empty list raises EUNSUPPORTED)."""

    def __init__(self, cell_width: float, device: int = 0):
        from sparkl_b200 import abi2d
        l = ref2d_lib()
        h = C.c_void_p()
        rc = l.ref_create(device, cell_width, REF2D_CUBIN.encode(), C.byref(h))
        if rc:
            raise abi.SparklError(rc, "ref_create (dim2)")
        super().__init__(l, "ref_", h, types=abi2d)
        self.cell_width, self.device = cell_width, device

    def __del__(self):
        try:
            self._lib.ref_destroy(self._h)
        except Exception:
            pass

    def step(self, params: abi.SolverParameters, gravity):
        g = (C.c_float * 3)(*gravity)
        t = abi.StepTimings()
        sim = C.c_float()
        rc = self._lib.ref_step(self._h, C.byref(params), C.byref(g), C.byref(t), C.byref(sim))
        if rc:
            raise abi.SparklError(rc, "ref_step (dim2)")
        return sim.value, t
