"""TEST INFRASTRUCTURE: a host emulation of the dim2 CUDA pipeline.

The arithmetic is NOT re-written here: tests/host/kernels2d_host.cpp compiles the product's own device sources
(sparkl_b200/csrc/spk_g2p2g2d_body.cuh, spk_grid2d.cu, spk_math2d.cuh) for the host.  This module only restates, in numpy, what
cannot be compiled without nvcc: the particle repack of spk_api.cu (pack_kernel / unpack_kernel), the binning / block ids /
neighbour tables / dispatch of spk_sort.cu, and the dt hand-over of spk_stage_estimate_dt.  `Emu2D` exposes the Engine stage API, so
the GPU parity test body (tests/test_dim2_gpu.py) runs against it on GPU-less CI.  Never imported by sparkl_b200/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from sparkl_b200 import abi, abi2d

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
vp = C.c_void_p

DEVMODEL = np.dtype([("ctag", "<u4"), ("ptag", "<u4"), ("ftag", "<u4"), ("flags", "<u4"), ("c", "<f4", (4,)), ("p", "<f4", (8,)), ("f", "<f4", (2,))])
DISPATCH_ENTRY = np.dtype([("block", "<u4"), ("first", "<u4"), ("count", "<u4"), ("pad", "<u4")])
DT_RAW_NONE = 0x7F7F7F7F
F_FAILED, F_STATIC, F_KINEMATIC, F_MODEL_SHIFT = 1, 2, 4, 8


class GridBuffers(C.Structure):  # sparkl_b200/csrc/spk_internal.cuh
    _fields_ = [("keys", vp), ("vals", vp), ("cap_hash", C.c_uint32), ("cap_blocks", C.c_uint32), ("blk_key", vp), ("blk_first", vp), ("blk_count", vp),
                ("blk_halo", vp), ("blk_dfirst", vp), ("nbr_self", vp), ("nbr_prev", vp), ("vel", vp), ("acc", vp), ("proj", vp), ("cacheN", vp), ("cacheD", vp),
                ("external_hash", C.c_uint32), ("nab_host", C.c_uint32), ("has_velocity", C.c_uint32)]


class ParticleBuffers(C.Structure):
    _fields_ = [(n, vp) for n in ("a0", "a1", "a2", "a3", "a4", "a5")]


class DevCollider(C.Structure):
    _fields_ = [("half_extents", C.c_float * 3), ("rot", C.c_float * 4), ("tr", C.c_float * 3), ("friction", C.c_float), ("handling", C.c_uint32),
                ("shape", C.c_uint32), ("nrows", C.c_uint32), ("ncols", C.c_uint32), ("heights", vp), ("scale", C.c_float * 3), ("hf_flags", C.c_uint32)]


def devmodel(model: abi.ParticleModel) -> np.ndarray:
    """mirror of spk_models_upload (sparkl_b200/csrc/spk_api.cu): SpkGpuParticleModel -> DevModel"""
    d = np.zeros(1, DEVMODEL)
    c, p, f = model.constitutive_model, model.plastic_model, model.failure_model
    d["ctag"], d["ptag"], d["ftag"] = c.tag, 5 if p is None else p.tag, 2 if f is None else f.tag
    if c.tag == 0:
        d["c"][0, :3] = [c.cfl_coeff, c.lam, c.mu]
        d["flags"] |= 1 if c.split_stress_on_failure else 0
    elif c.tag == 1:
        d["c"][0, :3] = [c.cfl_coeff, c.lam, c.mu]
    elif c.tag == 2:
        d["c"][0] = [c.pressure0, np.array([c.gamma], "<u4").view("<f4")[0], c.viscosity, c.max_neg_pressure]
    if p is not None and p.tag == 0:
        d["p"][0, :7] = [p.h0, p.h1, p.h2, p.h3, p.lam, p.mu, p.volume_correction]
        d["flags"] |= 2 if p.only_active_when_failed else 0
    elif p is not None and p.tag == 1:
        d["p"][0, :5] = [p.mu, p.kappa, p.hardening_factor, p.cohesion, p.friction]
        d["flags"] |= 4 if p.hardening_enabled else 0
    elif p is not None and p.tag == 2:
        d["p"][0, :4] = [p.mu, p.lam, p.tensile_strength, p.softening_rate]
    elif p is not None and p.tag == 3:
        d["p"][0, :3] = [p.min_epsilon, p.max_epsilon, p.hardening_coeff]
    if f is not None:
        d["f"][0] = [f.max_principal_stress, f.max_shear_stress]
    return d


def _build(name):
    src = os.path.join(HERE, name + ".cpp")
    out = os.path.join(HERE, "lib" + name + ".so")
    deps = [src] + [os.path.join(ROOT, "sparkl_b200", "csrc", f) for f in ("spk_math2d.cuh", "spk_g2p2g2d_body.cuh", "spk_grid2d.cu", "spk_device.cuh",
                                                                            "spk_internal.cuh")] + [os.path.join(ROOT, "include", "sparkl_abi.h")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-I" + cuda_inc, "-o", out, src])
    return C.CDLL(out)


_LIBS = {}


def math_lib():
    if "m" not in _LIBS:
        l = _build("math2d_host")
        assert l.hm_sizeof_devmodel() == DEVMODEL.itemsize
        l.hm_dt_candidate.restype = C.c_float
        l.hm_dt_candidate.argtypes = [vp] + [C.c_float] * 7
        l.hm_stress.argtypes = [vp, vp, C.c_float, C.c_float, vp]
        _LIBS["m"] = l
    return _LIBS["m"]


def kernel_lib():
    if "k" not in _LIBS:
        l = _build("kernels2d_host")
        assert l.hk_sizeof_gridbuffers() == C.sizeof(GridBuffers) and l.hk_sizeof_devcollider() == C.sizeof(DevCollider)
        assert l.hk_sizeof_particlebuffers() == C.sizeof(ParticleBuffers)
        l.hk_g2p2g.argtypes = [ParticleBuffers, ParticleBuffers, vp, vp, vp, vp, C.c_uint32, vp, vp, vp, vp, vp, C.c_float, C.c_float, vp]
        l.hk_reset_grid.argtypes = [GridBuffers, GridBuffers, C.c_uint32]
        l.hk_grid_update.argtypes = [GridBuffers, C.c_uint32, vp, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_float]
        l.hk_grid_export.argtypes = [GridBuffers, C.c_uint32, vp]
        _LIBS["k"] = l
    return _LIBS["k"]


def _p(a):
    return a.ctypes.data_as(vp)


def _round_half_away(x):
    return np.trunc(x + np.copysign(np.float32(0.5), x)).astype(np.float32)


class _Grid:
    def __init__(self, nb=0):
        self.nb = nb
        self.key = np.zeros(nb, np.uint64)
        self.first = np.zeros(nb, np.uint32)
        self.count = np.zeros(nb, np.uint32)
        self.halo = np.zeros(nb, np.uint32)
        self.nbr_self = np.full((max(nb, 1), 4), -1, np.int32)
        self.nbr_prev = np.full((max(nb, 1), 4), -1, np.int32)
        self.vel = np.zeros((max(nb, 1) * 16, 4), np.float32)
        self.acc = np.zeros(max(nb, 1) * 96, np.float32)
        self.proj = np.zeros(max(nb, 1) * 16, np.uint32)
        self.cacheN = np.zeros((max(nb, 1) * 16, 4), np.float32)
        self.cacheD = np.zeros((max(nb, 1) * 16, 4), np.float32)
        self.has_velocity = 0
        self.index = {}

    def blob(self) -> GridBuffers:
        g = GridBuffers()
        g.cap_blocks, g.nab_host, g.has_velocity = self.nb, self.nb, self.has_velocity
        g.blk_key, g.blk_first, g.blk_count, g.blk_halo = _p(self.key), _p(self.first), _p(self.count), _p(self.halo)
        g.nbr_self, g.nbr_prev = _p(self.nbr_self), _p(self.nbr_prev)
        g.vel, g.acc, g.proj, g.cacheN, g.cacheD = _p(self.vel), _p(self.acc), _p(self.proj), _p(self.cacheN), _p(self.cacheD)
        return g


class Emu2D:
    """Engine-compatible driver of the host-compiled dim2 device sources"""
    T = abi2d
    dim = 2

    def __init__(self, cell_width: float):
        self.h = np.float32(cell_width)
        self.cell_width = cell_width
        self.k, self.m = kernel_lib(), math_lib()
        self.chunk = self.k.hk_chunk()
        self.grids = [_Grid(0), _Grid(0)]
        self.gcur = 0
        self.models = np.zeros(0, DEVMODEL)
        self.cols = (DevCollider * 1)()
        self.ncol = 0
        self._keep = []
        self.n = 0
        self.dt_acc_host = 0
        self.dt_cache_valid = False
        self.dt_raw_bits = np.array([DT_RAW_NONE], np.uint32)
        self.sorted_valid = False

    # ---- uploads (pack_kernel, spk_models_upload, spk_colliders_upload of spk_api.cu) -----------------------------------------
    def upload(self, pc):
        n = len(pc)
        self.n = n
        z = lambda: np.zeros((max(n, 1), 4), np.float32)  # noqa: E731
        self.pb = [[z() for _ in range(6)], [z() for _ in range(6)]]
        self.cur = 0
        a = self.pb[0]
        st = pc.status
        flags = (st["failed"] != 0) * F_FAILED | (st["is_static"] != 0) * F_STATIC | (st["kinematic_vel_enabled"] != 0) * F_KINEMATIC | \
            (st["model_index"].astype(np.uint32) << F_MODEL_SHIFT)
        a[0][:n, :2] = pc.position["point"]
        a[0][:n, 3] = flags.astype(np.uint32).view(np.float32)
        a[1][:n] = pc.volume["deformation_gradient"]
        a[3][:n, 1], a[3][:n, 2], a[3][:n, 3] = pc.volume["plastic_deformation_gradient_det"], pc.volume["mass"], pc.volume["volume0"]
        a[4][:n, 0], a[4][:n, 1] = pc.phase["phase"], pc.phase["psi_pos"]
        a[4][:n, 2], a[4][:n, 3] = pc.data["elastic_hardening"], pc.data["plastic_hardening"]
        a[5][:n, 0], a[5][:n, 1] = pc.data["log_vol_gain"], pc.data["nacc_alpha"]
        a[5][:n, 2] = np.arange(n, dtype=np.uint32).view(np.float32)
        self.vel = z()
        self.vel[:n, :2] = pc.velocity["vector"]
        self.cold = z()
        self.cold[:n, :2] = st["kinematic_vel"]
        self.cold[:n, 3] = pc.volume["radius0"]
        self.perm = np.arange(max(n, 1), dtype=np.uint32)
        self.sorted_valid = False
        self.dt_cache_valid = False

    def models_upload(self, models):
        self.models = np.concatenate([devmodel(m) for m in models]) if models else np.zeros(0, DEVMODEL)
        self.dt_cache_valid = False

    def colliders_upload(self, colliders):
        n = len(colliders)
        self.cols = (DevCollider * max(n, 1))()
        self._keep = []
        for i, c in enumerate(colliders):
            d = self.cols[i]
            d.shape = 3 if isinstance(c, abi2d.PolylineCollider) else (1 if isinstance(c, abi2d.HeightFieldCollider) else 0)
            if d.shape == 0:
                d.half_extents[0], d.half_extents[1] = np.float32(c.half_extents[0]), np.float32(c.half_extents[1])
            elif d.shape == 1:
                v = np.ascontiguousarray(c.heights, np.float32)
                self._keep.append(v)
                d.heights, d.nrows, d.hf_flags = v.ctypes.data, len(v), 1 | (2 if c.flip_interior else 0)
                d.scale[0], d.scale[1] = np.float32(c.scale[0]), np.float32(c.scale[1])
            else:
                v = np.ascontiguousarray(c.vertices, np.float32)
                self._keep.append(v)
                d.heights, d.nrows, d.hf_flags = v.ctypes.data, len(v), 2 if c.flip_interior else 0
            d.rot[0], d.rot[1] = abi2d._rotation(c.rotation)
            d.tr[0], d.tr[1] = np.float32(c.translation[0]), np.float32(c.translation[1])
            d.friction, d.handling = np.float32(c.friction), c.grid_boundary_handling
        self.ncol = n
        for g in self.grids:
            g.proj[:] = 0

    # ---- stages -----------------------------------------------------------------------------------------------------------------
    def frame_begin(self, params):
        self.dt_acc_host = _dt_from_sec(np.float32(params.dt))

    def stage_sort(self):
        """bin_particles / activate_neighbours / assign_ids / build_tables / scatter_perm of spk_sort.cu (ids = rank of the key)"""
        n, h = self.n, self.h
        a0 = self.pb[self.cur][0]
        q = _round_half_away((a0[:n, :2] / h).astype(np.float32)).astype(np.int64)
        blk = ((q - 2 + (abi2d.PACK_ORIGIN * 4)) // 4).astype(np.uint64)
        pkey = (blk[:, 0] & np.uint64(0xFFFFFFFF)) | ((blk[:, 1] & np.uint64(0xFFFFFFFF)) << np.uint64(32))
        mains = np.unique(pkey)
        xy = abi2d.unpack_key(mains).astype(np.uint64)
        act = np.unique(np.concatenate([(xy[:, 0] + np.uint64(o & 1)) | ((xy[:, 1] + np.uint64((o >> 1) & 1)) << np.uint64(32)) for o in range(4)])) \
            if len(mains) else np.zeros(0, np.uint64)
        g = _Grid(len(act))
        prev = self.grids[self.gcur]
        g.key[:] = act
        g.index = {int(k): i for i, k in enumerate(act)}
        pblock = np.searchsorted(act, pkey)
        g.count[:] = np.bincount(pblock, minlength=len(act)).astype(np.uint32) if n else 0
        g.first[:] = np.concatenate([[0], np.cumsum(g.count)[:-1]]).astype(np.uint32) if len(act) else 0
        order = np.argsort(pblock, kind="stable")  # sorted slot -> index in the current physical order
        self.perm = order.astype(np.uint32) if n else np.zeros(1, np.uint32)
        axy = abi2d.unpack_key(act).astype(np.uint64)
        for o in range(4):
            keys = (axy[:, 0] + np.uint64(o & 1)) | ((axy[:, 1] + np.uint64((o >> 1) & 1)) << np.uint64(32))
            g.nbr_self[:len(act), o] = [g.index.get(int(k), -1) for k in keys]
            g.nbr_prev[:len(act), o] = [prev.index.get(int(k), -1) for k in keys]
        disp = []
        for b in range(len(act)):
            c, f = int(g.count[b]), int(g.first[b])
            for m in range((c + self.chunk - 1) // self.chunk):
                disp.append((b, f + m * self.chunk, min(self.chunk, c - m * self.chunk), 0))
        self._dispatch = np.array(disp, DISPATCH_ENTRY) if disp else np.zeros(0, DISPATCH_ENTRY)
        self.grids[1 - self.gcur] = g
        self.sorted_valid = True
        info = abi.SortInfo()
        info.num_active_blocks = len(act)
        info.num_dispatch_blocks = int(sum((int(c) + 63) // 64 for c in g.count))  # the reference-format count (64 / entry in dim2)
        return info

    def stage_reset_grid(self):
        nxt, prev = self.grids[1 - self.gcur], self.grids[self.gcur]
        nxt.has_velocity = 0
        self.k.hk_reset_grid(nxt.blob(), prev.blob(), nxt.nb)

    def stage_estimate_dt(self, min_dt, max_dt):
        """spk_stage_estimate_dt (spk_api.cu): the candidate fused into the previous g2p2g when valid, else estimate_dt2d_kernel"""
        min_dt, max_dt = np.float32(min_dt), np.float32(max_dt)
        if self.dt_cache_valid:
            bits = int(self.dt_raw_bits[0])
            if bits < DT_RAW_NONE:
                raw = np.array([bits], np.uint32).view(np.float32)[0]
                d = np.float32(min(max_dt, raw))
                if d < min_dt and max_dt > min_dt:
                    d = min_dt
                self.dt_acc_host = min(self.dt_acc_host, _dt_from_sec(d))
        else:
            a = self.pb[self.cur]
            best = self.dt_acc_host
            for i in range(self.n):
                flags = int(a[0][i, 3:4].view(np.uint32)[0])
                if flags & F_FAILED:
                    continue
                m = self.models[flags >> F_MODEL_SHIFT: (flags >> F_MODEL_SHIFT) + 1]
                cand = np.float32(self.m.hm_dt_candidate(_p(m), self.h, a[3][i, 2], a[3][i, 3], a[1][i, 0], a[4][i, 2], self.vel[i, 0], self.vel[i, 1]))
                d = max_dt if np.isnan(cand) else np.float32(min(max_dt, cand))
                if d < min_dt and max_dt > min_dt:
                    d = min_dt
                best = min(best, _dt_from_sec(d))
            self.dt_acc_host = best
        return float(np.float32(self.dt_acc_host) / np.float32(1e12))

    def dt_acc(self):
        return self.dt_acc_host

    def set_dt_acc(self, v):
        self.dt_acc_host = int(v)

    def stage_g2p2g(self, dt, damage_model=0):
        assert self.sorted_valid
        nxt, prev = self.grids[1 - self.gcur], self.grids[self.gcur]
        src, dst = ParticleBuffers(*[_p(x) for x in self.pb[self.cur]]), ParticleBuffers(*[_p(x) for x in self.pb[1 - self.cur]])
        self.dt_raw_bits[0] = DT_RAW_NONE
        vel_out = np.zeros_like(self.vel)
        if len(self._dispatch):
            self.k.hk_g2p2g(src, dst, _p(vel_out), _p(self.cold), _p(self.perm), _p(self._dispatch), len(self._dispatch), _p(self.models), _p(nxt.nbr_self),
                            _p(nxt.nbr_prev), _p(nxt.acc), _p(prev.vel), self.h, np.float32(dt), _p(self.dt_raw_bits))
            self.vel = vel_out
            self.cur = 1 - self.cur
            self.sorted_valid = False
            self.dt_cache_valid = True

    def stage_grid_update(self, dt, gravity):
        nxt = self.grids[1 - self.gcur]
        self.k.hk_grid_update(nxt.blob(), nxt.nb, C.cast(self.cols, vp), self.ncol, self.h, np.float32(dt), np.float32(gravity[0]), np.float32(gravity[1]))
        nxt.has_velocity = 1
        self.gcur = 1 - self.gcur

    def substep(self, params, gravity, remaining):
        min_dt = np.float32(params.dt) / np.float32(params.max_num_substeps)
        self.stage_sort()
        self.stage_reset_grid()
        dt = self.stage_estimate_dt(float(min_dt), float(min(np.float32(remaining), np.float32(params.max_substep_dt))))
        self.stage_g2p2g(dt)
        self.stage_grid_update(dt, gravity)
        return dt

    # ---- read-back (unpack_kernel of spk_api.cu) ---------------------------------------------------------------------------------
    def __len__(self):
        return self.n

    def read(self, which, n=None, out=None):
        n = self.n
        a = self.pb[self.cur]
        ids = a[5][:n, 2].copy().view(np.uint32)
        if which == abi.ATTR_SORTED_IDS:
            src = self.perm[:n] if self.sorted_valid else np.arange(n)
            return ids[src].copy()
        out = np.zeros(n, abi2d.ATTR_DTYPES[which])
        if which == abi.ATTR_POSITION:
            out["point"][ids] = a[0][:n, :2]
        elif which == abi.ATTR_VELOCITY:
            out["vector"][ids] = self.vel[:n, :2]
        elif which == abi.ATTR_STATUS:
            f = a[0][:n, 3].copy().view(np.uint32)
            out["failed"][ids], out["is_static"][ids], out["kinematic_vel_enabled"][ids] = (f & 1) != 0, (f & 2) != 0, (f & 4) != 0
            out["kinematic_vel"] = self.cold[:n, :2]
            out["model_index"][ids] = f >> F_MODEL_SHIFT
        elif which == abi.ATTR_VOLUME:
            out["mass"][ids], out["volume0"][ids], out["radius0"] = a[3][:n, 2], a[3][:n, 3], self.cold[:n, 3]
            out["deformation_gradient"][ids] = a[1][:n]
            out["plastic_deformation_gradient_det"][ids] = a[3][:n, 1]
        elif which == abi.ATTR_PHASE:
            out["phase"][ids], out["psi_pos"][ids] = a[4][:n, 0], a[4][:n, 1]
        elif which == abi.ATTR_DATA:
            out["nacc_alpha"][ids], out["plastic_hardening"][ids] = a[5][:n, 1], a[4][:n, 3]
            out["elastic_hardening"][ids], out["log_vol_gain"][ids] = a[4][:n, 2], a[5][:n, 0]
        return out

    def read_all(self):
        pc = abi2d.ParticleComponents(0)
        pc.position, pc.velocity = self.read(abi.ATTR_POSITION), self.read(abi.ATTR_VELOCITY)
        pc.status, pc.volume = self.read(abi.ATTR_STATUS), self.read(abi.ATTR_VOLUME)
        pc.phase, pc.data = self.read(abi.ATTR_PHASE), self.read(abi.ATTR_DATA)
        return pc

    def _grid(self, which):
        return self.grids[1 - self.gcur] if which else self.grids[self.gcur]

    def num_active_blocks(self, which):
        return self._grid(which).nb

    def grid_headers(self, which):
        g = self._grid(which)
        out = np.zeros(g.nb, abi.ACTIVE_BLOCK_HEADER)
        out["virtual_id"], out["first_particle"], out["num_particles"] = g.key, g.first, g.count
        return out

    def grid_nodes(self, which):
        g = self._grid(which)
        out = np.zeros((g.nb, 16), abi2d.GRID_NODE)
        self.k.hk_grid_export(g.blob(), g.nb, _p(out))
        return out

    def dispatch(self, n):
        """the dispatch table in the reference's format (64 particles per entry in dim2), as spk_grid_read_dispatch derives it"""
        g = self._grid(1 if self.sorted_valid else 0)
        out = []
        for b in range(g.nb):
            c = int(g.count[b])
            for m in range((c + 63) // 64):
                out.append((b, int(g.first[b]) + m * 64))
        assert len(out) == n
        return np.array(out, abi.DISPATCH) if out else np.zeros(0, abi.DISPATCH)


def _dt_from_sec(dt) -> int:
    v = np.float32(dt) * np.float32(1e12)
    if np.isnan(v) or v <= 0:
        return 0
    if v >= np.float32(1.8446744073709552e19):
        return 0xFFFFFFFFFFFFFFFF
    return int(v)
