"""Host-side mirror of the reference's `#[repr(C)]` types for the `dim2` build (crate sparkl2d) of the GPU MPM path.

Same role as abi.py for the 3D build: numpy structured dtypes are byte images of include/sparkl_abi.h compiled with
-DSPK_DIM=2 (which cites the reference file:line of every struct).  Everything that does not depend on the dimension (solver
parameters, material models, error codes, enums) is re-exported from abi.py unchanged: the reference compiles both crates
from one source tree and the model structs hold no vectors.
"""
from __future__ import annotations

import math

import numpy as np

from .abi import *  # noqa: F401,F403  (dimension-independent types)
from .abi import BoundaryHandling, f32, shear_modulus, bulk_modulus
from . import abi as _abi3

DIM = 2
CELLS_PER_BLOCK = 16  # NUM_CELL_PER_BLOCK, src_kernels/gpu_grid.rs:9-10
REF_G2P2G_THREADS = 64  # src_kernels/cuda/mod.rs:12-13

POSITION = np.dtype([("point", "<f4", (2,))])                                   # particle.rs:17-20
VELOCITY = np.dtype([("vector", "<f4", (2,))])                                  # particle.rs:54-57
VOLUME = np.dtype([("mass", "<f4"), ("volume0", "<f4"), ("radius0", "<f4"),     # particle.rs:65-72
                   ("deformation_gradient", "<f4", (4,)),                       # column-major Matrix2
                   ("plastic_deformation_gradient_det", "<f4")])
STATUS = np.dtype([("failed", "u1"), ("is_static", "u1"), ("kinematic_vel_enabled", "u1"), ("_pad0", "u1"),
                   ("kinematic_vel", "<f4", (2,)), ("_pad1", "<u4"), ("model_index", "<u8")])  # particle.rs:119-127
PHASE, DATA, ACTIVE_BLOCK_HEADER, DISPATCH = _abi3.PHASE, _abi3.DATA, _abi3.ACTIVE_BLOCK_HEADER, _abi3.DISPATCH
GRID_NODE = np.dtype([("mass", "<f4"), ("momentum_velocity", "<f4", (2,)), ("psi_momentum_velocity", "<f4"),
                      ("psi_mass", "<f4"), ("prev_mass", "<f4"),
                      ("projection_status_tag", "<u8"), ("projection_status_payload", "<u8"),
                      ("collision_normal", "<f4", (2,)), ("projection_scaled_dir", "<f4", (2,))])  # gpu_grid.rs:443-459
assert POSITION.itemsize == 8 and VELOCITY.itemsize == 8 and VOLUME.itemsize == 32
assert STATUS.itemsize == 24 and GRID_NODE.itemsize == 56

ATTR_DTYPES = {_abi3.ATTR_POSITION: POSITION, _abi3.ATTR_VELOCITY: VELOCITY, _abi3.ATTR_STATUS: STATUS, _abi3.ATTR_VOLUME: VOLUME,
               _abi3.ATTR_PHASE: PHASE, _abi3.ATTR_DATA: DATA, _abi3.ATTR_SORTED_IDS: np.dtype("<u4")}

PACK_ORIGIN = 0x8000_0000  # BlockVirtualId::PACK_ORIGIN (dim2), gpu_grid.rs:76-77


def unpack_key(keys):
    """BlockVirtualId::unpack (dim2), gpu_grid.rs:106-109: x | y << 32"""
    k = np.asarray(keys, np.uint64)
    return np.stack([(k & np.uint64(0xFFFFFFFF)).astype(np.int64), (k >> np.uint64(32)).astype(np.int64)], axis=-1)


class NaccPlasticity(_abi3.NaccPlasticity):  # plasticity_nacc.rs:20-40: `friction` depends on DIM
    def __init__(self, young_modulus, poisson_ratio, cohesion, hardening_enabled, hardening_factor, friction_angle):
        sin_f = f32(math.sin(f32(friction_angle)))
        d = f32(2.0)
        self.mu, self.kappa = shear_modulus(young_modulus, poisson_ratio), bulk_modulus(young_modulus, poisson_ratio)
        self.hardening_enabled, self.hardening_factor, self.cohesion = bool(hardening_enabled), f32(hardening_factor), f32(cohesion)
        self.friction = f32(np.sqrt(f32(2.0) / f32(3.0)) * f32(2.0) * sin_f / (f32(3.0) - sin_f) * d / np.sqrt(f32(2.0) / (f32(6.0) - d)))


# ---------------------------------------------------------------------------------------------- colliders
def _rotation(angle):
    a = f32(angle)
    return (f32(math.cos(a)), f32(math.sin(a)))  # UnitComplex::new(angle)


class CuboidCollider:
    """A rapier2d cuboid collider + `CudaColliderOptions` (cuda_collider_set.rs:22-39,63-70) -> 72-byte dim2 SpkGpuCollider."""

    def __init__(self, half_extents, translation=(0, 0), rotation=0.0, friction=0.5, penalty_stiffness=0.0,
                 grid_boundary_handling=BoundaryHandling.Friction):
        self.half_extents, self.translation, self.rotation = half_extents, translation, rotation
        self.friction, self.penalty_stiffness, self.grid_boundary_handling = friction, penalty_stiffness, grid_boundary_handling

    def _tail(self, buf):
        np.frombuffer(buf, "<f4", 2, 40)[:] = _rotation(self.rotation)
        np.frombuffer(buf, "<f4", 2, 48)[:] = self.translation
        np.frombuffer(buf, "<f4", 2, 56)[:] = [self.friction, self.penalty_stiffness]
        np.frombuffer(buf, "<u4", 1, 64)[0] = self.grid_boundary_handling

    def to_gpu_collider(self) -> bytes:
        buf = bytearray(72)
        np.frombuffer(buf, "<u4", 1, 0)[0] = 0
        np.frombuffer(buf, "<f4", 2, 8)[:] = self.half_extents
        self._tail(buf)
        return bytes(buf)


class PolylineCollider(CuboidCollider):
    """A rapier2d polyline collider (closed polygon; counter-clockwise = solid interior) -> 72-byte dim2 SpkGpuCollider with a
    HOST pointer to the n x 2 vertices.  Keep the object alive until the upload."""

    def __init__(self, vertices, translation=(0, 0), rotation=0.0, friction=0.5, penalty_stiffness=0.0,
                 grid_boundary_handling=BoundaryHandling.Friction, flip_interior=False):
        v = np.ascontiguousarray(np.asarray(vertices, np.float32))
        assert v.ndim == 2 and v.shape[1] == 2
        super().__init__((0, 0), translation, rotation, friction, penalty_stiffness, grid_boundary_handling)
        self.vertices, self.flip_interior = v, flip_interior

    def to_gpu_collider(self) -> bytes:
        buf = bytearray(72)
        np.frombuffer(buf, "<u4", 1, 0)[0] = 3
        np.frombuffer(buf, "<u8", 1, 8)[0] = self.vertices.ctypes.data
        np.frombuffer(buf, "<u4", 1, 16)[0] = len(self.vertices)
        np.frombuffer(buf, "u1", 1, 20)[0] = int(self.flip_interior)
        self._tail(buf)
        return bytes(buf)


class HeightFieldCollider(CuboidCollider):
    """A rapier2d `ColliderBuilder::heightfield(heights, scale)` (examples2d/basic2.rs:32-41) -> 72-byte dim2 SpkGpuCollider with a
    HOST pointer to the height samples; the field spans [-scale.x / 2, scale.x / 2].  Keep the object alive until the upload."""

    def __init__(self, heights, scale, translation=(0, 0), rotation=0.0, friction=0.5, penalty_stiffness=0.0,
                 grid_boundary_handling=BoundaryHandling.Friction, flip_interior=False):
        h = np.ascontiguousarray(np.asarray(heights, np.float32))
        assert h.ndim == 1 and len(h) >= 2
        super().__init__((0, 0), translation, rotation, friction, penalty_stiffness, grid_boundary_handling)
        self.heights, self.scale, self.flip_interior = h, tuple(scale), flip_interior

    def to_gpu_collider(self) -> bytes:
        buf = bytearray(72)
        np.frombuffer(buf, "<u4", 1, 0)[0] = 1
        np.frombuffer(buf, "<u8", 1, 8)[0] = self.heights.ctypes.data
        np.frombuffer(buf, "<u4", 1, 16)[0] = len(self.heights)
        np.frombuffer(buf, "<f4", 2, 20)[:] = self.scale
        np.frombuffer(buf, "u1", 2, 28)[:] = [1, int(self.flip_interior)]  # bellow_heightfield_is_solid: cuda_collider_set.rs:77
        self._tail(buf)
        return bytes(buf)


class TriMeshCollider(CuboidCollider):
    """A rapier2d `ColliderBuilder::trimesh(vertices, indices)`: parry2d's TriMesh, a soup of solid 2D triangles -> 72-byte dim2
    SpkGpuCollider with HOST pointers to the n x 2 vertices and the m x 3 indices.  Keep the object alive until the upload."""

    def __init__(self, vertices, indices, translation=(0, 0), rotation=0.0, friction=0.5, penalty_stiffness=0.0,
                 grid_boundary_handling=BoundaryHandling.Friction, flip_interior=False):
        v = np.ascontiguousarray(np.asarray(vertices, np.float32))
        t = np.ascontiguousarray(np.asarray(indices, np.uint32))
        assert v.ndim == 2 and v.shape[1] == 2 and t.ndim == 2 and t.shape[1] == 3
        super().__init__((0, 0), translation, rotation, friction, penalty_stiffness, grid_boundary_handling)
        self.vertices, self.indices, self.flip_interior = v, t, flip_interior

    def to_gpu_collider(self) -> bytes:
        buf = bytearray(72)
        np.frombuffer(buf, "<u4", 1, 0)[0] = 2
        np.frombuffer(buf, "<u8", 2, 8)[:] = [self.vertices.ctypes.data, self.indices.ctypes.data]
        np.frombuffer(buf, "<u4", 2, 24)[:] = [len(self.vertices), len(self.indices)]
        np.frombuffer(buf, "u1", 1, 32)[0] = int(self.flip_interior)
        self._tail(buf)
        return bytes(buf)


def pack_colliders(colliders) -> np.ndarray:
    out = np.zeros((len(colliders), 72), dtype=np.uint8)
    for i, c in enumerate(colliders):
        out[i] = np.frombuffer(c.to_gpu_collider(), dtype=np.uint8)
    return out


# ---------------------------------------------------------------------------------------------- particles
class ParticleComponents:
    """`ParticleComponents` of cuda_particle_set.rs:180-187 (dim2 layouts)."""

    def __init__(self, n=0):
        self.position = np.zeros(n, POSITION)
        self.velocity = np.zeros(n, VELOCITY)
        self.volume = np.zeros(n, VOLUME)
        self.status = np.zeros(n, STATUS)
        self.phase = np.zeros(n, PHASE)
        self.data = np.zeros(n, DATA)

    def __len__(self):
        return len(self.position)

    @staticmethod
    def concatenate(parts):
        out = ParticleComponents(0)
        for name in ("position", "velocity", "volume", "status", "phase", "data"):
            setattr(out, name, np.concatenate([getattr(p, name) for p in parts]))
        return out

    def take(self, idx):
        out = ParticleComponents(0)
        for name in ("position", "velocity", "volume", "status", "phase", "data"):
            setattr(out, name, np.ascontiguousarray(getattr(self, name)[idx]))
        return out


def particles_from_positions(points: np.ndarray, model_index: int, radius, density0) -> ParticleComponents:
    """`Particle::new` (src/dynamics/particle.rs:53-104, dim2: volume0 = (2 r)^2) + `extract_particles_components`."""
    n = len(points)
    pc = ParticleComponents(n)
    pc.position["point"] = points.astype(np.float32)
    d = f32(radius) * f32(2.0)
    volume0 = f32(d * d)
    pc.volume["mass"] = f32(volume0 * f32(density0))
    pc.volume["volume0"] = volume0
    pc.volume["radius0"] = f32(radius)
    pc.volume["deformation_gradient"] = np.array([1, 0, 0, 1], np.float32)
    pc.volume["plastic_deformation_gradient_det"] = 1.0
    pc.status["model_index"] = model_index
    pc.phase["phase"] = 1.0
    pc.data["nacc_alpha"] = -0.01
    pc.data["plastic_hardening"] = 1.0
    pc.data["elastic_hardening"] = 1.0
    return pc


def rect_particles(origin, ni, nj, model_index, particle_rad, particle_density) -> ParticleComponents:
    """examples2d/helper.rs cube_particles with randomize = false: origin + 2r*(i, j), i-major insertion order."""
    r2 = f32(f32(particle_rad) * f32(2.0))
    i = (np.arange(ni, dtype=np.float32) * r2).astype(np.float32)
    j = (np.arange(nj, dtype=np.float32) * r2).astype(np.float32)
    pts = np.empty((ni, nj, 2), np.float32)
    pts[..., 0] = (f32(origin[0]) + i)[:, None]
    pts[..., 1] = (f32(origin[1]) + j)[None, :]
    return particles_from_positions(pts.reshape(-1, 2), model_index, particle_rad, particle_density)
