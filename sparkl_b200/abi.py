"""Host-side mirror of the reference's `#[repr(C)]` types for the GPU MPM path (3D / f32).

numpy structured dtypes are byte images of include/sparkl_abi.h (which cites the reference file:line of every
struct); the small classes below mirror the reference's model constructors so scenes read like the reference's
examples (examples3d/sand3.rs:40-48 etc.).  All arithmetic that feeds device parameters is done in float32 so the
values are the ones the Rust f32 code would produce.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

f32 = np.float32

# ---------------------------------------------------------------------------------------------- particle PODs
POSITION = np.dtype([("point", "<f4", (3,))])                                   # particle.rs:17-20
VELOCITY = np.dtype([("vector", "<f4", (3,))])                                  # particle.rs:54-57
VOLUME = np.dtype([("mass", "<f4"), ("volume0", "<f4"), ("radius0", "<f4"),     # particle.rs:65-72
                   ("deformation_gradient", "<f4", (9,)),                       # column-major
                   ("plastic_deformation_gradient_det", "<f4")])
STATUS = np.dtype([("failed", "u1"), ("is_static", "u1"), ("kinematic_vel_enabled", "u1"), ("_pad0", "u1"),
                   ("kinematic_vel", "<f4", (3,)), ("model_index", "<u8")])     # particle.rs:119-127
PHASE = np.dtype([("phase", "<f4"), ("psi_pos", "<f4")])                        # particle.rs:144-148
DATA = np.dtype([("nacc_alpha", "<f4"), ("plastic_hardening", "<f4"),           # particle.rs:201-208
                 ("elastic_hardening", "<f4"), ("log_vol_gain", "<f4")])
ACTIVE_BLOCK_HEADER = np.dtype([("virtual_id", "<u8"), ("first_particle", "<u4"), ("num_particles", "<u4"),
                                ("halo_state", "<u4"), ("_pad0", "<u4")])       # sort.rs:29-38
DISPATCH = np.dtype([("active_block_id", "<u4"), ("first_particle", "<u4")])    # gpu_grid.rs:166-172
GRID_NODE = np.dtype([("mass", "<f4"), ("momentum_velocity", "<f4", (3,)), ("psi_momentum_velocity", "<f4"),
                      ("psi_mass", "<f4"), ("prev_mass", "<f4"), ("_pad0", "<u4"),
                      ("projection_status_tag", "<u8"), ("projection_status_payload", "<u8"),
                      ("collision_normal", "<f4", (3,)), ("projection_scaled_dir", "<f4", (3,))])  # gpu_grid.rs:443-459
assert POSITION.itemsize == 12 and VELOCITY.itemsize == 12 and VOLUME.itemsize == 52
assert STATUS.itemsize == 24 and PHASE.itemsize == 8 and DATA.itemsize == 16
assert ACTIVE_BLOCK_HEADER.itemsize == 24 and DISPATCH.itemsize == 8 and GRID_NODE.itemsize == 72

ATTR_POSITION, ATTR_VELOCITY, ATTR_STATUS, ATTR_VOLUME, ATTR_PHASE, ATTR_DATA, ATTR_SORTED_IDS = range(7)
ATTR_DTYPES = {ATTR_POSITION: POSITION, ATTR_VELOCITY: VELOCITY, ATTR_STATUS: STATUS, ATTR_VOLUME: VOLUME,
               ATTR_PHASE: PHASE, ATTR_DATA: DATA, ATTR_SORTED_IDS: np.dtype("<u4")}

PROJ_NOT_COMPUTED, PROJ_INSIDE, PROJ_OUTSIDE, PROJ_TOO_FAR = range(4)

# error codes
OK, EINVAL, EUNSUPPORTED, ESTATE, EOVERFLOW, EPEER = 0, -1, -2, -3, -4, -5


class SparklError(RuntimeError):
    """Mirrors `CudaResult::Err` / the reference's assert panics at the C-ABI boundary."""

    def __init__(self, code: int, what: str = ""):
        self.code = code
        super().__init__(f"sparkl_b200 error {code}: {what}")


# ---------------------------------------------------------------------------------------------- enums
class BoundaryHandling:  # solver_parameters.rs:17-25
    Stick, Friction, FrictionZUp, NONE = 0, 1, 2, 3


class DamageModel:  # solver_parameters.rs:27-36
    CdMpm, Eigenerosion, ModifiedEigenerosion, NONE = 0, 1, 2, 3


class SolverParameters(C.Structure):  # solver_parameters.rs:38-68
    _fields_ = [("dt", C.c_float), ("max_substep_dt", C.c_float), ("max_num_substeps", C.c_uint32),
                ("boundary_handling", C.c_uint32), ("damage_model", C.c_uint32),
                ("force_fluids_volume_recomputation", C.c_uint8), ("enable_boundary_particle_projection", C.c_uint8),
                ("stop_after_one_substep", C.c_uint8), ("_pad0", C.c_uint8), ("simulation_dofs", C.c_uint32)]

    def __init__(self, **kw):
        super().__init__()
        self.dt = 1.0 / 60.0
        self.max_substep_dt = float(np.finfo(np.float32).max)
        self.max_num_substeps = 1000
        self.boundary_handling = BoundaryHandling.Friction
        self.damage_model = DamageModel.NONE
        for k, v in kw.items():
            setattr(self, k, v)


assert C.sizeof(SolverParameters) == 28


class StepTimings(C.Structure):  # CudaDeviceTimings, cuda_mpm_pipeline.rs:56-85
    _fields_ = [("num_substeps", C.c_uint32), ("last_dt", C.c_float), ("grid_resize_and_sort_ms", C.c_float),
                ("reset_grid_ms", C.c_float), ("estimate_timestep_ms", C.c_float), ("g2p2g_ms", C.c_float),
                ("grid_update_ms", C.c_float), ("total_ms", C.c_float), ("num_active_blocks", C.c_uint32),
                ("num_dispatch_blocks", C.c_uint32), ("g2p2g_launches", C.c_uint32), ("kernel_launches", C.c_uint32)]


class SortInfo(C.Structure):
    _fields_ = [("num_active_blocks", C.c_uint32), ("num_dispatch_blocks", C.c_uint32),
                ("num_dispatch_halo_blocks", C.c_uint32), ("hashmap_capacity", C.c_uint32)]


# ---------------------------------------------------------------------------------------------- material models
def lame_lambda_mu(young_modulus, poisson_ratio):
    """src_core/utils/physics.rs:16-27, evaluated in f32."""
    e, nu = f32(young_modulus), f32(poisson_ratio)
    one, two = f32(1.0), f32(2.0)
    lam = e * nu / ((one + nu) * (one - two * nu))
    mu = e / (two * (one + nu))
    return f32(lam), f32(mu)


def shear_modulus(e, nu):  # physics.rs:25-27
    return f32(f32(e) / (f32(2.0) * (f32(1.0) + f32(nu))))


def bulk_modulus(e, nu):  # physics.rs:29-31
    return f32(f32(e) / (f32(3.0) * (f32(1.0) - f32(2.0) * f32(nu))))


def _radians(deg):  # f32::to_radians
    return f32(f32(deg) * f32(f32(math.pi) / f32(180.0)))


class CorotatedLinearElasticity:  # elasticity_corotated_linear.rs:12-29
    tag = 0

    def __init__(self, young_modulus, poisson_ratio):
        self.lam, self.mu = lame_lambda_mu(young_modulus, poisson_ratio)
        self.split_stress_on_failure = True
        self.cfl_coeff = f32(0.9)

    def pack(self, buf):
        np.frombuffer(buf, "u1", 1, 8)[0] = int(self.split_stress_on_failure)
        np.frombuffer(buf, "<f4", 3, 12)[:] = [self.cfl_coeff, self.lam, self.mu]


class NeoHookeanElasticity:  # elasticity_neo_hookean.rs:11-28
    tag = 1

    def __init__(self, young_modulus, poisson_ratio):
        self.lam, self.mu = lame_lambda_mu(young_modulus, poisson_ratio)
        self.cfl_coeff = f32(0.5)

    def pack(self, buf):
        np.frombuffer(buf, "<f4", 3, 8)[:] = [self.cfl_coeff, self.lam, self.mu]


class MonaghanSphEos:  # eos_monaghan_sph.rs:11-26
    tag = 2

    def __init__(self, pressure0, gamma, viscosity):
        self.pressure0, self.gamma, self.viscosity, self.max_neg_pressure = f32(pressure0), int(gamma), f32(viscosity), f32(1.0)

    def pack(self, buf):
        np.frombuffer(buf, "<f4", 1, 8)[0] = self.pressure0
        np.frombuffer(buf, "<u4", 1, 12)[0] = self.gamma
        np.frombuffer(buf, "<f4", 2, 16)[:] = [self.viscosity, self.max_neg_pressure]


class CustomConstitutiveModel:  # GpuConstitutiveModel::Custom(u32): zero stress, no dt bound
    tag = 3

    def __init__(self, id=0):
        self.id = id

    def pack(self, buf):
        np.frombuffer(buf, "<u4", 1, 8)[0] = self.id


class DruckerPragerPlasticity:  # plasticity_drucker_prager.rs:10-34
    tag = 0

    def __init__(self, young_modulus, poisson_ratio):
        self.lam, self.mu = lame_lambda_mu(young_modulus, poisson_ratio)
        self.h0, self.h1, self.h2, self.h3 = _radians(35.0), _radians(9.0), f32(0.2), _radians(10.0)
        self.only_active_when_failed = False
        self.volume_correction = f32(1.0)

    def pack(self, buf):
        np.frombuffer(buf, "<f4", 6, 40)[:] = [self.h0, self.h1, self.h2, self.h3, self.lam, self.mu]
        np.frombuffer(buf, "u1", 1, 64)[0] = int(self.only_active_when_failed)
        np.frombuffer(buf, "<f4", 1, 68)[0] = self.volume_correction


class NaccPlasticity:  # plasticity_nacc.rs:12-59
    tag = 1

    def __init__(self, young_modulus, poisson_ratio, cohesion, hardening_enabled, hardening_factor, friction_angle):
        sin_f = f32(math.sin(f32(friction_angle)))
        d = f32(3.0)
        self.mu, self.kappa = shear_modulus(young_modulus, poisson_ratio), bulk_modulus(young_modulus, poisson_ratio)
        self.hardening_enabled, self.hardening_factor, self.cohesion = bool(hardening_enabled), f32(hardening_factor), f32(cohesion)
        self.friction = f32(np.sqrt(f32(2.0) / f32(3.0)) * f32(2.0) * sin_f / (f32(3.0) - sin_f) * d / np.sqrt(f32(2.0) / (f32(6.0) - d)))

    def pack(self, buf):
        np.frombuffer(buf, "<f4", 2, 40)[:] = [self.mu, self.kappa]
        np.frombuffer(buf, "u1", 1, 48)[0] = int(self.hardening_enabled)
        np.frombuffer(buf, "<f4", 3, 52)[:] = [self.hardening_factor, self.cohesion, self.friction]


class RankinePlasticity:  # plasticity_rankine.rs:11-33
    tag = 2

    def __init__(self, young_modulus, poisson_ratio, tensile_strength, softening_rate):
        self.lam, self.mu = lame_lambda_mu(young_modulus, poisson_ratio)
        self.tensile_strength, self.softening_rate = f32(tensile_strength), f32(softening_rate)

    def pack(self, buf):
        np.frombuffer(buf, "<f4", 4, 40)[:] = [self.mu, self.lam, self.tensile_strength, self.softening_rate]


class SnowPlasticity:  # plasticity_snow.rs:11-24
    tag = 3

    def __init__(self):
        self.min_epsilon, self.max_epsilon, self.hardening_coeff = f32(2.5e-2), f32(4.5e-3), f32(10.0)

    def pack(self, buf):
        np.frombuffer(buf, "<f4", 3, 40)[:] = [self.min_epsilon, self.max_epsilon, self.hardening_coeff]


class MaximumStressFailure:  # failure_maximum_stress.rs:7-18
    tag = 0

    def __init__(self, max_principal_stress, max_shear_stress):
        self.max_principal_stress, self.max_shear_stress = f32(max_principal_stress), f32(max_shear_stress)

    def pack(self, buf):
        np.frombuffer(buf, "<f4", 2, 84)[:] = [self.max_principal_stress, self.max_shear_stress]


class ParticleModel:
    """`ParticleModel::new / with_plasticity / with_failure` (src/dynamics/particle_model.rs:20-82) reduced to the
    `Core*Model` triple the GPU path consumes (cuda_particle_model_set.rs:39-58)."""

    def __init__(self, constitutive_model, plastic_model=None, failure_model=None):
        self.constitutive_model, self.plastic_model, self.failure_model = constitutive_model, plastic_model, failure_model

    @classmethod
    def with_plasticity(cls, constitutive_model, plasticity):
        return cls(constitutive_model, plasticity)

    def to_gpu_model(self) -> bytes:
        """96-byte GpuParticleModel image (src_kernels/gpu_particle_model.rs:7-13); device-pointer slots left null."""
        buf = bytearray(96)
        np.frombuffer(buf, "<u4", 1, 0)[0] = self.constitutive_model.tag
        self.constitutive_model.pack(buf)
        np.frombuffer(buf, "<u4", 1, 32)[0] = 5 if self.plastic_model is None else self.plastic_model.tag
        if self.plastic_model is not None:
            self.plastic_model.pack(buf)
        np.frombuffer(buf, "<u4", 1, 80)[0] = 2 if self.failure_model is None else self.failure_model.tag
        if self.failure_model is not None:
            self.failure_model.pack(buf)
        return bytes(buf)


def pack_models(models) -> np.ndarray:
    out = np.zeros((len(models), 96), dtype=np.uint8)
    for i, m in enumerate(models):
        out[i] = np.frombuffer(m.to_gpu_model(), dtype=np.uint8)
    return out


# ---------------------------------------------------------------------------------------------- colliders
class CuboidCollider:
    """A rapier cuboid collider + `CudaColliderOptions` (cuda_collider_set.rs:22-39,63-70) -> 336-byte GpuCollider."""

    def __init__(self, half_extents, translation=(0, 0, 0), rotation=(0, 0, 0, 1), friction=0.5, penalty_stiffness=0.0,
                 grid_boundary_handling=BoundaryHandling.Friction):
        self.half_extents, self.translation, self.rotation = half_extents, translation, rotation
        self.friction, self.penalty_stiffness, self.grid_boundary_handling = friction, penalty_stiffness, grid_boundary_handling

    def to_gpu_collider(self) -> bytes:
        buf = bytearray(336)
        np.frombuffer(buf, "<u4", 1, 0)[0] = 0
        np.frombuffer(buf, "<f4", 3, 8)[:] = self.half_extents
        np.frombuffer(buf, "<f4", 4, 296)[:] = self.rotation
        np.frombuffer(buf, "<f4", 3, 312)[:] = self.translation
        np.frombuffer(buf, "<f4", 2, 324)[:] = [self.friction, self.penalty_stiffness]
        np.frombuffer(buf, "<u4", 1, 332)[0] = self.grid_boundary_handling
        return bytes(buf)


class HeightFieldCollider:
    """A rapier `ColliderBuilder::heightfield(heights, scale)` + `CudaColliderOptions` -> 336-byte GpuCollider image with the
    HeightField payload of include/sparkl_abi.h (a HOST pointer to the column-major heights matrix; rows <-> z, columns
    <-> x; the field spans [-scale/2, scale/2] in x and z, parry3d heightfield3.rs).  Keep the object alive until the upload."""

    def __init__(self, heights, scale, translation=(0, 0, 0), rotation=(0, 0, 0, 1), friction=0.5, penalty_stiffness=0.0,
                 grid_boundary_handling=BoundaryHandling.Friction, flip_interior=False):
        h = np.asarray(heights, np.float32)
        assert h.ndim == 2 and h.shape[0] >= 2 and h.shape[1] >= 2
        self.nrows, self.ncols = h.shape
        self.heights = np.asfortranarray(h)  # nalgebra DMatrix storage
        self.scale, self.translation, self.rotation = tuple(scale), translation, rotation
        self.friction, self.penalty_stiffness, self.grid_boundary_handling = friction, penalty_stiffness, grid_boundary_handling
        self.flip_interior = flip_interior

    def to_gpu_collider(self) -> bytes:
        buf = bytearray(336)
        np.frombuffer(buf, "<u4", 1, 0)[0] = 1
        np.frombuffer(buf, "<u8", 1, 8)[0] = self.heights.ctypes.data
        np.frombuffer(buf, "<u4", 2, 16)[:] = [self.nrows, self.ncols]
        np.frombuffer(buf, "<f4", 3, 24)[:] = self.scale
        np.frombuffer(buf, "u1", 2, 36)[:] = [1, int(self.flip_interior)]  # bellow_heightfield_is_solid: cuda_collider_set.rs:77
        np.frombuffer(buf, "<f4", 4, 296)[:] = self.rotation
        np.frombuffer(buf, "<f4", 3, 312)[:] = self.translation
        np.frombuffer(buf, "<f4", 2, 324)[:] = [self.friction, self.penalty_stiffness]
        np.frombuffer(buf, "<u4", 1, 332)[0] = self.grid_boundary_handling
        return bytes(buf)


class PolylineCollider:
    """A rapier polyline collider (closed polygon in the xy plane, extruded along z in 3D) -> 336-byte GpuCollider image with
    the Polyline payload of include/sparkl_abi.h (HOST pointer to the n x 3 vertices).  Counter-clockwise = solid interior
    (parry's Polyline::project_local_point_assuming_solid_interior_ccw).  Keep the object alive until the upload."""

    def __init__(self, vertices, translation=(0, 0, 0), rotation=(0, 0, 0, 1), friction=0.5, penalty_stiffness=0.0,
                 grid_boundary_handling=BoundaryHandling.Friction, flip_interior=False):
        v = np.ascontiguousarray(np.asarray(vertices, np.float32))
        if v.ndim == 2 and v.shape[1] == 2:
            v = np.ascontiguousarray(np.concatenate([v, np.zeros((len(v), 1), np.float32)], axis=1))
        assert v.ndim == 2 and v.shape[1] == 3
        self.vertices, self.translation, self.rotation = v, translation, rotation
        self.friction, self.penalty_stiffness, self.grid_boundary_handling = friction, penalty_stiffness, grid_boundary_handling
        self.flip_interior = flip_interior

    def to_gpu_collider(self) -> bytes:
        buf = bytearray(336)
        np.frombuffer(buf, "<u4", 1, 0)[0] = 3
        np.frombuffer(buf, "<u8", 1, 8)[0] = self.vertices.ctypes.data
        np.frombuffer(buf, "<u4", 1, 16)[0] = len(self.vertices)
        np.frombuffer(buf, "u1", 1, 20)[0] = int(self.flip_interior)
        np.frombuffer(buf, "<f4", 4, 296)[:] = self.rotation
        np.frombuffer(buf, "<f4", 3, 312)[:] = self.translation
        np.frombuffer(buf, "<f4", 2, 324)[:] = [self.friction, self.penalty_stiffness]
        np.frombuffer(buf, "<u4", 1, 332)[0] = self.grid_boundary_handling
        return bytes(buf)


class TriMeshCollider:
    """A rapier `ColliderBuilder::trimesh(vertices, indices)` collider (default flags: a surface without pseudo-normals) -> 336-byte
    GpuCollider image with the TriMesh payload of include/sparkl_abi.h (HOST pointers to the n x 3 vertices and the t x 3 indices).
    Keep the object alive until the upload."""

    def __init__(self, vertices, indices, translation=(0, 0, 0), rotation=(0, 0, 0, 1), friction=0.5, penalty_stiffness=0.0,
                 grid_boundary_handling=BoundaryHandling.Friction, flip_interior=False):
        self.vertices = np.ascontiguousarray(np.asarray(vertices, np.float32))
        self.indices = np.ascontiguousarray(np.asarray(indices, np.uint32))
        assert self.vertices.ndim == 2 and self.vertices.shape[1] == 3 and self.indices.ndim == 2 and self.indices.shape[1] == 3
        self.translation, self.rotation = translation, rotation
        self.friction, self.penalty_stiffness, self.grid_boundary_handling = friction, penalty_stiffness, grid_boundary_handling
        self.flip_interior = flip_interior

    def to_gpu_collider(self) -> bytes:
        buf = bytearray(336)
        np.frombuffer(buf, "<u4", 1, 0)[0] = 2
        np.frombuffer(buf, "<u8", 2, 8)[:] = [self.vertices.ctypes.data, self.indices.ctypes.data]
        np.frombuffer(buf, "<u4", 2, 24)[:] = [len(self.vertices), len(self.indices)]
        np.frombuffer(buf, "u1", 1, 32)[0] = int(self.flip_interior)
        np.frombuffer(buf, "<f4", 4, 296)[:] = self.rotation
        np.frombuffer(buf, "<f4", 3, 312)[:] = self.translation
        np.frombuffer(buf, "<f4", 2, 324)[:] = [self.friction, self.penalty_stiffness]
        np.frombuffer(buf, "<u4", 1, 332)[0] = self.grid_boundary_handling
        return bytes(buf)


def pack_colliders(colliders) -> np.ndarray:
    out = np.zeros((len(colliders), 336), dtype=np.uint8)
    for i, c in enumerate(colliders):
        out[i] = np.frombuffer(c.to_gpu_collider(), dtype=np.uint8)
    return out


# ---------------------------------------------------------------------------------------------- particles
class ParticleComponents:
    """`ParticleComponents` of cuda_particle_set.rs:180-187: one array per attribute, index = particle id."""

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
    """`Particle::new` (src/dynamics/particle.rs:53-104) + `extract_particles_components` (cuda_particle_set.rs:189-248)."""
    n = len(points)
    pc = ParticleComponents(n)
    pc.position["point"] = points.astype(np.float32)
    d = f32(radius) * f32(2.0)  # (radius * 2).powi(3) == d * (d * d) in f32
    volume0 = f32(f32(d * d) * d)
    pc.volume["mass"] = f32(volume0 * f32(density0))
    pc.volume["volume0"] = volume0
    pc.volume["radius0"] = f32(radius)
    pc.volume["deformation_gradient"] = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], np.float32)
    pc.volume["plastic_deformation_gradient_det"] = 1.0
    pc.status["model_index"] = model_index
    pc.phase["phase"] = 1.0
    pc.data["nacc_alpha"] = -0.01
    pc.data["plastic_hardening"] = 1.0
    pc.data["elastic_hardening"] = 1.0
    return pc


def cube_particles(origin, ni, nj, nk, model_index, particle_rad, particle_density) -> ParticleComponents:
    """examples3d/helper.rs:6-40 with randomize = false: origin + 2r*(i,j,k), i-major / k-minor insertion order."""
    r2 = f32(f32(particle_rad) * f32(2.0))
    i = (np.arange(ni, dtype=np.float32) * r2).astype(np.float32)
    j = (np.arange(nj, dtype=np.float32) * r2).astype(np.float32)
    k = (np.arange(nk, dtype=np.float32) * r2).astype(np.float32)
    pts = np.empty((ni, nj, nk, 3), np.float32)
    pts[..., 0] = (f32(origin[0]) + i)[:, None, None]
    pts[..., 1] = (f32(origin[1]) + j)[None, :, None]
    pts[..., 2] = (f32(origin[2]) + k)[None, None, :]
    return particles_from_positions(pts.reshape(-1, 3), model_index, particle_rad, particle_density)
