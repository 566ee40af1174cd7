//! `extern "C"` declarations of include/sparkl_abi.h and the drop-in replacement of `src/cuda/cuda_mpm_pipeline.rs`.
//! The `#[repr(C)]` types are sparkl's own (`sparkl_core::dynamics::*`, `kernels::GpuParticleModel`, `GpuCollider`):
//! the C ABI takes their byte images unchanged.  NOT compiled here (no Rust toolchain in the image).
use crate::core::dynamics::solver::SolverParameters;
use crate::core::prelude::{ParticleData, ParticlePhase, ParticlePosition, ParticleStatus, ParticleVelocity, ParticleVolume};
use crate::kernels::{GpuCollider, GpuParticleModel};
use crate::math::{Real, Vector};
use std::os::raw::{c_int, c_void};

#[repr(C)]
pub struct SpkCtx { _private: [u8; 0] }

#[repr(C)]
#[derive(Default, Copy, Clone, Debug)]
pub struct SpkStepTimings { // ≙ CudaDeviceTimings, cuda_mpm_pipeline.rs:56-85
    pub num_substeps: u32, pub last_dt: f32, pub grid_resize_and_sort_ms: f32, pub reset_grid_ms: f32, pub estimate_timestep_ms: f32,
    pub g2p2g_ms: f32, pub grid_update_ms: f32, pub total_ms: f32, pub num_active_blocks: u32, pub num_dispatch_blocks: u32,
    pub g2p2g_launches: u32, pub kernel_launches: u32,
}

extern "C" {
    pub fn spk_ctx_create(device: c_int, cell_width: f32, block_capacity: u32, out: *mut *mut SpkCtx) -> c_int;
    pub fn spk_ctx_destroy(ctx: *mut SpkCtx) -> c_int;
    pub fn spk_particles_upload(ctx: *mut SpkCtx, status: *const ParticleStatus, pos: *const ParticlePosition, vel: *const ParticleVelocity,
                                volume: *const ParticleVolume, phase: *const ParticlePhase, data: *const ParticleData, n: usize) -> c_int;
    pub fn spk_particles_append(ctx: *mut SpkCtx, status: *const ParticleStatus, pos: *const ParticlePosition, vel: *const ParticleVelocity,
                                volume: *const ParticleVolume, phase: *const ParticlePhase, data: *const ParticleData, n: usize) -> c_int;
    pub fn spk_particles_remove_range(ctx: *mut SpkCtx, begin: usize, end: usize) -> c_int;
    pub fn spk_particles_len(ctx: *const SpkCtx, out: *mut usize) -> c_int;
    pub fn spk_particles_truncate(ctx: *mut SpkCtx, new_len: usize) -> c_int;
    pub fn spk_particles_capacity(ctx: *const SpkCtx, out: *mut usize) -> c_int;
    pub fn spk_particles_read(ctx: *mut SpkCtx, which: c_int, dst: *mut c_void, n: usize) -> c_int;
    pub fn spk_models_upload(ctx: *mut SpkCtx, models: *const GpuParticleModel, n: usize) -> c_int;
    pub fn spk_colliders_upload(ctx: *mut SpkCtx, colliders: *const GpuCollider, n: usize) -> c_int;
    pub fn spk_step(ctxs: *mut *mut SpkCtx, nctx: usize, params: *const SolverParameters, gravity: *const f32,
                    timings: *mut SpkStepTimings, simulated_time_out: *mut f32) -> c_int;
    // multi-GPU over NVLink peer memory, one process per GPU (sparkl_b200/csrc/spk_dist.cu)
    pub fn spk_dist_export(ctx: *mut SpkCtx, out: *mut SpkDistDescriptor) -> c_int;
    pub fn spk_dist_attach(ctx: *mut SpkCtx, rank: u32, world: u32, all_ranks: *const SpkDistDescriptor) -> c_int;
    pub fn spk_dist_detach(ctx: *mut SpkCtx) -> c_int;
    // device-side particle migration between slabs (sparkl_b200/csrc/spk_migrate.cu): rows in DEVICE memory, transport left to the host
    pub fn spk_particles_set_gids(ctx: *mut SpkCtx, gids: *const u32, n: usize) -> c_int;
    pub fn spk_particles_read_gids(ctx: *mut SpkCtx, dst: *mut u32, n: usize) -> c_int;
    pub fn spk_particles_migrate_out(ctx: *mut SpkCtx, bounds: *const i64, world: u32, rank: u32, dev_records: *mut SpkParticleRecord, capacity: usize,
                                     counts_out: *mut u32) -> c_int;
    pub fn spk_particles_migrate_in(ctx: *mut SpkCtx, dev_records: *const SpkParticleRecord, n: usize) -> c_int;
    // MlsSolver::recompute_fluids_volumes as an explicit stage (sparkl_b200/csrc/spk_fluids.cu)
    pub fn spk_stage_recompute_fluids_volume(ctx: *mut SpkCtx) -> c_int;
    /// box / per-axis bit counts of the Morton curve the last particle sort ranked the active blocks along (diagnostic; out: [u32; 8])
    pub fn spk_grid_morton_box(ctx: *mut SpkCtx, out: *mut u32) -> c_int;
}

/// Wire format of one migrating particle (128 bytes in dim3): the six per-particle structs + a global id.
#[repr(C)]
#[derive(Copy, Clone)]
pub struct SpkParticleRecord {
    pub status: ParticleStatus, pub position: ParticlePosition, pub velocity: ParticleVelocity, pub volume: ParticleVolume,
    pub phase: ParticlePhase, pub data: ParticleData, pub gid: u32,
}

/// Opaque 128-byte blob (CUDA IPC handle of the rank's shared arena + its capacity); all-gather it with the host's transport.
#[repr(C)]
#[derive(Copy, Clone)]
pub struct SpkDistDescriptor { pub opaque: [u8; 128] }

/// The HeightField payload of the 336-byte collider image (include/sparkl_abi.h): a HOST description of rapier's
/// `HeightField` instead of parry's `CudaHeightFieldPtr` (which holds device pointers of cust's context).
#[repr(C)]
#[derive(Copy, Clone)]
pub struct SpkHeightFieldPayload {
    pub heights: *const f32, // column-major DMatrix storage: `heightfield.heights().as_slice().as_ptr()`
    pub nrows: u32, pub ncols: u32,
    pub scale: [f32; 3],
    pub bellow_heightfield_is_solid: u8, pub flip_interior: u8,
}

/// `CudaResult`-shaped error: positive = cudaError_t, negative = SPK_E* (API misuse).
#[derive(Debug, Copy, Clone, PartialEq, Eq)]
pub struct SpkError(pub i32);
fn ck(rc: c_int) -> Result<(), SpkError> { if rc == 0 { Ok(()) } else { Err(SpkError(rc)) } }

/// Owns what `SingleGpuMpmContext` borrows in the reference: context, streams, particles, models, colliders, grid.
pub struct B200MpmContext { raw: *mut SpkCtx }

impl B200MpmContext {
    pub fn new(device: i32, cell_width: Real) -> Result<Self, SpkError> {
        let mut raw = std::ptr::null_mut();
        ck(unsafe { spk_ctx_create(device, cell_width, 1024, &mut raw) })?; // CudaSparseGrid::new(cell_width)
        Ok(Self { raw })
    }
    /// ≙ CudaParticleSet::from_particles + SharedCudaVec::<ParticleData>::from_slice (testbed_plugin.rs:208-217)
    pub fn set_particles(&mut self, c: &crate::cuda::ParticleComponents) -> Result<(), SpkError> {
        assert_eq!(c.status.len(), c.position.len(), "All attribute buffer must have the same length.");
        ck(unsafe { spk_particles_upload(self.raw, c.status.as_ptr(), c.position.as_ptr(), c.velocity.as_ptr(), c.volume.as_ptr(),
                                         c.phase.as_ptr(), c.data.as_ptr(), c.status.len()) })
    }
    /// ≙ CudaParticleModelSet::from_models / ensure_models_are_uploaded; the DevicePointer slots may be null.
    pub fn set_models(&mut self, models: &[GpuParticleModel]) -> Result<(), SpkError> {
        ck(unsafe { spk_models_upload(self.raw, models.as_ptr(), models.len()) })
    }
    /// ≙ CudaColliderSet::from_collider_set after the rapier → GpuCollider conversion.  Cuboid images pass unchanged; for a
    /// heightfield write tag 1 at offset 0 and a `SpkHeightFieldPayload` at offset 8 of the 336-byte image
    /// (`cuda_collider_set.rs:71-86` builds a `CudaHeightFieldPtr` there instead); pose / friction / boundary handling stay
    /// at their offsets.  Polyline: tag 3, then { vertices: *const f32 (n x 3, host), num_vertices: u32, flip_interior: u8 } at
    /// offset 8.  TriMesh is rejected with SPK_EUNSUPPORTED.
    pub fn set_colliders(&mut self, colliders: &[GpuCollider]) -> Result<(), SpkError> {
        ck(unsafe { spk_colliders_upload(self.raw, colliders.as_ptr(), colliders.len()) })
    }
    pub fn read_positions(&mut self) -> Result<Vec<ParticlePosition>, SpkError> {
        let mut n = 0usize;
        ck(unsafe { spk_particles_len(self.raw, &mut n) })?;
        let mut out = vec![ParticlePosition::default(); n];
        ck(unsafe { spk_particles_read(self.raw, 0, out.as_mut_ptr() as *mut c_void, n) })?;
        Ok(out)
    }
}
impl B200MpmContext {
    /// Multi-GPU, one process per GPU: call on every rank after `set_particles`, all-gather the descriptors (MPI, files, ...),
    /// then `dist_attach`.  Afterwards `B200MpmPipeline::step` runs the halo accumulation and the dt minimum itself over NVLink.
    pub fn dist_export(&mut self) -> Result<SpkDistDescriptor, SpkError> {
        let mut d = SpkDistDescriptor { opaque: [0; 128] };
        ck(unsafe { spk_dist_export(self.raw, &mut d) })?;
        Ok(d)
    }
    pub fn dist_attach(&mut self, rank: u32, all_ranks: &[SpkDistDescriptor]) -> Result<(), SpkError> {
        ck(unsafe { spk_dist_attach(self.raw, rank, all_ranks.len() as u32, all_ranks.as_ptr()) })
    }
}
impl Drop for B200MpmContext { fn drop(&mut self) { unsafe { spk_ctx_destroy(self.raw); } } }

// ---- custom particle kernels: `CudaParticleKernelsLauncher` (src/cuda/cuda_particle_kernels.rs:7-27) over the C ABI.
// The trait object is boxed and reached through three `extern "C"` trampolines; the kernels it launches are the user's
// instantiation of `g2p2g_kernel<MODE, MyUpdater>` (sparkl_b200/csrc/spk_g2p2g_kernel.cuh), compiled by the user's build.rs.
#[repr(C)]
pub struct SpkParticleKernelsLauncher {
    pub user: *mut c_void,
    pub timestep_started: Option<unsafe extern "C" fn(user: *mut c_void) -> c_int>,
    pub launch_estimate_particle_timestep_length: Option<unsafe extern "C" fn(user: *mut c_void, ctx: *mut SpkCtx, min_dt: f32, max_dt: f32) -> c_int>,
    pub launch_g2p2g: Option<unsafe extern "C" fn(user: *mut c_void, ctx: *mut SpkCtx, timestep_length: f32) -> c_int>,
}
extern "C" {
    pub fn spk_set_particle_kernels_launcher(ctx: *mut SpkCtx, launcher: *const SpkParticleKernelsLauncher) -> c_int;
    pub fn spk_default_launch_estimate_particle_timestep_length(ctx: *mut SpkCtx, min_dt: f32, max_dt: f32) -> c_int;
    pub fn spk_default_launch_g2p2g(ctx: *mut SpkCtx, timestep_length: f32) -> c_int;
}
/// What a custom launcher implements on the Rust side (same three methods as the reference's trait, minus the arguments the
/// library already owns: grid / block shapes, the halo flag, the solver parameters).
pub trait B200ParticleKernelsLauncher {
    fn timestep_started(&mut self) -> Result<(), SpkError> { Ok(()) }
    fn launch_estimate_particle_timestep_length(&self, ctx: *mut SpkCtx, min_dt: f32, max_dt: f32) -> Result<(), SpkError> {
        ck(unsafe { spk_default_launch_estimate_particle_timestep_length(ctx, min_dt, max_dt) })
    }
    fn launch_g2p2g(&self, ctx: *mut SpkCtx, timestep_length: f32) -> Result<(), SpkError> { ck(unsafe { spk_default_launch_g2p2g(ctx, timestep_length) }) }
}
unsafe extern "C" fn tramp_started(user: *mut c_void) -> c_int {
    (*(user as *mut Box<dyn B200ParticleKernelsLauncher>)).timestep_started().map(|_| 0).unwrap_or_else(|e| e.0)
}
unsafe extern "C" fn tramp_estimate(user: *mut c_void, ctx: *mut SpkCtx, min_dt: f32, max_dt: f32) -> c_int {
    (*(user as *mut Box<dyn B200ParticleKernelsLauncher>)).launch_estimate_particle_timestep_length(ctx, min_dt, max_dt).map(|_| 0).unwrap_or_else(|e| e.0)
}
unsafe extern "C" fn tramp_g2p2g(user: *mut c_void, ctx: *mut SpkCtx, dt: f32) -> c_int {
    (*(user as *mut Box<dyn B200ParticleKernelsLauncher>)).launch_g2p2g(ctx, dt).map(|_| 0).unwrap_or_else(|e| e.0)
}
impl B200MpmContext {
    /// `step_generic(…, custom_launchers)`: the launcher stays registered (and alive: it is leaked into the context) until replaced.
    pub fn set_particle_kernels_launcher(&mut self, launcher: Box<dyn B200ParticleKernelsLauncher>) -> Result<(), SpkError> {
        let user = Box::into_raw(Box::new(launcher)) as *mut c_void;
        let vt = SpkParticleKernelsLauncher { user, timestep_started: Some(tramp_started),
                                              launch_estimate_particle_timestep_length: Some(tramp_estimate), launch_g2p2g: Some(tramp_g2p2g) };
        ck(unsafe { spk_set_particle_kernels_launcher(self.raw, &vt) })
    }
}

/// Drop-in for `CudaMpmPipeline` (cuda_mpm_pipeline.rs:144-271).
#[derive(Default)]
pub struct B200MpmPipeline { pub step_id: u64 }

impl B200MpmPipeline {
    pub fn new() -> Self { Self { step_id: 0 } }
    /// Same contract as `CudaMpmPipeline::step`: advances by `params.dt` with adaptive substeps, returns the simulated time.
    pub fn step(&mut self, timings: Option<&mut SpkStepTimings>, params: &SolverParameters, gravity: &Vector<Real>,
                contexts: &mut [B200MpmContext]) -> Result<f32, SpkError> {
        let mut raw: Vec<*mut SpkCtx> = contexts.iter().map(|c| c.raw).collect();
        let mut simulated = 0.0f32;
        let t = timings.map(|t| t as *mut _).unwrap_or(std::ptr::null_mut());
        ck(unsafe { spk_step(raw.as_mut_ptr(), raw.len(), params, gravity.as_ptr(), t, &mut simulated) })?;
        if !t.is_null() { self.step_id += unsafe { (*t).num_substeps } as u64; }
        Ok(simulated)
    }
}
