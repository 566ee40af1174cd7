/*
 * sparkl_abi.h — C ABI of libsparkl_b200.so, the B200-native replacement of sparkl's GPU MPM substep.
 *
 * Every struct below is the byte image of a `#[repr(C)]` type of the reference (dimforge/sparkl 0.4.0, 3D/f32);
 * the reference file:line it mirrors is cited next to it, and the layout is pinned by static asserts.
 * Every entry point replaces one piece of the reference's Rust host API for `CudaMpmPipeline::step`
 * (src/cuda/cuda_mpm_pipeline.rs:262-659); the piece is cited at the declaration.  The Rust-side binding
 * a sparkl maintainer would add is in INTEGRATION.md (and rust_shim/).
 *
 * Conventions: all functions return 0 (SPK_OK) on success, a negative SPK_E* code for API misuse, or a
 * positive cudaError_t value for CUDA failures (mirrors `CudaResult<T>`).  Handles own all device memory;
 * callers own host arrays.  One caller thread per context at a time (same as `&mut SingleGpuMpmContext`).
 * There is NO CPU fallback: every compute entry point fails with a CUDA error if no sm_100 device exists.
 */
#ifndef SPARKL_ABI_H
#define SPARKL_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPK_ABI_VERSION 1

/* Dimension of the build.  Like the reference (crates sparkl2d / sparkl3d compiled from one source tree with the `dim2` /
 * `dim3` features, src_kernels/lib.rs:12-20), the library is built twice from the same sources: libsparkl_b200.so
 * (SPK_DIM == 3, the default) and libsparkl_b200_2d.so (-DSPK_DIM=2).  Both export the same entry points; the structs whose
 * Rust definition holds a `Vector<Real>` / `Point<Real>` / `Matrix<Real>` change size with the dimension. */
#ifndef SPK_DIM
#define SPK_DIM 3
#endif
#if SPK_DIM != 2 && SPK_DIM != 3
#error "SPK_DIM must be 2 or 3"
#endif

/* ---------------------------------------------------------------- error codes */
#define SPK_OK 0
#define SPK_EINVAL (-1)       /* bad argument (replaces the reference's assert panics, cuda_particle_set.rs:40-59) */
#define SPK_EUNSUPPORTED (-2) /* feature outside the hot-path scope (e.g. heightfield colliders) */
#define SPK_ESTATE (-3)       /* call sequence error (e.g. stage call before sort) */
#define SPK_EOVERFLOW (-4)    /* u32 particle/block limits exceeded (cuda_sparse_grid.rs:197,332) */
#define SPK_EPEER (-5)        /* multi-GPU: a peer rank did not reach a device-side barrier in time (SPK_DIST_TIMEOUT_MS, default 3000); the halo sums
                                 and the time step of that substep are incomplete — the run must be abandoned.  Sticky until spk_dist_detach */

/* ---------------------------------------------------------------- particle PODs
 * src_core/dynamics/particle.rs:14-219 (one device array per struct in the reference, cuda_particle_set.rs:12-19) */
typedef struct SpkParticlePosition { float point[SPK_DIM]; } SpkParticlePosition;   /* particle.rs:17-20  */
typedef struct SpkParticleVelocity { float vector[SPK_DIM]; } SpkParticleVelocity;  /* particle.rs:54-57  */
typedef struct SpkParticleVolume {                                                   /* particle.rs:65-72  */
    float mass;
    float volume0;
    float radius0;
    float deformation_gradient[SPK_DIM * SPK_DIM]; /* column-major (nalgebra Matrix3 / Matrix2) */
    float plastic_deformation_gradient_det;
} SpkParticleVolume;
typedef struct SpkParticleStatus {                                                   /* particle.rs:119-127 */
    uint8_t failed;
    uint8_t is_static;
    uint8_t kinematic_vel_enabled;
    uint8_t _pad0;
    float kinematic_vel[SPK_DIM];
#if SPK_DIM == 2
    uint32_t _pad1;       /* usize alignment */
#endif
    uint64_t model_index;
} SpkParticleStatus;
typedef struct SpkParticlePhase { float phase; float psi_pos; } SpkParticlePhase;   /* particle.rs:144-148 */
typedef struct SpkParticleData {                                                     /* particle.rs:201-208 */
    float nacc_alpha;
    float plastic_hardening;
    float elastic_hardening;
    float log_vol_gain;
} SpkParticleData;

/* ---------------------------------------------------------------- solver parameters
 * src_core/dynamics/solver/solver_parameters.rs:17-68 */
enum { SPK_BOUNDARY_STICK = 0, SPK_BOUNDARY_FRICTION = 1, SPK_BOUNDARY_FRICTION_Z_UP = 2, SPK_BOUNDARY_NONE = 3 };
enum { SPK_DAMAGE_CDMPM = 0, SPK_DAMAGE_EIGENEROSION = 1, SPK_DAMAGE_MODIFIED_EIGENEROSION = 2, SPK_DAMAGE_NONE = 3 };
typedef struct SpkSolverParameters {
    float dt;                  /* default 1/60 */
    float max_substep_dt;      /* default FLT_MAX */
    uint32_t max_num_substeps; /* default 1000 */
    uint32_t boundary_handling; /* ignored on the GPU path (per-collider instead), kept for layout */
    uint32_t damage_model;      /* passed to g2p2g but unused there (g2p2g.rs:190) */
    uint8_t force_fluids_volume_recomputation;   /* ignored by the CUDA pipeline */
    uint8_t enable_boundary_particle_projection; /* ignored (particle_updater.rs:204 is `if false`) */
    uint8_t stop_after_one_substep;
    uint8_t _pad0;
    uint32_t simulation_dofs;   /* ignored by the CUDA pipeline */
} SpkSolverParameters;

/* ---------------------------------------------------------------- material models
 * GpuParticleModel, 96 bytes: src_kernels/gpu_particle_model.rs:7-13.  Byte image of the Rust enums
 * (tags/offsets recovered from the shipped PTX).  The DevicePointer<ParticleData> slots
 * are ignored by this library: per-particle ParticleData is owned by the context (spk_particles_upload). */
enum { SPK_CONSTITUTIVE_COROTATED = 0, SPK_CONSTITUTIVE_NEO_HOOKEAN = 1, SPK_CONSTITUTIVE_MONAGHAN = 2, SPK_CONSTITUTIVE_CUSTOM = 3 };
enum { SPK_PLASTIC_DRUCKER_PRAGER = 0, SPK_PLASTIC_NACC = 1, SPK_PLASTIC_RANKINE = 2, SPK_PLASTIC_SNOW = 3, SPK_PLASTIC_CUSTOM = 4, SPK_PLASTIC_NONE = 5 };
enum { SPK_FAILURE_MAXIMUM_STRESS = 0, SPK_FAILURE_CUSTOM = 1, SPK_FAILURE_NONE = 2 };

typedef struct SpkCorotatedLinearElasticity { /* elasticity_corotated_linear.rs:12-17 */
    uint8_t split_stress_on_failure; uint8_t _pad[3];
    float cfl_coeff; float lambda; float mu;
} SpkCorotatedLinearElasticity;
typedef struct SpkNeoHookeanElasticity { float cfl_coeff; float lambda; float mu; } SpkNeoHookeanElasticity; /* elasticity_neo_hookean.rs:11-15 */
typedef struct SpkMonaghanSphEos { float pressure0; uint32_t gamma; float viscosity; float max_neg_pressure; } SpkMonaghanSphEos; /* eos_monaghan_sph.rs:11-16 */
typedef struct SpkDruckerPragerPlasticity { /* plasticity_drucker_prager.rs:10-19 */
    float h0, h1, h2, h3, lambda, mu;
    uint8_t only_active_when_failed; uint8_t _pad[3];
    float volume_correction;
} SpkDruckerPragerPlasticity;
typedef struct SpkNaccPlasticity { /* plasticity_nacc.rs:12-19 */
    float mu, kappa;
    uint8_t hardening_enabled; uint8_t _pad[3];
    float hardening_factor, cohesion, friction;
} SpkNaccPlasticity;
typedef struct SpkRankinePlasticity { float mu, lambda, tensile_strength, softening_rate; } SpkRankinePlasticity; /* plasticity_rankine.rs:11-16 */
typedef struct SpkSnowPlasticity { float min_epsilon, max_epsilon, hardening_coeff; } SpkSnowPlasticity;         /* plasticity_snow.rs:11-15 */
typedef struct SpkMaximumStressFailure { float max_principal_stress, max_shear_stress; } SpkMaximumStressFailure; /* failure_maximum_stress.rs:7-10 */

typedef struct SpkGpuParticleModel {
    /* GpuConstitutiveModel: src_kernels/gpu_constitutive_model.rs:12-17 */
    uint32_t constitutive_tag; uint32_t _pad0;
    union {
        struct { SpkCorotatedLinearElasticity m; uint64_t data_ptr; } corotated;
        struct { SpkNeoHookeanElasticity m; uint32_t _p; uint64_t data_ptr; } neo_hookean;
        struct { SpkMonaghanSphEos m; } monaghan;
        struct { uint32_t id; } custom;
        uint8_t _bytes[24];
    } constitutive;
    /* Option<GpuPlasticModel>: src_kernels/gpu_plastic_model.rs:12-18, None encoded as tag 5 */
    uint32_t plastic_tag; uint32_t _pad1;
    union {
        struct { SpkDruckerPragerPlasticity m; uint64_t data_ptr; } drucker_prager;
        struct { SpkNaccPlasticity m; uint64_t data_ptr; } nacc;
        struct { SpkRankinePlasticity m; uint64_t data_ptr; } rankine;
        struct { SpkSnowPlasticity m; uint32_t _p; uint64_t data_ptr; } snow;
        struct { uint32_t id; } custom;
        uint8_t _bytes[40];
    } plastic;
    /* Option<CoreFailureModel>: src_core/dynamics/models/failure_model.rs:8-11, None encoded as tag 2 */
    uint32_t failure_tag;
    union {
        SpkMaximumStressFailure maximum_stress;
        uint32_t custom_id;
    } failure;
    uint32_t _pad2;
} SpkGpuParticleModel;

/* ---------------------------------------------------------------- colliders
 * GpuCollider, 336 bytes: src_kernels/gpu_collider.rs:203-212 (offsets from the PTX).
 * Scope: Cuboid (tag 0), HeightField (tag 1), TriMesh (tag 2) and Polyline (tag 3).
 * The reference's HeightField payload is parry's `CudaHeightFieldPtr` (device pointers of ITS CUDA context,
 * src_kernels/gpu_collider.rs:14-18, src/cuda/cuda_collider_set.rs:71-86); here the payload carries a HOST description of
 * the same rapier `HeightField` (heights matrix + scale) and the library owns the device copy. */
enum { SPK_SHAPE_CUBOID = 0, SPK_SHAPE_HEIGHTFIELD = 1, SPK_SHAPE_TRIMESH = 2, SPK_SHAPE_POLYLINE = 3 };
#if SPK_DIM == 3
typedef struct SpkGpuCollider {
    uint32_t shape_tag; uint32_t _pad0;
    union {
        struct { float half_extents[3]; } cuboid;
        struct {                      /* parry3d HeightField::new(heights, scale) + GpuColliderShape::HeightField flags */
            const float* heights;     /* HOST pointer, column-major nrows x ncols (nalgebra DMatrix): rows <-> z, columns <-> x */
            uint32_t nrows, ncols;    /* number of height samples per axis (>= 2); the field spans [-scale/2, scale/2] in x and z */
            float scale[3];
            uint8_t bellow_heightfield_is_solid; /* always true in cuda_collider_set.rs:77 */
            uint8_t flip_interior;
        } heightfield;
        struct {                      /* GpuColliderShape::Polyline (src_kernels/gpu_collider.rs:24-27,105-201): a closed polygon */
            const float* vertices;    /* HOST pointer, num_vertices x (x, y, z); segment i joins vertex i and (i + 1) % n */
            uint32_t num_vertices;
            uint8_t flip_interior;
        } polyline;
        struct {                      /* GpuColliderShape::TriMesh (src_kernels/gpu_collider.rs:19-23,58-68): a rapier `ColliderBuilder::trimesh(vertices,
                                       * indices)` (no ORIENTED flag, hence no pseudo-normals: a surface, a point is "inside" only when it lies on
                                       * it; flip_interior inverts that).  The reference's payload is parry's CudaTriMeshPtr (device QBVH); here a
                                       * HOST description of the same mesh, projected by exhaustive search over the triangles. */
            const float* vertices;    /* HOST pointer, num_vertices x (x, y, z) */
            const uint32_t* indices;  /* HOST pointer, num_triangles x 3 vertex indices */
            uint32_t num_vertices, num_triangles;
            uint8_t flip_interior;
        } trimesh;
        uint8_t _bytes[288];
    } shape;
    float rotation[4];    /* unit quaternion, nalgebra coordinate order (i, j, k, w) */
    float translation[3];
    float friction;
    float penalty_stiffness;
    uint32_t grid_boundary_handling; /* SPK_BOUNDARY_* */
} SpkGpuCollider;
#else
/* dim2 GpuCollider (same Rust definition, src_kernels/gpu_collider.rs:203-212, with parry2d shapes and an Isometry2).
 * Cuboid (tag 0), HeightField (tag 1), TriMesh (tag 2) and Polyline (tag 3).  The reference's own dim2 image is 280 bytes and holds parry2d's CUDA
 * pointer types (recovered from sparkl2d-kernels.ptx: oracle/ref_harness.cpp, ref_colliders_upload); this 72-byte layout is this
 * library's own: field order and meaning as in the Rust struct, heightfield / polyline payloads as HOST descriptions like the 3D ones. */
typedef struct SpkGpuCollider {
    uint32_t shape_tag; uint32_t _pad0;
    union {
        struct { float half_extents[2]; } cuboid;
        struct {                      /* parry2d HeightField::new(heights, scale) + GpuColliderShape::HeightField flags (examples2d/basic2.rs:32-41) */
            const float* heights;     /* HOST pointer, num_heights samples (>= 2); sample i sits at x = (-0.5 + i / (n - 1)) * scale[0], y = heights[i] * scale[1] */
            uint32_t num_heights;
            float scale[2];
            uint8_t bellow_heightfield_is_solid; /* always true in cuda_collider_set.rs:77 */
            uint8_t flip_interior;
        } heightfield;
        struct {
            const float* vertices;    /* HOST pointer, num_vertices x (x, y); segment i joins vertex i and (i + 1) % n */
            uint32_t num_vertices;
            uint8_t flip_interior;
        } polyline;
        struct {                      /* GpuColliderShape::TriMesh with parry2d's TriMesh: a soup of SOLID 2D triangles.  The closest triangle
                                       * decides; with `solid = false` (grid_update.rs:59-64) a point inside a triangle is projected onto that
                                       * triangle's nearest edge and flagged inside (parry2d point_triangle.rs).  HOST description, exhaustive search. */
            const float* vertices;    /* HOST pointer, num_vertices x (x, y) */
            const uint32_t* indices;  /* HOST pointer, num_triangles x 3 vertex indices */
            uint32_t num_vertices, num_triangles;
            uint8_t flip_interior;
        } trimesh;
        uint8_t _bytes[32];
    } shape;
    float rotation[2];    /* unit complex (re, im) = (cos, sin), nalgebra UnitComplex */
    float translation[2];
    float friction;
    float penalty_stiffness;
    uint32_t grid_boundary_handling; /* SPK_BOUNDARY_*; Friction and FrictionZUp act the same in 2D (grid_update.rs:159-160) */
    uint32_t _pad1;
} SpkGpuCollider;
#endif

/* ---------------------------------------------------------------- sparse grid read-back images
 * Used by the parity tests and by hosts that inspect the grid; internal device storage is SoA. */
typedef struct SpkActiveBlockHeader { /* src_kernels/cuda/sort.rs:29-38 */
    uint64_t virtual_id;  /* BlockVirtualId: x | y<<21 | z<<42, origin 2^20; dim2: x | y<<32, origin 2^31 (gpu_grid.rs:73-120) */
    uint32_t first_particle;
    uint32_t num_particles;
    uint32_t halo_state;  /* bit0 IS_HALO, bit1 HAS_HALO_NEIGHBOR (sort.rs:15-27) */
    uint32_t _pad0;
} SpkActiveBlockHeader;
typedef struct SpkDispatchBlock2ActiveBlock { uint32_t active_block_id; uint32_t first_particle; } SpkDispatchBlock2ActiveBlock; /* gpu_grid.rs:166-172 */
enum { SPK_PROJ_NOT_COMPUTED = 0, SPK_PROJ_INSIDE = 1, SPK_PROJ_OUTSIDE = 2, SPK_PROJ_TOO_FAR = 3 };
typedef struct SpkGpuGridNode { /* src_kernels/gpu_grid.rs:443-459, 72 bytes (dim2: 56 bytes) */
    float mass;
    float momentum_velocity[SPK_DIM];
    float psi_momentum_velocity;
    float psi_mass;
    float prev_mass;
#if SPK_DIM == 3
    uint32_t _pad0;
#endif
    uint64_t projection_status_tag;     /* SPK_PROJ_* */
    uint64_t projection_status_payload; /* collider index for Inside/Outside */
    float collision_normal[SPK_DIM];
    float projection_scaled_dir[SPK_DIM];
} SpkGpuGridNode;

/* ---------------------------------------------------------------- timings
 * CudaDeviceTimings / CudaSubstepTimings (cuda_mpm_pipeline.rs:28-85), milliseconds from CUDA events,
 * accumulated over the substeps of one spk_step call. */
typedef struct SpkStepTimings {
    uint32_t num_substeps;
    float last_dt;
    float grid_resize_and_sort_ms;
    float reset_grid_ms;
    float estimate_timestep_ms;
    float g2p2g_ms;
    float grid_update_ms;
    float total_ms;
    uint32_t num_active_blocks;   /* of the last substep */
    uint32_t num_dispatch_blocks; /* of the last substep */
    uint32_t g2p2g_launches;
    uint32_t kernel_launches;     /* all kernels launched by this call */
} SpkStepTimings;

typedef struct SpkSortInfo {
    uint32_t num_active_blocks;
    uint32_t num_dispatch_blocks;
    uint32_t num_dispatch_halo_blocks;
    uint32_t hashmap_capacity;
} SpkSortInfo;

#if defined(__cplusplus)
#if SPK_DIM == 3
static_assert(sizeof(SpkParticlePosition) == 12, "ParticlePosition");
static_assert(sizeof(SpkParticleVelocity) == 12, "ParticleVelocity");
static_assert(sizeof(SpkParticleVolume) == 52, "ParticleVolume");
static_assert(sizeof(SpkGpuCollider) == 336 && offsetof(SpkGpuCollider, rotation) == 296 &&
              offsetof(SpkGpuCollider, translation) == 312 && offsetof(SpkGpuCollider, friction) == 324 &&
              offsetof(SpkGpuCollider, grid_boundary_handling) == 332, "GpuCollider");
static_assert(sizeof(SpkGpuGridNode) == 72 && offsetof(SpkGpuGridNode, projection_status_tag) == 32 &&
              offsetof(SpkGpuGridNode, collision_normal) == 48 && offsetof(SpkGpuGridNode, projection_scaled_dir) == 60, "GpuGridNode");
#else
static_assert(sizeof(SpkParticlePosition) == 8, "ParticlePosition (dim2)");
static_assert(sizeof(SpkParticleVelocity) == 8, "ParticleVelocity (dim2)");
static_assert(sizeof(SpkParticleVolume) == 32, "ParticleVolume (dim2)");
static_assert(sizeof(SpkGpuCollider) == 72 && offsetof(SpkGpuCollider, rotation) == 40 && offsetof(SpkGpuCollider, translation) == 48 &&
              offsetof(SpkGpuCollider, friction) == 56 && offsetof(SpkGpuCollider, grid_boundary_handling) == 64, "GpuCollider (dim2)");
static_assert(sizeof(SpkGpuGridNode) == 56 && offsetof(SpkGpuGridNode, projection_status_tag) == 24 &&
              offsetof(SpkGpuGridNode, collision_normal) == 40 && offsetof(SpkGpuGridNode, projection_scaled_dir) == 48, "GpuGridNode (dim2)");
#endif
static_assert(sizeof(SpkParticleStatus) == 24 && offsetof(SpkParticleStatus, kinematic_vel) == 4 &&
              offsetof(SpkParticleStatus, model_index) == 16, "ParticleStatus");
static_assert(sizeof(SpkParticlePhase) == 8, "ParticlePhase");
static_assert(sizeof(SpkParticleData) == 16, "ParticleData");
static_assert(sizeof(SpkSolverParameters) == 28 && offsetof(SpkSolverParameters, simulation_dofs) == 24, "SolverParameters");
static_assert(sizeof(SpkGpuParticleModel) == 96, "GpuParticleModel");
static_assert(offsetof(SpkGpuParticleModel, constitutive) == 8, "constitutive payload @8");
static_assert(offsetof(SpkGpuParticleModel, plastic_tag) == 32 && offsetof(SpkGpuParticleModel, plastic) == 40, "plastic @32/@40");
static_assert(offsetof(SpkGpuParticleModel, failure_tag) == 80 && offsetof(SpkGpuParticleModel, failure) == 84, "failure @80/@84");
static_assert(sizeof(SpkActiveBlockHeader) == 24, "ActiveBlockHeader");
#endif

/* ================================================================ entry points */
typedef struct spk_ctx spk_ctx;

/* which-selectors of spk_particles_read */
enum { SPK_ATTR_POSITION = 0, SPK_ATTR_VELOCITY = 1, SPK_ATTR_STATUS = 2, SPK_ATTR_VOLUME = 3, SPK_ATTR_PHASE = 4,
       SPK_ATTR_DATA = 5, SPK_ATTR_SORTED_IDS = 6 };

int spk_abi_version(void);
int spk_dim(void);                /* SPK_DIM of the loaded library: 3 (libsparkl_b200.so) or 2 (libsparkl_b200_2d.so) */
const char* spk_error_string(int code);

/* Context + 2 streams + CudaSparseGrid::with_capacity (testbed_plugin.rs:191-236, cuda_sparse_grid.rs:53-90). */
int spk_ctx_create(int device, float cell_width, uint32_t block_capacity, spk_ctx** out);
int spk_ctx_destroy(spk_ctx* ctx);
int spk_ctx_cell_width(const spk_ctx* ctx, float* out);          /* CudaSparseGrid::cell_width, cuda_sparse_grid.rs:92 */
int spk_ctx_synchronize(spk_ctx* ctx);
/* the context's main stream (`SingleGpuMpmContext::stream`, cuda_mpm_pipeline.rs:179) as a cudaStream_t, so hosts can
 * record their own events / order their own copies against the pipeline */
int spk_ctx_stream(spk_ctx* ctx, void** cuda_stream_out);

/* CudaParticleSet::from_particles + SharedCudaVec<ParticleData>::from_slice (cuda_particle_set.rs:32-71,
 * testbed_plugin.rs:208-217).  Replaces the whole set.  `data` may be NULL (ParticleData::default()). */
int spk_particles_upload(spk_ctx* ctx, const SpkParticleStatus* status, const SpkParticlePosition* pos,
                         const SpkParticleVelocity* vel, const SpkParticleVolume* volume,
                         const SpkParticlePhase* phase, const SpkParticleData* data, size_t n);
/* CudaParticleSet::append (cuda_particle_set.rs:137-177). */
int spk_particles_append(spk_ctx* ctx, const SpkParticleStatus* status, const SpkParticlePosition* pos,
                         const SpkParticleVelocity* vel, const SpkParticleVolume* volume,
                         const SpkParticlePhase* phase, const SpkParticleData* data, size_t n);
/* Asynchronous variant of spk_particles_upload for hosts that stream a new particle state every frame: the arrays (which
 * must be PINNED and stay valid until the next blocking call or spk_ctx_synchronize) are copied on a separate copy stream
 * into a dedicated staging buffer, so the transfer overlaps the frame the GPU is still computing; only the repack kernel is
 * ordered after both.  Returns without synchronising; an out-of-range model index is reported (SPK_EINVAL) by the next
 * spk_step / spk_stage_sort instead. */
int spk_particles_upload_async(spk_ctx* ctx, const SpkParticleStatus* status, const SpkParticlePosition* pos,
                               const SpkParticleVelocity* vel, const SpkParticleVolume* volume,
                               const SpkParticlePhase* phase, const SpkParticleData* data, size_t n);
/* The two halves of spk_particles_upload_async, for hosts that prefetch: _stage_async starts the host->device copy of the
 * NEXT state and returns (the current particles are untouched, spk_step may run meanwhile: the copy overlaps it);
 * _commit_staged makes the staged state the current one (repack kernel, ordered on the context's stream after the copy).
 * One staged upload at a time; committing without a staged upload is SPK_ESTATE. */
int spk_particles_stage_async(spk_ctx* ctx, const SpkParticleStatus* status, const SpkParticlePosition* pos,
                              const SpkParticleVelocity* vel, const SpkParticleVolume* volume,
                              const SpkParticlePhase* phase, const SpkParticleData* data, size_t n);
int spk_particles_commit_staged(spk_ctx* ctx);
/* CudaParticleSet::remove_range (cuda_particle_set.rs:123-131): removes ids [begin, end); later ids shift down. */
int spk_particles_remove_range(spk_ctx* ctx, size_t begin, size_t end);
/* CudaParticleSet::len (cuda_particle_set.rs:73). */
int spk_particles_len(const spk_ctx* ctx, size_t* out);
/* CudaParticleSet::truncate / capacity (cuda_particle_set.rs:77-87): drop the particles with id >= new_len (no-op when the set is
 * shorter); number of particles the device buffers hold without reallocating */
int spk_particles_truncate(spk_ctx* ctx, size_t new_len);
int spk_particles_capacity(const spk_ctx* ctx, size_t* out);
/* CudaParticleSet::read_* (cuda_particle_set.rs:93-119): copies attribute `which` of the first n particles,
 * indexed by the caller's particle id (upload order), in the reference's struct layout, into host memory. */
int spk_particles_read(spk_ctx* ctx, int which, void* dst, size_t n);
/* Asynchronous frame read-back (the testbed reads positions / velocities / phases back after every frame with blocking
 * copies, src/bin/.../testbed_plugin.rs:533-588): the attribute is unpacked into its own device staging buffer on the
 * context stream and copied to `dst` on a separate copy stream; the call returns at once and the next spk_step can be
 * enqueued immediately, so the device->host transfer overlaps the following frame.  `dst` should be pinned host memory and
 * must stay valid until spk_particles_read_wait, which blocks until every outstanding read has landed.  One read per
 * attribute may be in flight (a second one of the same attribute waits for the first). */
int spk_particles_read_async(spk_ctx* ctx, int which, void* dst, size_t n);
int spk_particles_read_wait(spk_ctx* ctx);

/* CudaParticleModelSet::from_models / ensure_models_are_uploaded (cuda_particle_model_set.rs:27-84). */
int spk_models_upload(spk_ctx* ctx, const SpkGpuParticleModel* models, size_t n);
/* CudaColliderSet::from_collider_set after the rapier->GpuCollider conversion (cuda_collider_set.rs:45-126). */
int spk_colliders_upload(spk_ctx* ctx, const SpkGpuCollider* colliders, size_t n);

/* CudaMpmPipeline::step (cuda_mpm_pipeline.rs:262-659): advances all contexts by params->dt with adaptive
 * substeps; writes the simulated time (params.dt - remaining) to *simulated_time_out.  The fast multi-GPU path is one process
 * per GPU with nctx == 1 (attach the contexts with spk_dist_* and call spk_step on every rank).  nctx > 1 is the reference's own
 * calling convention — one host thread driving several contexts, which may sit on different devices or on one: every substep the
 * active block keys are gathered, shared blocks tagged, dt reduced over the contexts and the shared blocks exchanged with
 * device-to-device copies and merged (cuda_mpm_pipeline.rs:329-351, 393-398, 458-532); a compatibility path, synchronous, with
 * only num_substeps / last_dt / block counts filled in the timings.
 * With timings_or_null == NULL the call returns once the frame is ENQUEUED (every substep still reads its 64-byte counter
 * record back, so only the tail of the last substep is outstanding): consecutive frames then run back to back on the GPU.
 * All other entry points are ordered on the context's stream; spk_ctx_synchronize() waits for completion. */
int spk_step(spk_ctx** ctxs, size_t nctx, const SpkSolverParameters* params, const float gravity[SPK_DIM],
             SpkStepTimings* timings_or_null, float* simulated_time_out);

/* ---- custom particle kernels: `CudaParticleKernelsLauncher` + `CudaMpmPipeline::step_generic`
 * (src/cuda/cuda_particle_kernels.rs:7-27, cuda_mpm_pipeline.rs:273-281,299-302,381-392,479-498).  A host that compiles its own
 * G2P2G kernel — `g2p2g_kernel<MODE, MyUpdater>` from sparkl_b200/csrc/spk_g2p2g_kernel.cuh, the counterpart of the reference's
 * `g2p2g_generic::<MyUpdater>` — registers its launch functions here; spk_step / spk_stage_* then call them instead of the
 * default ones.  A NULL member keeps the default for that launch.  The callbacks run on the caller's thread, must enqueue on the
 * context's stream (spk_ctx_stream) and return 0 or a cudaError_t / SPK_E* code, which the calling entry point returns.
 * `spk_default_launch_*` are the `DefaultCudaParticleKernelsLauncher`: a custom launcher may fall back to them.
 * Fused CFL estimate: the library's G2P2G kernels (the default one and every `g2p2g_kernel<.., Updater>` instantiation) also reduce the
 * NEXT substep's time-step candidate, so with the DEFAULT estimate launcher the stand-alone estimate kernel only runs after a
 * particle / model edit.  A REGISTERED launch_estimate_particle_timestep_length is called every substep, like step_generic does. */
typedef struct SpkParticleKernelsLauncher {
    void* user;                                                                   /* passed back to every callback */
    int (*timestep_started)(void* user);                                          /* once per spk_step, before the first substep */
    int (*launch_estimate_particle_timestep_length)(void* user, spk_ctx* ctx, float min_dt, float max_dt);
    int (*launch_g2p2g)(void* user, spk_ctx* ctx, float timestep_length);
} SpkParticleKernelsLauncher;
int spk_set_particle_kernels_launcher(spk_ctx* ctx, const SpkParticleKernelsLauncher* launcher_or_null); /* copied; NULL restores the default */
int spk_default_launch_estimate_particle_timestep_length(spk_ctx* ctx, float min_dt, float max_dt);
int spk_default_launch_g2p2g(spk_ctx* ctx, float timestep_length);

/* ---- stage-level calls: the pieces step_generic sequences, exposed so that (a) the parity tests can compare
 * each stage with the oracle and (b) the multi-process pipeline can interleave the halo exchange. */
/* reset of the frame's dt accumulator: timestep_length.copy_from(from_sec(params.dt)) (cuda_mpm_pipeline.rs:294-299) */
int spk_frame_begin(spk_ctx* ctx, const SpkSolverParameters* params);
/* CudaSparseGrid::launch_sort (cuda_sparse_grid.rs:147-489): K1,K2,K5-K12 of SURVEY §2.2 */
int spk_stage_sort(spk_ctx* ctx, SpkSortInfo* info_or_null);
/* reset_grid + copy_grid_projection_data (cuda_mpm_pipeline.rs:371-376): K13,K14 */
int spk_stage_reset_grid(spk_ctx* ctx);
/* estimate_timestep_length + read-back (cuda_mpm_pipeline.rs:381-398): K15. Returns into_sec(dt_acc). */
int spk_stage_estimate_dt(spk_ctx* ctx, float min_dt, float max_dt, float* dt_out);
/* raw dt accumulator (GpuTimestepLength, gpu_timestep.rs:8-26) for the multi-process min-reduction */
int spk_dt_acc_get(spk_ctx* ctx, uint64_t* out);
int spk_dt_acc_set(spk_ctx* ctx, uint64_t value);
/* g2p2g (cuda_mpm_pipeline.rs:479-498): K16, all dispatch blocks (halo or not) */
int spk_stage_g2p2g(spk_ctx* ctx, float dt, uint32_t damage_model);
/* grid_update + swap_buffers (cuda_mpm_pipeline.rs:536-558): K19 */
int spk_stage_grid_update(spk_ctx* ctx, float dt, const float gravity[SPK_DIM]);
/* Fluids-volume recomputation: recompute_fluids_volume_p2g / _g2p (src_kernels/cuda/fluids_volume.rs:9-105; the reference comments the
 * two kernels out of its build, src_kernels/cuda/mod.rs:2,18, and runs the same pass on the CPU after the sort when
 * SolverParameters::force_fluids_volume_recomputation is set, src/dynamics/solver/mls_solver.rs:46-48, fluids_volume.rs:9-116).
 * F[0][0] of every Monaghan particle <- (mass / (sum_nodes w * node_mass / h^DIM)) / volume0.  Valid between spk_stage_reset_grid and
 * spk_stage_g2p2g (SPK_ESTATE before the first sort); the next grid is left zeroed.  spk_step never calls it: the reference's CUDA
 * pipeline ignores the flag. */
int spk_stage_recompute_fluids_volume(spk_ctx* ctx);

/* ---- halo exchange for multi-process runs (replaces K3,K4,K17,K18 and the D2D copies,
 * cuda_sparse_grid.rs:256-322, cuda_mpm_pipeline.rs:407-533).  Keys travel as u64 virtual ids; payload is the
 * 6 summed floats per node (mass, momentum xyz, psi_momentum, psi_mass) = 1536 B per block.  All buffers are DEVICE
 * pointers owned by the caller; every call is asynchronous on the context's stream except spk_halo_count.  Buffers
 * have a fixed capacity agreed between ranks and are padded with the EMPTY key (u64::MAX), which tag/merge skip, so
 * they can be all-gathered without exchanging lengths first. */
int spk_halo_begin(spk_ctx* ctx);                                                     /* num_halo_blocks <- 0 */
int spk_halo_active_keys(spk_ctx* ctx, uint64_t* dev_keys_out, uint32_t capacity);    /* next grid's block keys, padded */
int spk_halo_tag(spk_ctx* ctx, const uint64_t* dev_remote_keys, uint32_t n_remote);   /* tag_halo_blocks, sort.rs:74-96 */
int spk_halo_count(spk_ctx* ctx, uint32_t* num_halo_out);                             /* blocking read of num_halo_blocks */
int spk_halo_pack(spk_ctx* ctx, uint64_t* dev_keys_out, float* dev_payload_out, uint32_t capacity); /* copy_halo_to_staging */
int spk_halo_merge(spk_ctx* ctx, const uint64_t* dev_keys, const float* dev_payload, uint32_t n_blocks); /* merge_halo_blocks */

/* ---- multi-GPU over NVLink peer memory, one process per GPU (replaces the gather of active block headers,
 * src/cuda/cuda_sparse_grid.rs:256-322; K3/K4/K17/K18, src_kernels/cuda/sort.rs:74-164; the per-pair cuMemcpyDtoDAsync
 * and the host dt minimum, src/cuda/cuda_mpm_pipeline.rs:393-398,458-472,512-532).  Every rank exports one descriptor,
 * the caller all-gathers them with whatever transport it has (MPI, torch.distributed, a file) and attaches; from then on
 * spk_step() runs the exchange itself with device-initiated NVLink loads / atomics and two device-side barriers per
 * substep.  Collective: every rank must make the same sequence of spk_step / upload calls.  A rank may own no
 * particle (it takes part in the barriers and in the dt minimum with an empty grid); the shared arena does not grow, so a context that is to
 * receive migrants later must be created with a block capacity for them. */
typedef struct SpkDistDescriptor { unsigned char opaque[128]; } SpkDistDescriptor;
int spk_dist_export(spk_ctx* ctx, SpkDistDescriptor* out);   /* after the particle upload, before the first step */
int spk_dist_attach(spk_ctx* ctx, uint32_t rank, uint32_t world, const SpkDistDescriptor* all_ranks /* [world] */);
int spk_dist_detach(spk_ctx* ctx);
int spk_dist_info(spk_ctx* ctx, uint32_t* num_halo_blocks, uint32_t* barrier_epoch);

/* ---- particle migration between slabs, device side.  The reference keeps particle ownership static
 * (SingleGpuMpmContext, src/cuda/cuda_mpm_pipeline.rs:177-194: one CudaParticleSet per device for the whole run); with spatial slabs
 * a particle that leaves its slab is handed to the rank that owns its block column.  The wire format is one row per particle: the
 * reference's six per-particle structs (src_core/dynamics/particle.rs:14-219) + a 32-bit global id — 128 bytes in dim3. */
typedef struct SpkParticleRecord {
    SpkParticleStatus status;
    SpkParticlePosition position;
    SpkParticleVelocity velocity;
    SpkParticleVolume volume;
    SpkParticlePhase phase;
    SpkParticleData data;
    uint32_t gid;
} SpkParticleRecord;
#if defined(__cplusplus) && SPK_DIM == 3
static_assert(sizeof(SpkParticleRecord) == 128, "SpkParticleRecord");
#endif
/* optional global ids, indexed by particle id (HOST arrays; n must equal the particle count).  Default: gid == particle id.  Any
 * host-side edit of the set (upload, append, remove_range) resets them to the identity. */
int spk_particles_set_gids(spk_ctx* ctx, const uint32_t* gids, size_t n);
int spk_particles_read_gids(spk_ctx* ctx, uint32_t* dst, size_t n);
/* Owner of a particle = slab r with bounds[r] <= block_x < bounds[r + 1], block_x as in GpuGrid::block_associated_to_point
 * (gpu_grid.rs:271-279); `bounds` is a HOST array of world + 1 increasing block coordinates whose outer two are open ended.  Every
 * particle owned by another rank is written to dev_records (DEVICE memory, `capacity` rows): grouped by destination rank, ascending
 * particle id inside a group; counts_out[world] (HOST) receives the rows per destination.  The remaining particles keep their relative
 * order and are renumbered densely.  dev_records == NULL only counts.  Nothing changes when nobody leaves (the sort state stays
 * valid) or when `capacity` is too small (SPK_EOVERFLOW; counts_out is still filled). */
int spk_particles_migrate_out(spk_ctx* ctx, const int64_t* bounds, uint32_t world, uint32_t rank, SpkParticleRecord* dev_records, size_t capacity,
                              uint32_t* counts_out);
/* appends n rows from DEVICE memory; their particle ids continue after the current ones, in row order (grows on the device).  n == 0
 * only marks the set as edited: in a collective migration every rank calls this, so that all ranks take the same path (stand-alone dt
 * kernel + its device-side barrier) in the next step of the peer-memory pipeline. */
int spk_particles_migrate_in(spk_ctx* ctx, const SpkParticleRecord* dev_records, size_t n);

/* ---- read-back of the sparse grid in the reference's layouts (parity tests; ≙ the commented-out
 * check_active_blocks, cuda_sparse_grid.rs:491-652).  which_grid: 0 = curr, 1 = next. */
int spk_grid_num_active_blocks(spk_ctx* ctx, int which_grid, uint32_t* out);
int spk_grid_read_headers(spk_ctx* ctx, int which_grid, SpkActiveBlockHeader* dst, uint32_t n);
int spk_grid_read_nodes(spk_ctx* ctx, int which_grid, SpkGpuGridNode* dst, uint32_t n_blocks); /* 64 (dim2: 16) nodes per block */
int spk_grid_read_dispatch(spk_ctx* ctx, SpkDispatchBlock2ActiveBlock* dst, uint32_t n);
/* The box the LAST particle sort ranked its blocks in (block ids follow the Morton curve of the block coordinate relative to this box;
 * encoding of src_core/utils/morton3.rs:3-15 generalised to one bit count per axis: for l = 0, 1, ...: bit l of x, of y, of z — each
 * only while l < its bit count — appended from the least significant end).  out[0..2]: origin (raw BlockVirtualId coordinates),
 * out[3..5]: bits per axis, out[6]: 1 when the sort had a box at all, out[7]: 1 when a block fell outside it; with out[6] == 0 or
 * out[7] == 1 that sort ranked the blocks by hash slot (valid ids, no spatial order).  dim2: entries [2] and [5] are 0. */
int spk_grid_morton_box(spk_ctx* ctx, uint32_t out[8]);
/* raw device pointers for zero-copy wrappers (torch / NCCL plumbing); valid until the next upload/append */
int spk_device_ptr(spk_ctx* ctx, int which /* SPK_ATTR_* */, void** out);
/* Kernel timing without perturbing the run (≙ the g2p2g field of CudaDeviceTimings, cuda_mpm_pipeline.rs:56-85, which the
 * reference fills with blocking event reads): with `enable` != 0 every following g2p2g launch is bracketed by a pair of CUDA
 * events recorded on the context's stream WITHOUT synchronising (up to 256 launches); each call returns the summed duration
 * and the number of launches recorded since the previous call, and resets the ring. */
int spk_g2p2g_time(spk_ctx* ctx, int enable, float* total_ms_out, uint32_t* launches_out);

#ifdef __cplusplus
}
#endif
#endif /* SPARKL_ABI_H */
