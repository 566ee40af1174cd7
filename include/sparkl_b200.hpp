// sparkl_b200.hpp — C++17 host-side mirror of sparkl's CUDA pipeline API over the C ABI of sparkl_abi.h (header-only).
//
// sparkl's host code is Rust; this environment has no Rust toolchain, so the compiled-language host layer above the C ABI is this
// header (the Rust shim a sparkl maintainer would use instead is rust_shim/src/ffi.rs).  Names, argument meaning and error behaviour
// follow src/cuda/*.rs of the reference:
//   CudaParticleSet        cuda_particle_set.rs:12-177          CudaParticleModelSet   cuda_particle_model_set.rs:8-84
//   CudaColliderSet        cuda_collider_set.rs:13-131          CudaSparseGrid         cuda_sparse_grid.rs:26-144
//   SingleGpuMpmContext    cuda_mpm_pipeline.rs:177-194         CudaMpmPipeline::step  cuda_mpm_pipeline.rs:262-271
//   CudaParticleKernelsLauncher  cuda_particle_kernels.rs:7-27  model constructors     src_core/dynamics/models/*.rs
// `CudaResult::Err` and the reference's assert panics both surface as sparkl::CudaError (code > 0: cudaError_t, code < 0: SPK_E*).
// The reference keeps particles / models / colliders / grid as separate objects borrowed by a SingleGpuMpmContext; here one spk_ctx
// owns all device memory of a context and the classes below are views on it.  Works for both builds (SPK_DIM 3 and 2).
#pragma once
#include <cmath>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "sparkl_abi.h"

namespace sparkl {

struct CudaError : std::runtime_error {
    int code;
    CudaError(int c, const std::string& what) : std::runtime_error("sparkl_b200 error " + std::to_string(c) + ": " + what + " (" + spk_error_string(c) + ")"), code(c) {}
};
inline void check(int rc, const char* what) { if (rc != SPK_OK) throw CudaError(rc, what); }

// ---- material models (src_core/dynamics/models/*.rs; all arithmetic in f32 like the Rust code) --------------------------------
inline std::pair<float, float> lame_lambda_mu(float young_modulus, float poisson_ratio) { // utils/physics.rs:16-27
    const float lam = young_modulus * poisson_ratio / ((1.0f + poisson_ratio) * (1.0f - 2.0f * poisson_ratio));
    return {lam, young_modulus / (2.0f * (1.0f + poisson_ratio))};
}
struct ParticleModel { // ParticleModel::new / with_plasticity / with_failure (src/dynamics/particle_model.rs:20-82) -> GpuParticleModel
    SpkGpuParticleModel gpu;
    ParticleModel() { std::memset(&gpu, 0, sizeof(gpu)); gpu.constitutive_tag = SPK_CONSTITUTIVE_CUSTOM; gpu.plastic_tag = SPK_PLASTIC_NONE; gpu.failure_tag = SPK_FAILURE_NONE; }
    static ParticleModel corotated_linear_elasticity(float young_modulus, float poisson_ratio) { // elasticity_corotated_linear.rs:19-29
        ParticleModel m;
        const auto lm = lame_lambda_mu(young_modulus, poisson_ratio);
        m.gpu.constitutive_tag = SPK_CONSTITUTIVE_COROTATED;
        m.gpu.constitutive.corotated.m = SpkCorotatedLinearElasticity{1, {0, 0, 0}, 0.9f, lm.first, lm.second};
        return m;
    }
    static ParticleModel neo_hookean_elasticity(float young_modulus, float poisson_ratio) { // elasticity_neo_hookean.rs:17-28
        ParticleModel m;
        const auto lm = lame_lambda_mu(young_modulus, poisson_ratio);
        m.gpu.constitutive_tag = SPK_CONSTITUTIVE_NEO_HOOKEAN;
        m.gpu.constitutive.neo_hookean.m = SpkNeoHookeanElasticity{0.5f, lm.first, lm.second};
        return m;
    }
    static ParticleModel monaghan_sph_eos(float pressure0, uint32_t gamma, float viscosity) { // eos_monaghan_sph.rs:18-26
        ParticleModel m;
        m.gpu.constitutive_tag = SPK_CONSTITUTIVE_MONAGHAN;
        m.gpu.constitutive.monaghan.m = SpkMonaghanSphEos{pressure0, gamma, viscosity, 1.0f};
        return m;
    }
    ParticleModel& with_drucker_prager_plasticity(float young_modulus, float poisson_ratio) { // plasticity_drucker_prager.rs:21-34
        const auto lm = lame_lambda_mu(young_modulus, poisson_ratio);
        const float rad = 3.14159265358979323846f / 180.0f;
        gpu.plastic_tag = SPK_PLASTIC_DRUCKER_PRAGER;
        gpu.plastic.drucker_prager.m = SpkDruckerPragerPlasticity{35.0f * rad, 9.0f * rad, 0.2f, 10.0f * rad, lm.first, lm.second, 0, {0, 0, 0}, 1.0f};
        return *this;
    }
    ParticleModel& with_rankine_plasticity(float young_modulus, float poisson_ratio, float tensile_strength, float softening_rate) { // plasticity_rankine.rs:18-33
        const auto lm = lame_lambda_mu(young_modulus, poisson_ratio);
        gpu.plastic_tag = SPK_PLASTIC_RANKINE;
        gpu.plastic.rankine.m = SpkRankinePlasticity{lm.second, lm.first, tensile_strength, softening_rate};
        return *this;
    }
    ParticleModel& with_snow_plasticity() { // plasticity_snow.rs:17-24
        gpu.plastic_tag = SPK_PLASTIC_SNOW;
        gpu.plastic.snow.m = SpkSnowPlasticity{2.5e-2f, 4.5e-3f, 10.0f};
        return *this;
    }
    ParticleModel& with_maximum_stress_failure(float max_principal_stress, float max_shear_stress) { // failure_maximum_stress.rs:12-18
        gpu.failure_tag = SPK_FAILURE_MAXIMUM_STRESS;
        gpu.failure.maximum_stress = SpkMaximumStressFailure{max_principal_stress, max_shear_stress};
        return *this;
    }
};

// ---- particles: `ParticleComponents` of cuda_particle_set.rs:180-187 -----------------------------------------------------------------
struct ParticleComponents {
    std::vector<SpkParticleStatus> status;
    std::vector<SpkParticlePosition> position;
    std::vector<SpkParticleVelocity> velocity;
    std::vector<SpkParticleVolume> volume;
    std::vector<SpkParticlePhase> phase;
    std::vector<SpkParticleData> data;
    size_t len() const { return position.size(); }
    // Particle::new (src/dynamics/particle.rs:53-104) + extract_particles_components (cuda_particle_set.rs:189-248)
    void push(size_t model_index, const float* point, float radius, float density0) {
        SpkParticleStatus st; std::memset(&st, 0, sizeof(st)); st.model_index = model_index;
        SpkParticlePosition p; SpkParticleVelocity v; std::memset(&v, 0, sizeof(v));
        for (int a = 0; a < SPK_DIM; ++a) p.point[a] = point[a];
        SpkParticleVolume vo; std::memset(&vo, 0, sizeof(vo));
        float volume0 = 1.0f;
        for (int a = 0; a < SPK_DIM; ++a) volume0 *= radius * 2.0f; // (radius * 2).powi(DIM)
        vo.mass = volume0 * density0; vo.volume0 = volume0; vo.radius0 = radius; vo.plastic_deformation_gradient_det = 1.0f;
        for (int a = 0; a < SPK_DIM; ++a) vo.deformation_gradient[a * SPK_DIM + a] = 1.0f;
        status.push_back(st); position.push_back(p); velocity.push_back(v); volume.push_back(vo);
        phase.push_back(SpkParticlePhase{1.0f, 0.0f});
        data.push_back(SpkParticleData{-0.01f, 1.0f, 1.0f, 0.0f}); // ParticleData::default, particle.rs:211-219
    }
};

class SingleGpuMpmContext;

class CudaParticleSet { // cuda_particle_set.rs:12-177
  public:
    explicit CudaParticleSet(spk_ctx* c) : c_(c) {}
    void from_particles(const ParticleComponents& pc) {
        const size_t n = pc.len();
        if (pc.status.size() != n || pc.velocity.size() != n || pc.volume.size() != n || pc.phase.size() != n || (!pc.data.empty() && pc.data.size() != n))
            throw CudaError(SPK_EINVAL, "All attribute buffer must have the same length."); // the reference's assert, cuda_particle_set.rs:40-59
        check(spk_particles_upload(c_, pc.status.data(), pc.position.data(), pc.velocity.data(), pc.volume.data(), pc.phase.data(),
                                   pc.data.empty() ? nullptr : pc.data.data(), n), "CudaParticleSet::from_particles");
    }
    void append(const ParticleComponents& pc) {
        check(spk_particles_append(c_, pc.status.data(), pc.position.data(), pc.velocity.data(), pc.volume.data(), pc.phase.data(),
                                   pc.data.empty() ? nullptr : pc.data.data(), pc.len()), "CudaParticleSet::append");
    }
    void remove_range(size_t begin, size_t end) { check(spk_particles_remove_range(c_, begin, end), "CudaParticleSet::remove_range"); }
    size_t len() const { size_t n = 0; check(spk_particles_len(c_, &n), "CudaParticleSet::len"); return n; }
    size_t capacity() const { size_t n = 0; check(spk_particles_capacity(c_, &n), "CudaParticleSet::capacity"); return n; }
    void truncate(size_t new_len) { check(spk_particles_truncate(c_, new_len), "CudaParticleSet::truncate"); }
    bool is_empty() const { return len() == 0; }
    std::vector<SpkParticlePosition> read_positions() const { return read<SpkParticlePosition>(SPK_ATTR_POSITION); }
    std::vector<SpkParticleVelocity> read_velocities() const { return read<SpkParticleVelocity>(SPK_ATTR_VELOCITY); }
    std::vector<SpkParticleStatus> read_statuses() const { return read<SpkParticleStatus>(SPK_ATTR_STATUS); }
    std::vector<SpkParticleVolume> read_volumes() const { return read<SpkParticleVolume>(SPK_ATTR_VOLUME); }
    std::vector<SpkParticlePhase> read_phase() const { return read<SpkParticlePhase>(SPK_ATTR_PHASE); }
    std::vector<uint32_t> read_sorted_particle_ids() const { return read<uint32_t>(SPK_ATTR_SORTED_IDS); }

  private:
    template <typename T>
    std::vector<T> read(int which) const {
        std::vector<T> out(len());
        if (!out.empty()) check(spk_particles_read(c_, which, out.data(), out.size()), "CudaParticleSet::read_*");
        return out;
    }
    spk_ctx* c_;
};

class CudaParticleModelSet { // cuda_particle_model_set.rs:8-84
  public:
    explicit CudaParticleModelSet(spk_ctx* c) : c_(c) {}
    void from_models(const std::vector<ParticleModel>& models) {
        models_.clear();
        for (const auto& m : models) models_.push_back(m.gpu);
        ensure_models_are_uploaded();
    }
    void ensure_models_are_uploaded() { check(spk_models_upload(c_, models_.data(), models_.size()), "CudaParticleModelSet::ensure_models_are_uploaded"); }
    std::vector<SpkGpuParticleModel>& models() { return models_; }

  private:
    spk_ctx* c_;
    std::vector<SpkGpuParticleModel> models_;
};

class CudaColliderSet { // cuda_collider_set.rs:13-131
  public:
    explicit CudaColliderSet(spk_ctx* c) : c_(c) {}
    void from_collider_set(const std::vector<SpkGpuCollider>& colliders) { check(spk_colliders_upload(c_, colliders.data(), colliders.size()), "CudaColliderSet::from_collider_set"); }
    // a rapier cuboid collider + CudaColliderOptions (cuda_collider_set.rs:22-39,63-70)
    static SpkGpuCollider cuboid(const float* half_extents, const float* translation, float friction = 0.5f, uint32_t handling = SPK_BOUNDARY_FRICTION) {
        SpkGpuCollider c; std::memset(&c, 0, sizeof(c));
        c.shape_tag = SPK_SHAPE_CUBOID;
        for (int a = 0; a < SPK_DIM; ++a) { c.shape.cuboid.half_extents[a] = half_extents[a]; c.translation[a] = translation[a]; }
#if SPK_DIM == 3
        c.rotation[3] = 1.0f; // identity quaternion (i, j, k, w)
#else
        c.rotation[0] = 1.0f; // identity unit complex (re, im)
#endif
        c.friction = friction; c.grid_boundary_handling = handling;
        return c;
    }
    // the other shapes of GpuColliderShape (src_kernels/gpu_collider.rs:10-96).  The payload is a HOST description: the arrays must stay
    // alive until from_collider_set() returns (the library keeps its own device copy).
    // rapier `ColliderBuilder::trimesh(vertices, indices)` (no ORIENTED flag): num_vertices x SPK_DIM coordinates, num_triangles x 3 indices
    static SpkGpuCollider trimesh(const float* vertices, uint32_t num_vertices, const uint32_t* indices, uint32_t num_triangles, const float* translation,
                                  float friction = 0.5f, uint32_t handling = SPK_BOUNDARY_FRICTION, bool flip_interior = false) {
        const float zero[3] = {0.f, 0.f, 0.f};
        SpkGpuCollider c = cuboid(zero, translation, friction, handling);
        c.shape_tag = SPK_SHAPE_TRIMESH;
        c.shape.trimesh.vertices = vertices; c.shape.trimesh.indices = indices;
        c.shape.trimesh.num_vertices = num_vertices; c.shape.trimesh.num_triangles = num_triangles;
        c.shape.trimesh.flip_interior = flip_interior;
        return c;
    }
#if SPK_DIM == 3
    // rapier `ColliderBuilder::heightfield(heights, scale)` (examples3d/sand3.rs:23-38): column-major nrows x ncols samples
    static SpkGpuCollider heightfield(const float* heights, uint32_t nrows, uint32_t ncols, const float* scale, const float* translation, float friction = 0.5f,
                                      uint32_t handling = SPK_BOUNDARY_FRICTION, bool flip_interior = false) {
        const float zero[3] = {0.f, 0.f, 0.f};
        SpkGpuCollider c = cuboid(zero, translation, friction, handling);
        c.shape_tag = SPK_SHAPE_HEIGHTFIELD;
        c.shape.heightfield.heights = heights; c.shape.heightfield.nrows = nrows; c.shape.heightfield.ncols = ncols;
        for (int a = 0; a < 3; ++a) c.shape.heightfield.scale[a] = scale[a];
        c.shape.heightfield.bellow_heightfield_is_solid = 1; // cuda_collider_set.rs:77
        c.shape.heightfield.flip_interior = flip_interior;
        return c;
    }
#else
    // rapier2d `ColliderBuilder::heightfield(heights, scale)` (examples2d/basic2.rs:32-41)
    static SpkGpuCollider heightfield(const float* heights, uint32_t num_heights, const float* scale, const float* translation, float friction = 0.5f,
                                      uint32_t handling = SPK_BOUNDARY_FRICTION, bool flip_interior = false) {
        const float zero[2] = {0.f, 0.f};
        SpkGpuCollider c = cuboid(zero, translation, friction, handling);
        c.shape_tag = SPK_SHAPE_HEIGHTFIELD;
        c.shape.heightfield.heights = heights; c.shape.heightfield.num_heights = num_heights;
        for (int a = 0; a < 2; ++a) c.shape.heightfield.scale[a] = scale[a];
        c.shape.heightfield.bellow_heightfield_is_solid = 1;
        c.shape.heightfield.flip_interior = flip_interior;
        return c;
    }
#endif
    // GpuColliderShape::Polyline: a closed polygon of num_vertices x SPK_DIM coordinates (gpu_collider.rs:24-27,105-201)
    static SpkGpuCollider polyline(const float* vertices, uint32_t num_vertices, const float* translation, float friction = 0.5f,
                                   uint32_t handling = SPK_BOUNDARY_FRICTION, bool flip_interior = false) {
        const float zero[3] = {0.f, 0.f, 0.f};
        SpkGpuCollider c = cuboid(zero, translation, friction, handling);
        c.shape_tag = SPK_SHAPE_POLYLINE;
        c.shape.polyline.vertices = vertices; c.shape.polyline.num_vertices = num_vertices;
        c.shape.polyline.flip_interior = flip_interior;
        return c;
    }

  private:
    spk_ctx* c_;
};

class CudaSparseGrid { // cuda_sparse_grid.rs:26-144
  public:
    explicit CudaSparseGrid(spk_ctx* c) : c_(c) {}
    float cell_width() const { float h = 0; check(spk_ctx_cell_width(c_, &h), "CudaSparseGrid::cell_width"); return h; }
    uint32_t num_active_blocks() const { uint32_t n = 0; check(spk_grid_num_active_blocks(c_, 0, &n), "CudaSparseGrid::num_active_blocks"); return n; }

  private:
    spk_ctx* c_;
};

// CudaParticleKernelsLauncher (cuda_particle_kernels.rs:7-27): override to launch custom kernels (see examples/custom_updater)
class CudaParticleKernelsLauncher {
  public:
    virtual ~CudaParticleKernelsLauncher() = default;
    virtual int timestep_started() { return SPK_OK; }
    virtual int launch_estimate_particle_timestep_length(spk_ctx* ctx, float min_dt, float max_dt) { return spk_default_launch_estimate_particle_timestep_length(ctx, min_dt, max_dt); }
    virtual int launch_g2p2g(spk_ctx* ctx, float timestep_length) { return spk_default_launch_g2p2g(ctx, timestep_length); }
};

class SingleGpuMpmContext { // cuda_mpm_pipeline.rs:177-194
  public:
    explicit SingleGpuMpmContext(float cell_width, int device = 0, uint32_t block_capacity = 1024) {
        check(spk_ctx_create(device, cell_width, block_capacity, &c_), "SingleGpuMpmContext::new");
    }
    ~SingleGpuMpmContext() { if (c_) spk_ctx_destroy(c_); }
    SingleGpuMpmContext(const SingleGpuMpmContext&) = delete;
    SingleGpuMpmContext& operator=(const SingleGpuMpmContext&) = delete;
    SingleGpuMpmContext(SingleGpuMpmContext&& o) noexcept : c_(o.c_), launcher_(std::move(o.launcher_)) { o.c_ = nullptr; }
    CudaParticleSet particles() { return CudaParticleSet(c_); }
    CudaParticleModelSet models() { return CudaParticleModelSet(c_); }
    CudaColliderSet colliders() { return CudaColliderSet(c_); }
    CudaSparseGrid grid() { return CudaSparseGrid(c_); }
    void synchronize() { check(spk_ctx_synchronize(c_), "SingleGpuMpmContext::synchronize"); }
    // step_generic(.., custom_launchers): the launcher stays registered (and owned by the context) until replaced
    void set_particle_kernels_launcher(std::unique_ptr<CudaParticleKernelsLauncher> l) {
        launcher_ = std::move(l);
        if (!launcher_) { check(spk_set_particle_kernels_launcher(c_, nullptr), "set_particle_kernels_launcher"); return; }
        SpkParticleKernelsLauncher vt;
        vt.user = launcher_.get();
        vt.timestep_started = [](void* u) { return static_cast<CudaParticleKernelsLauncher*>(u)->timestep_started(); };
        vt.launch_estimate_particle_timestep_length = [](void* u, spk_ctx* c, float a, float b) { return static_cast<CudaParticleKernelsLauncher*>(u)->launch_estimate_particle_timestep_length(c, a, b); };
        vt.launch_g2p2g = [](void* u, spk_ctx* c, float dt) { return static_cast<CudaParticleKernelsLauncher*>(u)->launch_g2p2g(c, dt); };
        check(spk_set_particle_kernels_launcher(c_, &vt), "set_particle_kernels_launcher");
    }
    // MlsSolver::recompute_fluids_volumes (src/dynamics/solver/fluids_volume.rs:9-116) as an explicit stage: between the sort / grid
    // reset and the transfer of a hand-sequenced substep (spk_stage_* calls on raw())
    void recompute_fluids_volume() { check(spk_stage_recompute_fluids_volume(c_), "SingleGpuMpmContext::recompute_fluids_volume"); }
    // device-side particle migration between slabs (the reference keeps ownership static): see include/sparkl_abi.h
    std::vector<uint32_t> migrate_out(const std::vector<int64_t>& bounds, uint32_t rank, SpkParticleRecord* dev_records, size_t capacity) {
        std::vector<uint32_t> counts(bounds.size() > 1 ? bounds.size() - 1 : 0);
        check(spk_particles_migrate_out(c_, bounds.data(), (uint32_t)counts.size(), rank, dev_records, capacity, counts.data()), "SingleGpuMpmContext::migrate_out");
        return counts;
    }
    void migrate_in(const SpkParticleRecord* dev_records, size_t n) { check(spk_particles_migrate_in(c_, dev_records, n), "SingleGpuMpmContext::migrate_in"); }
    void set_gids(const std::vector<uint32_t>& gids) { check(spk_particles_set_gids(c_, gids.data(), gids.size()), "SingleGpuMpmContext::set_gids"); }
    std::vector<uint32_t> read_gids() {
        std::vector<uint32_t> out(particles().len());
        check(spk_particles_read_gids(c_, out.data(), out.size()), "SingleGpuMpmContext::read_gids");
        return out;
    }
    spk_ctx* raw() { return c_; }

  private:
    spk_ctx* c_ = nullptr;
    std::unique_ptr<CudaParticleKernelsLauncher> launcher_;
};

class CudaMpmPipeline { // cuda_mpm_pipeline.rs:144-271
  public:
    uint64_t step_id = 0;
    // Same contract as CudaMpmPipeline::step: advances the contexts by params.dt with adaptive substeps, returns the simulated time.
    float step(SpkStepTimings* timings, const SpkSolverParameters& params, const float* gravity, std::vector<SingleGpuMpmContext*>& contexts) {
        std::vector<spk_ctx*> raw;
        for (auto* c : contexts) raw.push_back(c->raw());
        float simulated = 0.0f;
        check(spk_step(raw.data(), raw.size(), &params, gravity, timings, &simulated), "CudaMpmPipeline::step");
        if (timings) step_id += timings->num_substeps;
        return simulated;
    }
};

inline SpkSolverParameters default_solver_parameters() { // SolverParameters::default, solver_parameters.rs:53-68
    SpkSolverParameters p; std::memset(&p, 0, sizeof(p));
    p.dt = 1.0f / 60.0f; p.max_substep_dt = 3.402823466e38f; p.max_num_substeps = 1000;
    p.boundary_handling = SPK_BOUNDARY_FRICTION; p.damage_model = SPK_DAMAGE_NONE;
    return p;
}

} // namespace sparkl
