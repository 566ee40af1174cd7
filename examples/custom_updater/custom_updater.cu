// custom_updater.cu — example of the custom particle-kernel extension point.
//
// What a sparkl user does with `ParticleUpdater` + `g2p2g_generic::<MyUpdater>` + `CudaParticleKernelsLauncher`
// (src_kernels/cuda/particle_updater.rs:11-38, src/cuda/cuda_particle_kernels.rs:7-27), done against libsparkl_b200:
//   1. define an updater type (two static device functions, see sparkl_b200/csrc/spk_updater.cuh);
//   2. instantiate the library's own G2P2G / dt kernels with it (spk_launch_g2p2g_with<U>, spk_launch_estimate_dt_with<U>);
//   3. hand the launch functions to the pipeline (spk_set_particle_kernels_launcher).
// `TableUpdater` gives a meaning to `GpuConstitutiveModel::Custom(id)` (zero stress in the default updater): particles whose model
// is Custom(id) are updated with entry `id` of a model table the host uploads, and every particle's velocity is damped by
// exp(-damping * dt).  With damping == 0 and a table entry equal to a built-in model the result is bit-identical to running that
// built-in model, which is what tests/test_custom_updater_gpu.py checks.
//
// Build (python -m sparkl_b200.build does it): nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -I<repo>/include
#include "../../sparkl_b200/csrc/spk_g2p2g_kernel.cuh"

struct TableUpdaterData { // device-resident `user` block
    const DevModel* table;
    uint32_t table_len;
    float damping;
};

struct TableUpdater {
    __device__ __forceinline__ static const DevModel& resolve(const DevModel& model, const TableUpdaterData* u) {
        if (model.ctag != SPK_CONSTITUTIVE_CUSTOM) return model;
        const uint32_t id = __float_as_uint(model.c[0]); // GpuConstitutiveModel::Custom(id)
        return id < u->table_len ? u->table[id] : model;
    }
    __device__ __forceinline__ static bool update(const DevModel& model, const float4* __restrict__ cold, float h, float dt, ParticleState& p, float (&vel)[3], Mat3& G,
                                                  Mat3& stress, const void* user) {
        const TableUpdaterData* u = static_cast<const TableUpdaterData*>(user);
        const bool p2g = DefaultParticleUpdater::update(resolve(model, u), cold, h, dt, p, vel, G, stress, nullptr);
        if (u->damping != 0.f) {
            const float k = expf(-u->damping * dt);
            vel[0] *= k; vel[1] *= k; vel[2] *= k;
        }
        return p2g;
    }
    __device__ __forceinline__ static float dt_candidate(const DevModel& model, float h, const ParticleState& p, const float (&vel)[3], const void* user) {
        return DefaultParticleUpdater::dt_candidate(resolve(model, static_cast<const TableUpdaterData*>(user)), h, p, vel, nullptr);
    }
};

struct CuLauncher {
    SpkParticleKernelsLauncher vtable;
    TableUpdaterData* d_data;
    DevModel* d_table;
    int calls[3]; // timestep_started, estimate, g2p2g
};

static int cb_started(void* user) { static_cast<CuLauncher*>(user)->calls[0] += 1; return 0; }
static int cb_estimate(void* user, spk_ctx* ctx, float min_dt, float max_dt) {
    CuLauncher* l = static_cast<CuLauncher*>(user);
    l->calls[1] += 1;
    return (int)spk_launch_estimate_dt_with<TableUpdater>(ctx, min_dt, max_dt, l->d_data);
}
static int cb_g2p2g(void* user, spk_ctx* ctx, float dt) {
    CuLauncher* l = static_cast<CuLauncher*>(user);
    l->calls[2] += 1;
    return (int)spk_launch_g2p2g_with<TableUpdater>(ctx, dt, l->d_data);
}

extern "C" {
// `table`: host array of `n` DevModel images (the library's compact model block, spk_internal.cuh); returns a handle
int cu_launcher_create(const void* table, uint32_t n, float damping, void** out) {
    CuLauncher* l = new CuLauncher();
    l->vtable = SpkParticleKernelsLauncher{l, cb_started, cb_estimate, cb_g2p2g};
    l->calls[0] = l->calls[1] = l->calls[2] = 0;
    cudaError_t e = cudaMalloc((void**)&l->d_table, (n ? n : 1) * sizeof(DevModel));
    if (e == cudaSuccess && n) e = cudaMemcpy(l->d_table, table, n * sizeof(DevModel), cudaMemcpyHostToDevice);
    TableUpdaterData h{l->d_table, n, damping};
    if (e == cudaSuccess) e = cudaMalloc((void**)&l->d_data, sizeof(h));
    if (e == cudaSuccess) e = cudaMemcpy(l->d_data, &h, sizeof(h), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { delete l; return (int)e; }
    *out = l;
    return 0;
}
const SpkParticleKernelsLauncher* cu_launcher_vtable(void* h) { return &static_cast<CuLauncher*>(h)->vtable; }
int cu_launcher_calls(void* h, int which) { return static_cast<CuLauncher*>(h)->calls[which]; }
void cu_launcher_destroy(void* h) {
    CuLauncher* l = static_cast<CuLauncher*>(h);
    cudaFree(l->d_data); cudaFree(l->d_table);
    delete l;
}
int cu_sizeof_devmodel(void) { return (int)sizeof(DevModel); }
}
