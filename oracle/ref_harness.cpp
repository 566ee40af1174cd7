// ref_harness.cpp — ORACLE-A: drives the REFERENCE'S OWN GPU KERNELS.  TEST INFRASTRUCTURE ONLY (same rule as the CPU
// oracle: only tests/, smoke() and bench.py's reference legs may load it).
//
// The reference ships its device code as PTX (resources/sparkl3d-kernels.ptx, generated from src_kernels/ by
// rustc_codegen_nvvm).  oracle/Makefile assembles that PTX for sm_100a with the register cap the reference's host
// code requests (MaxRegisters(130), src/cuda/cuda_mpm_pipeline.rs:21,255) into oracle/_ref/sparkl3d-kernels.cubin.
// This file restates ONLY the reference's Rust host sequencing (which cannot be compiled here: no cargo/rustc) and
// launches the unmodified kernels through the CUDA driver API with the reference's grid/block shapes and by-value
// parameter blocks:
//   CudaSparseGrid::with_capacity / launch_sort      src/cuda/cuda_sparse_grid.rs:53-90,147-489
//   PrefixSumWorkspace::{reserve,launch,read_max}    src/utils/prefix_sum.rs:38-151
//   CudaMpmPipeline::step_generic                    src/cuda/cuda_mpm_pipeline.rs:273-659
//   DefaultCudaParticleKernelsLauncher               src/cuda/cuda_particle_kernels.rs:29-99
// It exports the same stage-level C API as the product (prefix `ref_`), so sparkl_b200.engine.Engine drives it.  Colliders: Cuboid
// images pass through unchanged; HeightField / Polyline host descriptions are converted to the reference's own device images
// (parry's CudaHeightFieldPtr / CudaArrayPointer1 layouts recovered from the PTX, see ref_colliders_upload).
// libcuda is dlopen'ed so the file builds on a GPU-less host.
#include <cuda.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../include/sparkl_abi.h"

// The same file builds the dim2 harness (-DSPK_DIM=2 -> oracle/_ref/libref_harness2d.so driving resources/sparkl2d-kernels.ptx):
// the kernels take the same parameter blocks in both dimensions (checked against the .entry declarations of both PTX files);
// what changes is the block shape (4x4 threads, 16 nodes of 56 bytes per block), G2P2G_THREADS (64, src_kernels/cuda/mod.rs:12-13),
// the particle struct sizes, the 8-byte gravity vector and the 280-byte GpuCollider image (see ref_colliders_upload).
#if SPK_DIM == 2
#define REF_CELLS 16u
#define REF_BLOCK_Z 1u
#define REF_G2P2G_THREADS 64u
#else
#define REF_CELLS 64u
#define REF_BLOCK_Z 4u
#define REF_G2P2G_THREADS 128u
#endif
#define REF_NODE_BYTES sizeof(SpkGpuGridNode)

namespace {

struct Driver {
    void* lib = nullptr;
#define DRV(name) decltype(&name) p_##name = nullptr;
    DRV(cuInit) DRV(cuDeviceGet) DRV(cuDevicePrimaryCtxRetain) DRV(cuCtxSetCurrent) DRV(cuModuleLoadData) DRV(cuModuleGetFunction)
    DRV(cuMemAlloc_v2) DRV(cuMemFree_v2) DRV(cuMemcpyHtoD_v2) DRV(cuMemcpyDtoH_v2) DRV(cuMemcpyDtoD_v2) DRV(cuMemsetD8_v2) DRV(cuLaunchKernel)
    DRV(cuCtxSynchronize) DRV(cuEventCreate) DRV(cuEventRecord) DRV(cuEventSynchronize) DRV(cuEventElapsedTime) DRV(cuModuleUnload)
#undef DRV
    bool load() {
        if (lib) return true;
        lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
#define GET(name) p_##name = (decltype(p_##name))dlsym(lib, #name); if (!p_##name) return false;
        GET(cuInit) GET(cuDeviceGet) GET(cuDevicePrimaryCtxRetain) GET(cuCtxSetCurrent) GET(cuModuleLoadData) GET(cuModuleGetFunction)
        GET(cuMemAlloc_v2) GET(cuMemFree_v2) GET(cuMemcpyHtoD_v2) GET(cuMemcpyDtoH_v2) GET(cuMemcpyDtoD_v2) GET(cuMemsetD8_v2) GET(cuLaunchKernel)
        GET(cuCtxSynchronize) GET(cuEventCreate) GET(cuEventRecord) GET(cuEventSynchronize) GET(cuEventElapsedTime) GET(cuModuleUnload)
#undef GET
        return true;
    }
} D;

#define CU(expr)                                 \
    do {                                         \
        CUresult _r = (expr);                    \
        if (_r != CUDA_SUCCESS) return (int)_r;  \
    } while (0)

struct GpuGridBlob { // GpuGrid, src_kernels/gpu_grid.rs:183-196 (offsets from the PTX ld.param sequence)
    float cell_width; uint32_t _p0;
    uint64_t buffer, active_blocks, num_active_blocks, entries;
    uint32_t capacity; uint32_t _p1;
    uint64_t dispatch, dispatch_halo, capacity_nodes;
};
static_assert(sizeof(GpuGridBlob) == 72, "GpuGrid");
struct HashMapBlob { uint64_t entries; uint32_t capacity; uint32_t _p; };

struct DevBuf {
    CUdeviceptr p = 0; size_t bytes = 0;
    int ensure(size_t n, bool zero) {
        if (n <= bytes && p) return 0;
        if (p) D.p_cuMemFree_v2(p);
        p = 0; bytes = 0;
        n = std::max<size_t>(n, 16);
        CU(D.p_cuMemAlloc_v2(&p, n));
        bytes = n;
        if (zero) CU(D.p_cuMemsetD8_v2(p, 0, n));
        return 0;
    }
    void release() { if (p) D.p_cuMemFree_v2(p); p = 0; bytes = 0; }
};

struct GridData { // CudaSparseGridData, cuda_sparse_grid.rs:19-24
    DevBuf buffer, active_blocks, hashmap;
    uint32_t block_capacity = 0;
    uint32_t nab = 0; // host copy
};

struct ScanStage { uint32_t len, capacity; DevBuf buf; };

} // namespace

struct ref_ctx {
    CUcontext ctx = nullptr;
    CUmodule mod = nullptr;
    float cell_width = 0;
    size_t n = 0;
    DevBuf status, pos, vel, volume, phase, data, sorted_ids;
    std::vector<SpkGpuParticleModel> models;
    DevBuf d_models, d_colliders;
    std::vector<DevBuf> collider_bufs; // device arrays the HeightField / Polyline images point to
    size_t ncolliders = 0;
    GridData grid[2];
    int curr = 0;
    DevBuf dispatch, dispatch_halo, num_active_blocks, scan_values, halo_scan_values, timestep_length;
    std::vector<ScanStage> stages; size_t num_stages = 0;
    uint32_t num_dispatch_blocks = 0;
    CUfunction k_reset_hashmap, k_touch, k_count, k_copy_len, k_scan, k_add, k_copy_first, k_finalize, k_mult, k_dispatch, k_reset_grid, k_copy_proj,
        k_dt, k_g2p2g, k_grid_update;
    CUevent ev[8];
    GridData& cur() { return grid[curr]; }
    GridData& nxt() { return grid[1 - curr]; }
    GpuGridBlob blob(GridData& g) {
        GpuGridBlob b;
        std::memset(&b, 0, sizeof(b));
        b.cell_width = cell_width; b.buffer = g.buffer.p; b.active_blocks = g.active_blocks.p; b.num_active_blocks = num_active_blocks.p;
        b.entries = g.hashmap.p; b.capacity = g.block_capacity; b.dispatch = dispatch.p; b.dispatch_halo = dispatch_halo.p;
        b.capacity_nodes = (uint64_t)g.block_capacity * REF_CELLS; // cuda_sparse_grid.rs:116,133
        return b;
    }
};

namespace {

int launch(ref_ctx* c, CUfunction f, unsigned gx, unsigned bx, unsigned by, unsigned bz, void** args) {
    (void)c;
    CU(D.p_cuLaunchKernel(f, gx, 1, 1, bx, by, bz, 0, nullptr, args, nullptr));
    return 0;
}
int reset_hashmap(ref_ctx* c, GridData& g) {
    HashMapBlob hm{g.hashmap.p, g.block_capacity, 0};
    void* args[] = {&hm};
    return launch(c, c->k_reset_hashmap, g.block_capacity / 1024 + 1, 1024, 1, 1, args); // cuda_sparse_grid.rs:182-194
}
// PrefixSumWorkspace::reserve, src/utils/prefix_sum.rs:38-78
int scan_reserve(ref_ctx* c, uint32_t buffer_len) {
    uint32_t stage_len = buffer_len / 512 + ((buffer_len % 512) != 0);
    if (c->stages.empty() || c->stages[0].capacity < stage_len) {
        for (auto& s : c->stages) s.buf.release();
        c->stages.clear();
        while (stage_len != 1) {
            ScanStage s; s.len = s.capacity = stage_len;
            if (int r = s.buf.ensure((size_t)stage_len * 4, true)) return r;
            c->stages.push_back(s);
            stage_len = stage_len / 512 + ((stage_len % 512) != 0);
            if (stage_len == 0) break;
        }
        ScanStage s; s.len = s.capacity = 1;
        if (int r = s.buf.ensure(4, true)) return r;
        c->stages.push_back(s);
        c->num_stages = c->stages.size();
    } else if (c->stages[0].len != stage_len) {
        c->num_stages = 0;
        while (stage_len != 1) {
            c->stages[c->num_stages].len = stage_len;
            c->num_stages += 1;
            stage_len = stage_len / 512 + ((stage_len % 512) != 0);
        }
        c->stages[c->num_stages].len = 1;
        c->num_stages += 1;
    }
    return 0;
}
// PrefixSumWorkspace::launch, src/utils/prefix_sum.rs:93-151
int scan_launch(ref_ctx* c, CUdeviceptr buffer, uint32_t buffer_len) {
    if (int r = scan_reserve(c, std::max<uint32_t>(buffer_len, 1))) return r;
    uint32_t ngroups0 = c->stages[0].len;
    CUdeviceptr aux0 = c->stages[0].buf.p;
    { void* a[] = {&buffer, &buffer_len, &aux0}; if (int r = launch(c, c->k_scan, ngroups0, 512, 1, 1, a)) return r; }
    for (size_t i = 0; i + 1 < c->num_stages; ++i) {
        uint32_t len = c->stages[i].len, ngroups = c->stages[i + 1].len;
        CUdeviceptr buf = c->stages[i].buf.p, aux = c->stages[i + 1].buf.p;
        void* a[] = {&buf, &len, &aux};
        if (int r = launch(c, c->k_scan, ngroups, 512, 1, 1, a)) return r;
    }
    if (c->num_stages > 2) {
        for (size_t ii = c->num_stages - 2; ii-- > 0;) {
            uint32_t len = c->stages[ii].len, ngroups = c->stages[ii + 1].len;
            CUdeviceptr buf = c->stages[ii].buf.p, aux = c->stages[ii + 1].buf.p;
            void* a[] = {&buf, &len, &aux};
            if (int r = launch(c, c->k_add, ngroups, 512, 1, 1, a)) return r;
        }
    }
    if (c->num_stages > 1) {
        void* a[] = {&buffer, &buffer_len, &aux0};
        if (int r = launch(c, c->k_add, ngroups0, 512, 1, 1, a)) return r;
    }
    return 0;
}
int scan_read_max(ref_ctx* c, uint32_t* out) { // read_max_scan_value, prefix_sum.rs:80-91
    for (size_t i = 0; i < c->stages.size(); ++i)
        if (c->stages[i].len == 1) { CU(D.p_cuMemcpyDtoH_v2(out, c->stages[i].buf.p, 4)); return 0; }
    return SPK_ESTATE;
}
uint32_t next_pow2(uint64_t v) { uint64_t p = 1; while (p < v) p <<= 1; return (uint32_t)p; }

// patch the DevicePointer<ParticleData> slots of the 96-byte models (cuda_particle_model.rs:55-72,96-113) and upload
int upload_models(ref_ctx* c) {
    if (c->models.empty()) return 0;
    std::vector<SpkGpuParticleModel> m = c->models;
    for (auto& x : m) {
        const uint64_t ptr = c->data.p;
        unsigned char* b = reinterpret_cast<unsigned char*>(&x);
        if (x.constitutive_tag == SPK_CONSTITUTIVE_COROTATED || x.constitutive_tag == SPK_CONSTITUTIVE_NEO_HOOKEAN) std::memcpy(b + 24, &ptr, 8);
        switch (x.plastic_tag) {
        case SPK_PLASTIC_DRUCKER_PRAGER: std::memcpy(b + 72, &ptr, 8); break;
        case SPK_PLASTIC_NACC: std::memcpy(b + 64, &ptr, 8); break;
        case SPK_PLASTIC_RANKINE: case SPK_PLASTIC_SNOW: std::memcpy(b + 56, &ptr, 8); break;
        default: break;
        }
    }
    if (int r = c->d_models.ensure(m.size() * 96, false)) return r;
    CU(D.p_cuMemcpyHtoD_v2(c->d_models.p, m.data(), m.size() * 96));
    return 0;
}
int alloc_grid_data(ref_ctx* c, GridData& g, uint32_t block_capacity) { // CudaSparseGrid::with_capacity, cuda_sparse_grid.rs:53-90
    g.block_capacity = block_capacity;
    if (int r = g.buffer.ensure((size_t)block_capacity * REF_CELLS * REF_NODE_BYTES, true)) return r; // GpuGridNode::default() == all-zero bytes
    if (int r = g.active_blocks.ensure((size_t)block_capacity * 24, true)) return r;
    if (int r = g.hashmap.ensure((size_t)block_capacity * 16, true)) return r;
    return reset_hashmap(c, g);
}

const float DT_FACTOR = 1000000000000.0f;
uint64_t dt_from_sec(float dt) { float v = dt * DT_FACTOR; if (std::isnan(v) || v <= 0.f) return 0; if (v >= 1.8446744073709552e19f) return UINT64_MAX; return (uint64_t)v; }
float dt_into_sec(uint64_t v) { return (float)v / DT_FACTOR; }

} // namespace

extern "C" {

int ref_available(void) { return D.load() ? 1 : 0; }

int ref_create(int device, float cell_width, const char* cubin_path, ref_ctx** out) {
    if (!out || !cubin_path) return SPK_EINVAL;
    *out = nullptr;
    if (!D.load()) return SPK_EUNSUPPORTED;
    CU(D.p_cuInit(0));
    CUdevice dev;
    CU(D.p_cuDeviceGet(&dev, device));
    ref_ctx* c = new ref_ctx();
    CU(D.p_cuDevicePrimaryCtxRetain(&c->ctx, dev));
    CU(D.p_cuCtxSetCurrent(c->ctx));
    FILE* f = std::fopen(cubin_path, "rb");
    if (!f) { delete c; return SPK_EINVAL; }
    std::vector<char> image;
    char tmp[65536];
    size_t got;
    while ((got = std::fread(tmp, 1, sizeof(tmp), f)) > 0) image.insert(image.end(), tmp, tmp + got);
    std::fclose(f);
    image.push_back(0);
    CU(D.p_cuModuleLoadData(&c->mod, image.data()));
#define FN(field, name) CU(D.p_cuModuleGetFunction(&c->field, c->mod, name));
    FN(k_reset_hashmap, "reset_hashmap") FN(k_touch, "touch_particle_blocks") FN(k_count, "update_block_particle_count")
    FN(k_copy_len, "copy_particles_len_to_scan_value") FN(k_scan, "prefix_sum_512") FN(k_add, "add_data_grp")
    FN(k_copy_first, "copy_scan_values_to_first_particles") FN(k_finalize, "finalize_particles_sort") FN(k_mult, "write_blocks_multiplicity_to_scan_value")
    FN(k_dispatch, "init_gpu_dispatch_blocks_mapping") FN(k_reset_grid, "reset_grid") FN(k_copy_proj, "copy_grid_projection_data")
    FN(k_dt, "estimate_timestep_length") FN(k_g2p2g, "g2p2g") FN(k_grid_update, "grid_update")
#undef FN
    c->cell_width = cell_width;
    const uint32_t cap = 1024; // CudaSparseGrid::new, cuda_sparse_grid.rs:48-50
    if (int r = alloc_grid_data(c, c->grid[0], cap)) return r;
    if (int r = alloc_grid_data(c, c->grid[1], cap)) return r;
    if (int r = c->dispatch.ensure((size_t)cap * 8, true)) return r;
    if (int r = c->dispatch_halo.ensure((size_t)cap * 8, true)) return r;
    if (int r = c->num_active_blocks.ensure(4, true)) return r;
    if (int r = c->scan_values.ensure((size_t)cap * 4, true)) return r;
    if (int r = c->halo_scan_values.ensure((size_t)cap * 4, true)) return r;
    if (int r = c->timestep_length.ensure(8, true)) return r;
    if (int r = scan_reserve(c, cap)) return r;
    for (auto& e : c->ev) CU(D.p_cuEventCreate(&e, 0));
    CU(D.p_cuCtxSynchronize());
    *out = c;
    return SPK_OK;
}

int ref_destroy(ref_ctx* c) {
    if (!c) return SPK_EINVAL;
    D.p_cuCtxSetCurrent(c->ctx);
    D.p_cuCtxSynchronize();
    for (DevBuf* b : {&c->status, &c->pos, &c->vel, &c->volume, &c->phase, &c->data, &c->sorted_ids, &c->d_models, &c->d_colliders, &c->dispatch, &c->dispatch_halo,
                      &c->num_active_blocks, &c->scan_values, &c->halo_scan_values, &c->timestep_length})
        b->release();
    for (auto& g : c->grid) { g.buffer.release(); g.active_blocks.release(); g.hashmap.release(); }
    for (auto& b : c->collider_bufs) b.release();
    for (auto& s : c->stages) s.buf.release();
    if (c->mod) D.p_cuModuleUnload(c->mod);
    delete c;
    return SPK_OK;
}

int ref_particles_upload(ref_ctx* c, const SpkParticleStatus* status, const SpkParticlePosition* pos, const SpkParticleVelocity* vel,
                         const SpkParticleVolume* volume, const SpkParticlePhase* phase, const SpkParticleData* data, size_t n) {
    if (!c) return SPK_EINVAL;
    CU(D.p_cuCtxSetCurrent(c->ctx));
    c->n = n;
    struct { DevBuf* b; const void* src; size_t sz; } items[] = {{&c->status, status, sizeof(SpkParticleStatus)}, {&c->pos, pos, sizeof(SpkParticlePosition)},
                                                                   {&c->vel, vel, sizeof(SpkParticleVelocity)}, {&c->volume, volume, sizeof(SpkParticleVolume)},
                                                                   {&c->phase, phase, 8}, {&c->data, data, 16}};
    std::vector<SpkParticleData> dflt;
    if (!data) { dflt.assign(n, SpkParticleData{-0.01f, 1.0f, 1.0f, 0.0f}); items[5].src = dflt.data(); }
    for (auto& it : items) {
        if (int r = it.b->ensure(n * it.sz, false)) return r;
        if (n) CU(D.p_cuMemcpyHtoD_v2(it.b->p, it.src, n * it.sz));
    }
    if (int r = c->sorted_ids.ensure(n * 4, true)) return r;
    return upload_models(c); // ParticleData may have moved
}
int ref_particles_len(const ref_ctx* c, size_t* out) { *out = c->n; return SPK_OK; }
int ref_particles_read(ref_ctx* c, int which, void* dst, size_t n) {
    if (!c || n > c->n) return SPK_EINVAL;
    CU(D.p_cuCtxSetCurrent(c->ctx));
    DevBuf* b; size_t sz;
    switch (which) {
    case SPK_ATTR_POSITION: b = &c->pos; sz = sizeof(SpkParticlePosition); break;
    case SPK_ATTR_VELOCITY: b = &c->vel; sz = sizeof(SpkParticleVelocity); break;
    case SPK_ATTR_STATUS: b = &c->status; sz = sizeof(SpkParticleStatus); break;
    case SPK_ATTR_VOLUME: b = &c->volume; sz = sizeof(SpkParticleVolume); break;
    case SPK_ATTR_PHASE: b = &c->phase; sz = 8; break;
    case SPK_ATTR_DATA: b = &c->data; sz = 16; break;
    case SPK_ATTR_SORTED_IDS: b = &c->sorted_ids; sz = 4; break;
    default: return SPK_EINVAL;
    }
    CU(D.p_cuCtxSynchronize());
    if (n) CU(D.p_cuMemcpyDtoH_v2(dst, b->p, n * sz));
    return SPK_OK;
}
int ref_models_upload(ref_ctx* c, const SpkGpuParticleModel* models, size_t n) {
    if (!c) return SPK_EINVAL;
    CU(D.p_cuCtxSetCurrent(c->ctx));
    c->models.assign(models, models + n);
    return upload_models(c);
}
int ref_colliders_upload(ref_ctx* c, const SpkGpuCollider* cols, size_t n) {
    if (!c) return SPK_EINVAL;
    CU(D.p_cuCtxSetCurrent(c->ctx));
#if SPK_DIM == 2
    // The reference's dim2 GpuCollider image (280 bytes), recovered from the ld.global offsets of `grid_update` in
    // resources/sparkl2d-kernels.ptx (collider loop: `add.s64 ..., 280` :9562-9563; BoundaryHandling::None test `[c+272] == 3` :9564-9565;
    // tag dispatch u16 @0: 1 -> HeightField, 2 -> TriMesh, 3 -> Polyline, else Cuboid :9568-9576):
    //   position: UnitComplex re @248, im @252, translation @256 / @260 (:9580-9590); friction @264, penalty @268, handling @272
    //   Cuboid (tag 0): half_extents @8 / @12 (:10289-10290)
    //   HeightField (tag 1) = parry2d GenericHeightField<CudaStoragePtr>: heights {ptr @8, len @16}, status {ptr @24, len @32} (u8,
    //     0 = segment removed: `setp.eq.s16 ..., 0 -> skip` :9999-10002, so present segments are 1), scale @40, aabb mins @48 / maxs @56,
    //     bellow_heightfield_is_solid @64, flip_interior @65 (:9918-9919, :9961-9990, :10217-10229)
    //   Polyline (tag 3): vertices {ptr @8, len @16} (Point2<f32>), flip_interior @24 (:9578, :9593-9597)
    for (auto& b : c->collider_bufs) b.release();
    c->collider_bufs.clear();
    c->collider_bufs.reserve(2 * n + 1);
    std::vector<unsigned char> img(std::max<size_t>(n, 1) * 280, 0);
    for (size_t i = 0; i < n; ++i) {
        unsigned char* b = img.data() + i * 280;
        auto put64 = [&](size_t off, uint64_t v) { std::memcpy(b + off, &v, 8); };
        const uint16_t tag = (uint16_t)cols[i].shape_tag;
        std::memcpy(b, &tag, 2);
        std::memcpy(b + 248, cols[i].rotation, 8);
        std::memcpy(b + 256, cols[i].translation, 8);
        std::memcpy(b + 264, &cols[i].friction, 4);
        std::memcpy(b + 268, &cols[i].penalty_stiffness, 4);
        std::memcpy(b + 272, &cols[i].grid_boundary_handling, 4);
        if (cols[i].shape_tag == SPK_SHAPE_CUBOID) {
            std::memcpy(b + 8, cols[i].shape.cuboid.half_extents, 8);
        } else if (cols[i].shape_tag == SPK_SHAPE_HEIGHTFIELD) {
            const auto& hf = cols[i].shape.heightfield;
            if (!hf.heights || hf.num_heights < 2) return SPK_EINVAL;
            const size_t nh = hf.num_heights, ns = nh - 1;
            c->collider_bufs.emplace_back(); DevBuf& dh = c->collider_bufs.back();
            if (int r = dh.ensure(nh * 4, false)) return r;
            CU(D.p_cuMemcpyHtoD_v2(dh.p, hf.heights, nh * 4));
            c->collider_bufs.emplace_back(); DevBuf& ds = c->collider_bufs.back();
            if (int r = ds.ensure(ns, false)) return r;
            std::vector<unsigned char> present(ns, 1);
            CU(D.p_cuMemcpyHtoD_v2(ds.p, present.data(), ns));
            put64(8, dh.p); put64(16, nh);
            put64(24, ds.p); put64(32, ns);
            std::memcpy(b + 40, hf.scale, 8);
            float lo = hf.heights[0], hi = hf.heights[0];
            for (size_t k = 1; k < nh; ++k) { lo = std::min(lo, hf.heights[k]); hi = std::max(hi, hf.heights[k]); }
            const float mins[2] = {-0.5f * hf.scale[0], lo * hf.scale[1]}, maxs[2] = {0.5f * hf.scale[0], hi * hf.scale[1]};
            std::memcpy(b + 48, mins, 8); std::memcpy(b + 56, maxs, 8);
            b[64] = hf.bellow_heightfield_is_solid; b[65] = hf.flip_interior;
        } else if (cols[i].shape_tag == SPK_SHAPE_POLYLINE) {
            const auto& pl = cols[i].shape.polyline;
            if (!pl.vertices || pl.num_vertices == 0) return SPK_EINVAL;
            c->collider_bufs.emplace_back(); DevBuf& dv = c->collider_bufs.back();
            if (int r = dv.ensure((size_t)pl.num_vertices * 8, false)) return r;
            CU(D.p_cuMemcpyHtoD_v2(dv.p, pl.vertices, (size_t)pl.num_vertices * 8));
            put64(8, dv.p); put64(16, pl.num_vertices);
            b[24] = pl.flip_interior;
        } else {
            return SPK_EUNSUPPORTED;
        }
    }
    if (int r = c->d_colliders.ensure(std::max<size_t>(n, 1) * 280, true)) return r;
    if (n) CU(D.p_cuMemcpyHtoD_v2(c->d_colliders.p, img.data(), n * 280));
#endif
#if SPK_DIM == 3
    // The reference's own byte images of the non-cuboid shapes, recovered from the ld.global offsets of `grid_update` in
    // resources/sparkl3d-kernels.ptx (tag: u16 @0; all shapes share position @296, friction @324, handling @332):
    //   HeightField (tag 1) = parry3d GenericHeightField<CudaStoragePtr>: heights {ptr @8, nrows @16, ncols @24} (f32, column-major),
    //     status {ptr @32, nrows - 1 @40, ncols - 1 @48} (u8 HeightFieldCellStatus, 0 = both triangles, default diagonal),
    //     scale @56, aabb mins @68 / maxs @80, num_triangles @96, bellow_heightfield_is_solid @104, flip_interior @105
    //   Polyline (tag 3): vertices {ptr @8, len @16} (Point3<f32>), flip_interior @24
    // TriMesh (tag 2) holds a device QBVH and is not reproduced.
    for (auto& b : c->collider_bufs) b.release();
    c->collider_bufs.clear();
    c->collider_bufs.reserve(2 * n + 1);
    std::vector<unsigned char> img(std::max<size_t>(n, 1) * 336, 0);
    for (size_t i = 0; i < n; ++i) {
        unsigned char* b = img.data() + i * 336;
        std::memcpy(b, &cols[i], 336);
        auto put64 = [&](size_t off, uint64_t v) { std::memcpy(b + off, &v, 8); };
        if (cols[i].shape_tag == SPK_SHAPE_CUBOID) continue;
        std::memset(b + 8, 0, 288);
        if (cols[i].shape_tag == SPK_SHAPE_HEIGHTFIELD) {
            const auto& hf = cols[i].shape.heightfield;
            if (!hf.heights || hf.nrows < 2 || hf.ncols < 2) return SPK_EINVAL;
            const size_t nh = (size_t)hf.nrows * hf.ncols, ns = (size_t)(hf.nrows - 1) * (hf.ncols - 1);
            c->collider_bufs.emplace_back(); DevBuf& dh = c->collider_bufs.back();
            if (int r = dh.ensure(nh * 4, false)) return r;
            CU(D.p_cuMemcpyHtoD_v2(dh.p, hf.heights, nh * 4));
            c->collider_bufs.emplace_back(); DevBuf& ds = c->collider_bufs.back();
            if (int r = ds.ensure(ns, true)) return r;
            put64(8, dh.p); put64(16, hf.nrows); put64(24, hf.ncols);
            put64(32, ds.p); put64(40, hf.nrows - 1); put64(48, hf.ncols - 1);
            std::memcpy(b + 56, hf.scale, 12);
            float lo = hf.heights[0], hi = hf.heights[0];
            for (size_t k = 1; k < nh; ++k) { lo = std::min(lo, hf.heights[k]); hi = std::max(hi, hf.heights[k]); }
            const float mins[3] = {-0.5f * hf.scale[0], lo * hf.scale[1], -0.5f * hf.scale[2]}, maxs[3] = {0.5f * hf.scale[0], hi * hf.scale[1], 0.5f * hf.scale[2]};
            std::memcpy(b + 68, mins, 12); std::memcpy(b + 80, maxs, 12);
            put64(96, (uint64_t)ns * 2);
            b[104] = hf.bellow_heightfield_is_solid; b[105] = hf.flip_interior;
        } else if (cols[i].shape_tag == SPK_SHAPE_POLYLINE) {
            const auto& pl = cols[i].shape.polyline;
            if (!pl.vertices || pl.num_vertices == 0) return SPK_EINVAL;
            c->collider_bufs.emplace_back(); DevBuf& dv = c->collider_bufs.back();
            if (int r = dv.ensure((size_t)pl.num_vertices * 12, false)) return r;
            CU(D.p_cuMemcpyHtoD_v2(dv.p, pl.vertices, (size_t)pl.num_vertices * 12));
            put64(8, dv.p); put64(16, pl.num_vertices);
            b[24] = pl.flip_interior;
        } else {
            return SPK_EUNSUPPORTED;
        }
    }
    if (int r = c->d_colliders.ensure(std::max<size_t>(n, 1) * 336, true)) return r;
    if (n) CU(D.p_cuMemcpyHtoD_v2(c->d_colliders.p, img.data(), n * 336));
#endif
    c->ncolliders = n;
    return SPK_OK;
}

int ref_frame_begin(ref_ctx* c, const SpkSolverParameters* p) { // cuda_mpm_pipeline.rs:294-299
    CU(D.p_cuCtxSetCurrent(c->ctx));
    uint64_t v = dt_from_sec(p->dt);
    CU(D.p_cuMemcpyHtoD_v2(c->timestep_length.p, &v, 8));
    return SPK_OK;
}

// CudaSparseGrid::launch_sort (single context), cuda_sparse_grid.rs:147-489
int ref_stage_sort(ref_ctx* c, SpkSortInfo* info) {
    CU(D.p_cuCtxSetCurrent(c->ctx));
    const uint32_t nthreads = 512;
    uint32_t particles_len = (uint32_t)c->n;
    CUdeviceptr pos = c->pos.p;
    GridData& g = c->nxt();
    uint32_t nab = 0;
    for (;;) {
        CU(D.p_cuMemsetD8_v2(c->num_active_blocks.p, 0, 4));
        if (int r = reset_hashmap(c, g)) return r;
        GpuGridBlob next = c->blob(g);
        void* a[] = {&pos, &particles_len, &next};
        if (int r = launch(c, c->k_touch, particles_len / nthreads + 1, nthreads, 1, 1, a)) return r;
        CU(D.p_cuCtxSynchronize());
        CU(D.p_cuMemcpyDtoH_v2(&nab, c->num_active_blocks.p, 4));
        if ((uint64_t)nab > (uint64_t)g.block_capacity * 50 / 100) { // MAX_HASHMAP_LOAD_FACTOR, cuda_sparse_grid.rs:219-222
            size_t estimated = std::max<size_t>(next_pow2((c->n * 100) / (REF_CELLS * (1u << SPK_DIM) * 50)), c->cur().block_capacity); // 4^DIM * 2^DIM * load factor
            size_t new_capacity = std::max<size_t>(estimated, (size_t)g.block_capacity * 2);
            g.active_blocks.release(); g.hashmap.release();
            if (int r = g.active_blocks.ensure(new_capacity * 24, true)) return r;
            if (int r = g.hashmap.ensure(new_capacity * 16, true)) return r;
            g.block_capacity = (uint32_t)new_capacity;
            c->scan_values.release(); c->halo_scan_values.release();
            if (int r = c->scan_values.ensure(new_capacity * 4, true)) return r;
            if (int r = c->halo_scan_values.ensure(new_capacity * 4, true)) return r;
        } else break;
    }
    g.nab = nab;
    GpuGridBlob next = c->blob(g);
    CUdeviceptr scan = c->scan_values.p, halo_scan = c->halo_scan_values.p, sorted = c->sorted_ids.p;
    uint32_t ngroups = particles_len / nthreads + 1, nbg = nab / nthreads + 1, tpb = REF_G2P2G_THREADS;
    { void* a[] = {&pos, &particles_len, &next}; if (int r = launch(c, c->k_count, ngroups, nthreads, 1, 1, a)) return r; }
    { void* a[] = {&next, &scan}; if (int r = launch(c, c->k_copy_len, nbg, nthreads, 1, 1, a)) return r; }
    if (int r = scan_launch(c, scan, nab)) return r;
    { void* a[] = {&next, &scan}; if (int r = launch(c, c->k_copy_first, nbg, nthreads, 1, 1, a)) return r; }
    { void* a[] = {&pos, &particles_len, &next, &scan, &sorted}; if (int r = launch(c, c->k_finalize, ngroups, nthreads, 1, 1, a)) return r; }
    { void* a[] = {&next, &scan, &halo_scan, &tpb}; if (int r = launch(c, c->k_mult, nbg, nthreads, 1, 1, a)) return r; }
    if (int r = scan_launch(c, scan, nab)) return r;
    CU(D.p_cuCtxSynchronize());
    if (int r = scan_read_max(c, &c->num_dispatch_blocks)) return r;
    if (c->dispatch.bytes / 8 < c->num_dispatch_blocks) {
        c->dispatch.release();
        if (int r = c->dispatch.ensure((size_t)next_pow2(c->num_dispatch_blocks) * 8, true)) return r;
    }
    next = c->blob(g);
    if (nab) { void* a[] = {&next, &scan, &halo_scan, &tpb}; if (int r = launch(c, c->k_dispatch, nab, 32, 1, 1, a)) return r; }
    const size_t desired = (size_t)nab * REF_CELLS;
    if (g.buffer.bytes / REF_NODE_BYTES < desired) { // grown buffers are initialised with default nodes (cuda_sparse_grid.rs:455-461)
        g.buffer.release();
        if (int r = g.buffer.ensure(desired * REF_NODE_BYTES, true)) return r;
    }
    if (info) { info->num_active_blocks = nab; info->num_dispatch_blocks = c->num_dispatch_blocks; info->num_dispatch_halo_blocks = 0; info->hashmap_capacity = g.block_capacity; }
    return SPK_OK;
}

int ref_stage_reset_grid(ref_ctx* c) { // cuda_mpm_pipeline.rs:371-376
    CU(D.p_cuCtxSetCurrent(c->ctx));
    GridData& g = c->nxt();
    if (!g.nab) return SPK_OK;
    GpuGridBlob next = c->blob(g), curr = c->blob(c->cur());
    { void* a[] = {&next}; if (int r = launch(c, c->k_reset_grid, g.nab, 4, 4, REF_BLOCK_Z, a)) return r; }
    { void* a[] = {&curr, &next}; if (int r = launch(c, c->k_copy_proj, g.nab, 4, 4, REF_BLOCK_Z, a)) return r; }
    return SPK_OK;
}

int ref_stage_estimate_dt(ref_ctx* c, float min_dt, float max_dt, float* dt_out) { // cuda_particle_kernels.rs:38-62, cuda_mpm_pipeline.rs:381-398
    CU(D.p_cuCtxSetCurrent(c->ctx));
    CUdeviceptr status = c->status.p, volume = c->volume.p, vel = c->vel.p, models = c->d_models.p, out = c->timestep_length.p;
    uint64_t n = c->n;
    float h = c->cell_width;
    void* a[] = {&min_dt, &max_dt, &status, &volume, &vel, &n, &models, &h, &out};
    if (int r = launch(c, c->k_dt, (unsigned)(c->n / 512 + 1), 512, 1, 1, a)) return r;
    CU(D.p_cuCtxSynchronize());
    uint64_t v;
    CU(D.p_cuMemcpyDtoH_v2(&v, c->timestep_length.p, 8));
    if (dt_out) *dt_out = dt_into_sec(v);
    return SPK_OK;
}
int ref_dt_acc_get(ref_ctx* c, uint64_t* out) { CU(D.p_cuCtxSetCurrent(c->ctx)); CU(D.p_cuCtxSynchronize()); CU(D.p_cuMemcpyDtoH_v2(out, c->timestep_length.p, 8)); return SPK_OK; }
int ref_dt_acc_set(ref_ctx* c, uint64_t v) { CU(D.p_cuCtxSetCurrent(c->ctx)); CU(D.p_cuMemcpyHtoD_v2(c->timestep_length.p, &v, 8)); return SPK_OK; }

int ref_stage_g2p2g(ref_ctx* c, float dt, uint32_t damage_model) { // cuda_particle_kernels.rs:64-99
    CU(D.p_cuCtxSetCurrent(c->ctx));
    if (c->num_dispatch_blocks == 0) return SPK_OK; // cuda_mpm_pipeline.rs:485-486
    CUdeviceptr cols = c->d_colliders.p, status = c->status.p, pos = c->pos.p, vel = c->vel.p, volume = c->volume.p, phase = c->phase.p, sorted = c->sorted_ids.p,
                models = c->d_models.p;
    uint64_t ncol = c->ncolliders;
    GpuGridBlob curr = c->blob(c->cur()), next = c->blob(c->nxt());
    uint8_t halo = 0;
    void* a[] = {&dt, &cols, &ncol, &status, &pos, &vel, &volume, &phase, &sorted, &models, &curr, &next, &damage_model, &halo};
    return launch(c, c->k_g2p2g, c->num_dispatch_blocks, REF_G2P2G_THREADS, 1, 1, a);
}

int ref_stage_grid_update(ref_ctx* c, float dt, const float gravity[SPK_DIM]) { // cuda_mpm_pipeline.rs:536-558
    CU(D.p_cuCtxSetCurrent(c->ctx));
    GridData& g = c->nxt();
    if (g.nab) {
        GpuGridBlob next = c->blob(g);
        CUdeviceptr cols = c->d_colliders.p;
        uint64_t ncol = c->ncolliders;
        float grav[3] = {gravity[0], gravity[1], SPK_DIM == 3 ? gravity[SPK_DIM - 1] : 0.f}; // Vector<Real> by value: 12 bytes (dim2: 8)
        void* a[] = {&dt, &next, &cols, &ncol, grav};
        if (int r = launch(c, c->k_grid_update, g.nab, 4, 4, REF_BLOCK_Z, a)) return r;
    }
    c->curr = 1 - c->curr; // swap_buffers
    return SPK_OK;
}

// CudaMpmPipeline::step_generic (one context) with the reference's own per-phase event timers (CudaDeviceTimings)
int ref_step(ref_ctx* c, const SpkSolverParameters* params, const float gravity[SPK_DIM], SpkStepTimings* t, float* simulated_time_out) {
    if (!c || !params) return SPK_EINVAL;
    CU(D.p_cuCtxSetCurrent(c->ctx));
    if (t) std::memset(t, 0, sizeof(*t));
    int r;
    if ((r = ref_frame_begin(c, params))) return r;
    if (params->dt == 0.0f || c->n == 0) { if (simulated_time_out) *simulated_time_out = params->dt; return SPK_OK; }
    const float min_dt = params->dt / (float)params->max_num_substeps;
    float remaining = params->dt;
    uint32_t niter = 0;
    float acc[5] = {0, 0, 0, 0, 0};
    CU(D.p_cuEventRecord(c->ev[6], nullptr));
    while (remaining > 0.0f) {
        ++niter;
        SpkSortInfo info;
        CU(D.p_cuEventRecord(c->ev[0], nullptr));
        if ((r = ref_stage_sort(c, &info))) return r;
        CU(D.p_cuEventRecord(c->ev[1], nullptr));
        if ((r = ref_stage_reset_grid(c))) return r;
        CU(D.p_cuEventRecord(c->ev[2], nullptr));
        float dt;
        if ((r = ref_stage_estimate_dt(c, min_dt, std::fmin(remaining, params->max_substep_dt), &dt))) return r;
        CU(D.p_cuEventRecord(c->ev[3], nullptr));
        if ((r = ref_stage_g2p2g(c, dt, params->damage_model))) return r;
        CU(D.p_cuEventRecord(c->ev[4], nullptr));
        if ((r = ref_stage_grid_update(c, dt, gravity))) return r;
        CU(D.p_cuEventRecord(c->ev[5], nullptr));
        CU(D.p_cuEventSynchronize(c->ev[5]));
        for (int k = 0; k < 5; ++k) { float ms = 0; CU(D.p_cuEventElapsedTime(&ms, c->ev[k], c->ev[k + 1])); acc[k] += ms; }
        if (t) { t->last_dt = dt; t->num_active_blocks = info.num_active_blocks; t->num_dispatch_blocks = info.num_dispatch_blocks; }
        remaining -= dt;
        if (remaining <= 10.0f / DT_FACTOR || params->stop_after_one_substep) break;
    }
    CU(D.p_cuEventRecord(c->ev[7], nullptr));
    CU(D.p_cuEventSynchronize(c->ev[7]));
    if (t) {
        CU(D.p_cuEventElapsedTime(&t->total_ms, c->ev[6], c->ev[7]));
        t->num_substeps = niter;
        t->grid_resize_and_sort_ms = acc[0]; t->reset_grid_ms = acc[1]; t->estimate_timestep_ms = acc[2]; t->g2p2g_ms = acc[3]; t->grid_update_ms = acc[4];
        t->g2p2g_launches = niter;
    }
    if (simulated_time_out) *simulated_time_out = params->dt - remaining;
    return SPK_OK;
}

int ref_grid_num_active_blocks(ref_ctx* c, int which, uint32_t* out) { *out = (which ? c->nxt() : c->cur()).nab; return SPK_OK; }
int ref_grid_read_headers(ref_ctx* c, int which, SpkActiveBlockHeader* dst, uint32_t n) {
    CU(D.p_cuCtxSetCurrent(c->ctx));
    GridData& g = which ? c->nxt() : c->cur();
    if (n > g.nab) return SPK_EINVAL;
    CU(D.p_cuCtxSynchronize());
    if (n) CU(D.p_cuMemcpyDtoH_v2(dst, g.active_blocks.p, (size_t)n * 24));
    return SPK_OK;
}
int ref_grid_read_nodes(ref_ctx* c, int which, SpkGpuGridNode* dst, uint32_t n_blocks) {
    CU(D.p_cuCtxSetCurrent(c->ctx));
    GridData& g = which ? c->nxt() : c->cur();
    if (n_blocks > g.nab) return SPK_EINVAL;
    CU(D.p_cuCtxSynchronize());
    if (n_blocks) CU(D.p_cuMemcpyDtoH_v2(dst, g.buffer.p, (size_t)n_blocks * REF_CELLS * REF_NODE_BYTES));
    return SPK_OK;
}
int ref_grid_read_dispatch(ref_ctx* c, SpkDispatchBlock2ActiveBlock* dst, uint32_t n) {
    CU(D.p_cuCtxSetCurrent(c->ctx));
    if (n > c->num_dispatch_blocks) return SPK_EINVAL;
    CU(D.p_cuCtxSynchronize());
    if (n) CU(D.p_cuMemcpyDtoH_v2(dst, c->dispatch.p, (size_t)n * 8));
    return SPK_OK;
}

} // extern "C"
