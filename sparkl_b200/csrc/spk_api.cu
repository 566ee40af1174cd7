// spk_api.cu — the C ABI of include/sparkl_abi.h: context / buffers, particle pack & unpack, and the host sequencing
// of one `CudaMpmPipeline::step` (src/cuda/cuda_mpm_pipeline.rs:273-659, src/cuda/cuda_sparse_grid.rs:147-489).
//
// Host-side redesign: ONE blocking read-back per substep (a 64-byte pinned record holding the active-block count,
// the dispatch count, the overflow flag and the dt accumulator) instead of the reference's three
// (cuda_sparse_grid.rs:211-215,392-394, cuda_mpm_pipeline.rs:393-398); buffers only ever grow on the device (no
// host-initialised `vec![default; n]` + H2D on growth, cuda_sparse_grid.rs:459-475).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "spk_device.cuh"

#define CK(expr)                                   \
    do {                                           \
        cudaError_t _e = (expr);                   \
        if (_e != cudaSuccess) return (int)_e;     \
    } while (0)

namespace {

template <typename T>
cudaError_t dev_alloc(T*& p, size_t count) {
    p = nullptr;
    if (count == 0) count = 1;
    return cudaMalloc((void**)&p, count * sizeof(T));
}
template <typename T>
void dev_free(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}
// particles per cell of the reference's own grid-size estimate: 2^DIM (cuda_sparse_grid.rs:222-229)
#define SPK_PARTICLES_PER_CELL (1 << SPK_DIM)
uint32_t next_pow2(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return (uint32_t)std::min<uint64_t>(p, 1ull << 31);
}

void free_grid(GridBuffers& g) {
    if (g.external_hash) { g.keys = nullptr; g.vals = nullptr; g.external_hash = 0; } // owned by the multi-GPU arena
    dev_free(g.keys); dev_free(g.vals); dev_free(g.blk_key); dev_free(g.blk_first); dev_free(g.blk_count); dev_free(g.blk_halo); dev_free(g.blk_dfirst);
    dev_free(g.nbr_self); dev_free(g.nbr_prev); dev_free(g.vel); dev_free(g.acc); dev_free(g.proj); dev_free(g.cacheN); dev_free(g.cacheD);
}
// (re)allocate one grid buffer for `cap_hash` hash slots; contents are NOT preserved (only ever done on the next grid,
// which launch_sort rebuilds from scratch — cuda_sparse_grid.rs:235-243)
cudaError_t alloc_grid(spk_ctx* c, GridBuffers& g, uint32_t cap_hash) {
    free_grid(g);
    g.cap_hash = cap_hash;
    g.cap_blocks = cap_hash / 2;
    const size_t nb = g.cap_blocks, nn = nb * SPK_CELLS_PER_BLOCK;
    cudaError_t e;
    if (c->dist_arena) {
        // multi-GPU: the hash table lives in the arena shared with the peers and cannot move or grow
        if (!spk_dist_grid_storage(c, (int)(&g - c->grid), cap_hash, &g.keys, &g.vals)) return cudaErrorMemoryAllocation;
        g.external_hash = 1;
    } else {
        if ((e = dev_alloc(g.keys, cap_hash)) != cudaSuccess) return e;
        if ((e = dev_alloc(g.vals, cap_hash)) != cudaSuccess) return e;
    }
    if ((e = dev_alloc(g.blk_key, nb)) != cudaSuccess) return e;
    if ((e = dev_alloc(g.blk_first, nb)) != cudaSuccess) return e;
    if ((e = dev_alloc(g.blk_count, nb)) != cudaSuccess) return e;
    if ((e = dev_alloc(g.blk_halo, nb)) != cudaSuccess) return e;
    if ((e = dev_alloc(g.blk_dfirst, nb)) != cudaSuccess) return e;
    if ((e = dev_alloc(g.nbr_self, nb * SPK_NBR)) != cudaSuccess) return e;
    if ((e = dev_alloc(g.nbr_prev, nb * SPK_NBR)) != cudaSuccess) return e;
    if ((e = dev_alloc(g.vel, nn)) != cudaSuccess) return e;
    if ((e = dev_alloc(g.acc, nn * SPK_ACC_CHANNELS)) != cudaSuccess) return e;
    if ((e = dev_alloc(g.proj, nn)) != cudaSuccess) return e;
    if ((e = dev_alloc(g.cacheN, nn)) != cudaSuccess) return e;
    if ((e = dev_alloc(g.cacheD, nn)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(g.keys, 0xff, (size_t)cap_hash * 8, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(g.vel, 0, nn * sizeof(float4), c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(g.acc, 0, nn * SPK_ACC_CHANNELS * sizeof(float), c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(g.proj, 0, nn * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(g.cacheN, 0, nn * sizeof(float4), c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(g.cacheD, 0, nn * sizeof(float4), c->stream)) != cudaSuccess) return e;
    g.nab_host = 0;
    g.has_velocity = 0;
    return cudaSuccess;
}
cudaError_t ensure_slot_scratch(spk_ctx* c, uint32_t cap_hash) {
    if (c->slot_cnt_cap >= cap_hash) return cudaSuccess;
    dev_free(c->slot_cnt);
    dev_free(c->cell_cnt);
    dev_free(c->mslot);
    cudaError_t e = dev_alloc(c->slot_cnt, cap_hash);
    if (e == cudaSuccess) e = dev_alloc(c->cell_cnt, (size_t)cap_hash * SPK_CELLS_PER_BLOCK);
    if (e == cudaSuccess) e = dev_alloc(c->mslot, (size_t)cap_hash * 4); // SPK_MORTON_FACTOR entries per hash slot (spk_sort.cu)
    if (e == cudaSuccess) e = cudaMemsetAsync(c->dist_bbox, 0, 16 * sizeof(uint32_t), c->stream); // no Morton box yet
    c->slot_cnt_cap = e == cudaSuccess ? cap_hash : 0;
    return e;
}
cudaError_t ensure_dispatch(spk_ctx* c) {
    const GridBuffers& g = c->grid[1 - c->gcur];
    const uint64_t need = c->n / c->chunk + g.cap_blocks + 2;
    if (c->dispatch_cap >= need) return cudaSuccess;
    dev_free(c->dispatch);
    dev_free(c->dispatch_nbr);
    cudaError_t e = dev_alloc(c->dispatch, need);
    if (e == cudaSuccess) e = dev_alloc(c->dispatch_nbr, need * 2 * SPK_NBR);
    c->dispatch_cap = e == cudaSuccess ? (uint32_t)need : 0;
    return e;
}
cudaError_t ensure_staging(spk_ctx* c, size_t bytes) {
    bytes = std::max<size_t>(bytes, 1 << 20);
    if (c->staging_bytes >= bytes) return cudaSuccess;
    dev_free(c->staging);
    cudaError_t e = cudaMalloc(&c->staging, bytes);
    c->staging_bytes = e == cudaSuccess ? bytes : 0;
    return e;
}
void free_particles(spk_ctx* c) {
    for (int b = 0; b < 2; ++b) {
        dev_free(c->pb[b].a0); dev_free(c->pb[b].a1); dev_free(c->pb[b].a2); dev_free(c->pb[b].a3); dev_free(c->pb[b].a4); dev_free(c->pb[b].a5);
    }
    dev_free(c->vel); dev_free(c->cold); dev_free(c->slot_of); dev_free(c->rank_of); dev_free(c->perm); dev_free(c->gid);
    c->cap = 0;
}
cudaError_t alloc_particles(spk_ctx* c, size_t cap) {
    free_particles(c);
    cudaError_t e;
    for (int b = 0; b < 2; ++b) {
        if ((e = dev_alloc(c->pb[b].a0, cap)) != cudaSuccess) return e;
        if ((e = dev_alloc(c->pb[b].a1, cap)) != cudaSuccess) return e;
        if ((e = dev_alloc(c->pb[b].a2, cap)) != cudaSuccess) return e;
        if ((e = dev_alloc(c->pb[b].a3, cap)) != cudaSuccess) return e;
        if ((e = dev_alloc(c->pb[b].a4, cap)) != cudaSuccess) return e;
        if ((e = dev_alloc(c->pb[b].a5, cap)) != cudaSuccess) return e;
    }
    if ((e = dev_alloc(c->vel, cap)) != cudaSuccess) return e;
    if ((e = dev_alloc(c->cold, cap)) != cudaSuccess) return e;
    if ((e = dev_alloc(c->slot_of, cap)) != cudaSuccess) return e;
    if ((e = dev_alloc(c->rank_of, cap)) != cudaSuccess) return e;
    if ((e = dev_alloc(c->perm, cap)) != cudaSuccess) return e;
    c->cap = cap;
    return cudaSuccess;
}

// ---- pack: reference AoS-per-attribute arrays -> internal float4 streams (ids first..first+n) --------------------
__global__ void pack_kernel(ParticleBuffers pb, float4* vel, float4* cold, const SpkParticleStatus* status, const SpkParticlePosition* pos,
                            const SpkParticleVelocity* v, const SpkParticleVolume* volume, const SpkParticlePhase* phase, const SpkParticleData* data,
                            size_t first, size_t n, uint32_t* bad_flag) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const size_t i = first + k;
    const SpkParticleStatus st = status[k];
    if (st.model_index >= (1u << 24)) *bad_flag = 1u; // the flags word keeps 24 bits of model index
    const SpkParticleVolume vo = volume[k];
    const uint32_t flags = (st.failed ? SPK_F_FAILED : 0u) | (st.is_static ? SPK_F_STATIC : 0u) | (st.kinematic_vel_enabled ? SPK_F_KINEMATIC : 0u) |
                           ((uint32_t)st.model_index << SPK_F_MODEL_SHIFT);
    const float* F = vo.deformation_gradient;
#if SPK_DIM == 3
    pb.a0[i] = make_float4(pos[k].point[0], pos[k].point[1], pos[k].point[2], __uint_as_float(flags));
    pb.a1[i] = make_float4(F[0], F[1], F[2], F[3]);
    pb.a2[i] = make_float4(F[4], F[5], F[6], F[7]);
    pb.a3[i] = make_float4(F[8], vo.plastic_deformation_gradient_det, vo.mass, vo.volume0);
#else // dim2: a1 holds the whole column-major 2x2 F, a2 is unused
    pb.a0[i] = make_float4(pos[k].point[0], pos[k].point[1], 0.f, __uint_as_float(flags));
    pb.a1[i] = make_float4(F[0], F[1], F[2], F[3]);
    pb.a2[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    pb.a3[i] = make_float4(0.f, vo.plastic_deformation_gradient_det, vo.mass, vo.volume0);
#endif
    SpkParticleData d = data ? data[k] : SpkParticleData{-0.01f, 1.0f, 1.0f, 0.0f}; // ParticleData::default, particle.rs:211-219
    pb.a4[i] = make_float4(phase[k].phase, phase[k].psi_pos, d.elastic_hardening, d.plastic_hardening);
    pb.a5[i] = make_float4(d.log_vol_gain, d.nacc_alpha, __uint_as_float((uint32_t)i), 0.f);
#if SPK_DIM == 3
    vel[i] = make_float4(v[k].vector[0], v[k].vector[1], v[k].vector[2], 0.f);
    cold[i] = make_float4(st.kinematic_vel[0], st.kinematic_vel[1], st.kinematic_vel[2], vo.radius0);
#else
    vel[i] = make_float4(v[k].vector[0], v[k].vector[1], 0.f, 0.f);
    cold[i] = make_float4(st.kinematic_vel[0], st.kinematic_vel[1], 0.f, vo.radius0);
#endif
}
// ---- unpack: internal streams (physical order) -> reference layout indexed by original particle id ---------------
__global__ void unpack_kernel(ParticleBuffers pb, const float4* vel, const float4* cold, const uint32_t* perm_or_null, int which, void* out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (which == SPK_ATTR_SORTED_IDS) {
        const size_t src = perm_or_null ? perm_or_null[i] : i;
        ((uint32_t*)out)[i] = __float_as_uint(pb.a5[src].z);
        return;
    }
    const uint32_t id = __float_as_uint(pb.a5[i].z);
    if (id >= n) return;
    switch (which) {
#if SPK_DIM == 3
    case SPK_ATTR_POSITION: { const float4 r = pb.a0[i]; SpkParticlePosition o; o.point[0] = r.x; o.point[1] = r.y; o.point[2] = r.z; ((SpkParticlePosition*)out)[id] = o; break; }
    case SPK_ATTR_VELOCITY: { const float4 r = vel[i]; SpkParticleVelocity o; o.vector[0] = r.x; o.vector[1] = r.y; o.vector[2] = r.z; ((SpkParticleVelocity*)out)[id] = o; break; }
#else
    case SPK_ATTR_POSITION: { const float4 r = pb.a0[i]; SpkParticlePosition o; o.point[0] = r.x; o.point[1] = r.y; ((SpkParticlePosition*)out)[id] = o; break; }
    case SPK_ATTR_VELOCITY: { const float4 r = vel[i]; SpkParticleVelocity o; o.vector[0] = r.x; o.vector[1] = r.y; ((SpkParticleVelocity*)out)[id] = o; break; }
#endif
    case SPK_ATTR_STATUS: {
        const uint32_t f = __float_as_uint(pb.a0[i].w);
        const float4 cd = cold[id];
        SpkParticleStatus o;
        o.failed = (f & SPK_F_FAILED) ? 1 : 0; o.is_static = (f & SPK_F_STATIC) ? 1 : 0; o.kinematic_vel_enabled = (f & SPK_F_KINEMATIC) ? 1 : 0; o._pad0 = 0;
        o.kinematic_vel[0] = cd.x; o.kinematic_vel[1] = cd.y;
#if SPK_DIM == 3
        o.kinematic_vel[2] = cd.z;
#else
        o._pad1 = 0;
#endif
        o.model_index = f >> SPK_F_MODEL_SHIFT;
        ((SpkParticleStatus*)out)[id] = o;
        break;
    }
    case SPK_ATTR_VOLUME: {
        const float4 r1 = pb.a1[i], r2 = pb.a2[i], r3 = pb.a3[i];
        SpkParticleVolume o;
        o.mass = r3.z; o.volume0 = r3.w; o.radius0 = cold[id].w;
        o.deformation_gradient[0] = r1.x; o.deformation_gradient[1] = r1.y; o.deformation_gradient[2] = r1.z; o.deformation_gradient[3] = r1.w;
#if SPK_DIM == 3
        o.deformation_gradient[4] = r2.x; o.deformation_gradient[5] = r2.y; o.deformation_gradient[6] = r2.z; o.deformation_gradient[7] = r2.w;
        o.deformation_gradient[8] = r3.x;
#else
        (void)r2;
#endif
        o.plastic_deformation_gradient_det = r3.y;
        ((SpkParticleVolume*)out)[id] = o;
        break;
    }
    case SPK_ATTR_PHASE: { const float4 r = pb.a4[i]; SpkParticlePhase o; o.phase = r.x; o.psi_pos = r.y; ((SpkParticlePhase*)out)[id] = o; break; }
    case SPK_ATTR_DATA: {
        const float4 r4 = pb.a4[i], r5 = pb.a5[i];
        SpkParticleData o; o.nacc_alpha = r5.y; o.plastic_hardening = r4.w; o.elastic_hardening = r4.z; o.log_vol_gain = r5.x;
        ((SpkParticleData*)out)[id] = o;
        break;
    }
    default: break;
    }
}

size_t attr_size(int which) {
    switch (which) {
    case SPK_ATTR_POSITION: return sizeof(SpkParticlePosition);
    case SPK_ATTR_VELOCITY: return sizeof(SpkParticleVelocity);
    case SPK_ATTR_STATUS: return sizeof(SpkParticleStatus);
    case SPK_ATTR_VOLUME: return sizeof(SpkParticleVolume);
    case SPK_ATTR_PHASE: return sizeof(SpkParticlePhase);
    case SPK_ATTR_DATA: return sizeof(SpkParticleData);
    case SPK_ATTR_SORTED_IDS: return sizeof(uint32_t);
    default: return 0;
    }
}

// the 64-byte counter record reaches the host through MAPPED pinned memory written by a one-warp kernel, not through a DMA
// copy: a copy would queue on the device-to-host copy engine behind any bulk read-back in flight
// (spk_particles_read_async) and stall the substep loop until that transfer ends
__global__ void publish_counters_kernel(const DeviceCounters* __restrict__ src, DeviceCounters* dst_mapped) {
    reinterpret_cast<uint32_t*>(dst_mapped)[threadIdx.x] = reinterpret_cast<const uint32_t*>(src)[threadIdx.x];
    __threadfence_system();
}
// split version: publish now, keep enqueueing work that does not need the host's answer, wait later
int publish_counters(spk_ctx* c) {
    publish_counters_kernel<<<1, sizeof(DeviceCounters) / 4, 0, c->stream>>>(c->d_counters, c->h_counters_dev);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev_counters, c->stream));
    return SPK_OK;
}
int wait_counters(spk_ctx* c) {
    CK(cudaEventSynchronize(c->ev_counters));
    c->counters_fresh = true;
    return c->h_counters->pad[1] ? SPK_EPEER : SPK_OK; // a device-side barrier of the multi-GPU path gave up on a peer (sticky)
}
int read_counters(spk_ctx* c) {
    static_assert(sizeof(DeviceCounters) % 4 == 0 && sizeof(DeviceCounters) <= 128, "one warp copies the record");
    if (spk_dist_active(c) && c->gu_pending) { CK(cudaStreamWaitEvent(c->stream, c->ev_gu_done, 0)); c->gu_pending = false; } // the dt minimum may still be on its way there
    publish_counters_kernel<<<1, sizeof(DeviceCounters) / 4, 0, c->stream>>>(c->d_counters, c->h_counters_dev);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    c->counters_fresh = true;
    return c->h_counters->pad[1] ? SPK_EPEER : SPK_OK;
}

const float DT_FACTOR = 1000000000000.0f; // GpuTimestepLength::FACTOR, gpu_timestep.rs:18
unsigned long long host_dt_from_sec(float dt) {
    const float v = dt * DT_FACTOR;
    if (std::isnan(v) || v <= 0.f) return 0ull;
    if (v >= 1.8446744073709552e19f) return 0xffffffffffffffffull;
    return (unsigned long long)v;
}
float host_dt_into_sec(unsigned long long v) { return (float)v / DT_FACTOR; }

int upload_range(spk_ctx* c, const SpkParticleStatus* status, const SpkParticlePosition* pos, const SpkParticleVelocity* vel, const SpkParticleVolume* volume,
                 const SpkParticlePhase* phase, const SpkParticleData* data, size_t first, size_t n) {
    if (n == 0) return SPK_OK;
    // stage the host AoS arrays on the device, then repack with one kernel
    const size_t per = sizeof(SpkParticleStatus) + sizeof(SpkParticlePosition) + sizeof(SpkParticleVelocity) + sizeof(SpkParticleVolume) +
                       sizeof(SpkParticlePhase) + sizeof(SpkParticleData);
    CK(ensure_staging(c, n * per + 256));
    char* base = (char*)c->staging;
    auto put = [&](const void* src, size_t bytes, char*& cursor) -> cudaError_t {
        cudaError_t e = cudaMemcpyAsync(cursor, src, bytes, cudaMemcpyHostToDevice, c->stream);
        cursor += (bytes + 15) & ~size_t(15);
        return e;
    };
    char* cur = base;
    char* d_status = cur; CK(put(status, n * sizeof(SpkParticleStatus), cur));
    char* d_pos = cur; CK(put(pos, n * sizeof(SpkParticlePosition), cur));
    char* d_vel = cur; CK(put(vel, n * sizeof(SpkParticleVelocity), cur));
    char* d_volume = cur; CK(put(volume, n * sizeof(SpkParticleVolume), cur));
    char* d_phase = cur; CK(put(phase, n * sizeof(SpkParticlePhase), cur));
    char* d_data = nullptr;
    if (data) { d_data = cur; CK(put(data, n * sizeof(SpkParticleData), cur)); }
    pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->pb[c->cur], c->vel, c->cold, (const SpkParticleStatus*)d_status, (const SpkParticlePosition*)d_pos,
                                                                     (const SpkParticleVelocity*)d_vel, (const SpkParticleVolume*)d_volume,
                                                                     (const SpkParticlePhase*)d_phase, (const SpkParticleData*)d_data, first, n, &c->d_counters->pad[0]);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(c->h_upload + 1, &c->d_counters->pad[0], 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream)); // host arrays may be pageable: keep the call synchronous like CudaVec::from_slice
    if (*(uint32_t*)(c->h_upload + 1)) { // a model index does not fit (validated on the device: no host pass over the arrays)
        CK(cudaMemsetAsync(&c->d_counters->pad[0], 0, 4, c->stream));
        c->n = first;
        return SPK_EINVAL;
    }
    return SPK_OK;
}

} // namespace

// ---- internal entry points for spk_migrate.cu (device-resident counterparts of upload / read) ------------------------------
cudaError_t spk_launch_pack_upload(spk_ctx* c, const void* d_status, const void* d_pos, const void* d_vel, const void* d_volume, const void* d_phase,
                                   const void* d_data, size_t first, size_t n) {
    if (n == 0) return cudaSuccess;
    pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->pb[c->cur], c->vel, c->cold, (const SpkParticleStatus*)d_status, (const SpkParticlePosition*)d_pos,
                                                                     (const SpkParticleVelocity*)d_vel, (const SpkParticleVolume*)d_volume, (const SpkParticlePhase*)d_phase,
                                                                     (const SpkParticleData*)d_data, first, n, &c->d_counters->pad[0]);
    c->launches += 1;
    return cudaGetLastError();
}
cudaError_t spk_launch_unpack(spk_ctx* c, int which, void* d_out, size_t n) {
    if (n == 0) return cudaSuccess;
    unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->pb[c->cur], c->vel, c->cold, nullptr, which, d_out, n);
    c->launches += 1;
    return cudaGetLastError();
}
// the bookkeeping of spk_particles_upload without the copy: capacity (reallocating DISCARDS the particle streams), counts, grid sizing
int spk_particles_resize(spk_ctx* c, size_t n, bool fresh) {
    if (n >= 0xffffffffull) return SPK_EOVERFLOW;
    // every capacity is checked BEFORE the context changes: on SPK_EOVERFLOW the particle set is exactly what it was
    const uint32_t want = std::max<uint32_t>(next_pow2((uint64_t)n * 100 / (SPK_CELLS_PER_BLOCK * SPK_PARTICLES_PER_CELL * 50)) * 4u, 2048u);
    if (c->dist_arena && want > c->dist_cap_hash) return SPK_EOVERFLOW; // the arena shared with the peers cannot grow
    if (n > c->cap) { CK(cudaStreamSynchronize(c->stream)); CK(alloc_particles(c, n + n / 8 + 1024)); fresh = true; }
    c->n = n;
    if (fresh) c->cur = 0;
    c->sorted_valid = false;
    c->dt_cache_valid = false;
    GridBuffers& next = c->grid[1 - c->gcur];
    if (next.cap_hash < want) { CK(alloc_grid(c, next, want)); CK(ensure_slot_scratch(c, want)); }
    CK(ensure_dispatch(c));
    return SPK_OK;
}

// orders the main stream behind a grid update that spk_step left running on its own stream (no-op otherwise); called by every entry
// point that reads or writes the sparse grid outside spk_step's own loop
int spk_join_grid_update(spk_ctx* c) {
    if (c->gu_pending) { CK(cudaStreamWaitEvent(c->stream, c->ev_gu_done, 0)); c->gu_pending = false; }
    return SPK_OK;
}

// CudaParticleKernelsLauncher dispatch: the registered callbacks, else the default launches
static int launch_g2p2g_hooked(spk_ctx* c, float dt) {
    if (c->launcher.launch_g2p2g) return c->launcher.launch_g2p2g(c->launcher.user, c, dt);
    return (int)spk_launch_g2p2g(c, dt);
}
static int launch_estimate_dt_hooked(spk_ctx* c, float min_dt, float max_dt) {
    if (c->launcher.launch_estimate_particle_timestep_length) return c->launcher.launch_estimate_particle_timestep_length(c->launcher.user, c, min_dt, max_dt);
    return (int)spk_launch_estimate_dt(c, min_dt, max_dt);
}

int spk_dist_realloc_grids(spk_ctx* c, uint32_t cap_hash) {
    for (int g = 0; g < 2; ++g) CK(alloc_grid(c, c->grid[g], cap_hash));
    CK(ensure_slot_scratch(c, cap_hash));
    CK(ensure_dispatch(c));
    c->gcur = 0;
    c->sorted_valid = false;
    c->dt_cache_valid = false;
    CK(cudaStreamSynchronize(c->stream));
    return SPK_OK;
}

extern "C" {

int spk_abi_version(void) { return SPK_ABI_VERSION; }
int spk_dim(void) { return SPK_DIM; }

const char* spk_error_string(int code) {
    switch (code) {
    case SPK_OK: return "ok";
    case SPK_EINVAL: return "invalid argument";
    case SPK_EUNSUPPORTED: return "unsupported (outside the hot-path scope)";
    case SPK_ESTATE: return "invalid call sequence";
    case SPK_EOVERFLOW: return "u32 particle/block limit exceeded";
    case SPK_EPEER: return "a peer rank did not reach a device-side barrier in time (multi-GPU)";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

int spk_ctx_create(int device, float cell_width, uint32_t block_capacity, spk_ctx** out) {
    if (!out || !(cell_width > 0.f)) return SPK_EINVAL;
    *out = nullptr;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 9) return (int)cudaErrorInvalidDevice; // TMA bulk copies need sm_90+; the library is built for sm_100a only
    spk_ctx* c = new (std::nothrow) spk_ctx();
    if (!c) return (int)cudaErrorMemoryAllocation;
    std::memset(c, 0, sizeof(*c));
    c->device = device;
    c->cell_width = cell_width;
    c->sm_count = prop.multiProcessorCount;
    c->chunk = SPK_G2P2G_CHUNK;
#if SPK_DIM == 3
    // 2 (default): cell-coherent kernel, 512 particles per dispatch entry; 1: round-1 cell-binned kernel; 0: per-particle shared atomics
    // (the last two with 256-particle entries; kept for A/B profiling, SPK_G2P2G_MODE)
    c->g2p2g_mode = 2;
    if (const char* m = getenv("SPK_G2P2G_MODE")) { const int v = atoi(m); c->g2p2g_mode = v < 0 ? 0 : (v > 2 ? 2 : v); }
    if (c->g2p2g_mode == 2) c->chunk = SPK_CELL_THREADS;
    c->g2p2g_specialize = getenv("SPK_G2P2G_GENERIC") ? 0 : 1;
    // persistent CTAs that prefetch their next entry (kept for A/B: SPK_G2P2G_STREAM=1).  Measured slower on the B200 (0.88 vs 0.68 ms on the 4 M
    // sand scene): the two CTAs of an SM fall into step, so their LSU-bound P2G phases and their FMA-bound update phases no longer
    // overlap, and the entry loop costs ~5 live registers the 64-register body does not have (profiles/r2_g2p2g_experiments.md)
    c->g2p2g_prefetch = 0; // measured: 0.630 vs 0.624 ms with it on (the waiting threads were not idle issue slots after all)
    if (const char* m = getenv("SPK_G2P2G_PREFETCH")) c->g2p2g_prefetch = atoi(m) ? 1 : 0;
    c->g2p2g_stream = 0;
    if (const char* m = getenv("SPK_G2P2G_STREAM")) c->g2p2g_stream = atoi(m) ? 1 : 0;
#else
    c->g2p2g_mode = 1;
#endif
    {   // the main stream outranks the side stream (grid update, grid reset, multi-GPU exchange): where both have CTAs ready — the side
        // stream's bandwidth-bound kernels next to the sort's binning pass — the block scheduler serves the critical path first and the
        // side work fills what the sort's small latency-bound kernels leave idle (SPK_STREAM_PRIORITY=0: all streams equal, A/B)
        int least = 0, greatest = 0;
        CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        const bool prio = !(getenv("SPK_STREAM_PRIORITY") && atoi(getenv("SPK_STREAM_PRIORITY")) == 0);
        CK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio ? greatest : least));
    }
    CK(cudaStreamCreateWithFlags(&c->halo_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->gu_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_transfer_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_gu_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_tables, cudaEventDisableTiming));
    c->overlap_grid_update = getenv("SPK_NO_GU_OVERLAP") ? 0 : 1;
    CK(cudaEventCreateWithFlags(&c->up_ready, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->up_consumed, cudaEventDisableTiming));
    for (auto& e : c->ev) CK(cudaEventCreate(&e));
    CK(cudaEventCreateWithFlags(&c->ev_counters, cudaEventDisableTiming));
    CK(cudaMalloc((void**)&c->d_counters, sizeof(DeviceCounters)));
    CK(cudaMemset(c->d_counters, 0, sizeof(DeviceCounters)));
    CK(cudaMalloc((void**)&c->dist_bbox, 24 * sizeof(uint32_t))); // bounding box of the active blocks + Morton boxes (spk_sort.cu)
    CK(cudaMemset(c->dist_bbox, 0, 24 * sizeof(uint32_t)));
    CK(cudaHostAlloc((void**)&c->h_counters, sizeof(DeviceCounters), cudaHostAllocMapped));
    std::memset(c->h_counters, 0, sizeof(DeviceCounters));
    CK(cudaHostGetDevicePointer((void**)&c->h_counters_dev, c->h_counters, 0));
    CK(cudaHostAlloc((void**)&c->h_upload, 64, cudaHostAllocDefault));
    // CudaSparseGrid::with_capacity: next_power_of_two().max(1024) blocks (cuda_sparse_grid.rs:53-54); the hash map holds
    // twice as many slots (50 % maximum load factor, cuda_sparse_grid.rs:219)
    const uint32_t cap_hash = std::max<uint32_t>(next_pow2(block_capacity), 1024u) * 2u;
    CK(alloc_grid(c, c->grid[0], cap_hash));
    CK(alloc_grid(c, c->grid[1], cap_hash));
    CK(ensure_slot_scratch(c, cap_hash));
    CK(ensure_staging(c, 1 << 20));
    CK(alloc_particles(c, 1024));
    CK(ensure_dispatch(c));
    CK(cudaStreamSynchronize(c->stream));
    *out = c;
    return SPK_OK;
}

int spk_ctx_destroy(spk_ctx* c) {
    if (!c) return SPK_EINVAL;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    spk_dist_detach(c);
    if (c->dist_arena) { c->grid[0].external_hash = c->grid[1].external_hash = 1; }
    free_particles(c);
    free_grid(c->grid[0]); free_grid(c->grid[1]);
    dev_free(c->slot_cnt); dev_free(c->cell_cnt); dev_free(c->mslot); dev_free(c->dispatch); dev_free(c->dispatch_nbr); dev_free(c->models); dev_free(c->colliders); dev_free(c->collider_data); dev_free(c->d_counters);
    if (c->staging) cudaFree(c->staging);
    cudaStreamSynchronize(c->halo_stream);
    for (int k = 0; k < 8; ++k) {
        if (c->async_buf[k]) cudaFree(c->async_buf[k]);
        if (c->async_ready[k]) { cudaEventDestroy(c->async_ready[k]); cudaEventDestroy(c->async_done[k]); }
    }
    if (c->dist_arena) cudaFree(c->dist_arena);
    if (c->dist_peer_ids) cudaFree(c->dist_peer_ids);
    if (c->dist_push_ptrs) cudaFree(c->dist_push_ptrs);
    if (c->dist_bbox) cudaFree(c->dist_bbox);
    if (c->mg_scratch) cudaFree(c->mg_scratch);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->h_upload) cudaFreeHost(c->h_upload);
    for (auto& e : c->ev) cudaEventDestroy(e);
    cudaEventDestroy(c->ev_counters);
    if (c->kev[0]) for (auto& e : c->kev) cudaEventDestroy(e);
    cudaStreamSynchronize(c->h2d_stream);
    if (c->up_staging) cudaFree(c->up_staging);
    cudaEventDestroy(c->up_ready); cudaEventDestroy(c->up_consumed);
    cudaStreamDestroy(c->stream); cudaStreamDestroy(c->halo_stream); cudaStreamDestroy(c->h2d_stream); cudaStreamDestroy(c->gu_stream);
    cudaEventDestroy(c->ev_transfer_done); cudaEventDestroy(c->ev_gu_done); cudaEventDestroy(c->ev_tables);
    delete c;
    return SPK_OK;
}

int spk_ctx_cell_width(const spk_ctx* c, float* out) { if (!c || !out) return SPK_EINVAL; *out = c->cell_width; return SPK_OK; }
int spk_ctx_stream(spk_ctx* c, void** out) { if (!c || !out) return SPK_EINVAL; *out = (void*)c->stream; return SPK_OK; }
int spk_ctx_synchronize(spk_ctx* c) {
    if (!c) return SPK_EINVAL;
    CK(cudaSetDevice(c->device));
    { int jrc = spk_join_grid_update(c); if (jrc) return jrc; }
    CK(cudaStreamSynchronize(c->stream));
    return c->h_counters->pad[1] ? SPK_EPEER : SPK_OK;
}

int spk_particles_upload(spk_ctx* c, const SpkParticleStatus* status, const SpkParticlePosition* pos, const SpkParticleVelocity* vel,
                         const SpkParticleVolume* volume, const SpkParticlePhase* phase, const SpkParticleData* data, size_t n) {
    if (!c || (n && (!status || !pos || !vel || !volume || !phase))) return SPK_EINVAL;
    if (n >= 0xffffffffull) return SPK_EOVERFLOW;
    CK(cudaSetDevice(c->device));
    if (n > c->cap) CK(alloc_particles(c, n + n / 8 + 1024));
    dev_free(c->gid); // a new particle set: global ids are the identity again (spk_particles_set_gids)
    c->n = n;
    c->cur = 0;
    c->sorted_valid = false;
    c->dt_cache_valid = false;
    // size the sparse grid for the particle count up front: 2x2x2 particles per cell, 4x4x4 cells per block, 50 % load
    // (the reference's own estimate, cuda_sparse_grid.rs:222-229), doubled for the shell of neighbour blocks
    const uint32_t want = std::max<uint32_t>(next_pow2((uint64_t)n * 100 / (SPK_CELLS_PER_BLOCK * SPK_PARTICLES_PER_CELL * 50)) * 4u, 2048u);
    GridBuffers& next = c->grid[1 - c->gcur];
    GridBuffers& curr = c->grid[c->gcur];
    if (c->dist_arena && want > c->dist_cap_hash) return SPK_EOVERFLOW; // the arena shared with the peers cannot grow
    if (next.cap_hash < want) { CK(alloc_grid(c, next, want)); CK(ensure_slot_scratch(c, want)); }
    if (curr.cap_hash < want && curr.nab_host == 0) CK(alloc_grid(c, curr, want)); // empty: nothing to preserve
    CK(ensure_dispatch(c));
    return upload_range(c, status, pos, vel, volume, phase, data, 0, n);
}

// ---- asynchronous upload: stage (host -> device copy on its own stream) and commit (repack on the main stream) ------------
int spk_particles_stage_async(spk_ctx* c, const SpkParticleStatus* status, const SpkParticlePosition* pos, const SpkParticleVelocity* vel,
                              const SpkParticleVolume* volume, const SpkParticlePhase* phase, const SpkParticleData* data, size_t n) {
    if (!c || (n && (!status || !pos || !vel || !volume || !phase))) return SPK_EINVAL;
    if (n >= 0xffffffffull) return SPK_EOVERFLOW;
    CK(cudaSetDevice(c->device));
    const size_t per = sizeof(SpkParticleStatus) + sizeof(SpkParticlePosition) + sizeof(SpkParticleVelocity) + sizeof(SpkParticleVolume) +
                       sizeof(SpkParticlePhase) + sizeof(SpkParticleData);
    if (c->up_staging_bytes < n * per + 256) {
        CK(cudaStreamSynchronize(c->stream)); CK(cudaStreamSynchronize(c->h2d_stream));
        if (c->up_staging) cudaFree(c->up_staging);
        c->up_staging = nullptr; c->up_staging_bytes = 0;
        CK(cudaMalloc(&c->up_staging, n * per + 256));
        c->up_staging_bytes = n * per + 256;
    }
    // the repack kernel of the previous commit must have read the staging buffer before it is overwritten
    if (c->up_inflight) CK(cudaStreamWaitEvent(c->h2d_stream, c->up_consumed, 0));
    char* cur = (char*)c->up_staging;
    size_t* off = c->up_off;
    const void* src[6] = {status, pos, vel, volume, phase, data};
    const size_t bytes[6] = {n * sizeof(SpkParticleStatus), n * sizeof(SpkParticlePosition), n * sizeof(SpkParticleVelocity), n * sizeof(SpkParticleVolume),
                             n * sizeof(SpkParticlePhase), n * sizeof(SpkParticleData)};
    for (int k = 0; k < 6; ++k) {
        off[k] = (size_t)(cur - (char*)c->up_staging);
        if (src[k] && bytes[k]) CK(cudaMemcpyAsync(cur, src[k], bytes[k], cudaMemcpyHostToDevice, c->h2d_stream));
        cur += (bytes[k] + 15) & ~size_t(15);
    }
    CK(cudaEventRecord(c->up_ready, c->h2d_stream));
    c->up_staged_n = n;
    c->up_staged_data = data != nullptr;
    c->up_staged = true;
    return SPK_OK;
}

int spk_particles_commit_staged(spk_ctx* c) {
    if (!c) return SPK_EINVAL;
    if (!c->up_staged) return SPK_ESTATE;
    CK(cudaSetDevice(c->device));
    const size_t n = c->up_staged_n;
    const uint32_t want = std::max<uint32_t>(next_pow2((uint64_t)n * 100 / (SPK_CELLS_PER_BLOCK * SPK_PARTICLES_PER_CELL * 50)) * 4u, 2048u);
    if (c->dist_arena && want > c->dist_cap_hash) return SPK_EOVERFLOW; // checked before anything changes: the staged upload stays staged
    if (n > c->cap) { CK(cudaStreamSynchronize(c->stream)); CK(alloc_particles(c, n + n / 8 + 1024)); }
    if (c->gid) { CK(cudaStreamSynchronize(c->stream)); dev_free(c->gid); }
    c->n = n;
    c->cur = 0;
    c->sorted_valid = false;
    c->dt_cache_valid = false;
    GridBuffers& next = c->grid[1 - c->gcur];
    if (next.cap_hash < want) { CK(alloc_grid(c, next, want)); CK(ensure_slot_scratch(c, want)); }
    CK(ensure_dispatch(c));
    c->up_staged = false;
    if (n == 0) return SPK_OK;
    char* base = (char*)c->up_staging;
    CK(cudaStreamWaitEvent(c->stream, c->up_ready, 0));
    pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        c->pb[c->cur], c->vel, c->cold, (const SpkParticleStatus*)(base + c->up_off[0]), (const SpkParticlePosition*)(base + c->up_off[1]),
        (const SpkParticleVelocity*)(base + c->up_off[2]), (const SpkParticleVolume*)(base + c->up_off[3]), (const SpkParticlePhase*)(base + c->up_off[4]),
        c->up_staged_data ? (const SpkParticleData*)(base + c->up_off[5]) : nullptr, 0, n, &c->d_counters->pad[0]);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->up_consumed, c->stream));
    c->up_inflight = true;
    return SPK_OK; // no synchronisation: a bad model index is reported by the next sort (it travels in the counter record)
}

int spk_particles_upload_async(spk_ctx* c, const SpkParticleStatus* status, const SpkParticlePosition* pos, const SpkParticleVelocity* vel,
                               const SpkParticleVolume* volume, const SpkParticlePhase* phase, const SpkParticleData* data, size_t n) {
    int rc = spk_particles_stage_async(c, status, pos, vel, volume, phase, data, n);
    return rc ? rc : spk_particles_commit_staged(c);
}

int spk_particles_len(const spk_ctx* c, size_t* out) { if (!c || !out) return SPK_EINVAL; *out = c->n; return SPK_OK; }

int spk_particles_read(spk_ctx* c, int which, void* dst, size_t n) {
    if (!c || (n && !dst) || n > c->n) return SPK_EINVAL;
    const size_t sz = attr_size(which);
    if (!sz) return SPK_EINVAL;
    if (n == 0) return SPK_OK;
    CK(cudaSetDevice(c->device));
    CK(ensure_staging(c, c->n * sz));
    const uint32_t* perm = (which == SPK_ATTR_SORTED_IDS && c->sorted_valid) ? c->perm : nullptr;
    unpack_kernel<<<(unsigned)((c->n + 255) / 256), 256, 0, c->stream>>>(c->pb[c->cur], c->vel, c->cold, perm, which, c->staging, c->n);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(dst, c->staging, n * sz, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SPK_OK;
}

int spk_particles_read_async(spk_ctx* c, int which, void* dst, size_t n) {
    if (!c || (n && !dst) || n > c->n || which < 0 || which >= 8) return SPK_EINVAL;
    const size_t sz = attr_size(which);
    if (!sz) return SPK_EINVAL;
    if (n == 0) return SPK_OK;
    CK(cudaSetDevice(c->device));
    if (!c->async_ready[which]) {
        CK(cudaEventCreateWithFlags(&c->async_ready[which], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->async_done[which], cudaEventDisableTiming));
    }
    if (c->async_bytes[which] < c->n * sz) {
        if (c->async_inflight[which]) CK(cudaEventSynchronize(c->async_done[which]));
        if (c->async_buf[which]) cudaFree(c->async_buf[which]);
        c->async_buf[which] = nullptr;
        c->async_bytes[which] = 0;
        CK(cudaMalloc(&c->async_buf[which], c->n * sz));
        c->async_bytes[which] = c->n * sz;
    }
    // the previous copy out of this staging buffer must have finished before it is overwritten
    if (c->async_inflight[which]) CK(cudaStreamWaitEvent(c->stream, c->async_done[which], 0));
    const uint32_t* perm = (which == SPK_ATTR_SORTED_IDS && c->sorted_valid) ? c->perm : nullptr;
    unpack_kernel<<<(unsigned)((c->n + 255) / 256), 256, 0, c->stream>>>(c->pb[c->cur], c->vel, c->cold, perm, which, c->async_buf[which], c->n);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->async_ready[which], c->stream));
    CK(cudaStreamWaitEvent(c->halo_stream, c->async_ready[which], 0));
    CK(cudaMemcpyAsync(dst, c->async_buf[which], n * sz, cudaMemcpyDeviceToHost, c->halo_stream));
    CK(cudaEventRecord(c->async_done[which], c->halo_stream));
    c->async_inflight[which] = true;
    return SPK_OK;
}

int spk_particles_read_wait(spk_ctx* c) {
    if (!c) return SPK_EINVAL;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->halo_stream));
    for (auto& f : c->async_inflight) f = false;
    return SPK_OK;
}

int spk_particles_append(spk_ctx* c, const SpkParticleStatus* status, const SpkParticlePosition* pos, const SpkParticleVelocity* vel,
                         const SpkParticleVolume* volume, const SpkParticlePhase* phase, const SpkParticleData* data, size_t n) {
    if (!c || (n && (!status || !pos || !vel || !volume || !phase))) return SPK_EINVAL;
    if (n == 0) return SPK_OK;
    CK(cudaSetDevice(c->device));
    const size_t old_n = c->n, new_n = old_n + n;
    if (new_n >= 0xffffffffull) return SPK_EOVERFLOW;
    if (new_n > c->cap) {
        // grow: read everything back in id order, reallocate, upload again (cold path; CudaVec::append reallocates too)
        std::vector<SpkParticleStatus> s(new_n); std::vector<SpkParticlePosition> p(new_n); std::vector<SpkParticleVelocity> v(new_n);
        std::vector<SpkParticleVolume> vo(new_n); std::vector<SpkParticlePhase> ph(new_n); std::vector<SpkParticleData> d(new_n);
        int rc;
        if ((rc = spk_particles_read(c, SPK_ATTR_STATUS, s.data(), old_n))) return rc;
        if ((rc = spk_particles_read(c, SPK_ATTR_POSITION, p.data(), old_n))) return rc;
        if ((rc = spk_particles_read(c, SPK_ATTR_VELOCITY, v.data(), old_n))) return rc;
        if ((rc = spk_particles_read(c, SPK_ATTR_VOLUME, vo.data(), old_n))) return rc;
        if ((rc = spk_particles_read(c, SPK_ATTR_PHASE, ph.data(), old_n))) return rc;
        if ((rc = spk_particles_read(c, SPK_ATTR_DATA, d.data(), old_n))) return rc;
        std::copy(status, status + n, s.begin() + old_n); std::copy(pos, pos + n, p.begin() + old_n); std::copy(vel, vel + n, v.begin() + old_n);
        std::copy(volume, volume + n, vo.begin() + old_n); std::copy(phase, phase + n, ph.begin() + old_n);
        for (size_t i = 0; i < n; ++i) d[old_n + i] = data ? data[i] : SpkParticleData{-0.01f, 1.0f, 1.0f, 0.0f};
        return spk_particles_upload(c, s.data(), p.data(), v.data(), vo.data(), ph.data(), d.data(), new_n);
    }
    if (c->gid) { CK(cudaStreamSynchronize(c->stream)); dev_free(c->gid); } // host-side edits reset the global ids to the identity
    c->n = new_n;
    c->sorted_valid = false;
    c->dt_cache_valid = false;
    CK(ensure_dispatch(c));
    return upload_range(c, status, pos, vel, volume, phase, data, old_n, n);
}

int spk_particles_remove_range(spk_ctx* c, size_t begin, size_t end) {
    if (!c || begin > end || end > c->n) return SPK_EINVAL;
    if (begin == end) return SPK_OK;
    const size_t old_n = c->n;
    std::vector<SpkParticleStatus> s(old_n); std::vector<SpkParticlePosition> p(old_n); std::vector<SpkParticleVelocity> v(old_n);
    std::vector<SpkParticleVolume> vo(old_n); std::vector<SpkParticlePhase> ph(old_n); std::vector<SpkParticleData> d(old_n);
    int rc;
    if ((rc = spk_particles_read(c, SPK_ATTR_STATUS, s.data(), old_n))) return rc;
    if ((rc = spk_particles_read(c, SPK_ATTR_POSITION, p.data(), old_n))) return rc;
    if ((rc = spk_particles_read(c, SPK_ATTR_VELOCITY, v.data(), old_n))) return rc;
    if ((rc = spk_particles_read(c, SPK_ATTR_VOLUME, vo.data(), old_n))) return rc;
    if ((rc = spk_particles_read(c, SPK_ATTR_PHASE, ph.data(), old_n))) return rc;
    if ((rc = spk_particles_read(c, SPK_ATTR_DATA, d.data(), old_n))) return rc;
    s.erase(s.begin() + begin, s.begin() + end); p.erase(p.begin() + begin, p.begin() + end); v.erase(v.begin() + begin, v.begin() + end);
    vo.erase(vo.begin() + begin, vo.begin() + end); ph.erase(ph.begin() + begin, ph.begin() + end); d.erase(d.begin() + begin, d.begin() + end);
    return spk_particles_upload(c, s.data(), p.data(), v.data(), vo.data(), ph.data(), d.data(), s.size());
}

// CudaParticleSet::truncate / capacity, cuda_particle_set.rs:77-87
int spk_particles_truncate(spk_ctx* c, size_t new_len) {
    if (!c) return SPK_EINVAL;
    if (new_len >= c->n) return SPK_OK; // CudaVec::truncate: a no-op when the set is not longer than new_len
    return spk_particles_remove_range(c, new_len, c->n);
}
int spk_particles_capacity(const spk_ctx* c, size_t* out) {
    if (!c || !out) return SPK_EINVAL;
    *out = c->cap;
    return SPK_OK;
}

int spk_models_upload(spk_ctx* c, const SpkGpuParticleModel* models, size_t n) {
    if (!c || (n && !models)) return SPK_EINVAL;
    CK(cudaSetDevice(c->device));
    std::vector<DevModel> dm(n);
    for (size_t i = 0; i < n; ++i) {
        const SpkGpuParticleModel& m = models[i];
        DevModel& d = dm[i];
        std::memset(&d, 0, sizeof(d));
        d.ctag = m.constitutive_tag; d.ptag = m.plastic_tag; d.ftag = m.failure_tag;
        switch (m.constitutive_tag) {
        case SPK_CONSTITUTIVE_COROTATED:
            d.c[0] = m.constitutive.corotated.m.cfl_coeff; d.c[1] = m.constitutive.corotated.m.lambda; d.c[2] = m.constitutive.corotated.m.mu;
            if (m.constitutive.corotated.m.split_stress_on_failure) d.flags |= 1u;
            break;
        case SPK_CONSTITUTIVE_NEO_HOOKEAN:
            d.c[0] = m.constitutive.neo_hookean.m.cfl_coeff; d.c[1] = m.constitutive.neo_hookean.m.lambda; d.c[2] = m.constitutive.neo_hookean.m.mu;
            break;
        case SPK_CONSTITUTIVE_MONAGHAN: {
            d.c[0] = m.constitutive.monaghan.m.pressure0;
            uint32_t g = m.constitutive.monaghan.m.gamma;
            std::memcpy(&d.c[1], &g, 4);
            d.c[2] = m.constitutive.monaghan.m.viscosity; d.c[3] = m.constitutive.monaghan.m.max_neg_pressure;
            break;
        }
        case SPK_CONSTITUTIVE_CUSTOM: { // zero stress, no dt bound in the default updater (gpu_constitutive_model.rs:83,108)
            const uint32_t id = m.constitutive.custom.id; // kept (as bits) for custom updaters: GpuConstitutiveModel::Custom(id)
            std::memcpy(&d.c[0], &id, 4);
            break;
        }
        default: return SPK_EINVAL;
        }
        switch (m.plastic_tag) {
        case SPK_PLASTIC_DRUCKER_PRAGER: {
            const auto& p = m.plastic.drucker_prager.m;
            d.p[0] = p.h0; d.p[1] = p.h1; d.p[2] = p.h2; d.p[3] = p.h3; d.p[4] = p.lambda; d.p[5] = p.mu; d.p[6] = p.volume_correction;
            if (p.only_active_when_failed) d.flags |= 2u;
            break;
        }
        case SPK_PLASTIC_NACC: {
            const auto& p = m.plastic.nacc.m;
            d.p[0] = p.mu; d.p[1] = p.kappa; d.p[2] = p.hardening_factor; d.p[3] = p.cohesion; d.p[4] = p.friction;
            if (p.hardening_enabled) d.flags |= 4u;
            break;
        }
        case SPK_PLASTIC_RANKINE: {
            const auto& p = m.plastic.rankine.m;
            d.p[0] = p.mu; d.p[1] = p.lambda; d.p[2] = p.tensile_strength; d.p[3] = p.softening_rate;
            break;
        }
        case SPK_PLASTIC_SNOW: {
            const auto& p = m.plastic.snow.m;
            d.p[0] = p.min_epsilon; d.p[1] = p.max_epsilon; d.p[2] = p.hardening_coeff;
            break;
        }
        case SPK_PLASTIC_CUSTOM: case SPK_PLASTIC_NONE: break;
        default: return SPK_EINVAL;
        }
        switch (m.failure_tag) {
        case SPK_FAILURE_MAXIMUM_STRESS: d.f[0] = m.failure.maximum_stress.max_principal_stress; d.f[1] = m.failure.maximum_stress.max_shear_stress; break;
        case SPK_FAILURE_CUSTOM: case SPK_FAILURE_NONE: break;
        default: return SPK_EINVAL;
        }
    }
    dev_free(c->models);
    CK(dev_alloc(c->models, n));
    if (n) CK(cudaMemcpy(c->models, dm.data(), n * sizeof(DevModel), cudaMemcpyHostToDevice));
    c->nmodels = n;
    c->model_mask = 0;
    for (size_t i = 0; i < n; ++i) c->model_mask |= (1u << dm[i].ctag) | (1u << (4 + dm[i].ptag)) | (dm[i].ftag == SPK_FAILURE_MAXIMUM_STRESS ? (1u << 10) : 0u); // SPK_MM_*
    c->dt_cache_valid = false; // the cached CFL candidate was computed with the previous models
    return SPK_OK;
}

int spk_colliders_upload(spk_ctx* c, const SpkGpuCollider* cols, size_t n) {
    if (!c || (n && !cols)) return SPK_EINVAL;
    CK(cudaSetDevice(c->device));
    std::vector<DevCollider> dc(n);
    std::vector<float> data; // all heightfield matrices back to back
    std::vector<size_t> data_off(n, 0);
    for (size_t i = 0; i < n; ++i) {
        DevCollider& d = dc[i];
        std::memset(&d, 0, sizeof(d));
        if (cols[i].grid_boundary_handling > SPK_BOUNDARY_NONE) return SPK_EINVAL;
        d.shape = cols[i].shape_tag;
#if SPK_DIM == 2
        if (cols[i].shape_tag == SPK_SHAPE_CUBOID) {
            std::memcpy(d.half_extents, cols[i].shape.cuboid.half_extents, 8);
        } else if (cols[i].shape_tag == SPK_SHAPE_POLYLINE) {
            const auto& pl = cols[i].shape.polyline;
            if (pl.num_vertices && !pl.vertices) return SPK_EINVAL;
            d.nrows = pl.num_vertices;
            d.hf_flags = pl.flip_interior ? 2u : 0u;
            data_off[i] = data.size();
            data.insert(data.end(), pl.vertices, pl.vertices + (size_t)pl.num_vertices * 2);
        } else if (cols[i].shape_tag == SPK_SHAPE_HEIGHTFIELD) {
            const auto& hf = cols[i].shape.heightfield;
            if (!hf.heights || hf.num_heights < 2) return SPK_EINVAL; // HeightField::new asserts > 1 samples
            d.nrows = hf.num_heights;
            std::memcpy(d.scale, hf.scale, 8);
            d.hf_flags = (hf.bellow_heightfield_is_solid ? 1u : 0u) | (hf.flip_interior ? 2u : 0u);
            data_off[i] = data.size();
            data.insert(data.end(), hf.heights, hf.heights + hf.num_heights);
        } else if (cols[i].shape_tag == SPK_SHAPE_TRIMESH) {
            const auto& tm = cols[i].shape.trimesh;
            if (!tm.vertices || !tm.indices || tm.num_vertices == 0 || tm.num_triangles == 0) return SPK_EINVAL;
            for (size_t k = 0; k < (size_t)tm.num_triangles * 3; ++k)
                if (tm.indices[k] >= tm.num_vertices) return SPK_EINVAL;
            d.nrows = tm.num_vertices; d.ncols = tm.num_triangles;
            d.hf_flags = tm.flip_interior ? 2u : 0u;
            data_off[i] = data.size();
            data.insert(data.end(), tm.vertices, tm.vertices + (size_t)tm.num_vertices * 2); // [vertices | indices as raw bits]
            for (size_t k = 0; k < (size_t)tm.num_triangles * 3; ++k) { float bits; std::memcpy(&bits, &tm.indices[k], 4); data.push_back(bits); }
        } else {
            return SPK_EUNSUPPORTED;
        }
        std::memcpy(d.rot, cols[i].rotation, 8);
        std::memcpy(d.tr, cols[i].translation, 8);
#else
        if (cols[i].shape_tag == SPK_SHAPE_CUBOID) {
            std::memcpy(d.half_extents, cols[i].shape.cuboid.half_extents, 12);
        } else if (cols[i].shape_tag == SPK_SHAPE_HEIGHTFIELD) {
            const auto& hf = cols[i].shape.heightfield;
            if (!hf.heights || hf.nrows < 2 || hf.ncols < 2) return SPK_EINVAL; // HeightField::new asserts > 1 samples per axis
            d.nrows = hf.nrows; d.ncols = hf.ncols;
            std::memcpy(d.scale, hf.scale, 12);
            d.hf_flags = (hf.bellow_heightfield_is_solid ? 1u : 0u) | (hf.flip_interior ? 2u : 0u);
            data_off[i] = data.size();
            data.insert(data.end(), hf.heights, hf.heights + (size_t)hf.nrows * hf.ncols);
        } else if (cols[i].shape_tag == SPK_SHAPE_POLYLINE) {
            const auto& pl = cols[i].shape.polyline;
            if (pl.num_vertices && !pl.vertices) return SPK_EINVAL;
            d.nrows = pl.num_vertices; // number of vertices
            d.hf_flags = pl.flip_interior ? 2u : 0u;
            data_off[i] = data.size();
            data.insert(data.end(), pl.vertices, pl.vertices + (size_t)pl.num_vertices * 3);
        } else if (cols[i].shape_tag == SPK_SHAPE_TRIMESH) {
            const auto& tm = cols[i].shape.trimesh;
            if (!tm.vertices || !tm.indices || tm.num_vertices == 0 || tm.num_triangles == 0) return SPK_EINVAL;
            for (size_t k = 0; k < (size_t)tm.num_triangles * 3; ++k)
                if (tm.indices[k] >= tm.num_vertices) return SPK_EINVAL;
            d.nrows = tm.num_vertices; d.ncols = tm.num_triangles;
            d.hf_flags = tm.flip_interior ? 2u : 0u;
            data_off[i] = data.size();
            data.insert(data.end(), tm.vertices, tm.vertices + (size_t)tm.num_vertices * 3); // [vertices | indices as raw bits]
            for (size_t k = 0; k < (size_t)tm.num_triangles * 3; ++k) { float bits; std::memcpy(&bits, &tm.indices[k], 4); data.push_back(bits); }
        } else {
            return SPK_EUNSUPPORTED;
        }
        std::memcpy(d.rot, cols[i].rotation, 16);
        std::memcpy(d.tr, cols[i].translation, 12);
#endif
        d.friction = cols[i].friction;
        d.handling = cols[i].grid_boundary_handling;
    }
    CK(cudaStreamSynchronize(c->stream));
    dev_free(c->colliders);
    dev_free(c->collider_data);
    CK(dev_alloc(c->colliders, n));
    CK(dev_alloc(c->collider_data, data.size()));
    if (!data.empty()) CK(cudaMemcpy(c->collider_data, data.data(), data.size() * sizeof(float), cudaMemcpyHostToDevice));
    for (size_t i = 0; i < n; ++i)
        if (dc[i].shape == SPK_SHAPE_HEIGHTFIELD || dc[i].shape == SPK_SHAPE_POLYLINE || dc[i].shape == SPK_SHAPE_TRIMESH) dc[i].heights = c->collider_data + data_off[i];
    if (n) CK(cudaMemcpy(c->colliders, dc.data(), n * sizeof(DevCollider), cudaMemcpyHostToDevice));
    c->ncolliders = n;
    // colliders changed: invalidate the per-node projection cache of both grids (the reference uploads colliders once)
    for (int g = 0; g < 2; ++g) CK(cudaMemsetAsync(c->grid[g].proj, 0, (size_t)c->grid[g].cap_blocks * SPK_CELLS_PER_BLOCK * 4, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SPK_OK;
}

// ------------------------------------------------------------------------------------------------ custom particle kernels
int spk_set_particle_kernels_launcher(spk_ctx* c, const SpkParticleKernelsLauncher* l) {
    if (!c) return SPK_EINVAL;
    if (l) c->launcher = *l;
    else std::memset(&c->launcher, 0, sizeof(c->launcher));
    c->dt_cache_valid = false; // the cached CFL candidate was produced by the previous updater
    return SPK_OK;
}
int spk_default_launch_estimate_particle_timestep_length(spk_ctx* c, float min_dt, float max_dt) {
    if (!c) return SPK_EINVAL;
    CK(cudaSetDevice(c->device));
    return (int)spk_launch_estimate_dt(c, min_dt, max_dt);
}
int spk_default_launch_g2p2g(spk_ctx* c, float dt) {
    if (!c) return SPK_EINVAL;
    if (!c->sorted_valid) return SPK_ESTATE;
    CK(cudaSetDevice(c->device));
    return (int)spk_launch_g2p2g(c, dt);
}

// ------------------------------------------------------------------------------------------------ stages
int spk_frame_begin(spk_ctx* c, const SpkSolverParameters* params) {
    if (!c || !params) return SPK_EINVAL;
    c->dt_acc_host = host_dt_from_sec(params->dt); // timestep_length.copy_from(from_sec(params.dt)), cuda_mpm_pipeline.rs:294-299
    return SPK_OK;
}

int spk_stage_sort(spk_ctx* c, SpkSortInfo* info) {
    if (!c) return SPK_EINVAL;
    CK(cudaSetDevice(c->device));
    if (c->n && !c->models) return SPK_ESTATE;
    for (int attempt = 0; attempt < 16; ++attempt) {
        GridBuffers& g = c->grid[1 - c->gcur];
        CK(ensure_staging(c, ((size_t)g.cap_hash * 4 / 256 + 1) * 12)); // the scan partials of spk_launch_sort_bin: 12 bytes per scan tile (>= 256 entries of the Morton table, 4 per hash slot)
        CK(spk_launch_sort_bin(c));
        // The counter record (block / dispatch counts, overflow flag, dt candidate) is published to mapped host memory, then
        // the two kernels that finish the sort are enqueued for the grid's CAPACITY (they read the true block count on the
        // device), and only then does the host wait for the record: its wake-up and the launches of the following stages
        // overlap ~30 us of GPU work instead of leaving the GPU idle in the middle of every substep.
        // (multi-GPU: the record carries the dt minimum over the ranks, which the wait half of the last barrier — left on the grid-update
        // stream by spk_step — writes: the main stream meets that stream here, after the binning kernels ran next to it)
        if (spk_dist_active(c)) { int jrc = spk_join_grid_update(c); if (jrc) return jrc; }
        int rc = publish_counters(c);
        if (rc) return rc;
        CK(spk_launch_sort_finish(c));
        if ((rc = wait_counters(c))) return rc;
        const DeviceCounters& hc = *c->h_counters;
        if (hc.pad[0]) { // spk_particles_upload_async found a model index that does not fit the flags word
            CK(cudaMemsetAsync(&c->d_counters->pad[0], 0, 4, c->stream));
            return SPK_EINVAL;
        }
        // grow-and-retry until the load factor is <= 50 % (cuda_sparse_grid.rs:162-253)
        if ((hc.overflow || hc.nab > g.cap_blocks) && c->dist_arena) return SPK_EOVERFLOW; // the shared arena cannot grow
        if (hc.overflow || hc.nab > g.cap_blocks) {
            // jump straight to the size the particle count suggests (the reference's estimate, cuda_sparse_grid.rs:222-233)
            const uint64_t estimate = (uint64_t)next_pow2((uint64_t)c->n * 100 / (SPK_CELLS_PER_BLOCK * SPK_PARTICLES_PER_CELL * 50)) * 4u;
            const uint64_t want = std::max<uint64_t>({(uint64_t)g.cap_hash * 2, (uint64_t)next_pow2((uint64_t)hc.nab * 4), std::min<uint64_t>(estimate, 1ull << 26)});
            if (want > (1ull << 30)) return SPK_EOVERFLOW;
            CK(alloc_grid(c, g, (uint32_t)want));
            CK(ensure_slot_scratch(c, (uint32_t)want));
            CK(ensure_dispatch(c));
            continue;
        }
        g.nab_host = hc.nab;
        c->sorted_valid = true;
        // the fused dt candidate of the last g2p2g travelled in this record; the device word has since been reset for the next g2p2g
        // (build_tables_kernel), so the host keeps the value until that g2p2g runs
        c->dt_raw_host = hc.dt_raw_bits;
        c->dt_raw_host_valid = true;
        if (info) {
            info->num_active_blocks = hc.nab;
            info->num_dispatch_blocks = hc.ndispatch;
            info->num_dispatch_halo_blocks = 0;
            info->hashmap_capacity = g.cap_hash;
        }
        return SPK_OK;
    }
    return SPK_EOVERFLOW;
}

int spk_stage_reset_grid(spk_ctx* c) {
    if (!c) return SPK_EINVAL;
    CK(cudaSetDevice(c->device));
    { int jrc = spk_join_grid_update(c); if (jrc) return jrc; }
    CK(spk_launch_reset_grid(c));
    return SPK_OK;
}

int spk_stage_estimate_dt(spk_ctx* c, float min_dt, float max_dt, float* dt_out) {
    if (!c) return SPK_EINVAL;
    CK(cudaSetDevice(c->device));
    // A registered estimate launcher is called EVERY substep, like the reference's step_generic does (cuda_mpm_pipeline.rs:381-392); the
    // candidate fused into g2p2g replaces only the DEFAULT estimate kernel (its result is bit-identical, see below)
    if (c->dt_cache_valid && !c->launcher.launch_estimate_particle_timestep_length) {
        // K15 was fused into the previous g2p2g (the particle state has not changed since): the reduced un-clamped
        // candidate arrived with the counters record; apply timestep.rs:87-112 to it.  min / clamp / fixed-point
        // conversion are monotone, so min_p f(c_p) == f(min_p c_p), bit for bit.
        if (!c->dt_raw_host_valid && !c->counters_fresh) { int rc = read_counters(c); if (rc) return rc; }
        const uint32_t bits = c->dt_raw_host_valid ? c->dt_raw_host : c->h_counters->dt_raw_bits;
        if (bits < SPK_DT_RAW_NONE) { // at least one non-failed particle
            float raw;
            std::memcpy(&raw, &bits, 4);
            float d = std::fmin(max_dt, raw);
            if (d < min_dt && max_dt > min_dt) d = min_dt;
            c->dt_acc_host = std::min(c->dt_acc_host, host_dt_from_sec(d));
        }
    } else {
        { int rc = launch_estimate_dt_hooked(c, min_dt, max_dt); if (rc) return rc; }
        if (spk_dist_active(c)) CK(spk_dist_launch_barrier(c, 2)); // dt_acc <- min over the ranks
        int rc = read_counters(c);
        if (rc) return rc;
        if (c->n || spk_dist_active(c)) c->dt_acc_host = c->h_counters->dt_acc; // (multi-GPU: the minimum over the ranks, also for a rank without particles)
    }
    if (dt_out) *dt_out = host_dt_into_sec(c->dt_acc_host);
    return SPK_OK;
}

int spk_dt_acc_get(spk_ctx* c, uint64_t* out) {
    if (!c || !out) return SPK_EINVAL;
    *out = c->dt_acc_host;
    return SPK_OK;
}
int spk_dt_acc_set(spk_ctx* c, uint64_t v) {
    if (!c) return SPK_EINVAL;
    c->dt_acc_host = v;
    return SPK_OK;
}

int spk_stage_g2p2g(spk_ctx* c, float dt, uint32_t damage_model) {
    (void)damage_model; // unused by the reference kernel too (g2p2g.rs:190)
    if (!c) return SPK_EINVAL;
    if (!c->sorted_valid) return SPK_ESTATE;
    CK(cudaSetDevice(c->device));
    { int jrc = spk_join_grid_update(c); if (jrc) return jrc; }
    return launch_g2p2g_hooked(c, dt);
}

// recompute_fluids_volume_{p2g,g2p} (src_kernels/cuda/fluids_volume.rs, not compiled into the reference's module) with the sequencing of
// MlsSolver::recompute_fluids_volumes (src/dynamics/solver/fluids_volume.rs:9-116): after the sort, before the transfer
int spk_stage_recompute_fluids_volume(spk_ctx* c) {
    if (!c) return SPK_EINVAL;
    if (!c->sorted_valid) return SPK_ESTATE; // needs the particle -> block map of spk_stage_sort (and the zeroed next grid)
    if (spk_dist_active(c)) return SPK_EUNSUPPORTED; // the node masses of shared blocks would need their own halo exchange
    CK(cudaSetDevice(c->device));
    { int jrc = spk_join_grid_update(c); if (jrc) return jrc; }
    CK(spk_launch_fluids_volume(c));
    c->dt_cache_valid = false; // F[0][0] of the fluid particles changed: the cached CFL candidate is stale
    return SPK_OK;
}

int spk_stage_grid_update(spk_ctx* c, float dt, const float gravity[SPK_DIM]) {
    if (!c || !gravity) return SPK_EINVAL;
    CK(cudaSetDevice(c->device));
    { int jrc = spk_join_grid_update(c); if (jrc) return jrc; }
    CK(spk_launch_grid_update(c, dt, gravity));
    return SPK_OK;
}

// CudaMpmPipeline::step with SEVERAL contexts driven serially by one host thread — the reference's own multi-GPU mode
// (cuda_mpm_pipeline.rs:262-659: per substep every context sorts, the active block headers are gathered and shared blocks tagged
// :329-351 / cuda_sparse_grid.rs:256-322, dt is the minimum over the devices :393-398, the halo blocks are copied between the devices
// :458-472 and merged :512-532).  A compatibility path for hosts that keep that calling convention: it sequences the portable halo
// protocol (spk_halo_*) with device-to-device copies; the fast multi-GPU path is one process per GPU over NVLink peer memory
// (spk_dist_attach + plain spk_step).  Contexts may live on the same device.
namespace {
struct MultiScratch { // per context, on the context's device; grown on demand, freed when the call returns
    unsigned long long *keys = nullptr, *rkeys = nullptr, *hk = nullptr, *rhk = nullptr;
    float *hp = nullptr, *rhp = nullptr;
    uint32_t kcap = 0, hcap = 0;
};
struct MultiScratchSet {
    std::vector<MultiScratch> s;
    std::vector<spk_ctx*> ctx;
    ~MultiScratchSet() {
        for (size_t i = 0; i < s.size(); ++i) {
            cudaSetDevice(ctx[i]->device);
            cudaStreamSynchronize(ctx[i]->stream);
            cudaFree(s[i].keys); cudaFree(s[i].rkeys); cudaFree(s[i].hk); cudaFree(s[i].rhk); cudaFree(s[i].hp); cudaFree(s[i].rhp);
        }
    }
};
int sync_all(spk_ctx** ctxs, size_t nctx) {
    for (size_t i = 0; i < nctx; ++i) { CK(cudaSetDevice(ctxs[i]->device)); CK(cudaStreamSynchronize(ctxs[i]->stream)); }
    return SPK_OK;
}
int step_multi(spk_ctx** ctxs, size_t nctx, const SpkSolverParameters* params, const float* gravity, SpkStepTimings* timings, float* simulated_time_out) {
    constexpr size_t PAYLOAD = (size_t)SPK_ACC_CHANNELS * SPK_CELLS_PER_BLOCK;
    for (size_t i = 0; i < nctx; ++i) {
        if (!ctxs[i]) return SPK_EINVAL;
        if (spk_dist_active(ctxs[i])) return SPK_ESTATE; // attached to the peer-memory path: one context per process there
        if (ctxs[i]->cell_width != ctxs[0]->cell_width) return SPK_EINVAL;
        for (size_t j = 0; j < i; ++j) if (ctxs[j] == ctxs[i]) return SPK_EINVAL;
    }
    if (timings) std::memset(timings, 0, sizeof(*timings));
    int rc;
    for (size_t i = 0; i < nctx; ++i) {
        spk_ctx* c = ctxs[i];
        if ((rc = spk_frame_begin(c, params))) return rc;
        if (c->launcher.timestep_started && (rc = c->launcher.timestep_started(c->launcher.user))) return rc;
    }
    size_t total = 0;
    for (size_t i = 0; i < nctx; ++i) total += ctxs[i]->n;
    if (params->dt == 0.0f || total == 0) {
        if (simulated_time_out) *simulated_time_out = params->dt;
        return SPK_OK;
    }
    MultiScratchSet scratch;
    scratch.s.resize(nctx);
    scratch.ctx.assign(ctxs, ctxs + nctx);
    const float min_dt = params->dt / (float)params->max_num_substeps;
    float remaining = params->dt, dt = 0.f;
    uint32_t niter = 0;
    std::vector<SpkSortInfo> info(nctx);
    while (remaining > 0.0f) {
        ++niter;
        uint32_t max_nab = 0;
        for (size_t i = 0; i < nctx; ++i) {
            if ((rc = spk_stage_sort(ctxs[i], &info[i]))) return rc;
            max_nab = std::max(max_nab, info[i].num_active_blocks);
        }
        // ---- gather the active block keys and tag the shared blocks (cuda_sparse_grid.rs:256-322)
        const uint32_t kcap = (max_nab + 255u) / 256u * 256u;
        for (size_t i = 0; i < nctx; ++i) {
            spk_ctx* c = ctxs[i];
            MultiScratch& m = scratch.s[i];
            CK(cudaSetDevice(c->device));
            if (m.kcap < kcap) {
                CK(cudaStreamSynchronize(c->stream));
                cudaFree(m.keys); cudaFree(m.rkeys); m.keys = m.rkeys = nullptr; m.kcap = 0;
                CK(cudaMalloc((void**)&m.keys, (size_t)kcap * 8)); CK(cudaMalloc((void**)&m.rkeys, (size_t)kcap * 8));
                m.kcap = kcap;
            }
            if ((rc = spk_halo_begin(c))) return rc;
            if ((rc = spk_halo_active_keys(c, (uint64_t*)m.keys, kcap))) return rc;
        }
        if ((rc = sync_all(ctxs, nctx))) return rc;
        for (size_t i = 0; i < nctx; ++i) {
            spk_ctx* c = ctxs[i];
            CK(cudaSetDevice(c->device));
            for (size_t o = 0; o < nctx; ++o) {
                if (o == i) continue;
                CK(cudaMemcpyPeerAsync(scratch.s[i].rkeys, c->device, scratch.s[o].keys, ctxs[o]->device, (size_t)kcap * 8, c->stream));
                if ((rc = spk_halo_tag(c, (const uint64_t*)scratch.s[i].rkeys, kcap))) return rc;
            }
            CK(spk_launch_reset_grid(c));
        }
        // ---- dt: the minimum over the contexts (cuda_mpm_pipeline.rs:381-398)
        unsigned long long acc = ~0ull;
        for (size_t i = 0; i < nctx; ++i) {
            if ((rc = spk_stage_estimate_dt(ctxs[i], min_dt, std::fmin(remaining, params->max_substep_dt), nullptr))) return rc;
            acc = std::min(acc, ctxs[i]->dt_acc_host);
        }
        for (size_t i = 0; i < nctx; ++i) ctxs[i]->dt_acc_host = acc;
        dt = host_dt_into_sec(acc);
        uint32_t max_halo = 0;
        for (size_t i = 0; i < nctx; ++i) {
            uint32_t nh = 0;
            if ((rc = spk_halo_count(ctxs[i], &nh))) return rc;
            max_halo = std::max(max_halo, nh);
        }
        // ---- transfers, then exchange and merge the shared blocks (:407-533)
        for (size_t i = 0; i < nctx; ++i)
            if (info[i].num_dispatch_blocks > 0 && (rc = launch_g2p2g_hooked(ctxs[i], dt))) return rc;
        if (max_halo > 0) {
            const uint32_t hcap = (max_halo + 255u) / 256u * 256u;
            for (size_t i = 0; i < nctx; ++i) {
                spk_ctx* c = ctxs[i];
                MultiScratch& m = scratch.s[i];
                CK(cudaSetDevice(c->device));
                if (m.hcap < hcap) {
                    CK(cudaStreamSynchronize(c->stream));
                    cudaFree(m.hk); cudaFree(m.rhk); cudaFree(m.hp); cudaFree(m.rhp); m.hk = m.rhk = nullptr; m.hp = m.rhp = nullptr; m.hcap = 0;
                    CK(cudaMalloc((void**)&m.hk, (size_t)hcap * 8)); CK(cudaMalloc((void**)&m.rhk, (size_t)hcap * 8));
                    CK(cudaMalloc((void**)&m.hp, (size_t)hcap * PAYLOAD * 4)); CK(cudaMalloc((void**)&m.rhp, (size_t)hcap * PAYLOAD * 4));
                    m.hcap = hcap;
                }
                if ((rc = spk_halo_pack(c, (uint64_t*)m.hk, m.hp, hcap))) return rc;
            }
            if ((rc = sync_all(ctxs, nctx))) return rc;
            for (size_t i = 0; i < nctx; ++i) {
                spk_ctx* c = ctxs[i];
                CK(cudaSetDevice(c->device));
                for (size_t o = 0; o < nctx; ++o) {
                    if (o == i) continue;
                    CK(cudaMemcpyPeerAsync(scratch.s[i].rhk, c->device, scratch.s[o].hk, ctxs[o]->device, (size_t)hcap * 8, c->stream));
                    CK(cudaMemcpyPeerAsync(scratch.s[i].rhp, c->device, scratch.s[o].hp, ctxs[o]->device, (size_t)hcap * PAYLOAD * 4, c->stream));
                    if ((rc = spk_halo_merge(c, (const uint64_t*)scratch.s[i].rhk, scratch.s[i].rhp, hcap))) return rc;
                }
            }
        }
        for (size_t i = 0; i < nctx; ++i) {
            CK(cudaSetDevice(ctxs[i]->device));
            CK(spk_launch_grid_update(ctxs[i], dt, gravity));
        }
        remaining -= dt;
        if (remaining <= 10.0f / DT_FACTOR || params->stop_after_one_substep) break;
    }
    if ((rc = sync_all(ctxs, nctx))) return rc;
    if (timings) {
        timings->num_substeps = niter;
        timings->last_dt = dt;
        for (size_t i = 0; i < nctx; ++i) { timings->num_active_blocks += info[i].num_active_blocks; timings->num_dispatch_blocks += info[i].num_dispatch_blocks; }
    }
    if (simulated_time_out) *simulated_time_out = params->dt - remaining;
    return SPK_OK;
}
} // namespace

// CudaMpmPipeline::step_generic, cuda_mpm_pipeline.rs:273-659
int spk_step(spk_ctx** ctxs, size_t nctx, const SpkSolverParameters* params, const float gravity[SPK_DIM], SpkStepTimings* timings, float* simulated_time_out) {
    if (!ctxs || nctx == 0 || !params || !gravity) return SPK_EINVAL;
    if (nctx != 1) return step_multi(ctxs, nctx, params, gravity, timings, simulated_time_out);
    spk_ctx* c = ctxs[0];
    if (!c) return SPK_EINVAL;
    CK(cudaSetDevice(c->device));
    const uint32_t launches0 = c->launches, g2p2g0 = c->g2p2g_launches;
    if (timings) std::memset(timings, 0, sizeof(*timings));
    int rc;
    if ((rc = spk_frame_begin(c, params))) return rc;
    const bool dist = spk_dist_active(c);
    if (dist && c->h_counters->pad[1]) return SPK_EPEER; // a peer was given up on during an earlier call: the ranks have diverged
    if (c->launcher.timestep_started && (rc = c->launcher.timestep_started(c->launcher.user))) return rc; // :299-302
    // nothing to simulate (:305-309) — except that a rank of a multi-GPU run that owns no particle (yet) still walks through the substeps:
    // it takes part in the device-side barriers and in the dt minimum, with an empty grid
    if (params->dt == 0.0f || (c->n == 0 && !dist)) {
        if (simulated_time_out) *simulated_time_out = params->dt;
        return SPK_OK;
    }
    const float min_dt = params->dt / (float)params->max_num_substeps;
    float remaining = params->dt;
    uint32_t niter = 0;
    float acc_ms[5] = {0, 0, 0, 0, 0};
    cudaEvent_t* ev = c->ev;
    if (timings) CK(cudaEventRecord(ev[10], c->stream));
    while (remaining > 0.0f) {
        ++niter;
        if (timings) CK(cudaEventRecord(ev[0], c->stream));
        SpkSortInfo info;
        if ((rc = spk_stage_sort(c, &info))) return rc;
        if (timings) CK(cudaEventRecord(ev[1], c->stream));
        const bool side = !timings && c->overlap_grid_update && (!dist || c->dist_split_barriers);
        if (side) {
            // the grid reset goes to the side stream: there it follows the last grid update by stream order, needs only this sort's block
            // tables (ev_tables, recorded before the permutation scatter) and runs next to the scatter; g2p2g joins the side stream below
            CK(cudaStreamWaitEvent(c->gu_stream, c->ev_tables, 0));
            cudaStream_t main_stream = c->stream;
            c->stream = c->gu_stream;
            const cudaError_t re = spk_launch_reset_grid(c);
            c->stream = main_stream;
            CK(re);
            CK(cudaEventRecord(c->ev_gu_done, c->gu_stream));
            c->gu_pending = true;
        } else {
            if ((rc = spk_join_grid_update(c))) return rc; // reset / g2p2g read what the last grid update wrote
            CK(spk_launch_reset_grid(c));
        }
        if (timings) CK(cudaEventRecord(ev[2], c->stream));
        float dt;
        if ((rc = spk_stage_estimate_dt(c, min_dt, std::fmin(remaining, params->max_substep_dt), &dt))) return rc;
        if (dist) { // every rank has built its hash table: find the blocks shared with the peers
            // With the push kernel (the default) only the side-stream work after g2p2g reads what the detection writes (halo bits, the peers'
            // block ids): the wait for the peers and the probes of their hash tables over NVLink then run on the side stream, next to this
            // substep's grid reset and g2p2g, and are off the critical path altogether
            const bool detect_aside = side && !c->dist_fused_push;
            cudaStream_t main_stream = c->stream;
            if (detect_aside) {
                CK(cudaEventRecord(c->ev_transfer_done, c->stream));
                CK(cudaStreamWaitEvent(c->gu_stream, c->ev_transfer_done, 0));
                c->stream = c->gu_stream;
            }
            cudaError_t de = spk_dist_launch_barrier(c, 0, c->dist_s1_arrived ? 2 : 3); // (the sort already told the peers: only the wait half is left)
            c->dist_s1_arrived = 0;
            if (de == cudaSuccess) de = spk_dist_launch_detect(c);
            c->stream = main_stream;
            CK(de);
        }
        if (timings) CK(cudaEventRecord(ev[3], c->stream));
        if ((rc = spk_join_grid_update(c))) return rc; // (side: the grid reset; the detection queued behind it is not waited for)
        if ((rc = launch_g2p2g_hooked(c, dt))) return rc;
        if (timings) CK(cudaEventRecord(ev[4], c->stream));
        if (dist && !side) { // NVLink red.add of the shared blocks into the peers, barrier (+ global dt minimum), fold what arrived
            if (!c->dist_fused_push) CK(spk_dist_launch_push(c)); // (SPK_DIST_PUSH=fused: pushed by g2p2g's own flush)
            CK(spk_dist_launch_barrier(c, 1, 3));
            CK(spk_dist_launch_merge(c));
        }
        if (side) {
            // the grid update only touches the grid g2p2g just filled; the next substep's sort only reads the particles g2p2g just wrote
            // and builds the OTHER grid buffer: the two run side by side, and meet again before the next reset_grid (multi-GPU: before
            // the next sort publishes its counter record, which carries the dt minimum the barrier computes).  Multi-GPU: the halo push and
            // the whole post-g2p2g barrier sit on the side stream too: the NVLink reductions and the peers' flags travel while the next
            // sort bins particles on the main stream
            CK(cudaEventRecord(c->ev_transfer_done, c->stream));
            CK(cudaStreamWaitEvent(c->gu_stream, c->ev_transfer_done, 0));
            cudaStream_t main_stream = c->stream;
            c->stream = c->gu_stream;
            cudaError_t ge = cudaSuccess;
            if (dist) {
                if (!c->dist_fused_push) ge = spk_dist_launch_push(c);
                if (ge == cudaSuccess) ge = spk_dist_launch_barrier(c, 1, 3);
                if (ge == cudaSuccess) ge = spk_dist_launch_merge(c);
            }
            if (ge == cudaSuccess) ge = spk_launch_grid_update(c, dt, gravity);
            c->stream = main_stream;
            CK(ge);
            CK(cudaEventRecord(c->ev_gu_done, c->gu_stream));
            c->gu_pending = true;
        } else {
            CK(spk_launch_grid_update(c, dt, gravity));
        }
        if (timings) {
            CK(cudaEventRecord(ev[5], c->stream));
            CK(cudaEventSynchronize(ev[5]));
            for (int k = 0; k < 5; ++k) { float ms = 0; CK(cudaEventElapsedTime(&ms, ev[k], ev[k + 1])); acc_ms[k] += ms; }
            timings->last_dt = dt;
            timings->num_active_blocks = info.num_active_blocks;
            timings->num_dispatch_blocks = info.num_dispatch_blocks;
        }
        remaining -= dt;
        if (remaining <= 10.0f / DT_FACTOR || params->stop_after_one_substep) break; // :630-638
    }
    // Without timings the call returns as soon as the frame is enqueued: the next spk_step (or any other entry point, all of
    // them are ordered on the context's stream) follows without the GPU going idle in between; spk_ctx_synchronize() waits.
    if (timings) {
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaEventRecord(ev[11], c->stream));
        CK(cudaEventSynchronize(ev[11]));
        CK(cudaEventElapsedTime(&timings->total_ms, ev[10], ev[11]));
        timings->num_substeps = niter;
        timings->grid_resize_and_sort_ms = acc_ms[0]; timings->reset_grid_ms = acc_ms[1]; timings->estimate_timestep_ms = acc_ms[2];
        timings->g2p2g_ms = acc_ms[3]; timings->grid_update_ms = acc_ms[4];
        timings->g2p2g_launches = c->g2p2g_launches - g2p2g0;
        timings->kernel_launches = c->launches - launches0;
    }
    // (a grid update left on its own stream is joined by the next call that touches the grid: the next spk_step does so only before
    // its reset_grid, i.e. after its particle sort ran next to it)
    if (simulated_time_out) *simulated_time_out = params->dt - remaining;
    // (the barriers of the frame's last substep run after its counter record was published: a time-out there is written straight into
    // the mapped host record by the barrier kernel and is seen by the next call that looks — the next spk_step, spk_ctx_synchronize)
    if (dist && c->h_counters->pad[1]) return SPK_EPEER;
    return SPK_OK;
}

// ------------------------------------------------------------------------------------------------ read-back
int spk_grid_num_active_blocks(spk_ctx* c, int which_grid, uint32_t* out) {
    if (!c || !out) return SPK_EINVAL;
    *out = (which_grid ? c->grid[1 - c->gcur] : c->grid[c->gcur]).nab_host;
    return SPK_OK;
}
int spk_grid_read_headers(spk_ctx* c, int which_grid, SpkActiveBlockHeader* dst, uint32_t n) {
    if (!c || (n && !dst)) return SPK_EINVAL;
    { int jrc = spk_join_grid_update(c); if (jrc) return jrc; }
    GridBuffers& g = which_grid ? c->grid[1 - c->gcur] : c->grid[c->gcur];
    if (n > g.nab_host) return SPK_EINVAL;
    if (!n) return SPK_OK;
    CK(cudaSetDevice(c->device));
    std::vector<unsigned long long> key(n);
    std::vector<uint32_t> first(n), count(n), halo(n);
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(key.data(), g.blk_key, n * 8ull, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(first.data(), g.blk_first, n * 4ull, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(count.data(), g.blk_count, n * 4ull, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(halo.data(), g.blk_halo, n * 4ull, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < n; ++i) dst[i] = SpkActiveBlockHeader{key[i], first[i], count[i], halo[i], 0};
    return SPK_OK;
}
int spk_grid_read_nodes(spk_ctx* c, int which_grid, SpkGpuGridNode* dst, uint32_t n_blocks) {
    if (!c || (n_blocks && !dst)) return SPK_EINVAL;
    { int jrc = spk_join_grid_update(c); if (jrc) return jrc; }
    GridBuffers& g = which_grid ? c->grid[1 - c->gcur] : c->grid[c->gcur];
    if (n_blocks > g.nab_host) return SPK_EINVAL;
    if (!n_blocks) return SPK_OK;
    CK(cudaSetDevice(c->device));
    const size_t bytes = (size_t)n_blocks * SPK_CELLS_PER_BLOCK * sizeof(SpkGpuGridNode);
    CK(ensure_staging(c, bytes));
    CK(spk_launch_grid_export(c, which_grid, (SpkGpuGridNode*)c->staging, n_blocks));
    CK(cudaMemcpyAsync(dst, c->staging, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SPK_OK;
}
int spk_grid_morton_box(spk_ctx* c, uint32_t out[8]) {
    if (!c || !out) return SPK_EINVAL;
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(out, c->dist_bbox + 16, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SPK_OK;
}

// dispatch table in the reference's format: multiplicity = ceil(num_particles / 128) entries per block
// (write_blocks_multiplicity_to_scan_value + init_gpu_dispatch_blocks_mapping, sort.rs:226-280), derived from the headers
int spk_grid_read_dispatch(spk_ctx* c, SpkDispatchBlock2ActiveBlock* dst, uint32_t n) {
    if (!c || (n && !dst)) return SPK_EINVAL;
    { int jrc = spk_join_grid_update(c); if (jrc) return jrc; }
    GridBuffers& g = c->sorted_valid ? c->grid[1 - c->gcur] : c->grid[c->gcur];
    std::vector<SpkActiveBlockHeader> h(g.nab_host);
    int rc = spk_grid_read_headers(c, c->sorted_valid ? 1 : 0, h.data(), g.nab_host);
    if (rc) return rc;
    uint32_t k = 0;
    for (uint32_t b = 0; b < g.nab_host; ++b) {
        const uint32_t mult = h[b].num_particles / SPK_REF_G2P2G_THREADS + (h[b].num_particles % SPK_REF_G2P2G_THREADS > 0);
        for (uint32_t m = 0; m < mult; ++m) {
            if (k >= n) return SPK_OK;
            dst[k++] = SpkDispatchBlock2ActiveBlock{b, h[b].first_particle + m * SPK_REF_G2P2G_THREADS};
        }
    }
    return k == n ? SPK_OK : SPK_EINVAL;
}

int spk_g2p2g_time(spk_ctx* c, int enable, float* total_ms_out, uint32_t* launches_out) {
    if (!c) return SPK_EINVAL;
    CK(cudaSetDevice(c->device));
    float total = 0.f;
    if (c->kev_count) {
        CK(cudaEventSynchronize(c->kev[2 * c->kev_count - 1]));
        for (uint32_t k = 0; k < c->kev_count; ++k) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, c->kev[2 * k], c->kev[2 * k + 1])); total += ms; }
    }
    if (total_ms_out) *total_ms_out = total;
    if (launches_out) *launches_out = c->kev_count;
    c->kev_count = 0;
    if (enable && !c->kev[0]) for (auto& e : c->kev) CK(cudaEventCreate(&e));
    c->kev_on = enable != 0;
    return SPK_OK;
}

int spk_device_ptr(spk_ctx* c, int which, void** out) {
    if (!c || !out) return SPK_EINVAL;
    switch (which) {
    case SPK_ATTR_POSITION: *out = c->pb[c->cur].a0; break;
    case SPK_ATTR_VELOCITY: *out = c->vel; break;
    default: return SPK_EINVAL;
    }
    return SPK_OK;
}

// ------------------------------------------------------------------------------------------------ halo (multi-process)
int spk_halo_begin(spk_ctx* c) {
    if (!c) return SPK_EINVAL;
    { int jrc = spk_join_grid_update(c); if (jrc) return jrc; }
    CK(cudaSetDevice(c->device));
    CK(cudaMemsetAsync(&c->d_counters->num_halo, 0, 4, c->stream));
    return SPK_OK;
}
int spk_halo_active_keys(spk_ctx* c, uint64_t* dev_keys_out, uint32_t capacity) {
    if (!c || (capacity && !dev_keys_out)) return SPK_EINVAL;
    { int jrc = spk_join_grid_update(c); if (jrc) return jrc; }
    if (c->grid[1 - c->gcur].nab_host > capacity) return SPK_EOVERFLOW;
    CK(cudaSetDevice(c->device));
    CK(spk_launch_halo_keys(c, (unsigned long long*)dev_keys_out, capacity));
    return SPK_OK;
}
int spk_halo_tag(spk_ctx* c, const uint64_t* dev_remote_keys, uint32_t n_remote) {
    if (!c) return SPK_EINVAL;
    { int jrc = spk_join_grid_update(c); if (jrc) return jrc; }
    CK(cudaSetDevice(c->device));
    CK(spk_launch_halo_tag(c, (const unsigned long long*)dev_remote_keys, n_remote));
    return SPK_OK;
}
int spk_halo_count(spk_ctx* c, uint32_t* num_halo_out) {
    if (!c || !num_halo_out) return SPK_EINVAL;
    { int jrc = spk_join_grid_update(c); if (jrc) return jrc; }
    CK(cudaSetDevice(c->device));
    int rc = read_counters(c);
    if (rc) return rc;
    *num_halo_out = c->h_counters->num_halo;
    return SPK_OK;
}
int spk_halo_pack(spk_ctx* c, uint64_t* dev_keys_out, float* dev_payload_out, uint32_t capacity) {
    if (!c) return SPK_EINVAL;
    { int jrc = spk_join_grid_update(c); if (jrc) return jrc; }
    CK(cudaSetDevice(c->device));
    CK(spk_launch_halo_pack(c, (unsigned long long*)dev_keys_out, dev_payload_out, capacity));
    return SPK_OK;
}
int spk_halo_merge(spk_ctx* c, const uint64_t* dev_keys, const float* dev_payload, uint32_t n_blocks) {
    if (!c) return SPK_EINVAL;
    { int jrc = spk_join_grid_update(c); if (jrc) return jrc; }
    CK(cudaSetDevice(c->device));
    CK(spk_launch_halo_merge(c, (const unsigned long long*)dev_keys, dev_payload, n_blocks));
    return SPK_OK;
}

} // extern "C"
