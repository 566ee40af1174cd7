// spk_g2p2g2d_kernel.cuh — the dim2 fused G2P2G kernel and the dt kernel as templates over the particle updater
// (DefaultParticleUpdater2 in spk_g2p2g2d_body.cuh), with their launch functions; the dim2 counterpart of spk_g2p2g_kernel.cuh.
// libsparkl_b200_2d.so instantiates them with the default updater (spk_g2p2g2d.cu); a user's .cu can instantiate them with its own
// updater and register the launches through spk_set_particle_kernels_launcher.
#pragma once
#include "spk_device.cuh"
#include "spk_math2d.cuh"
#include "spk_g2p2g2d_body.cuh"

#if SPK_DIM != 2
#error "spk_g2p2g2d.cu is the dim2 kernel (compile with -DSPK_DIM=2)"
#endif

namespace spk_g2p2g2d_detail {

constexpr int G2P2G_THREADS = SPK_G2P2G_CHUNK;

struct G2P2GParams {
    ParticleBuffers src, dst;
    float4* vel_out;
    const float4* cold;
    const uint32_t* perm;
    const DispatchEntry* dispatch;
    const DevModel* models;
    const int32_t* nbr_self;
    const int32_t* nbr_prev;
    float* acc;             // next grid accumulators [block][6][16]
    const float4* prev_vel; // previous grid velocities [block][16]
    float h, inv_h, dt;
    uint32_t* dt_raw_bits;
    const void* user;       // opaque device pointer handed to the updater (NULL for the default)
};

#ifdef SPK_EMU // host emulation (tests/host/emu/): the bulk copies / reductions become plain loops, the mbarrier disappears (the
// kernel has a __syncthreads between the loads and their consumers)
__device__ __forceinline__ void mbar_init(uint64_t*, uint32_t) {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t*, uint32_t) {}
__device__ __forceinline__ void mbar_wait(uint64_t*, uint32_t) {}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t*) { memcpy(smem_dst, gmem_src, bytes); }
__device__ __forceinline__ void tma_reduce_add_f32_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    for (uint32_t i = 0; i < bytes / 4; ++i) static_cast<float*>(gmem_dst)[i] += static_cast<const float*>(smem_src)[i];
}
__device__ __forceinline__ void tma_commit_and_wait_read() {}
__device__ __forceinline__ void fence_proxy_async() {}
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_f32_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif // SPK_EMU

template <class Updater>
__global__ void __launch_bounds__(G2P2G_THREADS) g2p2g2d_kernel(const G2P2GParams P) {
    __shared__ __align__(128) float4 s_vel[SPK_TILE_CELLS];             // [quadrant][cell]
    __shared__ __align__(128) float s_acc[SPK_NBR * G2P2G2D_ACC_BLOCK]; // [quadrant][channel][cell]
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_nbr_self[SPK_NBR], s_nbr_prev[SPK_NBR];
    __shared__ unsigned s_dtmin;

    const int tid = threadIdx.x;
    const DispatchEntry ent = P.dispatch[blockIdx.x];
    if (tid < SPK_NBR) {
        s_nbr_self[tid] = P.nbr_self[ent.block * SPK_NBR + tid];
        s_nbr_prev[tid] = P.nbr_prev[ent.block * SPK_NBR + tid];
    }
    if (tid == 0) { mbar_init(&s_bar, 1); s_dtmin = SPK_DT_RAW_NONE; }
    for (int i = tid; i < SPK_NBR * G2P2G2D_ACC_BLOCK; i += G2P2G_THREADS) s_acc[i] = 0.f;
    __syncthreads();
    // ---- stage 0: velocity tile of the 2x2 blocks (transfer_global_blocks_to_shared_memory, g2p2g.rs:353-474)
    if (tid == 0) {
        uint32_t bytes = 0;
        for (int o = 0; o < SPK_NBR; ++o) bytes += s_nbr_prev[o] >= 0 ? (uint32_t)(SPK_CELLS_PER_BLOCK * 16) : 0u;
        fence_proxy_async();
        mbar_expect_tx(&s_bar, bytes);
        for (int o = 0; o < SPK_NBR; ++o)
            if (s_nbr_prev[o] >= 0)
                tma_load_1d(&s_vel[o * SPK_CELLS_PER_BLOCK], P.prev_vel + (size_t)s_nbr_prev[o] * SPK_CELLS_PER_BLOCK, SPK_CELLS_PER_BLOCK * 16, &s_bar);
    }
    for (int i = tid; i < SPK_TILE_CELLS; i += G2P2G_THREADS)
        if (s_nbr_prev[i >> SPK_CELL_SHIFT] < 0) s_vel[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    const bool has_particle = (uint32_t)tid < ent.count;

    // ---- stage 1: one particle per thread (spk_g2p2g2d_body.cuh): gather while the tile is in flight, then G2P + update
    Particle2dWork k;
    if (has_particle) g2p2g2d_load(P.src, P.perm[ent.first + tid], P.h, P.inv_h, k);
    mbar_wait(&s_bar, 0);
    __syncthreads();
    if (has_particle) g2p2g2d_update<Updater>(s_vel, P.models, P.cold, P.h, P.dt, k, P.user);

    // ---- fused K15 (timestep.rs:38-113): un-clamped dt candidate of this particle for the NEXT substep; the clamps and the
    // fixed-point conversion are monotone, so they are applied once to the reduced value (spk_stage_estimate_dt)
    {
        unsigned cand_bits = has_particle ? g2p2g2d_dt_bits<Updater>(P.models, P.h, k, P.user) : SPK_DT_RAW_NONE;
        cand_bits = __reduce_min_sync(0xffffffffu, cand_bits);
        if ((tid & 31) == 0 && cand_bits < SPK_DT_RAW_NONE) atomicMin(&s_dtmin, cand_bits);
    }

    // ---- store the particle into the next buffer at its sorted slot (coalesced), then P2G into the shared tile
    if (has_particle) {
        g2p2g2d_store(P.dst, P.vel_out, ent.first + tid, k);
        g2p2g2d_scatter(s_acc, P.h, P.inv_h, P.dt, k);
    }

    // ---- stage 3: flush the accumulation tile into the next grid (transfer_shared_blocks_to_grid, g2p2g.rs:476-553)
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        for (int o = 0; o < SPK_NBR; ++o) {
            const int nid = s_nbr_self[o];
            if (nid >= 0) tma_reduce_add_f32_1d(P.acc + (size_t)nid * G2P2G2D_ACC_BLOCK, &s_acc[o * G2P2G2D_ACC_BLOCK], G2P2G2D_ACC_BLOCK * 4);
        }
        if (s_dtmin < SPK_DT_RAW_NONE) atomicMin(P.dt_raw_bits, s_dtmin);
        tma_commit_and_wait_read();
    }
}

// ---------------------------------------------------------------------------------------------------- dt (K15)
constexpr int DT_THREADS = 256;
__device__ __forceinline__ unsigned long long dt_from_sec(float dt) { // GpuTimestepLength::from_sec, gpu_timestep.rs:20-22
    const float v = dt * 1000000000000.0f;
    if (v != v) return 0ull;
    return __float2ull_rz(v);
}
template <class Updater>
__global__ void __launch_bounds__(DT_THREADS) estimate_dt2d_kernel(const float4* __restrict__ a0, const float4* __restrict__ a1, const float4* __restrict__ a3,
                                                                    const float4* __restrict__ a4, const float4* __restrict__ vel, uint32_t n,
                                                                    const DevModel* __restrict__ models, float h, float min_dt, float max_dt,
                                                                    DeviceCounters* counters, const void* user) {
    unsigned long long best = 0xffffffffffffffffull;
    for (uint32_t i = blockIdx.x * DT_THREADS + threadIdx.x; i < n; i += gridDim.x * DT_THREADS) {
        const uint32_t flags = __float_as_uint(a0[i].w);
        if (flags & SPK_F_FAILED) continue; // timestep.rs:72-74
        const DevModel& m = models[flags >> SPK_F_MODEL_SHIFT];
        const float4 r3 = a3[i], r4 = a4[i], v = vel[i];
        float dt = max_dt;
        ParticleState2 p; // the fields a dt candidate may read
        p.flags = flags; p.mass = r3.z; p.volume0 = r3.w; p.jp = r3.y; p.F.m[0] = a1[i].x; p.elastic_hardening = r4.z; p.plastic_hardening = r4.w;
        p.phase = r4.x; p.psi_pos = r4.y;
        const float pv[2] = {v.x, v.y};
        dt = fminf(dt, Updater::dt_candidate(m, h, p, pv, user));
        if (dt < min_dt && max_dt > min_dt) dt = min_dt; // timestep.rs:107-109
        const unsigned long long c = dt_from_sec(dt);
        best = c < best ? c : best;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
        best = o < best ? o : best;
    }
    __shared__ unsigned long long s_best[DT_THREADS / 32];
    if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < DT_THREADS / 32; ++w) best = s_best[w] < best ? s_best[w] : best;
        if (best != 0xffffffffffffffffull) atomicMin(&counters->dt_acc, best);
    }
}

} // namespace spk_g2p2g2d_detail

template <class Updater>
inline cudaError_t spk_launch_g2p2g_with(spk_ctx* c, float dt, const void* user = nullptr) {
    using namespace spk_g2p2g2d_detail;
    GridBuffers& next = c->grid[1 - c->gcur];
    GridBuffers& prev = c->grid[c->gcur];
    const uint32_t nd = c->h_counters->ndispatch;
    if (nd == 0) return cudaSuccess; // never launch with 0 CTAs (cuda_mpm_pipeline.rs:485-486)
    G2P2GParams P;
    P.src = c->pb[c->cur];
    P.dst = c->pb[1 - c->cur];
    P.vel_out = c->vel;
    P.cold = c->cold;
    P.perm = c->perm;
    P.dispatch = c->dispatch;
    P.models = c->models;
    P.nbr_self = next.nbr_self;
    P.nbr_prev = next.nbr_prev;
    P.acc = next.acc;
    P.prev_vel = prev.vel;
    P.h = c->cell_width;
    P.inv_h = 1.0f / c->cell_width;
    P.dt = dt;
    P.dt_raw_bits = &c->d_counters->dt_raw_bits;
    P.user = user;
    { cudaError_t e = cudaMemsetAsync(&c->d_counters->dt_raw_bits, 0x7f, 4, c->stream); if (e != cudaSuccess) return e; }
    const bool timed = c->kev_on && c->kev_count < 256;
    if (timed) cudaEventRecord(c->kev[2 * c->kev_count], c->stream);
    g2p2g2d_kernel<Updater><<<nd, G2P2G_THREADS, 0, c->stream>>>(P);
    if (timed) { cudaEventRecord(c->kev[2 * c->kev_count + 1], c->stream); c->kev_count += 1; }
    c->launches += 1;
    c->g2p2g_launches += 1;
    c->cur = 1 - c->cur; // the next buffer now holds the particles, physically in sorted order
    c->sorted_valid = false;
    c->dt_cache_valid = true;
    c->dt_raw_host_valid = false;
    c->counters_fresh = false;
    return cudaGetLastError();
}

template <class Updater>
inline cudaError_t spk_launch_estimate_dt_with(spk_ctx* c, float min_dt, float max_dt, const void* user = nullptr) {
    using namespace spk_g2p2g2d_detail;
    const uint32_t n = (uint32_t)c->n;
    if (!n) return cudaSuccess;
    const ParticleBuffers& pb = c->pb[c->cur];
    int blocks = (int)((n + DT_THREADS - 1) / DT_THREADS);
    const int maxb = c->sm_count * 8;
    if (blocks > maxb) blocks = maxb;
    *c->h_upload = c->dt_acc_host;
    { cudaError_t e = cudaMemcpyAsync(&c->d_counters->dt_acc, c->h_upload, 8, cudaMemcpyHostToDevice, c->stream); if (e != cudaSuccess) return e; }
    c->counters_fresh = false;
    estimate_dt2d_kernel<Updater><<<blocks, DT_THREADS, 0, c->stream>>>(pb.a0, pb.a1, pb.a3, pb.a4, c->vel, n, c->models, c->cell_width, min_dt, max_dt, c->d_counters, user);
    c->launches += 1;
    return cudaGetLastError();
}
