// spk_g2p2g_kernel.cuh — the fused G2P2G kernel as a template over the particle updater (see spk_updater.cuh) and its launch
// function.  libsparkl_b200.so instantiates it with DefaultParticleUpdater (spk_g2p2g.cu); a user who needs the reference's
// `g2p2g_generic::<MyUpdater>` includes this header in their own .cu, defines `MyUpdater` and registers
// `spk_launch_g2p2g_with<MyUpdater>` through spk_set_particle_kernels_launcher (examples/custom_updater/).
#pragma once
#include <type_traits>

#include "spk_device.cuh"
#include "spk_math.cuh"
#include "spk_updater.cuh"

namespace spk_g2p2g_detail {

#ifndef SPK_G2P2G_MIN_CTAS
#define SPK_G2P2G_MIN_CTAS 3
#endif
constexpr int G2P2G_THREADS = 256;
constexpr int G2P2G_CHUNK = SPK_G2P2G_CHUNK;
constexpr int NBINS = 216;         // 6x6x6 possible stencil-base cells of a block's particles after advection
constexpr int PAY_FLOATS = 28;     // staged P2G payload per particle (7 float4)

struct G2P2GParams {
    ParticleBuffers src, dst;
    float4* vel_out;
    const float4* cold;
    const uint32_t* perm;
    const DispatchEntry* dispatch;
    const int32_t* dispatch_nbr; // [ndispatch][16]: nbr_self | nbr_prev of the entry's block
    const DevModel* models;
    const int32_t* nbr_self;
    const int32_t* nbr_prev;
    float* acc;             // next grid accumulators [block][6][64]
    const float4* prev_vel; // previous grid velocities [block][64]
    float h, inv_h, dt;
    uint32_t ndispatch;
    uint32_t* dt_raw_bits; // CFL candidate of the NEXT substep (fused K15)
    // multi-GPU (world == 0: single GPU): shared blocks are ALSO reduced into every holder's halo accumulator over NVLink
    uint32_t world;
    const uint32_t* blk_halo;   // [block] bit0: active on another rank too (spk_dist.cu: dist_detect_kernel)
    const int32_t* peer_ids;    // [block][world] the peer's id of the block, -1 if absent
    float* const* peer_hacc;    // [world] peer-mapped `hacc` of each rank's next grid
    uint32_t prefetch_dist;     // cell kernel: CTA i warms the L2 with the particles of dispatch entry i + prefetch_dist (0 = off)
    uint32_t push_wait;         // 1: a CTA that pushed waits for the remote completion of its reductions (A/B); 0: the kernel boundary orders them
    const void* user;           // opaque device pointer handed to the updater (custom updaters; NULL for the default)
};

// shared memory carve-up (dynamic)
constexpr int SM_VEL = 0;                                   // float4[512]            [octant][cell]
constexpr int SM_ACC = SM_VEL + 512 * 16;                   // float[8*6*64]          [octant][channel][cell]
constexpr int SM_PAY = SM_ACC + 8 * 6 * 64 * 4;             // float[256*28]          staged P2G payloads, cell-sorted
constexpr int SM_BINCNT = SM_PAY + G2P2G_CHUNK * PAY_FLOATS * 4;
constexpr int SM_BINSTART = SM_BINCNT + 224 * 4;
constexpr int SM_BINLIST = SM_BINSTART + 224 * 4;
constexpr int SM_MISC = SM_BINLIST + 224 * 2;               // mbarrier (8) + nbr_self[8] + nbr_prev[8] + nnz
constexpr int SM_HALO = SM_MISC + 8 + 32 + 32 + 16;          // uint32[8]: bit0 = octant block is shared with another rank (multi-GPU)
constexpr int SM_TOTAL = SM_HALO + 32;

// ---- TMA bulk helpers (sm_90+; SASS: UBLKCP / UBLKRED) ------------------------------------------------------------
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
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
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
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif // SPK_EMU

// an updater with `static constexpr bool kPhased = true` provides advect / constitutive (spk_updater.cuh) next to update
template <class U, class = void> struct updater_is_phased { static constexpr bool value = false; };
template <class U> struct updater_is_phased<U, std::enable_if_t<U::kPhased>> { static constexpr bool value = true; };
// keeps the compiler from forwarding values across a park / un-park pair (the point is to free the registers in between)
__device__ __forceinline__ void spk_compiler_fence() {
#ifndef SPK_EMU
    asm volatile("" ::: "memory");
#endif
}

// tile index helpers: tile coordinate t in [0,8) per axis -> (cell part, octant part) of the [octant][...][cell] layout
__device__ __forceinline__ int cell_part(int t, int axis) { return (t & 3) << (2 * axis); }
__device__ __forceinline__ int oct_part(int t, int axis) { return (t >> 2) << axis; }

// QuadraticKernel::eval_all, src_core/dynamics/solver/kernel.rs:15-21
__device__ __forceinline__ void quad_weights(float x, float w[3]) {
    w[0] = 0.5f * ((1.5f - x) * (1.5f - x));
    w[1] = 0.75f - ((x - 1.f) * (x - 1.f));
    w[2] = 0.5f * ((x - 0.5f) * (x - 0.5f));
}

// MODE 0: every particle adds its 27x6 contributions to the shared tile with float atomics (CAS loops on sm_100).
// MODE 1: particles are counting-sorted by stencil-base cell inside the CTA, their P2G payloads staged in shared
//         memory, and 9-lane groups (one lane per (sy, sz) stencil column) accumulate a whole cell in registers before
//         touching the tile: shared float atomics drop from 162 per particle to 162 per occupied cell, with no
//         same-address conflicts inside a group.
template <int MODE, class Updater>
__global__ void __launch_bounds__(G2P2G_THREADS, SPK_G2P2G_MIN_CTAS) g2p2g_kernel(const G2P2GParams P) {
#ifdef SPK_EMU
    unsigned char* smem = emu::dynamic_smem();
#else
    extern __shared__ __align__(128) unsigned char smem[];
#endif
    float4* s_vel = reinterpret_cast<float4*>(smem + SM_VEL);
    float* s_acc = reinterpret_cast<float*>(smem + SM_ACC);
    float* s_pay = reinterpret_cast<float*>(smem + SM_PAY);
    int* s_bin_cnt = reinterpret_cast<int*>(smem + SM_BINCNT);
    int* s_bin_start = reinterpret_cast<int*>(smem + SM_BINSTART);
    unsigned short* s_binlist = reinterpret_cast<unsigned short*>(smem + SM_BINLIST);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + SM_MISC);
    int* s_nbr_self = reinterpret_cast<int*>(smem + SM_MISC + 8);
    int* s_nbr_prev = reinterpret_cast<int*>(smem + SM_MISC + 40);
    int* s_nnz = reinterpret_cast<int*>(smem + SM_MISC + 72);
    unsigned* s_dtmin = reinterpret_cast<unsigned*>(smem + SM_MISC + 76);
    uint32_t* s_halo = reinterpret_cast<uint32_t*>(smem + SM_HALO);

    const int tid = threadIdx.x;
    const DispatchEntry ent = P.dispatch[blockIdx.x];
    if (tid < 8) {
        const int nid = P.nbr_self[ent.block * 8 + tid];
        s_nbr_self[tid] = nid;
        s_nbr_prev[tid] = P.nbr_prev[ent.block * 8 + tid];
        // multi-GPU: is the octant's block shared with a peer?  Fetched here by 8 threads in parallel, long before the flush needs
        // it (a serial chain of 8 L2 loads by thread 0 at the very end kept every CTA's slot busy for ~2.5 us)
        s_halo[tid] = (P.world > 1 && nid >= 0) ? P.blk_halo[nid] : 0u;
    }
    if (tid == 0) { mbar_init(s_bar, 1); *s_dtmin = SPK_DT_RAW_NONE; }
    for (int i = tid; i < 8 * SPK_ACC_CHANNELS * 64; i += G2P2G_THREADS) s_acc[i] = 0.f;
    if (MODE == 1 && tid < 224) s_bin_cnt[tid] = 0;
    __syncthreads();
    // ---- stage 0: velocity tile of the 2x2x2 blocks (transfer_global_blocks_to_shared_memory, g2p2g.rs:353-474)
    if (tid == 0) {
        uint32_t bytes = 0;
        for (int o = 0; o < 8; ++o) bytes += s_nbr_prev[o] >= 0 ? 1024u : 0u;
        fence_proxy_async();
        mbar_expect_tx(s_bar, bytes);
        for (int o = 0; o < 8; ++o)
            if (s_nbr_prev[o] >= 0) tma_load_1d(&s_vel[o * 64], P.prev_vel + (size_t)s_nbr_prev[o] * 64, 1024u, s_bar);
    }
    for (int i = tid; i < SPK_TILE_CELLS; i += G2P2G_THREADS)
        if (s_nbr_prev[i >> 6] < 0) s_vel[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    const float h = P.h, inv_h = P.inv_h, dt = P.dt;
    const float inv_d = 4.f / (h * h);
    const bool has_particle = (uint32_t)tid < ent.count;

    // ---- stage 1: one particle per thread ------------------------------------------------------------------------
    ParticleState p;
    float vel[3] = {0.f, 0.f, 0.f};
    Mat3 G = mat_zero();
    Mat3 stress = mat_zero();
    bool do_p2g = false;
    int tcell[3] = {1, 1, 1};
    float cell_before[3] = {0.f, 0.f, 0.f};
    float ref[3], w[3][3];
    if (has_particle) {
        const uint32_t si = P.perm[ent.first + tid];
        const float4 r0 = P.src.a0[si], r1 = P.src.a1[si], r2 = P.src.a2[si], r3 = P.src.a3[si], r4 = P.src.a4[si], r5 = P.src.a5[si];
        p.x[0] = r0.x; p.x[1] = r0.y; p.x[2] = r0.z; p.flags = __float_as_uint(r0.w);
        p.F.m[0] = r1.x; p.F.m[1] = r1.y; p.F.m[2] = r1.z; p.F.m[3] = r1.w;
        p.F.m[4] = r2.x; p.F.m[5] = r2.y; p.F.m[6] = r2.z; p.F.m[7] = r2.w;
        p.F.m[8] = r3.x; p.jp = r3.y; p.mass = r3.z; p.volume0 = r3.w;
        p.phase = r4.x; p.psi_pos = r4.y; p.elastic_hardening = r4.z; p.plastic_hardening = r4.w;
        p.log_vol_gain = r5.x; p.nacc_alpha = r5.y; p.orig_id = __float_as_uint(r5.z);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float q = roundf(spk_div_h(p.x[a], h, inv_h));
            cell_before[a] = q;
            ref[a] = (q - 1.f) * h - p.x[a];               // dir_to_associated_grid_node, particle.rs:37-47
            quad_weights(spk_div_h(-ref[a], h, inv_h), w[a]);
            const int assoc_cell = (q != q) ? -2 : (int)__float2ll_rz(q) - 2;
            tcell[a] = (assoc_cell & 3) + 1;                // associated_cell_index_in_block_off_by_two, particle.rs:28-35
        }
    }
    mbar_wait(s_bar, 0);
    __syncthreads();

    if (has_particle) {
        const DevModel& model = P.models[p.flags >> SPK_F_MODEL_SHIFT];
        // ---- G2P (particle_g2p2g, g2p2g.rs:192-245), separable accumulation: rows along x, then planes, then volume.
        //   V = sum w v,  G[:,a] = inv_d * sum w dpt_a v   with w = wx wy wz and dpt = ref + s h
        {
            float V[3] = {0.f, 0.f, 0.f}, Gx[3] = {0.f, 0.f, 0.f}, Gy[3] = {0.f, 0.f, 0.f}, Gz[3] = {0.f, 0.f, 0.f};
            const float wpx[3] = {w[0][0] * ref[0], w[0][1] * (ref[0] + h), w[0][2] * (ref[0] + 2.f * h)};
#pragma unroll
            for (int dz = 0; dz < 3; ++dz) {
                const int iz = cell_part(tcell[2] + dz, 2) | (oct_part(tcell[2] + dz, 2) << 6);
                float P0[3] = {0.f, 0.f, 0.f}, P1[3] = {0.f, 0.f, 0.f}, P2[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const int iy = iz | cell_part(tcell[1] + dy, 1) | (oct_part(tcell[1] + dy, 1) << 6);
                    float R0[3] = {0.f, 0.f, 0.f}, R1[3] = {0.f, 0.f, 0.f};
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const float4 c = s_vel[iy | cell_part(tcell[0] + dx, 0) | (oct_part(tcell[0] + dx, 0) << 6)];
                        R0[0] += w[0][dx] * c.x; R0[1] += w[0][dx] * c.y; R0[2] += w[0][dx] * c.z;
                        R1[0] += wpx[dx] * c.x; R1[1] += wpx[dx] * c.y; R1[2] += wpx[dx] * c.z;
                    }
                    const float wy = w[1][dy], wpy = wy * (ref[1] + (float)dy * h);
#pragma unroll
                    for (int a = 0; a < 3; ++a) { P0[a] += wy * R0[a]; P1[a] += wy * R1[a]; P2[a] += wpy * R0[a]; }
                }
                const float wz = w[2][dz], wpz = wz * (ref[2] + (float)dz * h);
#pragma unroll
                for (int a = 0; a < 3; ++a) { V[a] += wz * P0[a]; Gx[a] += wz * P1[a]; Gy[a] += wz * P2[a]; Gz[a] += wpz * P0[a]; }
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                vel[a] = V[a];
                M_(G, a, 0) = Gx[a] * inv_d; M_(G, a, 1) = Gy[a] * inv_d; M_(G, a, 2) = Gz[a] * inv_d;
            }
        }
        do_p2g = Updater::update(model, P.cold, h, dt, p, vel, G, stress, P.user);
    }

    // ---- fused K15 (timestep.rs:38-113): un-clamped dt candidate of this particle for the NEXT substep; the clamps and the
    // fixed-point conversion are monotone, so they are applied once to the reduced value (spk_stage_estimate_dt)
    {
        unsigned cand_bits = SPK_DT_RAW_NONE;
        if (has_particle && !(p.flags & SPK_F_FAILED)) {
            const DevModel& model = P.models[p.flags >> SPK_F_MODEL_SHIFT];
            const float cand = Updater::dt_candidate(model, h, p, vel, P.user);
            cand_bits = __float_as_uint(cand != cand ? 1.0e30f : fminf(fmaxf(cand, 0.f), 1.0e30f)); // NaN imposes no bound (Rust f32::min ignores NaN)
        }
        cand_bits = __reduce_min_sync(0xffffffffu, cand_bits);
        if ((tid & 31) == 0 && cand_bits < SPK_DT_RAW_NONE) atomicMin(s_dtmin, cand_bits);
    }

    // ---- P2G preparation with the UPDATED position (g2p2g.rs:280-309)
    float ref2[3], w2[3][3];
    int t2[3] = {0, 0, 0};
    Mat3 A = mat_zero(); // affine = mass * G - (volume0 * inv_d * dt) * stress
    float mom[3] = {0.f, 0.f, 0.f}, psi_mass = 0.f, psi_mom = 0.f;
    if (has_particle && do_p2g) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float q = roundf(spk_div_h(p.x[a], h, inv_h));
            ref2[a] = (q - 1.f) * h - p.x[a];
            quad_weights(spk_div_h(-ref2[a], h, inv_h), w2[a]);
            const float mvf = q - cell_before[a];
            const int mv = (mvf != mvf) ? 0 : (int)__float2ll_rz(mvf);
            t2[a] = tcell[a] + mv; // in [0, 5] after the velocity clamp
            do_p2g = do_p2g && t2[a] >= 0 && t2[a] <= 5;
        }
        const float ks = p.volume0 * inv_d * dt;
#pragma unroll
        for (int i = 0; i < 9; ++i) A.m[i] = p.mass * G.m[i] - ks * stress.m[i];
#pragma unroll
        for (int a = 0; a < 3; ++a) mom[a] = p.mass * vel[a];
        psi_mass = (p.phase > 0.f && !(p.flags & SPK_F_FAILED)) ? p.mass : 0.f;
        psi_mom = psi_mass * p.psi_pos;
    }

    uint32_t out_slot = tid; // slot inside the dispatch entry's range of the next buffer
    if (MODE == 1) {
        // ---- counting sort of the CTA's particles by stencil-base cell (bin 216 = particles without P2G)
        const int bin = (has_particle && do_p2g) ? (t2[0] + 6 * t2[1] + 36 * t2[2]) : NBINS;
        int rank = 0;
        if (has_particle) rank = atomicAdd(&s_bin_cnt[bin], 1);
        __syncthreads();
        if (tid < 32) { // warp 0: exclusive scan of 217 counters (7 per lane) + list of the occupied cell bins
            int c[7], sum = 0, nz = 0;
#pragma unroll
            for (int k = 0; k < 7; ++k) { c[k] = s_bin_cnt[tid * 7 + k]; sum += c[k]; nz += (c[k] != 0 && tid * 7 + k < NBINS) ? 1 : 0; }
            int inc = sum, inz = nz;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, inc, d), oz = __shfl_up_sync(0xffffffffu, inz, d);
                if (tid >= d) { inc += o; inz += oz; }
            }
            int run = inc - sum, runz = inz - nz;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                s_bin_start[tid * 7 + k] = run;
                run += c[k];
                if (c[k] != 0 && tid * 7 + k < NBINS) s_binlist[runz++] = (unsigned short)(tid * 7 + k);
            }
            if (tid == 31) *s_nnz = inz;
        }
        __syncthreads();
        out_slot = (uint32_t)(s_bin_start[bin] + rank);
    }

    // ---- store the particle into the next buffer (coalesced up to the in-CTA cell sort) -----------------------------
    if (has_particle) {
        const uint32_t di = ent.first + out_slot;
        P.dst.a0[di] = make_float4(p.x[0], p.x[1], p.x[2], __uint_as_float(p.flags));
        P.dst.a1[di] = make_float4(p.F.m[0], p.F.m[1], p.F.m[2], p.F.m[3]);
        P.dst.a2[di] = make_float4(p.F.m[4], p.F.m[5], p.F.m[6], p.F.m[7]);
        P.dst.a3[di] = make_float4(p.F.m[8], p.jp, p.mass, p.volume0);
        P.dst.a4[di] = make_float4(p.phase, p.psi_pos, p.elastic_hardening, p.plastic_hardening);
        P.dst.a5[di] = make_float4(p.log_vol_gain, p.nacc_alpha, __uint_as_float(p.orig_id), 0.f);
        P.vel_out[di] = make_float4(vel[0], vel[1], vel[2], 0.f);
    }

    if (MODE == 0) {
        // ---- P2G, one particle per thread, 27 x 6 shared float atomics (g2p2g.rs:311-349 without the spin lock)
        if (has_particle && do_p2g) {
#pragma unroll
            for (int dz = 0; dz < 3; ++dz) {
                const int bz = cell_part(t2[2] + dz, 2) + oct_part(t2[2] + dz, 2) * 384;
                const float pz = ref2[2] + (float)dz * h;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const int by = bz + cell_part(t2[1] + dy, 1) + oct_part(t2[1] + dy, 1) * 384;
                    const float py = ref2[1] + (float)dy * h;
                    const float wyz = w2[1][dy] * w2[2][dz];
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const int base = by + (cell_part(t2[0] + dx, 0) ^ ((t2[2] + dz) & 2)) + oct_part(t2[0] + dx, 0) * 384; // SPK_ACC_SWZ
                        const float px = ref2[0] + (float)dx * h;
                        const float wt = w2[0][dx] * wyz;
                        float* a = &s_acc[base];
                        atomicAdd(a + 0 * 64, wt * p.mass);
                        atomicAdd(a + 1 * 64, wt * (M_(A, 0, 0) * px + M_(A, 0, 1) * py + M_(A, 0, 2) * pz + mom[0]));
                        atomicAdd(a + 2 * 64, wt * (M_(A, 1, 0) * px + M_(A, 1, 1) * py + M_(A, 1, 2) * pz + mom[1]));
                        atomicAdd(a + 3 * 64, wt * (M_(A, 2, 0) * px + M_(A, 2, 1) * py + M_(A, 2, 2) * pz + mom[2]));
                        if (psi_mass != 0.f) {
                            atomicAdd(a + 4 * 64, wt * psi_mom);
                            atomicAdd(a + 5 * 64, wt * psi_mass);
                        }
                    }
                }
            }
        }
    } else {
        // ---- stage the payload at the particle's cell-sorted slot:
        //   q0 = {m, c.xyz}  c = A ref2 + mom          q1..q3 = {h A[:,k], (psi_mass, psi_mom, 0)}
        //   q4 = {wx0, wx1, wx2, 0}   q5 = {wy0, wy1, wy2, wz0}   q6 = {wz1, wz2, 0, 0}
        if (has_particle && do_p2g) {
            float4* q = reinterpret_cast<float4*>(s_pay + (size_t)out_slot * PAY_FLOATS);
            const float cx = M_(A, 0, 0) * ref2[0] + M_(A, 0, 1) * ref2[1] + M_(A, 0, 2) * ref2[2] + mom[0];
            const float cy = M_(A, 1, 0) * ref2[0] + M_(A, 1, 1) * ref2[1] + M_(A, 1, 2) * ref2[2] + mom[1];
            const float cz = M_(A, 2, 0) * ref2[0] + M_(A, 2, 1) * ref2[1] + M_(A, 2, 2) * ref2[2] + mom[2];
            q[0] = make_float4(p.mass, cx, cy, cz);
            q[1] = make_float4(h * M_(A, 0, 0), h * M_(A, 1, 0), h * M_(A, 2, 0), psi_mass);
            q[2] = make_float4(h * M_(A, 0, 1), h * M_(A, 1, 1), h * M_(A, 2, 1), psi_mom);
            q[3] = make_float4(h * M_(A, 0, 2), h * M_(A, 1, 2), h * M_(A, 2, 2), 0.f);
            q[4] = make_float4(w2[0][0], w2[0][1], w2[0][2], 0.f);
            q[5] = make_float4(w2[1][0], w2[1][1], w2[1][2], w2[2][0]);
            q[6] = make_float4(w2[2][1], w2[2][2], 0.f, 0.f);
        }
        __syncthreads();
        // ---- stage 2: 9-lane groups, lane = stencil column (sy, sz), 3 nodes (sx = 0..2) per lane in registers
        const int lane = tid & 31, warp = tid >> 5;
        const int grp = lane / 9, l9 = lane - grp * 9;
        const int sy = l9 % 3, sz = l9 / 3;
        const float fsy = (float)sy, fsz = (float)sz;
        const int nnz = *s_nnz;
        if (grp < 3) {
            for (int e = warp * 3 + grp; e < nnz; e += (G2P2G_THREADS / 32) * 3) {
                const int bin = s_binlist[e];
                const int start = s_bin_start[bin], cnt = s_bin_cnt[bin];
                float aM[3] = {0.f, 0.f, 0.f}, aX[3] = {0.f, 0.f, 0.f}, aY[3] = {0.f, 0.f, 0.f}, aZ[3] = {0.f, 0.f, 0.f}, aPm[3] = {0.f, 0.f, 0.f},
                      aPp[3] = {0.f, 0.f, 0.f};
                for (int i = 0; i < cnt; ++i) {
                    const float* pay = s_pay + (size_t)(start + i) * PAY_FLOATS;
                    const float4 q0 = *reinterpret_cast<const float4*>(pay), q1 = *reinterpret_cast<const float4*>(pay + 4),
                                 q2 = *reinterpret_cast<const float4*>(pay + 8), q3 = *reinterpret_cast<const float4*>(pay + 12),
                                 q4 = *reinterpret_cast<const float4*>(pay + 16);
                    const float wyz = pay[20 + sy] * pay[23 + sz];
                    // node momentum contribution at sx = 0: c + sy * hA[:,1] + sz * hA[:,2]
                    float bx = q0.y + fsy * q2.x + fsz * q3.x, by = q0.z + fsy * q2.y + fsz * q3.y, bz = q0.w + fsy * q2.z + fsz * q3.z;
                    const float wx[3] = {q4.x, q4.y, q4.z};
#pragma unroll
                    for (int sx = 0; sx < 3; ++sx) {
                        const float wt = wx[sx] * wyz;
                        aM[sx] += wt * q0.x;
                        aX[sx] += wt * bx; aY[sx] += wt * by; aZ[sx] += wt * bz;
                        aPm[sx] += wt * q1.w; aPp[sx] += wt * q2.w;
                        bx += q1.x; by += q1.y; bz += q1.z;
                    }
                }
                // flush the 3 x 6 sums of this (cell, column) into the tile
                const int bxc = bin % 6, byc = (bin / 6) % 6, bzc = bin / 36;
                const int byz = cell_part(byc + sy, 1) + oct_part(byc + sy, 1) * 384 + cell_part(bzc + sz, 2) + oct_part(bzc + sz, 2) * 384;
                const int zsw = (bzc + sz) & 2; // SPK_ACC_SWZ
#pragma unroll
                for (int sx = 0; sx < 3; ++sx) {
                    float* a = &s_acc[byz + (cell_part(bxc + sx, 0) ^ zsw) + oct_part(bxc + sx, 0) * 384];
                    atomicAdd(a + 0 * 64, aM[sx]);
                    atomicAdd(a + 1 * 64, aX[sx]);
                    atomicAdd(a + 2 * 64, aY[sx]);
                    atomicAdd(a + 3 * 64, aZ[sx]);
                    if (aPp[sx] != 0.f) atomicAdd(a + 4 * 64, aPp[sx]);
                    if (aPm[sx] != 0.f) atomicAdd(a + 5 * 64, aPm[sx]);
                }
            }
        }
    }

    // ---- stage 3: flush the accumulation tile into the next grid (transfer_shared_blocks_to_grid, g2p2g.rs:476-553)
    fence_proxy_async(); // make the generic-proxy smem writes visible to the async (TMA) proxy
    __syncthreads();
    if (tid == 0) {
        for (int o = 0; o < 8; ++o) {
            const int nid = s_nbr_self[o];
            if (nid >= 0) tma_reduce_add_f32_1d(P.acc + (size_t)nid * (SPK_ACC_CHANNELS * 64), &s_acc[o * SPK_ACC_CHANNELS * 64], SPK_ACC_CHANNELS * 64 * 4);
        }
        // fused halo push (multi-GPU): the same tile chunks are bulk-reduced straight into the peers' halo accumulators,
        // addressed by the peer's own block ids — the transfer rides on NVLink while other CTAs still compute
        bool pushed = false;
        if (P.world > 1) {
            for (int o = 0; o < 8; ++o) {
                const int nid = s_nbr_self[o];
                if (nid < 0 || !(s_halo[o] & 1u)) continue;
                for (uint32_t r = 0; r < P.world; ++r) {
                    const int32_t pid = P.peer_ids[(size_t)nid * P.world + r];
                    if (pid >= 0) {
                        tma_reduce_add_f32_1d(P.peer_hacc[r] + (size_t)pid * (SPK_ACC_CHANNELS * 64), &s_acc[o * SPK_ACC_CHANNELS * 64], SPK_ACC_CHANNELS * 64 * 4);
                        pushed = true;
                    }
                }
            }
        }
        if (*s_dtmin < SPK_DT_RAW_NONE) atomicMin(P.dt_raw_bits, *s_dtmin);
        // Remote reductions must have landed before the peer is told so.  The telling is done by the barrier kernel that FOLLOWS this
        // kernel on the stream (system fence, then the flag): a kernel is complete only when its bulk operations are, so the CTA
        // only has to wait until its shared memory has been read (push_wait = 1 restores the per-CTA wait for completion, A/B)
        if (pushed && P.push_wait) {
#ifndef SPK_EMU
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
            __threadfence_system();
        } else {
            tma_commit_and_wait_read();
        }
    }
}


// ================================================================================================================================
// MODE 2 (default): the cell-coherent kernel.  The sort orders particles by (block, cell) (spk_sort.cu), a dispatch entry holds up to
// 512 particles of one block, and a CTA of 512 threads processes it with one particle per thread:
//   * a warp's 32 consecutive particles sit in a handful of cells, so the 27 velocity loads of G2P are shared-memory broadcasts;
//   * every particle owns a 36-float slot in shared memory.  The tail of its record (F, volume, phase, plasticity state: 80 bytes)
//     is brought there by cp.async while G2P runs and is read field by field when a phase needs it; the interpolated velocity and
//     velocity gradient are parked there across the SVD / return mapping / stress evaluation; at the end the slot holds the P2G
//     payload (20 floats).  Registers only ever hold what the current phase computes on (64 per thread, 2 CTAs = 32 warps per SM);
//   * the particle is stored to the next buffer as soon as its fields are final (nothing of it stays in registers across the
//     in-CTA cell sort, which only moves 16-bit slot numbers);
//   * the per-CTA fixed work (tile staging, zeroing, scan, flush: 8 TMA loads + 8 TMA reduce-adds) is paid once per 512 particles and
//     the P2G bins hold twice the particles of a 256-particle chunk (the per-bin flush is the expensive part of P2G).
#ifndef SPK_PSI_SHARE
#define SPK_PSI_SHARE 1 /* 1: a CTA whose particles all have psi_mass == mass accumulates the mass channel once and flushes it twice */
#endif
#ifndef SPK_P2G_LANES8
#define SPK_P2G_LANES8 0 /* 1: four 8-lane P2G groups per warp (quarter-warp aligned payload broadcasts); 0: three 9-lane groups */
#endif
#ifndef SPK_CELL_MIN_CTAS
#define SPK_CELL_MIN_CTAS (1024 / SPK_CELL_THREADS)
#endif
#ifndef SPK_CELL_ALIAS
#define SPK_CELL_ALIAS (SPK_CELL_THREADS < 512) /* 1: the accumulation tile reuses the shared memory of the velocity tile (zeroed after G2P) */
#endif
constexpr int C_THREADS = SPK_CELL_THREADS;   // == particles per dispatch entry
// per-particle shared-memory slot, stored as 9 arrays of float4 indexed by the particle's thread ([k][tid]: every per-thread access is
// a conflict-free 16-byte access): arrays 0-4 hold the record tail (a1..a5), later the P2G payload; arrays 5-8 the parked G, velocity, ref2
constexpr int C_TAIL = 5, C_PARKED = 4;
constexpr int CS_VEL = 0;                                    // float4[512]
constexpr int CS_ACC = SPK_CELL_ALIAS ? CS_VEL : CS_VEL + 512 * 16;   // float[8][6][64]
constexpr int CS_SLOT = CS_ACC + 8 * 6 * 64 * 4;             // float4[9][C_THREADS]
constexpr int CS_BINCNT = CS_SLOT + (C_TAIL + C_PARKED) * C_THREADS * 16;  // int[224]
constexpr int CS_BINSTART = CS_BINCNT + 224 * 4;             // int[224]
constexpr int CS_BINLIST = CS_BINSTART + 224 * 4;            // u16[224] occupied bins
constexpr int CS_SORTED = CS_BINLIST + 224 * 2;              // u16[512] slot numbers in bin order
constexpr int CS_MISC = CS_SORTED + C_THREADS * 2;           // mbarrier (8) + nbr_self[8] + nbr_prev[8] + nnz + dtmin + next work item
constexpr int CS_HALO = CS_MISC + 8 + 32 + 32 + 16;
constexpr int CS_WTOT = CS_HALO + 32;                        // int2[8]: per-warp totals of the bin scan
constexpr int CS_TOTAL = CS_WTOT + 64;
// the streaming variant (one CTA walks many dispatch entries): a second accumulation tile (the flush of entry k is still being read
// by the TMA engine while entry k + 1 accumulates) and the prefetched dispatch record of the next entry
constexpr int CS_ENTN = (CS_TOTAL + 127) / 128 * 128;                            // DispatchEntry (16) + int[16] block ids of the next entry's tile
constexpr int CS_ACC2 = CS_ENTN + 128;                       // float[8][6][64]
constexpr int CS_PARAMS = CS_ACC2 + 8 * 6 * 64 * 4;          // G2P2GParams (the kernel parameters, see the top of the kernel)
constexpr int CS_TOTAL_STREAM = CS_PARAMS + 384;
static_assert(CS_ENTN % 16 == 0 && CS_ACC2 % 128 == 0 && CS_PARAMS % 16 == 0 && sizeof(G2P2GParams) <= 384, "cp.async / TMA alignment of the streaming variant's extra shared memory");

__device__ __forceinline__ void prefetch_l2(const void* p) {
#ifndef SPK_EMU
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// the velocity gradient of a particle where the cell kernel parks it: G.m[i] = float i of the float4 sequence base[0], base[stride], ...
struct ParkedGradient {
    const float4* base;
    __device__ __forceinline__ float operator[](int i) const { // i is a compile-time constant after unrolling
        const float4 v = base[(i >> 2) * SPK_CELL_THREADS];
        return (i & 3) == 0 ? v.x : ((i & 3) == 1 ? v.y : ((i & 3) == 2 ? v.z : v.w));
    }
};

#ifdef SPK_EMU
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) { memcpy(smem_dst, gmem_src, 16); }
__device__ __forceinline__ void cp_async_wait_all() {}
__device__ __forceinline__ void cp_async_commit() {}
template <int N> __device__ __forceinline__ void cp_async_wait_group() {}
__device__ __forceinline__ void tma_commit_group() {}
template <int N> __device__ __forceinline__ void tma_wait_group_read() {}
#else
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
#endif

// STREAM: the CTA is persistent — it walks the dispatch entries blockIdx.x, blockIdx.x + gridDim.x, ... and, while entry k is in its
// scan / P2G / flush phases, already brings in entry k + 1: its dispatch record and tile block ids (cp.async at the top of k), its
// velocity tile (TMA, as soon as every thread has finished G2P of k), the sort permutation and the positions (cp.async into the
// parking arrays, idle during P2G) and, once the payloads of k are consumed, the record tails.  The ~3 serial memory round trips that
// open a CTA's life (dispatch -> tile ids -> tile, dispatch -> permutation -> particle) are then paid once per CTA, not once per entry.
template <class Updater, bool STREAM>
__global__ void __launch_bounds__(C_THREADS, SPK_CELL_MIN_CTAS) g2p2g_cell_kernel(const __grid_constant__ G2P2GParams P) {
#ifdef SPK_EMU
    unsigned char* smem = emu::dynamic_smem();
#else
    extern __shared__ __align__(128) unsigned char smem[];
#endif
    float4* s_vel = reinterpret_cast<float4*>(smem + CS_VEL);
    float* s_acc = reinterpret_cast<float*>(smem + CS_ACC); // (STREAM: re-pointed at the top of every entry)
    DispatchEntry* s_entn = reinterpret_cast<DispatchEntry*>(smem + CS_ENTN);
    volatile uint32_t* s_state = reinterpret_cast<volatile uint32_t*>(smem + CS_ENTN + 96); // STREAM: {entry, turn, first particle, count} of the current entry
    int* s_nbrn = reinterpret_cast<int*>(smem + CS_ENTN + 16);
    float4* s_tail = reinterpret_cast<float4*>(smem + CS_SLOT);                  // [5][512]
    float4* s_park = s_tail + C_TAIL * C_THREADS;                                // [4][512]
    int* s_bin_cnt = reinterpret_cast<int*>(smem + CS_BINCNT);
    int* s_bin_start = reinterpret_cast<int*>(smem + CS_BINSTART);
    unsigned short* s_binlist = reinterpret_cast<unsigned short*>(smem + CS_BINLIST);
    unsigned short* s_sorted = reinterpret_cast<unsigned short*>(smem + CS_SORTED);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + CS_MISC);
    int* s_nbr_self = reinterpret_cast<int*>(smem + CS_MISC + 8);
    int* s_nbr_prev = reinterpret_cast<int*>(smem + CS_MISC + 40);
    int* s_nnz = reinterpret_cast<int*>(smem + CS_MISC + 72);
    unsigned* s_dtmin = reinterpret_cast<unsigned*>(smem + CS_MISC + 76);
    int* s_next = reinterpret_cast<int*>(smem + CS_MISC + 80);
    int* s_psi_generic = reinterpret_cast<int*>(smem + CS_MISC + 84); // 1: some particle of this CTA has psi_mass != mass (phase <= 0)
    uint32_t* s_halo = reinterpret_cast<uint32_t*>(smem + CS_HALO);
    int2* s_wtot = reinterpret_cast<int2*>(smem + CS_WTOT);

    const int tid = threadIdx.x;
    // STREAM: the kernel parameters are read from a shared-memory copy.  Inside the entry loop the compiler would otherwise keep every
    // pointer of the parameter block in registers across the whole body (loop-invariant loads), i.e. spill them: shared-memory loads
    // cannot be moved across the barriers of the loop and are re-issued where they are used
    if (STREAM) {
        if (tid < (int)(sizeof(G2P2GParams) / 4)) reinterpret_cast<uint32_t*>(smem + CS_PARAMS)[tid] = reinterpret_cast<const uint32_t*>(&P)[tid];
        __syncthreads();
    }
    const G2P2GParams& Q = STREAM ? *reinterpret_cast<const G2P2GParams*>(smem + CS_PARAMS) : P;
    // the dispatch entry this CTA works on.  STREAM keeps it (and the turn counter) in shared memory, re-read where needed: loop-carried
    // registers would be live across the whole particle update, which has none to spare
    const DispatchEntry ent0 = Q.dispatch[blockIdx.x];
    if (STREAM && tid == 0) { s_state[0] = blockIdx.x; s_state[1] = 0; s_state[2] = ent0.first; s_state[3] = ent0.count; }
#define CUR_E (STREAM ? s_state[0] : blockIdx.x)
#define CUR_ITER (STREAM ? s_state[1] : 0u)
#define CUR_FIRST (STREAM ? s_state[2] : ent0.first)
#define CUR_COUNT (STREAM ? s_state[3] : ent0.count)
#define NEXT_VALID (STREAM && CUR_E + gridDim.x < Q.ndispatch)
    // ---- stage 0: velocity tile of the 2x2x2 blocks (transfer_global_blocks_to_shared_memory, g2p2g.rs:353-474).  Warp 0 reads the
    // block ids of the tile (one coalesced 64-byte load that depends on blockIdx only) and launches the 8 TMA bulk copies by itself,
    // before the CTA-wide barrier: the tile travels while the other warps zero the accumulators and fetch their particles
    if (tid < 32) {
        int nid = -1;
        if (tid < 16) {
            nid = Q.dispatch_nbr[(size_t)blockIdx.x * 16 + tid];
            s_nbr_self[tid] = nid; // s_nbr_prev follows s_nbr_self
        }
        __syncwarp();
        if (tid == 0) {
            mbar_init(s_bar, 1);
            *s_dtmin = SPK_DT_RAW_NONE;
            *s_next = 0;
            *s_psi_generic = 0;
            uint32_t bytes = 0;
            for (int o = 0; o < 8; ++o) bytes += s_nbr_prev[o] >= 0 ? 1024u : 0u;
            fence_proxy_async();
            mbar_expect_tx(s_bar, bytes);
            for (int o = 0; o < 8; ++o)
                if (s_nbr_prev[o] >= 0) tma_load_1d(&s_vel[o * 64], Q.prev_vel + (size_t)s_nbr_prev[o] * 64, 1024u, s_bar);
        }
        // (multi-GPU: the halo bits of the tile's blocks are only read by the flush; fetched AFTER the tile copies are on their way, so
        // that this dependent load does not sit on the critical path of every CTA)
        if (tid < 8) s_halo[tid] = (Q.world > 1 && nid >= 0) ? Q.blk_halo[nid] : 0u;
    }
    bool has_particle = (uint32_t)tid < ent0.count;
    float4* const tail = s_tail + tid; // tail[k * C_THREADS]: a(k+1) of this particle, later payload vector k
    float4* const park = s_park + tid;
    // ---- the particle: position now (registers), the rest of the record on its way into the slot (cp.async, no registers)
    float x[3] = {0.f, 0.f, 0.f};
    uint32_t flags = 0;
    if (has_particle) {
        const uint32_t si = Q.perm[ent0.first + tid];
        cp_async16(tail + 0 * C_THREADS, &Q.src.a1[si]);
        cp_async16(tail + 1 * C_THREADS, &Q.src.a2[si]);
        cp_async16(tail + 2 * C_THREADS, &Q.src.a3[si]);
        cp_async16(tail + 3 * C_THREADS, &Q.src.a4[si]);
        cp_async16(tail + 4 * C_THREADS, &Q.src.a5[si]);
        const float4 r0 = Q.src.a0[si];
        x[0] = r0.x; x[1] = r0.y; x[2] = r0.z; flags = __float_as_uint(r0.w);
    }
    if (STREAM) cp_async_commit(); // group "tails of the first entry"
    static_assert(!(STREAM && SPK_CELL_ALIAS), "the streaming variant keeps its accumulation tiles apart from the velocity tile");
    if (!SPK_CELL_ALIAS)
        for (int i = tid; i < (STREAM ? 2 : 1) * 8 * SPK_ACC_CHANNELS * 64 / 4; i += C_THREADS) {
            float4* a = reinterpret_cast<float4*>(i < 768 ? smem + CS_ACC : smem + CS_ACC2 - 768 * 16);
            a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    if (tid < 224) s_bin_cnt[tid] = 0;

    const float h = Q.h, inv_h = Q.inv_h, dt = Q.dt;
    const float inv_d = 4.f / (h * h);

#pragma unroll 1
    for (;;) { // ================================================================ one dispatch entry per turn (!STREAM: exactly one turn)
    __syncthreads();
    const uint32_t iter = CUR_ITER;
    if (STREAM) {
        s_acc = reinterpret_cast<float*>(smem + ((iter & 1u) ? CS_ACC2 : CS_ACC));
        has_particle = (uint32_t)tid < CUR_COUNT;
    }
    for (int i = tid; i < SPK_TILE_CELLS; i += C_THREADS)
        if (s_nbr_prev[i >> 6] < 0) s_vel[i] = make_float4(0.f, 0.f, 0.f, 0.f); // blocks absent from the previous grid read as zero
    if (STREAM) {
        // the dispatch record of the next entry (5 x 16 bytes), on its way while this entry is updated
        if (tid < 5 && NEXT_VALID) {
            const uint32_t e_next = CUR_E + gridDim.x;
            if (tid == 0) cp_async16(s_entn, &Q.dispatch[e_next]);
            else cp_async16(s_nbrn + 4 * (tid - 1), &Q.dispatch_nbr[(size_t)e_next * 16 + 4 * (tid - 1)]);
        }
        cp_async_commit(); // group "next record" (empty for most threads: every thread commits the same sequence of groups)
        // this turn's accumulation tile was flushed two turns ago; thread 0 made sure at the end of the last turn that the TMA engine
        // has read it (the first two turns use the tiles zeroed in the prologue)
        if (iter >= 2)
            for (int i = tid; i < 8 * SPK_ACC_CHANNELS * 64 / 4; i += C_THREADS) reinterpret_cast<float4*>(s_acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (STREAM && iter > 0) {
        // position and flags of this entry's particle were prefetched into the first parking array during the previous P2G
        cp_async_wait_group<2>(); // all but the two youngest groups ("tails of this entry", "next record"): the positions have landed
        const float4 r0 = park[0];
        x[0] = r0.x; x[1] = r0.y; x[2] = r0.z; flags = __float_as_uint(r0.w);
        if (!has_particle) { x[0] = x[1] = x[2] = 0.f; flags = 0; }
    }

    int tcell[3] = {1, 1, 1};
    float cell_before[3] = {0.f, 0.f, 0.f};
    float ref[3] = {0.f, 0.f, 0.f}, w[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float q = roundf(spk_div_h(x[a], h, inv_h));
        cell_before[a] = q;
        ref[a] = (q - 1.f) * h - x[a];                  // dir_to_associated_grid_node, particle.rs:37-47
        quad_weights(spk_div_h(-ref[a], h, inv_h), w[a]);
        const int assoc_cell = (q != q) ? -2 : (int)__float2ll_rz(q) - 2;
        tcell[a] = (assoc_cell & 3) + 1;                 // associated_cell_index_in_block_off_by_two, particle.rs:28-35
    }
    if (!STREAM && Q.prefetch_dist) {
        // While this CTA waits for its tile anyway: warm the L2 with the particle records of the dispatch entry that will start about one
        // CTA lifetime from now (the CTAs resident right now cover entries up to blockIdx + prefetch_dist).  Its dependent chain
        // dispatch -> permutation -> record then ends in L2 hits instead of DRAM round trips.
        const uint32_t e2 = blockIdx.x + Q.prefetch_dist;
        if (e2 < Q.ndispatch) {
            const DispatchEntry ent2 = Q.dispatch[e2];
            if ((uint32_t)tid < ent2.count) {
                const uint32_t s2 = Q.perm[ent2.first + tid];
                prefetch_l2(&Q.src.a0[s2]); prefetch_l2(&Q.src.a1[s2]); prefetch_l2(&Q.src.a2[s2]);
                prefetch_l2(&Q.src.a3[s2]); prefetch_l2(&Q.src.a4[s2]); prefetch_l2(&Q.src.a5[s2]);
            }
        }
    }
    mbar_wait(s_bar, iter & 1u); // the tile is complete (TMA) ...
    __syncthreads();             // ... and the absent octants are zero-filled

    int bin = NBINS;
    unsigned cand_bits = SPK_DT_RAW_NONE;
    if (has_particle) {
        const DevModel& model = Q.models[flags >> SPK_F_MODEL_SHIFT];
        float vel[3];
        Mat3 G;
        // ---- G2P (particle_g2p2g, g2p2g.rs:192-245): rows along x are accumulated separably, then folded into the 12 sums
        //   V = sum w v,  G[:,a] = inv_d * sum w dpt_a v   with w = wx wy wz and dpt = ref + s h
        {
            float V[3] = {0.f, 0.f, 0.f}, Gx[3] = {0.f, 0.f, 0.f}, Gy[3] = {0.f, 0.f, 0.f}, Gz[3] = {0.f, 0.f, 0.f};
            const float wpx[3] = {w[0][0] * ref[0], w[0][1] * (ref[0] + h), w[0][2] * (ref[0] + 2.f * h)};
#pragma unroll
            for (int dz = 0; dz < 3; ++dz) {
                const int iz = cell_part(tcell[2] + dz, 2) | (oct_part(tcell[2] + dz, 2) << 6);
                const float pz = ref[2] + (float)dz * h;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const int iy = iz | cell_part(tcell[1] + dy, 1) | (oct_part(tcell[1] + dy, 1) << 6);
                    float R0[3] = {0.f, 0.f, 0.f}, R1[3] = {0.f, 0.f, 0.f};
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const float4 c = s_vel[iy | cell_part(tcell[0] + dx, 0) | (oct_part(tcell[0] + dx, 0) << 6)];
                        R0[0] += w[0][dx] * c.x; R0[1] += w[0][dx] * c.y; R0[2] += w[0][dx] * c.z;
                        R1[0] += wpx[dx] * c.x; R1[1] += wpx[dx] * c.y; R1[2] += wpx[dx] * c.z;
                    }
                    const float wyz = w[1][dy] * w[2][dz];
                    const float wpy = wyz * (ref[1] + (float)dy * h), wpz = wyz * pz;
#pragma unroll
                    for (int a = 0; a < 3; ++a) { V[a] += wyz * R0[a]; Gx[a] += wyz * R1[a]; Gy[a] += wpy * R0[a]; Gz[a] += wpz * R0[a]; }
                }
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                vel[a] = V[a];
                M_(G, a, 0) = Gx[a] * inv_d; M_(G, a, 1) = Gy[a] * inv_d; M_(G, a, 2) = Gz[a] * inv_d;
            }
        }
        cp_async_wait_all(); // this thread's record tail is in its slot
        const uint32_t di = CUR_FIRST + (uint32_t)tid; // slot order == sorted order of this substep; the next sort re-orders by cell
        ParticleState p;
        p.x[0] = x[0]; p.x[1] = x[1]; p.x[2] = x[2]; p.flags = flags;
        Mat3 stress = mat_zero();
        bool do_p2g;
        float ref2[3], f2[3];
        if constexpr (updater_is_phased<Updater>::value) {
            // phase 1: advection and F update; position, velocity and velocity gradient are final afterwards: stored / parked
            {
                const float4 r1 = tail[0], r2 = tail[C_THREADS];
                p.F.m[0] = r1.x; p.F.m[1] = r1.y; p.F.m[2] = r1.z; p.F.m[3] = r1.w;
                p.F.m[4] = r2.x; p.F.m[5] = r2.y; p.F.m[6] = r2.z; p.F.m[7] = r2.w;
                p.F.m[8] = tail[2 * C_THREADS].x;
                p.orig_id = __float_as_uint(tail[4 * C_THREADS].z);
            }
            Updater::advect(model, Q.cold, h, dt, p, vel, G, Q.user);
            bool inside = true;
            int t2[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) { // P2G geometry with the UPDATED position (g2p2g.rs:280-309)
                const float q = roundf(spk_div_h(p.x[a], h, inv_h));
                ref2[a] = (q - 1.f) * h - p.x[a];
                const float mvf = q - cell_before[a];
                const int mv = (mvf != mvf) ? 0 : (int)__float2ll_rz(mvf);
                t2[a] = tcell[a] + mv; // in [0, 5] after the velocity clamp
                inside = inside && t2[a] >= 0 && t2[a] <= 5;
            }
            if (inside) bin = t2[0] + 6 * t2[1] + 36 * t2[2];
            Q.dst.a0[di] = make_float4(p.x[0], p.x[1], p.x[2], __uint_as_float(p.flags));
            Q.vel_out[di] = make_float4(vel[0], vel[1], vel[2], 0.f);
            park[0] = make_float4(G.m[0], G.m[1], G.m[2], G.m[3]);
            park[C_THREADS] = make_float4(G.m[4], G.m[5], G.m[6], G.m[7]);
            park[2 * C_THREADS] = make_float4(G.m[8], vel[0], vel[1], vel[2]);
            park[3 * C_THREADS] = make_float4(ref2[0], ref2[1], ref2[2], 0.f);
            spk_compiler_fence();
            // phase 2: the decomposition, alone in the register file
            Svd3 svd;
            if (Updater::needs_decomposition(model)) svd = Updater::decompose(model, p.F, Q.user);
            spk_compiler_fence();
            // phase 3: plasticity, failure, energies, stress on the scalars of the record (read from the slot only now)
            {
                const float4 r3 = tail[2 * C_THREADS], r4 = tail[3 * C_THREADS], r5 = tail[4 * C_THREADS];
                p.jp = r3.y; p.mass = r3.z; p.volume0 = r3.w;
                p.phase = r4.x; p.psi_pos = r4.y; p.elastic_hardening = r4.z; p.plastic_hardening = r4.w;
                p.log_vol_gain = r5.x; p.nacc_alpha = r5.y; p.orig_id = __float_as_uint(r5.z);
            }
            const uint32_t flags0 = p.flags;
            do_p2g = Updater::constitutive(model, p, svd, ParkedGradient{park}, stress, Q.user);
            if (p.flags != flags0) Q.dst.a0[di].w = __uint_as_float(p.flags); // the particle failed in phase 3 (rare)
            spk_compiler_fence();
            {
                const float4 g0 = park[0], g1 = park[C_THREADS], g2 = park[2 * C_THREADS], g3 = park[3 * C_THREADS];
                G.m[0] = g0.x; G.m[1] = g0.y; G.m[2] = g0.z; G.m[3] = g0.w; G.m[4] = g1.x; G.m[5] = g1.y; G.m[6] = g1.z; G.m[7] = g1.w; G.m[8] = g2.x;
                vel[0] = g2.y; vel[1] = g2.z; vel[2] = g2.w;
                ref2[0] = g3.x; ref2[1] = g3.y; ref2[2] = g3.z;
            }
        } else {
            {
                const float4 r1 = tail[0], r2 = tail[C_THREADS], r3 = tail[2 * C_THREADS], r4 = tail[3 * C_THREADS], r5 = tail[4 * C_THREADS];
                p.F.m[0] = r1.x; p.F.m[1] = r1.y; p.F.m[2] = r1.z; p.F.m[3] = r1.w;
                p.F.m[4] = r2.x; p.F.m[5] = r2.y; p.F.m[6] = r2.z; p.F.m[7] = r2.w;
                p.F.m[8] = r3.x; p.jp = r3.y; p.mass = r3.z; p.volume0 = r3.w;
                p.phase = r4.x; p.psi_pos = r4.y; p.elastic_hardening = r4.z; p.plastic_hardening = r4.w;
                p.log_vol_gain = r5.x; p.nacc_alpha = r5.y; p.orig_id = __float_as_uint(r5.z);
            }
            do_p2g = Updater::update(model, Q.cold, h, dt, p, vel, G, stress, Q.user);
            bool inside = true;
            int t2[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float q = roundf(spk_div_h(p.x[a], h, inv_h));
                ref2[a] = (q - 1.f) * h - p.x[a];
                const float mvf = q - cell_before[a];
                const int mv = (mvf != mvf) ? 0 : (int)__float2ll_rz(mvf);
                t2[a] = tcell[a] + mv;
                inside = inside && t2[a] >= 0 && t2[a] <= 5;
            }
            if (inside) bin = t2[0] + 6 * t2[1] + 36 * t2[2];
            Q.dst.a0[di] = make_float4(p.x[0], p.x[1], p.x[2], __uint_as_float(p.flags));
            Q.vel_out[di] = make_float4(vel[0], vel[1], vel[2], 0.f);
        }
        if (!do_p2g) bin = NBINS;
        // ---- fused K15 (timestep.rs:38-113): un-clamped dt candidate of this particle for the NEXT substep; the clamps and the
        // fixed-point conversion are monotone, so they are applied once to the reduced value (spk_stage_estimate_dt)
        if (!(p.flags & SPK_F_FAILED)) {
            const float cand = Updater::dt_candidate(model, h, p, vel, Q.user);
            cand_bits = __float_as_uint(cand != cand ? 1.0e30f : fminf(fmaxf(cand, 0.f), 1.0e30f)); // NaN imposes no bound (Rust f32::min ignores NaN)
        }
        Q.dst.a1[di] = make_float4(p.F.m[0], p.F.m[1], p.F.m[2], p.F.m[3]);
        Q.dst.a2[di] = make_float4(p.F.m[4], p.F.m[5], p.F.m[6], p.F.m[7]);
        Q.dst.a3[di] = make_float4(p.F.m[8], p.jp, p.mass, p.volume0);
        Q.dst.a4[di] = make_float4(p.phase, p.psi_pos, p.elastic_hardening, p.plastic_hardening);
        Q.dst.a5[di] = make_float4(p.log_vol_gain, p.nacc_alpha, __uint_as_float(p.orig_id), 0.f);
        // ---- P2G payload (overwrites the record tail in the slot)
        //   q0 = {m, c.xyz}  c = A ref2 + m v        q1 = {h A[:,0], psi_mass}   q2 = {h A[:,1], psi_mom}   q3 = {h A[:,2], fy}
        //   q4 = {wx0, wx1, wx2, fz}                 A = m G - (volume0 inv_d dt) stress,  fy / fz: B-spline arguments along y / z
        if (bin != NBINS) {
            const float ks = p.volume0 * inv_d * dt;
            Mat3 A;
#pragma unroll
            for (int i = 0; i < 9; ++i) A.m[i] = p.mass * G.m[i] - ks * stress.m[i];
            const float psi_mass = (p.phase > 0.f && !(p.flags & SPK_F_FAILED)) ? p.mass : 0.f;
            if (!SPK_PSI_SHARE || psi_mass != p.mass) *s_psi_generic = 1; // (benign race: every writer stores 1)
#pragma unroll
            for (int a = 0; a < 3; ++a) f2[a] = spk_div_h(-ref2[a], h, inv_h);
            float wx[3];
            quad_weights(f2[0], wx);
            tail[0] = make_float4(p.mass, M_(A, 0, 0) * ref2[0] + M_(A, 0, 1) * ref2[1] + M_(A, 0, 2) * ref2[2] + p.mass * vel[0],
                               M_(A, 1, 0) * ref2[0] + M_(A, 1, 1) * ref2[1] + M_(A, 1, 2) * ref2[2] + p.mass * vel[1],
                               M_(A, 2, 0) * ref2[0] + M_(A, 2, 1) * ref2[1] + M_(A, 2, 2) * ref2[2] + p.mass * vel[2]);
            tail[C_THREADS] = make_float4(h * M_(A, 0, 0), h * M_(A, 1, 0), h * M_(A, 2, 0), psi_mass);
            tail[2 * C_THREADS] = make_float4(h * M_(A, 0, 1), h * M_(A, 1, 1), h * M_(A, 2, 1), psi_mass * p.psi_pos);
            tail[3 * C_THREADS] = make_float4(h * M_(A, 0, 2), h * M_(A, 1, 2), h * M_(A, 2, 2), f2[1]);
            tail[4 * C_THREADS] = make_float4(wx[0], wx[1], wx[2], f2[2]);
        }
    }
    cand_bits = __reduce_min_sync(0xffffffffu, cand_bits);
    if ((tid & 31) == 0 && cand_bits < SPK_DT_RAW_NONE) atomicMin(s_dtmin, cand_bits);
    int rank = 0;
    if (has_particle) rank = atomicAdd(&s_bin_cnt[bin], 1);
    if (STREAM && tid < 5) cp_async_wait_all(); // the next entry's dispatch record is in shared memory (its 5 loaders make sure)
    __syncthreads();
    if (SPK_CELL_ALIAS) // every thread is done with the velocity tile: its shared memory becomes the (zeroed) accumulation tile
        for (int i = tid; i < 8 * SPK_ACC_CHANNELS * 64 / 4; i += C_THREADS) reinterpret_cast<float4*>(s_acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    // ---- STREAM: every thread is done with the velocity tile -> the next entry's tile may overwrite it; its particles are looked up
    uint32_t si_next = 0;
    bool has_next = false;
    if (NEXT_VALID) {
        if (tid == 0) {
            uint32_t bytes = 0;
            for (int o = 0; o < 8; ++o) bytes += s_nbrn[8 + o] >= 0 ? 1024u : 0u;
            fence_proxy_async();
            mbar_expect_tx(s_bar, bytes);
            for (int o = 0; o < 8; ++o)
                if (s_nbrn[8 + o] >= 0) tma_load_1d(&s_vel[o * 64], Q.prev_vel + (size_t)s_nbrn[8 + o] * 64, 1024u, s_bar);
        }
        has_next = (uint32_t)tid < s_entn->count;
        if (has_next) si_next = Q.perm[s_entn->first + tid];
    }

    // ---- counting sort of the slots by stencil-base cell (bin 216 = particles without P2G): exclusive scan of the 217 counters by
    // 7 warps (one bin per thread: warp scan, then the per-warp totals), and the list of the occupied bins in bin order
    {
        const int lane = tid & 31, warp = tid >> 5;
        int c = 0, inc = 0, nzb = 0;
        bool occ = false;
        if (tid < 224) {
            c = s_bin_cnt[tid];
            occ = c != 0 && tid < NBINS;
            const unsigned ball = __ballot_sync(0xffffffffu, occ);
            nzb = __popc(ball & spk_lanemask_lt());
            inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += o;
            }
            if (lane == 31) s_wtot[warp] = make_int2(inc, __popc(ball));
        }
        __syncthreads();
        if (tid < 224) {
            int base = 0, nzbase = 0;
#pragma unroll
            for (int wv = 0; wv < 7; ++wv) {
                const int2 t = s_wtot[wv];
                if (wv < warp) { base += t.x; nzbase += t.y; }
            }
            if (tid == 223) *s_nnz = nzbase + s_wtot[6].y;
            s_bin_start[tid] = base + inc - c;
            if (occ) s_binlist[nzbase + nzb] = (unsigned short)tid;
        }
    }
    __syncthreads();
    if (has_particle) s_sorted[s_bin_start[bin] + rank] = (unsigned short)tid;
    if (STREAM) { // the parking arrays are idle from here to the next entry's update: they receive its positions and keep its permutation
        if (has_next) {
            cp_async16(park, &Q.src.a0[si_next]);
            reinterpret_cast<uint32_t*>(park + C_THREADS)[0] = si_next;
        }
        reinterpret_cast<uint32_t*>(park + C_THREADS)[1] = has_next ? 1u : 0u;
        cp_async_commit(); // group "positions of the next entry"
    }
    __syncthreads();

#if SPK_P2G_LANES8
    // ---- stage 2: 8-lane groups (a quarter warp each: the 128-bit payload reads of a group are ONE shared-memory wavefront), one occupied
    // bin at a time, four bins per warp.  Lane l = stencil column (sy, sz) = (l % 3, l / 3) for l = 0..7 with its 3 nodes (sx = 0..2) x 6
    // channels in registers over the bin's particles; the ninth column (2, 2) is shared out: lanes 0..2 each carry ONE more node of it
    // (sx = l).  One flush of the 27 x 6 sums into the tile per bin.
    {
        const int lane = tid & 31;
        const int grp = lane >> 3, l8 = lane & 7;
        const int sy = l8 % 3, sz = l8 / 3;
        const float fsy = (float)sy, fsz = (float)sz, fl8 = (float)l8;
        // QuadraticKernel::eval_all (kernel.rs:15-21) for ONE node: w_s(x) = A_s (x - B_s)^2 + C_s
        const float Ay = sy == 1 ? -1.f : 0.5f, By = sy == 0 ? 1.5f : (sy == 1 ? 1.f : 0.5f), Cy = sy == 1 ? 0.75f : 0.f;
        const float Az = sz == 1 ? -1.f : 0.5f, Bz = sz == 0 ? 1.5f : (sz == 1 ? 1.f : 0.5f), Cz = sz == 1 ? 0.75f : 0.f;
        const int nnz = *s_nnz;
        // psi_mass == mass for every particle of the CTA (no failed / phase <= 0 particle: the usual case): the psi_mass sums are then
        // the mass sums, term by term — they are not accumulated a second time, the flush sends the mass channel to both places
        const bool psi_generic = *s_psi_generic != 0;
        for (;;) { // a warp takes the next four occupied bins until the list is exhausted
            int e = 0;
            if (lane == 0) e = atomicAdd(s_next, 4);
            e = __shfl_sync(0xffffffffu, e, 0) + grp;
            if (e - grp >= nnz) break;
            if (e >= nnz) continue;
            const int b = s_binlist[e];
            const int start = s_bin_start[b], cnt = s_bin_cnt[b];
            float aM[3] = {0.f, 0.f, 0.f}, aX[3] = {0.f, 0.f, 0.f}, aY[3] = {0.f, 0.f, 0.f}, aZ[3] = {0.f, 0.f, 0.f}, aPm[3] = {0.f, 0.f, 0.f},
                  aPp[3] = {0.f, 0.f, 0.f};
            float eM = 0.f, eX = 0.f, eY = 0.f, eZ = 0.f, ePm = 0.f, ePp = 0.f; // the extra node (sx = l8, 2, 2) of lanes 0..2
            for (int i = 0; i < cnt; ++i) {
                const float4* pay = s_tail + s_sorted[start + i];
                const float4 q0 = pay[0], q1 = pay[C_THREADS], q2 = pay[2 * C_THREADS], q3 = pay[3 * C_THREADS], q4 = pay[4 * C_THREADS];
                const float dy = q3.w - By, dz = q4.w - Bz;
                const float wyz = fmaf(Ay * dy, dy, Cy) * fmaf(Az * dz, dz, Cz);
                // node momentum contribution at sx = 0: c + sy * hA[:,1] + sz * hA[:,2]
                float bx = q0.y + fsy * q2.x + fsz * q3.x, by = q0.z + fsy * q2.y + fsz * q3.y, bz = q0.w + fsy * q2.z + fsz * q3.z;
                const float wx[3] = {q4.x, q4.y, q4.z};
#pragma unroll
                for (int sx = 0; sx < 3; ++sx) {
                    const float wt = wx[sx] * wyz;
                    aM[sx] += wt * q0.x;
                    aX[sx] += wt * bx; aY[sx] += wt * by; aZ[sx] += wt * bz;
                    aPp[sx] += wt * q2.w;
                    if (psi_generic) aPm[sx] += wt * q1.w;
                    bx += q1.x; by += q1.y; bz += q1.z;
                }
                // the shared-out column (sy, sz) = (2, 2): weight 0.5 (fy - 0.5)^2 * 0.5 (fz - 0.5)^2, node sx = l8 (weight 0 for lanes 3..7)
                const float ey = q3.w - 0.5f, ez = q4.w - 0.5f;
                const float wte = (l8 == 0 ? q4.x : (l8 == 1 ? q4.y : (l8 == 2 ? q4.z : 0.f))) * ((0.5f * (ey * ey)) * (0.5f * (ez * ez)));
                eM += wte * q0.x;
                eX += wte * (q0.y + 2.f * (q2.x + q3.x) + fl8 * q1.x);
                eY += wte * (q0.z + 2.f * (q2.y + q3.y) + fl8 * q1.y);
                eZ += wte * (q0.w + 2.f * (q2.z + q3.z) + fl8 * q1.z);
                ePp += wte * q2.w;
                if (psi_generic) ePm += wte * q1.w;
            }
            const int bxc = b % 6, byc = (b / 6) % 6, bzc = b / 36;
            {
                const int byz = cell_part(byc + sy, 1) + oct_part(byc + sy, 1) * 384 + cell_part(bzc + sz, 2) + oct_part(bzc + sz, 2) * 384;
                const int zsw = (bzc + sz) & 2; // SPK_ACC_SWZ: bit 1 of the node's z inside its block flips bit 1 of its x
#pragma unroll
                for (int sx = 0; sx < 3; ++sx) {
                    float* a = &s_acc[byz + (cell_part(bxc + sx, 0) ^ zsw) + oct_part(bxc + sx, 0) * 384];
                    atomicAdd(a + 0 * 64, aM[sx]);
                    atomicAdd(a + 1 * 64, aX[sx]);
                    atomicAdd(a + 2 * 64, aY[sx]);
                    atomicAdd(a + 3 * 64, aZ[sx]);
                    if (aPp[sx] != 0.f) atomicAdd(a + 4 * 64, aPp[sx]);
                    if (psi_generic && aPm[sx] != 0.f) atomicAdd(a + 5 * 64, aPm[sx]);
                }
            }
            if (l8 < 3) {
                const int byz = cell_part(byc + 2, 1) + oct_part(byc + 2, 1) * 384 + cell_part(bzc + 2, 2) + oct_part(bzc + 2, 2) * 384;
                const int zsw = (bzc + 2) & 2;
                float* a = &s_acc[byz + (cell_part(bxc + l8, 0) ^ zsw) + oct_part(bxc + l8, 0) * 384];
                atomicAdd(a + 0 * 64, eM);
                atomicAdd(a + 1 * 64, eX);
                atomicAdd(a + 2 * 64, eY);
                atomicAdd(a + 3 * 64, eZ);
                if (ePp != 0.f) atomicAdd(a + 4 * 64, ePp);
                if (psi_generic && ePm != 0.f) atomicAdd(a + 5 * 64, ePm);
            }
        }
    }
#else
    // ---- stage 2: 9-lane groups, one occupied bin at a time; lane = stencil column (sy, sz), 3 nodes (sx = 0..2) x 6 channels in
    // registers over the bin's particles (payload reads are broadcasts), then one flush of the 18 sums into the tile
    {
        const int lane = tid & 31;
        const int grp = lane / 9, l9 = lane - grp * 9;
        const int sy = l9 % 3, sz = l9 / 3;
        const float fsy = (float)sy, fsz = (float)sz;
        // QuadraticKernel::eval_all (kernel.rs:15-21) for ONE node: w_s(x) = A_s (x - B_s)^2 + C_s
        const float Ay = sy == 1 ? -1.f : 0.5f, By = sy == 0 ? 1.5f : (sy == 1 ? 1.f : 0.5f), Cy = sy == 1 ? 0.75f : 0.f;
        const float Az = sz == 1 ? -1.f : 0.5f, Bz = sz == 0 ? 1.5f : (sz == 1 ? 1.f : 0.5f), Cz = sz == 1 ? 0.75f : 0.f;
        const int nnz = *s_nnz;
        // psi_mass == mass for every particle of the CTA (no failed / phase <= 0 particle: the usual case): the psi_mass sums are then
        // the mass sums, term by term — they are not accumulated a second time, the flush sends the mass channel to both places
        const bool psi_generic = *s_psi_generic != 0;
        if (grp < 3) {
            for (;;) { // a warp takes the next three occupied bins (in bin order: mostly x-neighbours) until the list is exhausted
                int e = 0;
                if (lane == 0) e = atomicAdd(s_next, 3);
                e = __shfl_sync(0x07ffffffu, e, 0) + grp;
                if (e - grp >= nnz) break;
                if (e >= nnz) continue;
                const int b = s_binlist[e];
                const int start = s_bin_start[b], cnt = s_bin_cnt[b];
                float aM[3] = {0.f, 0.f, 0.f}, aX[3] = {0.f, 0.f, 0.f}, aY[3] = {0.f, 0.f, 0.f}, aZ[3] = {0.f, 0.f, 0.f}, aPm[3] = {0.f, 0.f, 0.f},
                      aPp[3] = {0.f, 0.f, 0.f};
                for (int i = 0; i < cnt; ++i) {
                    const float4* pay = s_tail + s_sorted[start + i];
                    const float4 q0 = pay[0], q1 = pay[C_THREADS], q2 = pay[2 * C_THREADS], q3 = pay[3 * C_THREADS], q4 = pay[4 * C_THREADS];
                    const float dy = q3.w - By, dz = q4.w - Bz;
                    const float wyz = fmaf(Ay * dy, dy, Cy) * fmaf(Az * dz, dz, Cz);
                    // node momentum contribution at sx = 0: c + sy * hA[:,1] + sz * hA[:,2]
                    float bx = q0.y + fsy * q2.x + fsz * q3.x, by = q0.z + fsy * q2.y + fsz * q3.y, bz = q0.w + fsy * q2.z + fsz * q3.z;
                    const float wx[3] = {q4.x, q4.y, q4.z};
#pragma unroll
                    for (int sx = 0; sx < 3; ++sx) {
                        const float wt = wx[sx] * wyz;
                        aM[sx] += wt * q0.x;
                        aX[sx] += wt * bx; aY[sx] += wt * by; aZ[sx] += wt * bz;
                        aPp[sx] += wt * q2.w;
                        if (psi_generic) aPm[sx] += wt * q1.w;
                        bx += q1.x; by += q1.y; bz += q1.z;
                    }
                }
                const int bxc = b % 6, byc = (b / 6) % 6, bzc = b / 36;
                const int byz = cell_part(byc + sy, 1) + oct_part(byc + sy, 1) * 384 + cell_part(bzc + sz, 2) + oct_part(bzc + sz, 2) * 384;
                const int zsw = (bzc + sz) & 2; // SPK_ACC_SWZ: bit 1 of the node's z inside its block flips bit 1 of its x
#pragma unroll
                for (int sx = 0; sx < 3; ++sx) {
                    float* a = &s_acc[byz + (cell_part(bxc + sx, 0) ^ zsw) + oct_part(bxc + sx, 0) * 384];
                    atomicAdd(a + 0 * 64, aM[sx]);
                    atomicAdd(a + 1 * 64, aX[sx]);
                    atomicAdd(a + 2 * 64, aY[sx]);
                    atomicAdd(a + 3 * 64, aZ[sx]);
                    if (aPp[sx] != 0.f) atomicAdd(a + 4 * 64, aPp[sx]);
                    if (psi_generic && aPm[sx] != 0.f) atomicAdd(a + 5 * 64, aPm[sx]);
                }
            }
        }
    }

#endif

    // ---- stage 3: flush the accumulation tile into the next grid (transfer_shared_blocks_to_grid, g2p2g.rs:476-553)
    fence_proxy_async(); // make the generic-proxy smem writes visible to the async (TMA) proxy
    __syncthreads();
    if (tid == 0) {
        // one block chunk of the tile -> one block chunk of an accumulator array (this rank's grid or a peer's halo accumulator);
        // when the psi_mass channel was not accumulated separately it is the mass channel, sent a second time
        const bool psi_generic = *s_psi_generic != 0;
        auto reduce_chunk = [&](float* dst, int o) {
            const float* src = &s_acc[o * SPK_ACC_CHANNELS * 64];
            if (psi_generic) {
                tma_reduce_add_f32_1d(dst, src, SPK_ACC_CHANNELS * 64 * 4);
            } else {
                tma_reduce_add_f32_1d(dst, src, 5 * 64 * 4);
                tma_reduce_add_f32_1d(dst + 5 * 64, src, 64 * 4);
            }
        };
        for (int o = 0; o < 8; ++o) {
            const int nid = s_nbr_self[o];
            if (nid >= 0) reduce_chunk(Q.acc + (size_t)nid * (SPK_ACC_CHANNELS * 64), o);
        }
        bool pushed = false;
        if (Q.world > 1) { // fused halo push (multi-GPU), see g2p2g_kernel
            for (int o = 0; o < 8; ++o) {
                const int nid = s_nbr_self[o];
                if (nid < 0 || !(s_halo[o] & 1u)) continue;
                for (uint32_t r = 0; r < Q.world; ++r) {
                    const int32_t pid = Q.peer_ids[(size_t)nid * Q.world + r];
                    if (pid >= 0) {
                        reduce_chunk(Q.peer_hacc[r] + (size_t)pid * (SPK_ACC_CHANNELS * 64), o);
                        pushed = true;
                    }
                }
            }
        }
        if (!STREAM) {
            if (*s_dtmin < SPK_DT_RAW_NONE) atomicMin(Q.dt_raw_bits, *s_dtmin);
            if (pushed && Q.push_wait) {
#ifndef SPK_EMU
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
                __threadfence_system();
            } else {
                tma_commit_and_wait_read();
            }
        } else {
            // one bulk group per entry; the group of the PREVIOUS entry (the other accumulation tile) must have been read before that
            // tile is zeroed again at the top of the next turn (and everything must have been read before the CTA retires)
            tma_commit_group();
            if (NEXT_VALID) {
                tma_wait_group_read<1>();
            } else {
                if (*s_dtmin < SPK_DT_RAW_NONE) atomicMin(Q.dt_raw_bits, *s_dtmin);
                if (pushed && Q.push_wait) {
#ifndef SPK_EMU
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
                    __threadfence_system();
                } else {
                    tma_wait_group_read<0>();
                }
            }
        }
    }
    if (!NEXT_VALID) break;
    // ---- STREAM: the payloads of this entry are consumed; the next entry takes over
    if (reinterpret_cast<const uint32_t*>(park + C_THREADS)[1]) {
        const uint32_t si = reinterpret_cast<const uint32_t*>(park + C_THREADS)[0];
        cp_async16(tail + 0 * C_THREADS, &Q.src.a1[si]);
        cp_async16(tail + 1 * C_THREADS, &Q.src.a2[si]);
        cp_async16(tail + 2 * C_THREADS, &Q.src.a3[si]);
        cp_async16(tail + 3 * C_THREADS, &Q.src.a4[si]);
        cp_async16(tail + 4 * C_THREADS, &Q.src.a5[si]);
    }
    cp_async_commit(); // group "tails of the next entry"
    if (tid < 224) s_bin_cnt[tid] = 0;
    if (tid < 32) {
        __syncwarp(); // thread 0 has issued the flush with this entry's block ids
        int nid = -1;
        if (tid < 16) { nid = s_nbrn[tid]; s_nbr_self[tid] = nid; }
        if (tid < 8) s_halo[tid] = (Q.world > 1 && nid >= 0) ? Q.blk_halo[nid] : 0u;
        if (tid == 0) { *s_next = 0; *s_psi_generic = 0; }
    }
    __syncthreads(); // every thread has read the current entry's state (NEXT_VALID above) before thread 0 moves it on
    if (tid == 0) { s_state[0] = s_state[0] + gridDim.x; s_state[1] = s_state[1] + 1u; s_state[2] = s_entn->first; s_state[3] = s_entn->count; }
    } // ======================================================================================================== next dispatch entry
}

#undef CUR_E
#undef CUR_ITER
#undef CUR_FIRST
#undef CUR_COUNT
#undef NEXT_VALID

} // namespace spk_g2p2g_detail

// DefaultCudaParticleKernelsLauncher::launch_g2p2g (src/cuda/cuda_particle_kernels.rs:64-99) for any updater; `user` reaches the
// updater's functions unchanged (device pointer or NULL)
template <class Updater>
inline cudaError_t spk_launch_g2p2g_with(spk_ctx* c, float dt, const void* user = nullptr) {
    using namespace spk_g2p2g_detail;
    GridBuffers& next = c->grid[1 - c->gcur];
    GridBuffers& prev = c->grid[c->gcur];
    const uint32_t nd = c->h_counters->ndispatch;
    if (nd == 0) { // never launch with 0 CTAs (cuda_mpm_pipeline.rs:485-486)
        // a rank of a multi-GPU run without particles: it still follows the same path as its peers in the next substep (fused dt
        // candidate — it contributes "no bound" — rather than the stand-alone dt kernel, which ends in a barrier of its own)
        if (c->dist.world > 1) { c->sorted_valid = false; c->dt_cache_valid = true; c->dt_raw_host_valid = false; c->counters_fresh = false; }
        return cudaSuccess;
    }
    G2P2GParams P;
    P.src = c->pb[c->cur];
    P.dst = c->pb[1 - c->cur];
    P.vel_out = c->vel;
    P.cold = c->cold;
    P.perm = c->perm;
    P.dispatch = c->dispatch;
    P.dispatch_nbr = c->dispatch_nbr;
    P.models = c->models;
    P.nbr_self = next.nbr_self;
    P.nbr_prev = next.nbr_prev;
    P.acc = next.acc;
    P.prev_vel = prev.vel;
    P.h = c->cell_width;
    P.inv_h = 1.0f / c->cell_width;
    P.dt = dt;
    P.ndispatch = nd;
    P.dt_raw_bits = &c->d_counters->dt_raw_bits;
    P.world = (c->dist.world > 1 && c->dist_fused_push) ? c->dist.world : 0;
    P.blk_halo = next.blk_halo;
    P.peer_ids = c->dist_peer_ids;
    P.peer_hacc = c->dist_push_ptrs;
    P.push_wait = (uint32_t)c->dist_push_wait;
    P.prefetch_dist = c->g2p2g_prefetch ? (uint32_t)(c->sm_count * (1024 / SPK_CELL_THREADS)) : 0u;
    P.user = user;
    // (the dt candidate word was reset to SPK_DT_RAW_NONE by this substep's sort: build_tables_kernel, after the counter record was published)
    static bool configured[64] = {}; // per instantiation and device: the > 48 KB dynamic shared memory opt-in is a per-device function attribute
    if (c->device < 0 || c->device >= 64 || !configured[c->device]) {
        cudaError_t e = cudaFuncSetAttribute(g2p2g_kernel<0, Updater>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(g2p2g_kernel<1, Updater>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(g2p2g_cell_kernel<Updater, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_TOTAL);
#if !SPK_CELL_ALIAS
        if (e == cudaSuccess) e = cudaFuncSetAttribute(g2p2g_cell_kernel<Updater, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_TOTAL_STREAM);
#endif
        if (e != cudaSuccess) return e;
        if (c->device >= 0 && c->device < 64) configured[c->device] = true;
        c->g2p2g_configured = true;
    }
    const bool timed = c->kev_on && c->kev_count < 256;
    if (timed) cudaEventRecord(c->kev[2 * c->kev_count], c->stream);
    if (c->g2p2g_mode == 0) g2p2g_kernel<0, Updater><<<nd, G2P2G_THREADS, SM_TOTAL, c->stream>>>(P);
    else if (c->g2p2g_mode == 1) g2p2g_kernel<1, Updater><<<nd, G2P2G_THREADS, SM_TOTAL, c->stream>>>(P);
    else if (!c->g2p2g_stream || SPK_CELL_ALIAS) g2p2g_cell_kernel<Updater, false><<<nd, C_THREADS, CS_TOTAL, c->stream>>>(P);
#if !SPK_CELL_ALIAS
    else { // persistent CTAs: two per SM (64 registers x 512 threads), each walking every (2 x SMs)-th dispatch entry
        const uint32_t resident = 2u * (uint32_t)c->sm_count;
        g2p2g_cell_kernel<Updater, true><<<nd < resident ? nd : resident, C_THREADS, CS_TOTAL_STREAM, c->stream>>>(P);
    }
#endif
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

namespace spk_g2p2g_detail {

// ---------------------------------------------------------------------------------------------------- dt (K15)
constexpr int DT_THREADS = 256;
__device__ __forceinline__ unsigned long long dt_from_sec(float dt) { // GpuTimestepLength::from_sec, gpu_timestep.rs:20-22
    const float v = dt * 1000000000000.0f;
    if (v != v) return 0ull;
    return __float2ull_rz(v); // saturating
}
template <class Updater>
__global__ void __launch_bounds__(DT_THREADS) estimate_dt_kernel(const float4* __restrict__ a0, const float4* __restrict__ a1, const float4* __restrict__ a3,
                                                                  const float4* __restrict__ a4, const float4* __restrict__ vel, uint32_t n,
                                                                  const DevModel* __restrict__ models, float h, float min_dt, float max_dt,
                                                                  DeviceCounters* counters, const void* user) {
    unsigned long long best = 0xffffffffffffffffull;
    for (uint32_t i = blockIdx.x * DT_THREADS + threadIdx.x; i < n; i += gridDim.x * DT_THREADS) {
        const uint32_t flags = __float_as_uint(a0[i].w);
        if (flags & SPK_F_FAILED) continue; // timestep.rs:72-74
        const DevModel& m = models[flags >> SPK_F_MODEL_SHIFT];
        const float4 r3 = a3[i], r4 = a4[i], v = vel[i];
        const float f00 = a1[i].x;
        float dt = max_dt;
        ParticleState p; // the fields a dt candidate may read (timestep.rs:38-70: status, volume, velocity)
        p.flags = flags; p.mass = r3.z; p.volume0 = r3.w; p.jp = r3.y; p.F.m[0] = f00; p.elastic_hardening = r4.z; p.plastic_hardening = r4.w;
        p.phase = r4.x; p.psi_pos = r4.y;
        const float pv[3] = {v.x, v.y, v.z};
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
        if (best != 0xffffffffffffffffull) atomicMin(&counters->dt_acc, best); // one atomic per CTA (the reference: one per particle)
    }
}

} // namespace spk_g2p2g_detail

// DefaultCudaParticleKernelsLauncher::launch_estimate_particle_timestep_length (cuda_particle_kernels.rs:38-62) for any updater
template <class Updater>
inline cudaError_t spk_launch_estimate_dt_with(spk_ctx* c, float min_dt, float max_dt, const void* user = nullptr) {
    using namespace spk_g2p2g_detail;
    const uint32_t n = (uint32_t)c->n;
    if (!n && !(c->dist.world > 1)) return cudaSuccess;
    const ParticleBuffers& pb = c->pb[c->cur];
    int blocks = (int)((n + DT_THREADS - 1) / DT_THREADS);
    const int maxb = c->sm_count * 8;
    if (blocks > maxb) blocks = maxb;
    *c->h_upload = c->dt_acc_host;
    { cudaError_t e = cudaMemcpyAsync(&c->d_counters->dt_acc, c->h_upload, 8, cudaMemcpyHostToDevice, c->stream); if (e != cudaSuccess) return e; }
    c->counters_fresh = false;
    if (!n) return cudaSuccess; // (multi-GPU rank without particles: the frame's dt is its entry in the minimum over the ranks)
    estimate_dt_kernel<Updater><<<blocks, DT_THREADS, 0, c->stream>>>(pb.a0, pb.a1, pb.a3, pb.a4, c->vel, n, c->models, c->cell_width, min_dt, max_dt, c->d_counters, user);
    c->launches += 1;
    return cudaGetLastError();
}
