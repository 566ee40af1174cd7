// spk_g2p2g.cu — the fused G2P -> particle update -> P2G kernel (replaces K16 `g2p2g`, src_kernels/cuda/g2p2g.rs:64-553
// and the DefaultParticleUpdater, src_kernels/cuda/particle_updater.rs:73-236) and the CFL reduction (K15,
// src_kernels/cuda/timestep.rs:13-113).
//
// One CTA processes one dispatch entry = up to `chunk` particles of ONE particle block (4x4x4 cells).
//   stage 0  the 2x2x2-block velocity tile (8 x 1 KB, contiguous per block) is staged in shared memory with TMA bulk
//            copies through the neighbour table (no hash probes); blocks absent from the previous grid read as zero
//            (g2p2g.rs:441-472).
//   stage 1  one thread per particle: gather through the sort permutation, 27-node G2P from the tile, constitutive
//            update in registers (one shared SVD), coalesced store into the NEXT particle buffer at the sorted index.
//   stage 2  P2G into a shared-memory accumulation tile laid out [octant][channel][cell] (see p2g_* below).
//   stage 3  the tile is flushed to the next grid with 8 TMA bulk reduce-adds (1.5 KB each, f32 add at L2).
#include "spk_device.cuh"
#include "spk_math.cuh"

namespace {

constexpr int G2P2G_THREADS = 256;

struct G2P2GParams {
    ParticleBuffers src, dst;
    float4* vel_out;
    const float4* cold;
    const uint32_t* perm;
    const DispatchEntry* dispatch;
    const DevModel* models;
    // next grid
    const uint32_t* blk_first;
    const uint32_t* blk_count;
    const int32_t* nbr_self;
    const int32_t* nbr_prev;
    float* acc;             // next grid accumulators [block][6][64]
    const float4* prev_vel; // previous grid velocities [block][64]
    float h, dt;
    uint32_t ndispatch;
};

// ---- TMA bulk helpers (sm_90+; SASS: UBLKCP / UBLKRED) ------------------------------------------------------------
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

// tile index helpers: tile coordinate t in [0,8) per axis -> (cell part, octant part) of the [octant][...][cell] layout
__device__ __forceinline__ int cell_part(int t, int axis) { return (t & 3) << (2 * axis); }
__device__ __forceinline__ int oct_part(int t, int axis) { return (t >> 2) << axis; }

// QuadraticKernel::eval_all, src_core/dynamics/solver/kernel.rs:15-21
__device__ __forceinline__ void quad_weights(float x, float w[3]) {
    w[0] = 0.5f * ((1.5f - x) * (1.5f - x));
    w[1] = 0.75f - ((x - 1.f) * (x - 1.f));
    w[2] = 0.5f * ((x - 0.5f) * (x - 0.5f));
}

__global__ void __launch_bounds__(G2P2G_THREADS, 2) g2p2g_kernel(const G2P2GParams P) {
    __shared__ __align__(128) float4 s_vel[SPK_TILE_CELLS];                  // [octant][cell]            8 KB
    __shared__ __align__(128) float s_acc[8 * SPK_ACC_CHANNELS * 64];        // [octant][channel][cell]  12 KB
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_nbr_self[8], s_nbr_prev[8];

    const int tid = threadIdx.x;
    const DispatchEntry ent = P.dispatch[blockIdx.x];
    if (tid < 8) {
        s_nbr_self[tid] = P.nbr_self[ent.block * 8 + tid];
        s_nbr_prev[tid] = P.nbr_prev[ent.block * 8 + tid];
    }
    if (tid == 0) mbar_init(&s_bar, 1);
    for (int i = tid; i < 8 * SPK_ACC_CHANNELS * 64; i += G2P2G_THREADS) s_acc[i] = 0.f;
    __syncthreads();
    // ---- stage 0: velocity tile
    if (tid == 0) {
        uint32_t bytes = 0;
        for (int o = 0; o < 8; ++o) bytes += s_nbr_prev[o] >= 0 ? 1024u : 0u;
        fence_proxy_async();
        mbar_expect_tx(&s_bar, bytes);
        for (int o = 0; o < 8; ++o)
            if (s_nbr_prev[o] >= 0) tma_load_1d(&s_vel[o * 64], P.prev_vel + (size_t)s_nbr_prev[o] * 64, 1024u, &s_bar);
    }
    for (int i = tid; i < SPK_TILE_CELLS; i += G2P2G_THREADS)
        if (s_nbr_prev[i >> 6] < 0) s_vel[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    mbar_wait(&s_bar, 0);
    __syncthreads();

    const float h = P.h, dt = P.dt;
    const float inv_d = 4.f / (h * h);

    // ---- stage 1 + 2: particles
    for (uint32_t k = tid; k < ent.count; k += G2P2G_THREADS) {
        const uint32_t di = ent.first + k;
        const uint32_t si = P.perm[di];
        ParticleState p;
        {
            const float4 r0 = P.src.a0[si], r1 = P.src.a1[si], r2 = P.src.a2[si], r3 = P.src.a3[si], r4 = P.src.a4[si], r5 = P.src.a5[si];
            p.x[0] = r0.x; p.x[1] = r0.y; p.x[2] = r0.z; p.flags = __float_as_uint(r0.w);
            p.F.m[0] = r1.x; p.F.m[1] = r1.y; p.F.m[2] = r1.z; p.F.m[3] = r1.w;
            p.F.m[4] = r2.x; p.F.m[5] = r2.y; p.F.m[6] = r2.z; p.F.m[7] = r2.w;
            p.F.m[8] = r3.x; p.jp = r3.y; p.mass = r3.z; p.volume0 = r3.w;
            p.phase = r4.x; p.psi_pos = r4.y; p.elastic_hardening = r4.z; p.plastic_hardening = r4.w;
            p.log_vol_gain = r5.x; p.nacc_alpha = r5.y; p.orig_id = __float_as_uint(r5.z);
        }
        const DevModel& model = P.models[p.flags >> SPK_F_MODEL_SHIFT];

        // ---- G2P (particle_g2p2g, g2p2g.rs:192-245)
        float ref[3], w[3][3];
        int tcell[3]; // tile coordinate of the first stencil node: cell_in_block + 1
        float cell_before[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float q = roundf(__fdiv_rn(p.x[a], h));
            cell_before[a] = q;
            ref[a] = (q - 1.f) * h - p.x[a];               // dir_to_associated_grid_node, particle.rs:37-47
            quad_weights(__fdiv_rn(-ref[a], h), w[a]);
            const int assoc_cell = (q != q) ? -2 : (int)__float2ll_rz(q) - 2;
            tcell[a] = (assoc_cell & 3) + 1;                // associated_cell_index_in_block_off_by_two, particle.rs:28-35
        }
        float vel[3] = {0.f, 0.f, 0.f};
        Mat3 G = mat_zero();
        float div = 0.f;
#pragma unroll
        for (int dz = 0; dz < 3; ++dz) {
            const int iz = cell_part(tcell[2] + dz, 2) | (oct_part(tcell[2] + dz, 2) << 6);
            const float pz = ref[2] + (float)dz * h;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int iy = iz | cell_part(tcell[1] + dy, 1) | (oct_part(tcell[1] + dy, 1) << 6);
                const float py = ref[1] + (float)dy * h;
                const float wyz = w[1][dy] * w[2][dz];
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int idx = iy | cell_part(tcell[0] + dx, 0) | (oct_part(tcell[0] + dx, 0) << 6);
                    const float px = ref[0] + (float)dx * h;
                    const float wt = w[0][dx] * wyz;
                    const float4 c = s_vel[idx];
                    vel[0] += wt * c.x; vel[1] += wt * c.y; vel[2] += wt * c.z;
                    const float wi = wt * inv_d;
                    const float gx = wi * c.x, gy = wi * c.y, gz = wi * c.z;
                    M_(G, 0, 0) += gx * px; M_(G, 0, 1) += gx * py; M_(G, 0, 2) += gx * pz;
                    M_(G, 1, 0) += gy * px; M_(G, 1, 1) += gy * py; M_(G, 1, 2) += gy * pz;
                    M_(G, 2, 0) += gz * px; M_(G, 2, 1) += gz * py; M_(G, 2, 2) += gz * pz;
                    div += wt * (c.x * px + c.y * py + c.z * pz) * inv_d;
                }
            }
        }

        // ---- particle update (particle_updater.rs:73-236)
        const bool is_fluid = model_is_fluid(model);
        if (p.flags & SPK_F_KINEMATIC) {
            const float4 kv = P.cold[p.orig_id];
            vel[0] = kv.x; vel[1] = kv.y; vel[2] = kv.z;
        }
        if (fabsf(vel[0]) * dt >= h || fabsf(vel[1]) * dt >= h || fabsf(vel[2]) * dt >= h) {
#pragma unroll
            for (int a = 0; a < 3; ++a) vel[a] = signum_f(vel[a]) * h / dt;
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) p.x[a] += vel[a] * dt;
        if (!is_fluid) {
            Mat3 Gdt;
#pragma unroll
            for (int i = 0; i < 9; ++i) Gdt.m[i] = G.m[i] * dt;
            const Mat3 dF = mat_mul(Gdt, p.F);
#pragma unroll
            for (int i = 0; i < 9; ++i) p.F.m[i] += dF.m[i];
        } else {
            p.F.m[0] += (div * dt) * p.F.m[0];
        }
        Svd3 svd;
        const bool need_svd = model_needs_svd(model);
        if (need_svd) {
            svd = svd3(p.F);
            plastic_update(model, p, svd);
        }
        if (p.flags & SPK_F_STATIC) {
            vel[0] = vel[1] = vel[2] = 0.f;
            G = mat_zero();
        }
        const float j = mat_det(p.F);
        const float density_def_grad = (p.mass / p.volume0) / j;
        bool do_p2g = true;
        if (density_def_grad == 0.f || (p.flags & SPK_F_FAILED) || (!is_fluid && fabsf(p.F.m[0]) > 1.0e4f)) {
            p.flags |= SPK_F_FAILED;
            p.F = mat_identity();
            do_p2g = false;
        }
        Mat3 stress = mat_zero();
        if (do_p2g) {
            // pos energy (gpu_constitutive_model.rs:27-48)
            float energy = 0.f;
            if (model.ctag == SPK_CONSTITUTIVE_COROTATED) { // elasticity_corotated_linear.rs:104-118
                const float a0 = fmaxf(svd.s[0] - 1.f, 0.f), a1 = fmaxf(svd.s[1] - 1.f, 0.f), a2 = fmaxf(svd.s[2] - 1.f, 0.f);
                const float pos_dev = model.c[2] * p.elastic_hardening * (a0 * a0 + a1 * a1 + a2 * a2);
                const float sph = model.c[1] * p.elastic_hardening / 2.f * (j - 1.f) * (j - 1.f);
                energy = j < 1.f ? pos_dev : pos_dev + sph;
            } else if (model.ctag == SPK_CONSTITUTIVE_NEO_HOOKEAN) { // elasticity_neo_hookean.rs:117-141
                const float pc = (1.f - 0.001f) * p.phase * p.phase + 0.001f;
                const float hard = p.elastic_hardening;
                const float kk = 2.f / 3.f * model.c[2] * hard + model.c[1] * hard;
                float tr = 0.f;
#pragma unroll
                for (int i = 0; i < 9; ++i) tr += p.F.m[i] * p.F.m[i];
                const float dev = hard * model.c[2] / 2.f * (tr * powf(j, -2.f / 3.f) - 3.f);
                const float vol = kk / 2.f * ((j * j - 1.f) / 2.f - logf(j));
                energy = j < 1.f ? dev * pc : (dev + vol) * p.phase;
            }
            p.psi_pos = fmaxf(p.psi_pos, energy);
            // stress (gpu_constitutive_model.rs:50-86); with a failure model the reference evaluates it twice (before/after
            // the phase change, particle_updater.rs:185-235)
            for (int pass = (model.ftag == SPK_FAILURE_MAXIMUM_STRESS ? 0 : 1); pass < 2; ++pass) {
                if (model.ctag == SPK_CONSTITUTIVE_COROTATED) stress = corotated_stress(model, p.phase, p.elastic_hardening, p.F, svd, j);
                else if (model.ctag == SPK_CONSTITUTIVE_NEO_HOOKEAN) stress = neo_hookean_stress(model, p.phase, p.elastic_hardening, p.F, j);
                else if (model.ctag == SPK_CONSTITUTIVE_MONAGHAN) { // eos_monaghan_sph.rs:44-63, quirk 7
                    const float density_fluid = (p.mass / p.volume0) / p.F.m[0];
                    const float pr = -monaghan_pressure(model, p.mass, p.volume0, density_fluid) * p.jp;
                    stress = mat_zero();
                    stress.m[0] = pr; stress.m[4] = pr; stress.m[8] = pr;
                    if (model.c[2] != 0.f) {
                        Mat3 sr;
#pragma unroll
                        for (int r = 0; r < 3; ++r)
#pragma unroll
                            for (int c = 0; c < 3; ++c) M_(sr, r, c) = (M_(G, r, c) + M_(G, c, r)) * 0.5f;
                        const float sph = (sr.m[0] + sr.m[4] + sr.m[8]) / 3.f;
                        sr.m[0] -= sph; sr.m[4] -= sph; sr.m[8] -= sph;
                        const float kv = 2.f * model.c[2] * p.jp;
#pragma unroll
                        for (int i = 0; i < 9; ++i) stress.m[i] += kv * sr.m[i];
                    }
                }
                if (pass == 0) { // failure_maximum_stress.rs:20-28
                    bool finite = true;
#pragma unroll
                    for (int i = 0; i < 9; ++i) finite = finite && isfinite(stress.m[i]);
                    if (finite) {
                        float mn, mx;
                        sym_eig_minmax(stress, mn, mx);
                        if (mx > model.f[0] || (mx - mn) / 2.f > model.f[1]) p.phase = 0.f;
                    }
                }
            }
        }

        // ---- store the particle at its sorted slot of the next buffer (coalesced)
        P.dst.a0[di] = make_float4(p.x[0], p.x[1], p.x[2], __uint_as_float(p.flags));
        P.dst.a1[di] = make_float4(p.F.m[0], p.F.m[1], p.F.m[2], p.F.m[3]);
        P.dst.a2[di] = make_float4(p.F.m[4], p.F.m[5], p.F.m[6], p.F.m[7]);
        P.dst.a3[di] = make_float4(p.F.m[8], p.jp, p.mass, p.volume0);
        P.dst.a4[di] = make_float4(p.phase, p.psi_pos, p.elastic_hardening, p.plastic_hardening);
        P.dst.a5[di] = make_float4(p.log_vol_gain, p.nacc_alpha, __uint_as_float(p.orig_id), 0.f);
        P.vel_out[di] = make_float4(vel[0], vel[1], vel[2], 0.f);

        // ---- P2G (g2p2g.rs:280-349)
        if (do_p2g) {
            float ref2[3], w2[3][3];
            int t2[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float q = roundf(__fdiv_rn(p.x[a], h));
                ref2[a] = (q - 1.f) * h - p.x[a];
                quad_weights(__fdiv_rn(-ref2[a], h), w2[a]);
                const float mvf = q - cell_before[a];
                const int mv = (mvf != mvf) ? 0 : (int)__float2ll_rz(mvf);
                t2[a] = tcell[a] + mv; // in [0, 5] after the velocity clamp
            }
            bool in_tile = true;
#pragma unroll
            for (int a = 0; a < 3; ++a) in_tile = in_tile && t2[a] >= 0 && t2[a] <= 5;
            if (in_tile) {
                Mat3 A; // affine = mass * G - (volume0 * inv_d * dt) * stress
                const float ks = p.volume0 * inv_d * dt;
#pragma unroll
                for (int i = 0; i < 9; ++i) A.m[i] = p.mass * G.m[i] - ks * stress.m[i];
                const float mom[3] = {p.mass * vel[0], p.mass * vel[1], p.mass * vel[2]};
                const float psi_mass = (p.phase > 0.f && !(p.flags & SPK_F_FAILED)) ? p.mass : 0.f;
                const float psi_mom = psi_mass * p.psi_pos;
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
                            const int base = by + cell_part(t2[0] + dx, 0) + oct_part(t2[0] + dx, 0) * 384;
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
        tma_commit_and_wait_read();
    }
}

// ---------------------------------------------------------------------------------------------------- dt (K15)
constexpr int DT_THREADS = 256;
__device__ __forceinline__ unsigned long long dt_from_sec(float dt) { // GpuTimestepLength::from_sec, gpu_timestep.rs:20-22
    const float v = dt * 1000000000000.0f;
    if (v != v) return 0ull;
    return __float2ull_rz(v); // saturating
}
__global__ void __launch_bounds__(DT_THREADS) estimate_dt_kernel(const float4* __restrict__ a0, const float4* __restrict__ a1, const float4* __restrict__ a3,
                                                                  const float4* __restrict__ a4, const float4* __restrict__ vel, uint32_t n,
                                                                  const DevModel* __restrict__ models, float h, float min_dt, float max_dt,
                                                                  DeviceCounters* counters) {
    unsigned long long best = 0xffffffffffffffffull;
    for (uint32_t i = blockIdx.x * DT_THREADS + threadIdx.x; i < n; i += gridDim.x * DT_THREADS) {
        const uint32_t flags = __float_as_uint(a0[i].w);
        if (flags & SPK_F_FAILED) continue; // timestep.rs:72-74
        const DevModel& m = models[flags >> SPK_F_MODEL_SHIFT];
        const float4 r3 = a3[i], r4 = a4[i], v = vel[i];
        const float f00 = a1[i].x;
        float dt = max_dt;
        dt = fminf(dt, particle_dt_candidate(m, h, r3.z, r3.w, f00, r4.z, v.x, v.y, v.z));
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

} // namespace

cudaError_t spk_launch_g2p2g(spk_ctx* c, float dt) {
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
    P.blk_first = next.blk_first;
    P.blk_count = next.blk_count;
    P.nbr_self = next.nbr_self;
    P.nbr_prev = next.nbr_prev;
    P.acc = next.acc;
    P.prev_vel = prev.vel;
    P.h = c->cell_width;
    P.dt = dt;
    P.ndispatch = nd;
    g2p2g_kernel<<<nd, G2P2G_THREADS, 0, c->stream>>>(P);
    c->launches += 1;
    c->g2p2g_launches += 1;
    c->cur = 1 - c->cur; // the next buffer now holds the particles, physically in sorted order
    c->sorted_valid = false;
    return cudaGetLastError();
}

cudaError_t spk_launch_estimate_dt(spk_ctx* c, float min_dt, float max_dt) {
    const uint32_t n = (uint32_t)c->n;
    if (!n) return cudaSuccess;
    const ParticleBuffers& pb = c->pb[c->cur];
    int blocks = (int)((n + DT_THREADS - 1) / DT_THREADS);
    const int maxb = c->sm_count * 8;
    if (blocks > maxb) blocks = maxb;
    estimate_dt_kernel<<<blocks, DT_THREADS, 0, c->stream>>>(pb.a0, pb.a1, pb.a3, pb.a4, c->vel, n, c->models, c->cell_width, min_dt, max_dt, c->d_counters);
    c->launches += 1;
    return cudaGetLastError();
}
