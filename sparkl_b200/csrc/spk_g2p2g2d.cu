// spk_g2p2g2d.cu — dim2 build of the fused G2P -> particle update -> P2G kernel (K16 `g2p2g`, src_kernels/cuda/g2p2g.rs:64-553
// with the `dim2` feature: 4x4-cell blocks, a 2x2-block / 8x8-node shared tile, 9-node stencil) and of the CFL reduction (K15,
// src_kernels/cuda/timestep.rs:13-113).  Compiled only into libsparkl_b200_2d.so (-DSPK_DIM=2); spk_g2p2g.cu is its 3D sibling.
//
// One CTA = one dispatch entry = up to 128 particles of ONE particle block.  Same data flow as the 3D kernel, sized for 2D:
//   stage 0  the 2x2-block velocity tile (4 x 256 B, contiguous per block) is staged with TMA bulk copies through the neighbour
//            table; blocks absent from the previous grid read as zero (g2p2g.rs:441-472)
//   stage 1  one thread per particle: gather through the sort permutation, 9-node G2P, constitutive update in registers (one
//            shared 2x2 SVD), coalesced store into the NEXT particle buffer at the sorted index
//   stage 2  P2G with shared-memory float atomics into a [quadrant][channel][cell] tile (9 nodes x 5 channels per particle:
//            2D blocks hold ~64 particles, the cell-binned register accumulation of the 3D kernel does not pay here)
//   stage 3  the tile is flushed with 4 TMA bulk reduce-adds (384 B each); the CFL candidate of the next substep rides along
#include "spk_device.cuh"
#include "spk_math2d.cuh"

#if SPK_DIM != 2
#error "spk_g2p2g2d.cu is the dim2 kernel (compile with -DSPK_DIM=2)"
#endif

namespace {

constexpr int G2P2G_THREADS = SPK_G2P2G_CHUNK;
constexpr int ACC_BLOCK = SPK_ACC_CHANNELS * SPK_CELLS_PER_BLOCK; // floats per block in `acc`: [6][16]

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
};

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

// tile coordinate t in [0, 8) per axis -> cell part (x + 4 y inside the block) and quadrant part (bit0 = x, bit1 = y)
__device__ __forceinline__ int cell_part(int t, int axis) { return (t & 3) << (2 * axis); }
__device__ __forceinline__ int quad_part(int t, int axis) { return (t >> 2) << axis; }

__device__ __forceinline__ void quad_weights(float x, float w[3]) { // QuadraticKernel::eval_all, kernel.rs:15-21
    w[0] = 0.5f * ((1.5f - x) * (1.5f - x));
    w[1] = 0.75f - ((x - 1.f) * (x - 1.f));
    w[2] = 0.5f * ((x - 0.5f) * (x - 0.5f));
}

__global__ void __launch_bounds__(G2P2G_THREADS) g2p2g2d_kernel(const G2P2GParams P) {
    __shared__ __align__(128) float4 s_vel[SPK_TILE_CELLS];             // [quadrant][cell]
    __shared__ __align__(128) float s_acc[SPK_NBR * ACC_BLOCK];         // [quadrant][channel][cell]
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
    for (int i = tid; i < SPK_NBR * ACC_BLOCK; i += G2P2G_THREADS) s_acc[i] = 0.f;
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

    const float h = P.h, inv_h = P.inv_h, dt = P.dt;
    const float inv_d = 4.f / (h * h); // Kernel::inv_d
    const bool has_particle = (uint32_t)tid < ent.count;

    // ---- stage 1: one particle per thread
    ParticleState2 p;
    float vel[2] = {0.f, 0.f};
    Mat2 G = mat2_zero(), stress = mat2_zero();
    bool do_p2g = false;
    int tcell[2] = {1, 1};
    float cell_before[2] = {0.f, 0.f};
    float ref[2] = {0.f, 0.f}, w[2][3];
    if (has_particle) {
        const uint32_t si = P.perm[ent.first + tid];
        const float4 r0 = P.src.a0[si], r1 = P.src.a1[si], r3 = P.src.a3[si], r4 = P.src.a4[si], r5 = P.src.a5[si];
        p.x[0] = r0.x; p.x[1] = r0.y; p.flags = __float_as_uint(r0.w);
        p.F.m[0] = r1.x; p.F.m[1] = r1.y; p.F.m[2] = r1.z; p.F.m[3] = r1.w;
        p.jp = r3.y; p.mass = r3.z; p.volume0 = r3.w;
        p.phase = r4.x; p.psi_pos = r4.y; p.elastic_hardening = r4.z; p.plastic_hardening = r4.w;
        p.log_vol_gain = r5.x; p.nacc_alpha = r5.y; p.orig_id = __float_as_uint(r5.z);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const float q = roundf(spk_div_h(p.x[a], h, inv_h));
            cell_before[a] = q;
            ref[a] = (q - 1.f) * h - p.x[a];                // dir_to_associated_grid_node, particle.rs:37-47
            quad_weights(spk_div_h(-ref[a], h, inv_h), w[a]);
            const int assoc_cell = (q != q) ? -2 : (int)__float2ll_rz(q) - 2;
            tcell[a] = (assoc_cell & 3) + 1;                 // associated_cell_index_in_block_off_by_two, particle.rs:28-35
        }
    }
    mbar_wait(&s_bar, 0);
    __syncthreads();

    if (has_particle) {
        const DevModel& model = P.models[p.flags >> SPK_F_MODEL_SHIFT];
        // ---- G2P (particle_g2p2g, g2p2g.rs:192-245): V = sum w v, G[:, a] = inv_d * sum w dpt_a v
        {
            float V[2] = {0.f, 0.f}, Gx[2] = {0.f, 0.f}, Gy[2] = {0.f, 0.f};
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int iy = cell_part(tcell[1] + dy, 1) | (quad_part(tcell[1] + dy, 1) << SPK_CELL_SHIFT);
                const float py = ref[1] + (float)dy * h;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float4 c = s_vel[iy | cell_part(tcell[0] + dx, 0) | (quad_part(tcell[0] + dx, 0) << SPK_CELL_SHIFT)];
                    const float wt = w[0][dx] * w[1][dy];
                    const float px = ref[0] + (float)dx * h;
                    V[0] += wt * c.x; V[1] += wt * c.y;
                    Gx[0] += (wt * px) * c.x; Gx[1] += (wt * px) * c.y;
                    Gy[0] += (wt * py) * c.x; Gy[1] += (wt * py) * c.y;
                }
            }
#pragma unroll
            for (int a = 0; a < 2; ++a) { vel[a] = V[a]; M2_(G, a, 0) = Gx[a] * inv_d; M2_(G, a, 1) = Gy[a] * inv_d; }
        }
        const float div = M2_(G, 0, 0) + M2_(G, 1, 1); // velocity_gradient_det

        // ---- particle update (particle_updater.rs:73-236)
        const bool is_fluid = model_is_fluid(model);
        if (p.flags & SPK_F_KINEMATIC) {
            const float4 kv = P.cold[p.orig_id];
            vel[0] = kv.x; vel[1] = kv.y;
        }
        if (fabsf(vel[0]) * dt >= h || fabsf(vel[1]) * dt >= h) {
#pragma unroll
            for (int a = 0; a < 2; ++a) vel[a] = signum_f(vel[a]) * h / dt;
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) p.x[a] += vel[a] * dt;
        if (!is_fluid) {
            Mat2 Gdt;
#pragma unroll
            for (int i = 0; i < 4; ++i) Gdt.m[i] = G.m[i] * dt;
            const Mat2 dF = mat2_mul(Gdt, p.F);
#pragma unroll
            for (int i = 0; i < 4; ++i) p.F.m[i] += dF.m[i];
        } else {
            p.F.m[0] += (div * dt) * p.F.m[0];
        }
        Svd2 svd;
        svd.u = mat2_identity(); svd.v = mat2_identity(); svd.s[0] = svd.s[1] = 1.f;
        if (model_needs_svd(model)) {
            svd = svd2(p.F);
            plastic_update2(model, p, svd);
        }
        if (p.flags & SPK_F_STATIC) {
            vel[0] = vel[1] = 0.f;
            G = mat2_zero();
        }
        const float j = mat2_det(p.F);
        const float density_def_grad = (p.mass / p.volume0) / j;
        do_p2g = true;
        if (density_def_grad == 0.f || (p.flags & SPK_F_FAILED) || (!is_fluid && fabsf(p.F.m[0]) > 1.0e4f)) {
            p.flags |= SPK_F_FAILED;
            p.F = mat2_identity();
            do_p2g = false;
        }
        if (do_p2g) {
            float energy = 0.f; // pos energy (gpu_constitutive_model.rs:27-48)
            if (model.ctag == SPK_CONSTITUTIVE_COROTATED) { // elasticity_corotated_linear.rs:104-118
                const float a0 = fmaxf(svd.s[0] - 1.f, 0.f), a1 = fmaxf(svd.s[1] - 1.f, 0.f);
                const float pos_dev = model.c[2] * p.elastic_hardening * (a0 * a0 + a1 * a1);
                const float sph = model.c[1] * p.elastic_hardening / 2.f * (j - 1.f) * (j - 1.f);
                energy = j < 1.f ? pos_dev : pos_dev + sph;
            } else if (model.ctag == SPK_CONSTITUTIVE_NEO_HOOKEAN) { // elasticity_neo_hookean.rs:117-141
                const float pc = (1.f - 0.001f) * p.phase * p.phase + 0.001f;
                const float hard = p.elastic_hardening;
                const float kk = 2.f / 3.f * model.c[2] * hard + model.c[1] * hard;
                float tr = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) tr += p.F.m[i] * p.F.m[i];
                const float dev = hard * model.c[2] / 2.f * (tr * powf(j, -2.f / 2.f) - 2.f);
                const float vol = kk / 2.f * ((j * j - 1.f) / 2.f - logf(j));
                energy = j < 1.f ? dev * pc : (dev + vol) * p.phase;
            }
            p.psi_pos = fmaxf(p.psi_pos, energy);
            // stress (gpu_constitutive_model.rs:50-86); with a failure model it is evaluated before and after the phase change
            for (int pass = (model.ftag == SPK_FAILURE_MAXIMUM_STRESS ? 0 : 1); pass < 2; ++pass) {
                if (model.ctag == SPK_CONSTITUTIVE_COROTATED) stress = corotated_stress2(model, p.phase, p.elastic_hardening, p.F, svd, j);
                else if (model.ctag == SPK_CONSTITUTIVE_NEO_HOOKEAN) stress = neo_hookean_stress2(model, p.phase, p.elastic_hardening, p.F, j);
                else if (model.ctag == SPK_CONSTITUTIVE_MONAGHAN) { // eos_monaghan_sph.rs:44-63
                    const float density_fluid = (p.mass / p.volume0) / p.F.m[0];
                    const float pr = -monaghan_pressure(model, p.mass, p.volume0, density_fluid) * p.jp;
                    stress = mat2_zero();
                    stress.m[0] = pr; stress.m[3] = pr;
                    if (model.c[2] != 0.f) {
                        Mat2 sr;
#pragma unroll
                        for (int r = 0; r < 2; ++r)
#pragma unroll
                            for (int c = 0; c < 2; ++c) M2_(sr, r, c) = (M2_(G, r, c) + M2_(G, c, r)) * 0.5f;
                        const float sph = (sr.m[0] + sr.m[3]) / 2.f;
                        sr.m[0] -= sph; sr.m[3] -= sph;
                        const float kv = 2.f * model.c[2] * p.jp;
#pragma unroll
                        for (int i = 0; i < 4; ++i) stress.m[i] += kv * sr.m[i];
                    }
                }
                if (pass == 0) { // failure_maximum_stress.rs:20-28
                    const bool finite = isfinite(stress.m[0]) && isfinite(stress.m[1]) && isfinite(stress.m[2]) && isfinite(stress.m[3]);
                    if (finite) {
                        float mn, mx;
                        sym2_eig_minmax(stress, mn, mx);
                        if (mx > model.f[0] || (mx - mn) / 2.f > model.f[1]) p.phase = 0.f;
                    }
                }
            }
        }
    }

    // ---- fused K15 (timestep.rs:38-113): un-clamped dt candidate of this particle for the NEXT substep
    {
        unsigned cand_bits = SPK_DT_RAW_NONE;
        if (has_particle && !(p.flags & SPK_F_FAILED)) {
            const DevModel& model = P.models[p.flags >> SPK_F_MODEL_SHIFT];
            const float cand = particle_dt_candidate2(model, h, p.mass, p.volume0, p.F.m[0], p.elastic_hardening, vel[0], vel[1]);
            cand_bits = __float_as_uint(cand != cand ? 1.0e30f : fminf(fmaxf(cand, 0.f), 1.0e30f)); // NaN imposes no bound
        }
        cand_bits = __reduce_min_sync(0xffffffffu, cand_bits);
        if ((tid & 31) == 0 && cand_bits < SPK_DT_RAW_NONE) atomicMin(&s_dtmin, cand_bits);
    }

    // ---- store the particle into the next buffer at its sorted slot (coalesced)
    if (has_particle) {
        const uint32_t di = ent.first + tid;
        P.dst.a0[di] = make_float4(p.x[0], p.x[1], 0.f, __uint_as_float(p.flags));
        P.dst.a1[di] = make_float4(p.F.m[0], p.F.m[1], p.F.m[2], p.F.m[3]);
        P.dst.a3[di] = make_float4(0.f, p.jp, p.mass, p.volume0);
        P.dst.a4[di] = make_float4(p.phase, p.psi_pos, p.elastic_hardening, p.plastic_hardening);
        P.dst.a5[di] = make_float4(p.log_vol_gain, p.nacc_alpha, __uint_as_float(p.orig_id), 0.f);
        P.vel_out[di] = make_float4(vel[0], vel[1], 0.f, 0.f);
    }

    // ---- stage 2: P2G with the UPDATED position (g2p2g.rs:280-349)
    if (has_particle && do_p2g) {
        float ref2[2], w2[2][3];
        int t2[2];
        bool ok = true;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const float q = roundf(spk_div_h(p.x[a], h, inv_h));
            ref2[a] = (q - 1.f) * h - p.x[a];
            quad_weights(spk_div_h(-ref2[a], h, inv_h), w2[a]);
            const float mvf = q - cell_before[a];
            const int mv = (mvf != mvf) ? 0 : (int)__float2ll_rz(mvf);
            t2[a] = tcell[a] + mv; // in [0, 5] after the velocity clamp
            ok = ok && t2[a] >= 0 && t2[a] <= 5;
        }
        if (ok) {
            const float ks = p.volume0 * inv_d * dt;
            Mat2 A; // affine = mass * G - (volume0 * inv_d * dt) * stress
#pragma unroll
            for (int i = 0; i < 4; ++i) A.m[i] = p.mass * G.m[i] - ks * stress.m[i];
            const float mom[2] = {p.mass * vel[0], p.mass * vel[1]};
            const float psi_mass = (p.phase > 0.f && !(p.flags & SPK_F_FAILED)) ? p.mass : 0.f;
            const float psi_mom = psi_mass * p.psi_pos;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int by = cell_part(t2[1] + dy, 1) + quad_part(t2[1] + dy, 1) * ACC_BLOCK;
                const float py = ref2[1] + (float)dy * h;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int base = by + cell_part(t2[0] + dx, 0) + quad_part(t2[0] + dx, 0) * ACC_BLOCK;
                    const float px = ref2[0] + (float)dx * h;
                    const float wt = w2[0][dx] * w2[1][dy];
                    float* a = &s_acc[base];
                    atomicAdd(a + 0 * SPK_CELLS_PER_BLOCK, wt * p.mass);
                    atomicAdd(a + 1 * SPK_CELLS_PER_BLOCK, wt * (M2_(A, 0, 0) * px + M2_(A, 0, 1) * py + mom[0]));
                    atomicAdd(a + 2 * SPK_CELLS_PER_BLOCK, wt * (M2_(A, 1, 0) * px + M2_(A, 1, 1) * py + mom[1]));
                    if (psi_mass != 0.f) {
                        atomicAdd(a + 4 * SPK_CELLS_PER_BLOCK, wt * psi_mom);
                        atomicAdd(a + 5 * SPK_CELLS_PER_BLOCK, wt * psi_mass);
                    }
                }
            }
        }
    }

    // ---- stage 3: flush the accumulation tile into the next grid (transfer_shared_blocks_to_grid, g2p2g.rs:476-553)
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        for (int o = 0; o < SPK_NBR; ++o) {
            const int nid = s_nbr_self[o];
            if (nid >= 0) tma_reduce_add_f32_1d(P.acc + (size_t)nid * ACC_BLOCK, &s_acc[o * ACC_BLOCK], ACC_BLOCK * 4);
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
__global__ void __launch_bounds__(DT_THREADS) estimate_dt2d_kernel(const float4* __restrict__ a0, const float4* __restrict__ a1, const float4* __restrict__ a3,
                                                                    const float4* __restrict__ a4, const float4* __restrict__ vel, uint32_t n,
                                                                    const DevModel* __restrict__ models, float h, float min_dt, float max_dt,
                                                                    DeviceCounters* counters) {
    unsigned long long best = 0xffffffffffffffffull;
    for (uint32_t i = blockIdx.x * DT_THREADS + threadIdx.x; i < n; i += gridDim.x * DT_THREADS) {
        const uint32_t flags = __float_as_uint(a0[i].w);
        if (flags & SPK_F_FAILED) continue; // timestep.rs:72-74
        const DevModel& m = models[flags >> SPK_F_MODEL_SHIFT];
        const float4 r3 = a3[i], r4 = a4[i], v = vel[i];
        float dt = max_dt;
        dt = fminf(dt, particle_dt_candidate2(m, h, r3.z, r3.w, a1[i].x, r4.z, v.x, v.y));
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
    P.nbr_self = next.nbr_self;
    P.nbr_prev = next.nbr_prev;
    P.acc = next.acc;
    P.prev_vel = prev.vel;
    P.h = c->cell_width;
    P.inv_h = 1.0f / c->cell_width;
    P.dt = dt;
    P.dt_raw_bits = &c->d_counters->dt_raw_bits;
    { cudaError_t e = cudaMemsetAsync(&c->d_counters->dt_raw_bits, 0x7f, 4, c->stream); if (e != cudaSuccess) return e; }
    const bool timed = c->kev_on && c->kev_count < 256;
    if (timed) cudaEventRecord(c->kev[2 * c->kev_count], c->stream);
    g2p2g2d_kernel<<<nd, G2P2G_THREADS, 0, c->stream>>>(P);
    if (timed) { cudaEventRecord(c->kev[2 * c->kev_count + 1], c->stream); c->kev_count += 1; }
    c->launches += 1;
    c->g2p2g_launches += 1;
    c->cur = 1 - c->cur; // the next buffer now holds the particles, physically in sorted order
    c->sorted_valid = false;
    c->dt_cache_valid = true;
    c->counters_fresh = false;
    return cudaGetLastError();
}

cudaError_t spk_launch_estimate_dt(spk_ctx* c, float min_dt, float max_dt) {
    const uint32_t n = (uint32_t)c->n;
    if (!n) return cudaSuccess;
    const ParticleBuffers& pb = c->pb[c->cur];
    int blocks = (int)((n + DT_THREADS - 1) / DT_THREADS);
    const int maxb = c->sm_count * 8;
    if (blocks > maxb) blocks = maxb;
    *c->h_upload = c->dt_acc_host;
    { cudaError_t e = cudaMemcpyAsync(&c->d_counters->dt_acc, c->h_upload, 8, cudaMemcpyHostToDevice, c->stream); if (e != cudaSuccess) return e; }
    c->counters_fresh = false;
    estimate_dt2d_kernel<<<blocks, DT_THREADS, 0, c->stream>>>(pb.a0, pb.a1, pb.a3, pb.a4, c->vel, n, c->models, c->cell_width, min_dt, max_dt, c->d_counters);
    c->launches += 1;
    return cudaGetLastError();
}
