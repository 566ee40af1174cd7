// spk_g2p2g2d_body.cuh — the per-particle body of the dim2 G2P2G kernel (spk_g2p2g2d.cu): gather, 9-node G2P, particle update
// (src_kernels/cuda/particle_updater.rs:73-236), CFL candidate, store, 9-node P2G (src_kernels/cuda/g2p2g.rs:179-351 with the
// `dim2` feature).  Kept free of TMA / barrier / warp intrinsics so that the SAME source is also compiled for the host by the
// test suite (tests/host/kernels2d_host.cpp emulates the CTA glue and compares with the dim2 CPU restatement on GPU-less CI).
#pragma once
#include "spk_device.cuh"
#include "spk_math2d.cuh"

constexpr int G2P2G2D_ACC_BLOCK = SPK_ACC_CHANNELS * SPK_CELLS_PER_BLOCK; // floats per block in `acc`: [6][16]

// tile coordinate t in [0, 8) per axis -> cell part (x + 4 y inside the block) and quadrant part (bit0 = x, bit1 = y)
__device__ __forceinline__ int cell_part(int t, int axis) { return (t & 3) << (2 * axis); }
__device__ __forceinline__ int quad_part(int t, int axis) { return (t >> 2) << axis; }

__device__ __forceinline__ void quad_weights(float x, float w[3]) { // QuadraticKernel::eval_all, kernel.rs:15-21
    w[0] = 0.5f * ((1.5f - x) * (1.5f - x));
    w[1] = 0.75f - ((x - 1.f) * (x - 1.f));
    w[2] = 0.5f * ((x - 0.5f) * (x - 0.5f));
}

struct Particle2dWork { // everything one thread keeps in registers for its particle
    ParticleState2 p;
    float vel[2];
    Mat2 G, stress;
    bool do_p2g;
    int tcell[2];         // stencil-base tile coordinate before advection, in [1, 4]
    float cell_before[2]; // round(x / h) before advection
    float ref[2];
    float w[2][3];
};

// gather through the sort permutation + stencil base and weights at the OLD position
__device__ __forceinline__ void g2p2g2d_load(const ParticleBuffers& src, uint32_t si, float h, float inv_h, Particle2dWork& k) {
    ParticleState2& p = k.p;
    const float4 r0 = src.a0[si], r1 = src.a1[si], r3 = src.a3[si], r4 = src.a4[si], r5 = src.a5[si];
    p.x[0] = r0.x; p.x[1] = r0.y; p.flags = __float_as_uint(r0.w);
    p.F.m[0] = r1.x; p.F.m[1] = r1.y; p.F.m[2] = r1.z; p.F.m[3] = r1.w;
    p.jp = r3.y; p.mass = r3.z; p.volume0 = r3.w;
    p.phase = r4.x; p.psi_pos = r4.y; p.elastic_hardening = r4.z; p.plastic_hardening = r4.w;
    p.log_vol_gain = r5.x; p.nacc_alpha = r5.y; p.orig_id = __float_as_uint(r5.z);
    k.vel[0] = k.vel[1] = 0.f;
    k.G = mat2_zero(); k.stress = mat2_zero();
    k.do_p2g = false;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const float q = roundf(spk_div_h(p.x[a], h, inv_h));
        k.cell_before[a] = q;
        k.ref[a] = (q - 1.f) * h - p.x[a];                  // dir_to_associated_grid_node, particle.rs:37-47
        quad_weights(spk_div_h(-k.ref[a], h, inv_h), k.w[a]);
        const int assoc_cell = (q != q) ? -2 : (int)__float2ll_rz(q) - 2;
        k.tcell[a] = (assoc_cell & 3) + 1;                   // associated_cell_index_in_block_off_by_two, particle.rs:28-35
    }
}

// ---- the particle-updater extension point: see spk_updater.cuh for the contract.  An Updater is a type with
// these two static device functions; the kernel template of spk_g2p2g2d.cu (`g2p2g2d_kernel<Updater>`) is instantiated with
// DefaultParticleUpdater2 by the library and can be instantiated with a user's type through spk_launch_g2p2g2d_with<U>.
struct DefaultParticleUpdater2 {
    // DefaultParticleUpdater::update_particle_and_compute_kirchhoff_stress, src_kernels/cuda/particle_updater.rs:73-236 (dim2)
    __device__ __forceinline__ static bool update(const DevModel& model, const float4* __restrict__ cold, float h, float dt, ParticleState2& p, float (&vel)[2],
                                                  Mat2& G, Mat2& stress, const void* /*user*/) {
    const float div = M2_(G, 0, 0) + M2_(G, 1, 1); // velocity_gradient_det

    // ---- particle update (particle_updater.rs:73-236)
    const bool is_fluid = model_is_fluid(model);
    if (p.flags & SPK_F_KINEMATIC) {
        const float4 kv = cold[p.orig_id];
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
    bool do_p2g = true;
    if (density_def_grad == 0.f || (p.flags & SPK_F_FAILED) || (!is_fluid && fabsf(p.F.m[0]) > 1.0e4f)) {
        p.flags |= SPK_F_FAILED;
        p.F = mat2_identity();
        do_p2g = false;
        return do_p2g;
    }
    float energy = 0.f; // pos energy (gpu_constitutive_model.rs:27-48)
    if (model.ctag == SPK_CONSTITUTIVE_COROTATED) { // elasticity_corotated_linear.rs:104-118
        const float a0 = fmaxf(svd.s[0] - 1.f, 0.f), a1 = fmaxf(svd.s[1] - 1.f, 0.f);
        const float pos_dev = model.c[2] * p.elastic_hardening * (a0 * a0 + a1 * a1);
        const float sph = model.c[1] * p.elastic_hardening / 2.f * (j - 1.f) * (j - 1.f);
        energy = j < 1.f ? pos_dev : pos_dev + sph;
    } else if (model.ctag == SPK_CONSTITUTIVE_NEO_HOOKEAN) { // elasticity_neo_hookean.rs:117-141 (DIM = 2 in the exponent and the offset)
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
    // (particle_updater.rs:185-235)
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
    return do_p2g;
    }
    // estimate_particle_timestep_length (particle_updater.rs:47-70) + timestep.rs:74-112, before the clamps
    __device__ __forceinline__ static float dt_candidate(const DevModel& model, float h, const ParticleState2& p, const float (&vel)[2], const void* /*user*/) {
        return particle_dt_candidate2(model, h, p.mass, p.volume0, p.F.m[0], p.elastic_hardening, vel[0], vel[1]);
    }
};

// G2P from the staged velocity tile + DefaultParticleUpdater::update_particle_and_compute_kirchhoff_stress
template <class Updater>
__device__ inline void g2p2g2d_update(const float4* s_vel, const DevModel* models, const float4* cold, float h, float dt, Particle2dWork& k, const void* user) {
    ParticleState2& p = k.p;
    float* vel = k.vel;
    Mat2& G = k.G;
    const float inv_d = 4.f / (h * h); // Kernel::inv_d
    const DevModel& model = models[p.flags >> SPK_F_MODEL_SHIFT];
    // ---- G2P (particle_g2p2g, g2p2g.rs:192-245): V = sum w v, G[:, a] = inv_d * sum w dpt_a v
    {
        float V[2] = {0.f, 0.f}, Gx[2] = {0.f, 0.f}, Gy[2] = {0.f, 0.f};
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int iy = cell_part(k.tcell[1] + dy, 1) | (quad_part(k.tcell[1] + dy, 1) << SPK_CELL_SHIFT);
            const float py = k.ref[1] + (float)dy * h;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const float4 c = s_vel[iy | cell_part(k.tcell[0] + dx, 0) | (quad_part(k.tcell[0] + dx, 0) << SPK_CELL_SHIFT)];
                const float wt = k.w[0][dx] * k.w[1][dy];
                const float px = k.ref[0] + (float)dx * h;
                V[0] += wt * c.x; V[1] += wt * c.y;
                Gx[0] += (wt * px) * c.x; Gx[1] += (wt * px) * c.y;
                Gy[0] += (wt * py) * c.x; Gy[1] += (wt * py) * c.y;
            }
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) { vel[a] = V[a]; M2_(G, a, 0) = Gx[a] * inv_d; M2_(G, a, 1) = Gy[a] * inv_d; }
    }
    k.do_p2g = Updater::update(model, cold, h, dt, p, k.vel, G, k.stress, user);
}

// fused K15 (timestep.rs:38-113): un-clamped dt candidate (float bits) of this particle for the NEXT substep
template <class Updater>
__device__ __forceinline__ unsigned g2p2g2d_dt_bits(const DevModel* models, float h, const Particle2dWork& k, const void* user) {
    const ParticleState2& p = k.p;
    if (p.flags & SPK_F_FAILED) return SPK_DT_RAW_NONE;
    const DevModel& model = models[p.flags >> SPK_F_MODEL_SHIFT];
    const float cand = Updater::dt_candidate(model, h, p, k.vel, user);
    return __float_as_uint(cand != cand ? 1.0e30f : fminf(fmaxf(cand, 0.f), 1.0e30f)); // NaN imposes no bound (Rust f32::min ignores NaN)
}

__device__ __forceinline__ void g2p2g2d_store(const ParticleBuffers& dst, float4* vel_out, uint32_t di, const Particle2dWork& k) {
    const ParticleState2& p = k.p;
    dst.a0[di] = make_float4(p.x[0], p.x[1], 0.f, __uint_as_float(p.flags));
    dst.a1[di] = make_float4(p.F.m[0], p.F.m[1], p.F.m[2], p.F.m[3]);
    dst.a3[di] = make_float4(0.f, p.jp, p.mass, p.volume0);
    dst.a4[di] = make_float4(p.phase, p.psi_pos, p.elastic_hardening, p.plastic_hardening);
    dst.a5[di] = make_float4(p.log_vol_gain, p.nacc_alpha, __uint_as_float(p.orig_id), 0.f);
    vel_out[di] = make_float4(k.vel[0], k.vel[1], 0.f, 0.f);
}

// P2G with the UPDATED position (g2p2g.rs:280-349) into the [quadrant][channel][cell] tile; 9 nodes x (3 or 5) float atomics
__device__ inline void g2p2g2d_scatter(float* s_acc, float h, float inv_h, float dt, const Particle2dWork& k) {
    if (!k.do_p2g) return;
    const ParticleState2& p = k.p;
    const float inv_d = 4.f / (h * h);
    float ref2[2], w2[2][3];
    int t2[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const float q = roundf(spk_div_h(p.x[a], h, inv_h));
        ref2[a] = (q - 1.f) * h - p.x[a];
        quad_weights(spk_div_h(-ref2[a], h, inv_h), w2[a]);
        const float mvf = q - k.cell_before[a];
        const int mv = (mvf != mvf) ? 0 : (int)__float2ll_rz(mvf);
        t2[a] = k.tcell[a] + mv; // in [0, 5] after the velocity clamp
        if (t2[a] < 0 || t2[a] > 5) return; // guards NaN positions
    }
    const float ks = p.volume0 * inv_d * dt;
    Mat2 A; // affine = mass * G - (volume0 * inv_d * dt) * stress
#pragma unroll
    for (int i = 0; i < 4; ++i) A.m[i] = p.mass * k.G.m[i] - ks * k.stress.m[i];
    const float mom[2] = {p.mass * k.vel[0], p.mass * k.vel[1]};
    const float psi_mass = (p.phase > 0.f && !(p.flags & SPK_F_FAILED)) ? p.mass : 0.f;
    const float psi_mom = psi_mass * p.psi_pos;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int by = cell_part(t2[1] + dy, 1) + quad_part(t2[1] + dy, 1) * G2P2G2D_ACC_BLOCK;
        const float py = ref2[1] + (float)dy * h;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int base = by + cell_part(t2[0] + dx, 0) + quad_part(t2[0] + dx, 0) * G2P2G2D_ACC_BLOCK;
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
