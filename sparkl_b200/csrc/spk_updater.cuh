// spk_updater.cuh — the particle-updater extension point of the G2P2G kernel.
//
// The reference lets a downstream crate replace the per-particle update of its `g2p2g` kernel: it implements the `ParticleUpdater`
// trait (src_kernels/cuda/particle_updater.rs:11-38), instantiates `g2p2g_generic::<MyUpdater>` in its own kernel crate and plugs the
// launch into the pipeline through `CudaParticleKernelsLauncher` (src/cuda/cuda_particle_kernels.rs:7-27,
// `CudaMpmPipeline::step_generic`, cuda_mpm_pipeline.rs:273-281).  Here the same two halves are:
//   * device side (this file + spk_g2p2g_kernel.cuh): `g2p2g_kernel<MODE, Updater>` is a template; an Updater is a type with the two
//     static device functions below.  `DefaultParticleUpdater` is the reference's `DefaultParticleUpdater` (particle_updater.rs:40-236)
//     and is what libsparkl_b200.so instantiates.  A custom updater lives in the USER's .cu file, which includes these headers and
//     defines its own launch function with SPK_DEFINE_G2P2G_LAUNCHER (examples/custom_updater/).
//   * host side (include/sparkl_abi.h): `spk_set_particle_kernels_launcher` registers the user's launch functions, which the
//     pipeline calls instead of the default ones; `spk_default_launch_*` are the `DefaultCudaParticleKernelsLauncher`.
#pragma once
#include "spk_math.cuh"

template <uint32_t MM>
struct DefaultParticleUpdaterT {
    // The update comes in two phases so that the fused kernel can keep the interpolated velocity and velocity gradient in shared
    // memory (not in registers) across the SVD / return mapping / stress evaluation (spk_g2p2g_kernel.cuh, g2p2g_cell_kernel):
    //   advect        velocity handling, position, deformation-gradient update        particle_updater.rs:87-138 (+ :151-155)
    //                 (reads x, flags, F, orig_id of the particle; the other fields need not be loaded yet)
    //   decompose     the shared SVD of F
    //   constitutive  plasticity, failure, energies, Kirchhoff stress                 particle_updater.rs:140-236
    // `update` runs them in order and is the interface a custom updater implements; an updater that also wants the phased schedule declares
    // `static constexpr bool kPhased = true` and provides the three functions below.
    static constexpr bool kPhased = true;

    __device__ __forceinline__ static void advect(const DevModel& model, const float4* __restrict__ cold, float h, float dt, ParticleState& p, float (&vel)[3], Mat3& G,
                                                  const void* /*user*/) {
        const float div = M_(G, 0, 0) + M_(G, 1, 1) + M_(G, 2, 2); // velocity_gradient_det == trace of the affine matrix
        if (p.flags & SPK_F_KINEMATIC) {
            const float4 kv = cold[p.orig_id];
            vel[0] = kv.x; vel[1] = kv.y; vel[2] = kv.z;
        }
        if (fabsf(vel[0]) * dt >= h || fabsf(vel[1]) * dt >= h || fabsf(vel[2]) * dt >= h) {
#pragma unroll
            for (int a = 0; a < 3; ++a) vel[a] = signum_f(vel[a]) * h / dt;
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) p.x[a] += vel[a] * dt;
        if (!model_is_fluid<MM>(model)) {
            Mat3 Gdt;
#pragma unroll
            for (int i = 0; i < 9; ++i) Gdt.m[i] = G.m[i] * dt;
            const Mat3 dF = mat_mul(Gdt, p.F);
#pragma unroll
            for (int i = 0; i < 9; ++i) p.F.m[i] += dF.m[i];
        } else {
            p.F.m[0] += (div * dt) * p.F.m[0];
        }
        // particle_updater.rs:151-155 (after the plastic update there; the return mapping reads neither, so the order is free)
        if (p.flags & SPK_F_STATIC) {
            vel[0] = vel[1] = vel[2] = 0.f;
            G = mat_zero();
        }
    }

    __device__ __forceinline__ static bool needs_decomposition(const DevModel& model) { return model_needs_svd<MM>(model); }
    // the one decomposition F = U diag(s) V^T that plasticity, energy and stress share (called when model_needs_svd(model))
    __device__ __forceinline__ static Svd3 decompose(const DevModel& /*model*/, const Mat3& F, const void* /*user*/) { return svd3(F); }

    // `svd`: decompose(F) when model_needs_svd(model), unspecified otherwise.  `Gp`: the (final) velocity gradient, Gp[i] =
    // G.m[i] column-major, wherever the caller keeps it — only the Monaghan viscosity reads it
    template <class GradientView>
    __device__ __forceinline__ static bool constitutive(const DevModel& model, ParticleState& p, Svd3& svd, const GradientView& Gp, Mat3& stress, const void* /*user*/) {
        const bool is_fluid = model_is_fluid<MM>(model);
        if (model_needs_svd<MM>(model)) plastic_update<MM>(model, p, svd);
        const float j = mat_det(p.F);
        const float density_def_grad = (p.mass / p.volume0) / j;
        bool do_p2g = true;
        if (density_def_grad == 0.f || (p.flags & SPK_F_FAILED) || (!is_fluid && fabsf(p.F.m[0]) > 1.0e4f)) {
            p.flags |= SPK_F_FAILED;
            p.F = mat_identity();
            do_p2g = false;
        }
        if (do_p2g) {
            // pos energy (gpu_constitutive_model.rs:27-48)
            float energy = 0.f;
            if ((MM & SPK_MM_C(SPK_CONSTITUTIVE_COROTATED)) && model.ctag == SPK_CONSTITUTIVE_COROTATED) { // elasticity_corotated_linear.rs:104-118
                const float a0 = fmaxf(svd.s[0] - 1.f, 0.f), a1 = fmaxf(svd.s[1] - 1.f, 0.f), a2 = fmaxf(svd.s[2] - 1.f, 0.f);
                const float pos_dev = model.c[2] * p.elastic_hardening * (a0 * a0 + a1 * a1 + a2 * a2);
                const float sph = model.c[1] * p.elastic_hardening / 2.f * (j - 1.f) * (j - 1.f);
                energy = j < 1.f ? pos_dev : pos_dev + sph;
            } else if ((MM & SPK_MM_C(SPK_CONSTITUTIVE_NEO_HOOKEAN)) && model.ctag == SPK_CONSTITUTIVE_NEO_HOOKEAN) { // elasticity_neo_hookean.rs:117-141
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
            for (int pass = (((MM & SPK_MM_F_MAX_STRESS) && model.ftag == SPK_FAILURE_MAXIMUM_STRESS) ? 0 : 1); pass < 2; ++pass) {
                if ((MM & SPK_MM_C(SPK_CONSTITUTIVE_COROTATED)) && model.ctag == SPK_CONSTITUTIVE_COROTATED) stress = corotated_stress(model, p.phase, p.elastic_hardening, p.F, svd, j);
                else if ((MM & SPK_MM_C(SPK_CONSTITUTIVE_NEO_HOOKEAN)) && model.ctag == SPK_CONSTITUTIVE_NEO_HOOKEAN) stress = neo_hookean_stress(model, p.phase, p.elastic_hardening, p.F, j);
                else if (is_fluid) { // eos_monaghan_sph.rs:44-63, quirk 7
                    const float density_fluid = (p.mass / p.volume0) / p.F.m[0];
                    const float pr = -monaghan_pressure(model, p.mass, p.volume0, density_fluid) * p.jp;
                    stress = mat_zero();
                    stress.m[0] = pr; stress.m[4] = pr; stress.m[8] = pr;
                    if (model.c[2] != 0.f) {
                        Mat3 sr;
#pragma unroll
                        for (int r = 0; r < 3; ++r)
#pragma unroll
                            for (int c = 0; c < 3; ++c) M_(sr, r, c) = (Gp[c * 3 + r] + Gp[r * 3 + c]) * 0.5f;
                        const float sph = (sr.m[0] + sr.m[4] + sr.m[8]) / 3.f;
                        sr.m[0] -= sph; sr.m[4] -= sph; sr.m[8] -= sph;
                        const float kv = 2.f * model.c[2] * p.jp;
#pragma unroll
                        for (int i = 0; i < 9; ++i) stress.m[i] += kv * sr.m[i];
                    }
                }
                if ((MM & SPK_MM_F_MAX_STRESS) && pass == 0) { // failure_maximum_stress.rs:20-28
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
        return do_p2g;
    }

    // DefaultParticleUpdater::update_particle_and_compute_kirchhoff_stress, src_kernels/cuda/particle_updater.rs:73-236.
    // In: the particle `p`, the interpolated velocity `vel` and velocity gradient `G` (InterpolatedParticleData).  Out: the advected /
    // updated particle, its new velocity, `G` (zeroed for static particles), the Kirchhoff `stress`.  Returns false when the reference
    // returns None (no P2G for this particle).
    __device__ __forceinline__ static bool update(const DevModel& model, const float4* __restrict__ cold, float h, float dt, ParticleState& p, float (&vel)[3], Mat3& G,
                                                  Mat3& stress, const void* user) {
        advect(model, cold, h, dt, p, vel, G, user);
        Svd3 svd;
        if (model_needs_svd<MM>(model)) svd = decompose(model, p.F, user);
        return constitutive(model, p, svd, G.m, stress, user);
    }

    // estimate_particle_timestep_length (particle_updater.rs:47-70) + timestep.rs:74-112: the particle's dt candidate BEFORE the
    // max_dt / min_dt clamps; FLT_MAX-like when it imposes no bound
    __device__ __forceinline__ static float dt_candidate(const DevModel& model, float h, const ParticleState& p, const float (&vel)[3], const void* /*user*/) {
        return particle_dt_candidate<MM>(model, h, p.mass, p.volume0, p.F.m[0], p.elastic_hardening, vel[0], vel[1], vel[2]);
    }
};
using DefaultParticleUpdater = DefaultParticleUpdaterT<SPK_MM_ALL>; // every model of src_core/dynamics/models compiled in
