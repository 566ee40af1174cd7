// updater3d_host.cpp — TEST INFRASTRUCTURE: compiles the dim3 DEVICE updater (sparkl_b200/csrc/spk_updater.cuh: DefaultParticleUpdater
// = advection, F update, the shared one-sided-Jacobi SVD, every plastic return mapping, failure, stress; spk_math.cuh) for the host
// with g++, so that the exact source the fused G2P2G kernel inlines is compared with the CPU restatement of the reference on the
// GPU-less CI (tests/test_device_math3d_cpu.py).  MUFU approximations (rsqrtf, __fdividef) are replaced by their IEEE
// counterparts: the kernel only uses them where a Newton step or the Jacobi iteration absorbs the difference.
#define SPK_DIM 3
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline float __fdividef(float a, float b) { return a / b; }
using std::isfinite;
#include "../../sparkl_b200/csrc/spk_updater.cuh"

extern "C" {
int hu_sizeof_devmodel(void) { return (int)sizeof(DevModel); }
void hu_svd3(const float* f, float* u, float* s, float* v) {
    Mat3 F; std::memcpy(F.m, f, 36);
    Svd3 r = svd3(F);
    std::memcpy(u, r.u.m, 36); std::memcpy(v, r.v.m, 36); std::memcpy(s, r.s, 12);
}
// particle: x[3], F[9], jp, mass, volume0, phase, psi_pos, elastic_hardening, plastic_hardening, log_vol_gain, nacc_alpha (21 floats, in/out)
int hu_update(const DevModel* model, uint32_t* flags, float* particle, float* vel3, float* G9, float* stress9, const float* kinematic_vel, float h, float dt) {
    ParticleState p;
    std::memcpy(p.x, particle, 12); std::memcpy(p.F.m, particle + 3, 36);
    p.jp = particle[12]; p.mass = particle[13]; p.volume0 = particle[14]; p.phase = particle[15]; p.psi_pos = particle[16];
    p.elastic_hardening = particle[17]; p.plastic_hardening = particle[18]; p.log_vol_gain = particle[19]; p.nacc_alpha = particle[20];
    p.flags = *flags; p.orig_id = 0;
    const float4 cold = make_float4(kinematic_vel[0], kinematic_vel[1], kinematic_vel[2], 0.f);
    float vel[3] = {vel3[0], vel3[1], vel3[2]};
    Mat3 G, stress = mat_zero();
    std::memcpy(G.m, G9, 36);
    const bool ok = DefaultParticleUpdater::update(*model, &cold, h, dt, p, vel, G, stress, nullptr);
    std::memcpy(particle, p.x, 12); std::memcpy(particle + 3, p.F.m, 36);
    particle[12] = p.jp; particle[15] = p.phase; particle[16] = p.psi_pos; particle[17] = p.elastic_hardening; particle[18] = p.plastic_hardening;
    particle[19] = p.log_vol_gain; particle[20] = p.nacc_alpha;
    *flags = p.flags;
    std::memcpy(vel3, vel, 12); std::memcpy(G9, G.m, 36); std::memcpy(stress9, stress.m, 36);
    return ok ? 1 : 0;
}
float hu_dt_candidate(const DevModel* model, float h, float mass, float volume0, float f00, float hardening, float vx, float vy, float vz) {
    return particle_dt_candidate(*model, h, mass, volume0, f00, hardening, vx, vy, vz);
}
}
