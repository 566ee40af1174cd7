// math2d_host.cpp — TEST INFRASTRUCTURE: compiles the dim2 DEVICE math header (sparkl_b200/csrc/spk_math2d.cuh) for the host with
// g++ (cuda_runtime.h turns __device__ into nothing outside nvcc) so that the exact source the CUDA kernel inlines — 2x2 SVD,
// plastic return mappings, stresses, CFL candidate — is checked against the dim2 oracle on the GPU-less CI
// (tests/test_dim2_cpu.py::test_device_math_*).  Nothing in the product links this file.
#define SPK_DIM 2
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
#include "../../sparkl_b200/csrc/spk_math2d.cuh"

extern "C" {
int hm_sizeof_devmodel(void) { return (int)sizeof(DevModel); }
void hm_svd2(const float* f, float* u, float* s, float* v) {
    Mat2 F; std::memcpy(F.m, f, 16);
    Svd2 r = svd2(F);
    std::memcpy(u, r.u.m, 16); std::memcpy(v, r.v.m, 16); s[0] = r.s[0]; s[1] = r.s[1];
}
// state: F[4], jp, phase, elastic_hardening, plastic_hardening, log_vol_gain, nacc_alpha (10 floats), updated in place
int hm_plastic(const DevModel* m, float* st) {
    ParticleState2 p;
    std::memcpy(p.F.m, st, 16);
    p.jp = st[4]; p.phase = st[5]; p.elastic_hardening = st[6]; p.plastic_hardening = st[7]; p.log_vol_gain = st[8]; p.nacc_alpha = st[9];
    p.mass = p.volume0 = 1.f; p.flags = 0; p.psi_pos = 0.f; p.orig_id = 0; p.x[0] = p.x[1] = 0.f;
    Svd2 svd = svd2(p.F);
    const bool changed = plastic_update2(*m, p, svd);
    std::memcpy(st, p.F.m, 16);
    st[4] = p.jp; st[5] = p.phase; st[6] = p.elastic_hardening; st[7] = p.plastic_hardening; st[8] = p.log_vol_gain; st[9] = p.nacc_alpha;
    return changed ? 1 : 0;
}
void hm_stress(const DevModel* m, const float* f, float phase, float hardening, float* out) {
    Mat2 F; std::memcpy(F.m, f, 16);
    const float j = mat2_det(F);
    Mat2 t = mat2_zero();
    if (m->ctag == SPK_CONSTITUTIVE_COROTATED) { Svd2 svd = svd2(F); t = corotated_stress2(*m, phase, hardening, F, svd, j); }
    else if (m->ctag == SPK_CONSTITUTIVE_NEO_HOOKEAN) t = neo_hookean_stress2(*m, phase, hardening, F, j);
    std::memcpy(out, t.m, 16);
}
float hm_dt_candidate(const DevModel* m, float h, float mass, float volume0, float f00, float hardening, float vx, float vy) {
    return particle_dt_candidate2(*m, h, mass, volume0, f00, hardening, vx, vy);
}
void hm_eig_minmax(const float* a, float* mn, float* mx) { Mat2 A; std::memcpy(A.m, a, 16); sym2_eig_minmax(A, *mn, *mx); }
}
