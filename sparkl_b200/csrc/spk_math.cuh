// spk_math.cuh — register-resident 3x3 math, SVD and the constitutive / plasticity / failure models of the
// reference's `src_core/dynamics/models/*.rs`, written for one thread per particle.
//
// Design difference with the reference: the reference calls nalgebra's `svd_unordered` at every consumer
// (Drucker-Prager projection, pos_energy, kirchhoff_stress: 3 SVDs per sand particle per substep).
// Here ONE decomposition F = U diag(s) V^T is computed per particle and shared: the plastic return mapping edits
// `s` in place, and energy / stress are evaluated from (U, s, V) directly.
#pragma once
#include <cfloat>

#include "spk_internal.cuh"

struct Mat3 { // column-major like nalgebra: m[c*3 + r]
    float m[9];
};
#define M_(a, r, c) ((a).m[(c) * 3 + (r)])

__device__ __forceinline__ Mat3 mat_identity() { Mat3 a; a.m[0] = 1; a.m[1] = 0; a.m[2] = 0; a.m[3] = 0; a.m[4] = 1; a.m[5] = 0; a.m[6] = 0; a.m[7] = 0; a.m[8] = 1; return a; }
__device__ __forceinline__ Mat3 mat_zero() { Mat3 a;
#pragma unroll
    for (int i = 0; i < 9; ++i) a.m[i] = 0.f; return a; }
__device__ __forceinline__ Mat3 mat_mul(const Mat3& a, const Mat3& b) {
    Mat3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) M_(r, rr, c) = M_(a, rr, 0) * M_(b, 0, c) + M_(a, rr, 1) * M_(b, 1, c) + M_(a, rr, 2) * M_(b, 2, c);
    return r;
}
// a * b^T
__device__ __forceinline__ Mat3 mat_mul_bt(const Mat3& a, const Mat3& b) {
    Mat3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) M_(r, rr, c) = M_(a, rr, 0) * M_(b, c, 0) + M_(a, rr, 1) * M_(b, c, 1) + M_(a, rr, 2) * M_(b, c, 2);
    return r;
}
// nalgebra Matrix3::determinant (cofactor expansion along the first row)
__device__ __forceinline__ float mat_det(const Mat3& a) {
    float m12_23 = M_(a, 1, 1) * M_(a, 2, 2) - M_(a, 2, 1) * M_(a, 1, 2);
    float m11_23 = M_(a, 1, 0) * M_(a, 2, 2) - M_(a, 2, 0) * M_(a, 1, 2);
    float m11_22 = M_(a, 1, 0) * M_(a, 2, 1) - M_(a, 2, 0) * M_(a, 1, 1);
    return M_(a, 0, 0) * m12_23 - M_(a, 0, 1) * m11_23 + M_(a, 0, 2) * m11_22;
}
__device__ __forceinline__ float inv_exact(float e) { return e == 0.f ? 0.f : 1.f / e; } // physics.rs:7-14
__device__ __forceinline__ float signum_f(float x) { return (x != x) ? x : copysignf(1.f, x); } // Rust f32::signum

// ---------------------------------------------------------------------------------------------------- SVD
struct Svd3 { Mat3 u; float s[3]; Mat3 v; }; // F = u diag(s) v^T, s >= 0 (unordered)

// One Hestenes rotation of the column pair (p, q) of U (and V).  al / be are the squared column norms (tracked
// incrementally inside a sweep: a Jacobi rotation moves exactly t * ga from one to the other), ga their dot product.
//   t = tan(theta) = sign(d) 2 ga / (|d| + sqrt(d^2 + 4 ga^2)),  d = be - al     (the smaller root, |theta| <= pi/4)
// Approximate MUFU reciprocal / rsqrt are enough for t (the angle only has to be accurate enough for quadratic
// convergence); c = 1 / sqrt(1 + t^2) gets one Newton step so the rotation stays orthonormal to f32 round-off.
#define SPK_JACOBI_ROT(p, q)                                                                                            \
    {                                                                                                                   \
        const float ga = U.m[3 * p] * U.m[3 * q] + U.m[3 * p + 1] * U.m[3 * q + 1] + U.m[3 * p + 2] * U.m[3 * q + 2];     \
        const float g2 = ga * ga, ab = nrm[p] * nrm[q];                                                                 \
        big = big || (g2 > 1.0e-9f * ab);                                                                               \
        if (g2 > 1.0e-14f * ab) {                                                                                       \
            const float d = nrm[q] - nrm[p];                                                                            \
            const float x = fmaf(d, d, 4.f * g2);                                                                       \
            const float den = fabsf(d) + x * rsqrtf(x);                                                                 \
            float t = __fdividef(ga + ga, den);                                                                         \
            t = d < 0.f ? -t : t;                                                                                       \
            const float tt = fmaf(t, t, 1.f);                                                                           \
            float cc = rsqrtf(tt);                                                                                      \
            cc = cc * fmaf(-0.5f * tt, cc * cc, 1.5f);                                                                  \
            const float ss = cc * t;                                                                                    \
            nrm[p] = fmaf(-t, ga, nrm[p]);                                                                              \
            nrm[q] = fmaf(t, ga, nrm[q]);                                                                               \
            _Pragma("unroll") for (int r = 0; r < 3; ++r) {                                                             \
                const float up = U.m[3 * p + r], uq = U.m[3 * q + r];                                                   \
                U.m[3 * p + r] = cc * up - ss * uq;                                                                     \
                U.m[3 * q + r] = ss * up + cc * uq;                                                                     \
                const float vp = V.m[3 * p + r], vq = V.m[3 * q + r];                                                   \
                V.m[3 * p + r] = cc * vp - ss * vq;                                                                     \
                V.m[3 * q + r] = ss * vp + cc * vq;                                                                     \
            }                                                                                                           \
        }                                                                                                               \
    }

// rank-deficient F: complete U to an orthonormal basis (the reference's nalgebra SVD also returns a full U)
__device__ __forceinline__ void svd3_complete_basis(Mat3& U, float* sv, bool* ok) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (ok[c]) continue;
        const int a1 = (c + 1) % 3, a2 = (c + 2) % 3;
        float w0, w1, w2;
        if (ok[a1] && ok[a2]) {
            w0 = U.m[3 * a1 + 1] * U.m[3 * a2 + 2] - U.m[3 * a1 + 2] * U.m[3 * a2 + 1];
            w1 = U.m[3 * a1 + 2] * U.m[3 * a2 + 0] - U.m[3 * a1 + 0] * U.m[3 * a2 + 2];
            w2 = U.m[3 * a1 + 0] * U.m[3 * a2 + 1] - U.m[3 * a1 + 1] * U.m[3 * a2 + 0];
        } else {
            // at most one valid column k: any unit vector orthogonal to it
            // (selects on compile-time indices only: a run-time column index would force U into local memory on the hot path too)
            const bool use1 = ok[a1], use2 = !ok[a1] && ok[a2];
            const float kx = use1 ? U.m[3 * a1] : (use2 ? U.m[3 * a2] : 0.f);
            const float ky = use1 ? U.m[3 * a1 + 1] : (use2 ? U.m[3 * a2 + 1] : 0.f);
            const float kz = use1 ? U.m[3 * a1 + 2] : (use2 ? U.m[3 * a2 + 2] : 0.f);
            float ex = fabsf(kx) <= fabsf(ky) && fabsf(kx) <= fabsf(kz) ? 1.f : 0.f;
            float ey = (ex == 0.f && fabsf(ky) <= fabsf(kz)) ? 1.f : 0.f;
            float ez = (ex == 0.f && ey == 0.f) ? 1.f : 0.f;
            float d = ex * kx + ey * ky + ez * kz;
            w0 = ex - d * kx; w1 = ey - d * ky; w2 = ez - d * kz;
            float inv = rsqrtf(w0 * w0 + w1 * w1 + w2 * w2);
            w0 *= inv; w1 *= inv; w2 *= inv;
        }
        U.m[3 * c] = w0; U.m[3 * c + 1] = w1; U.m[3 * c + 2] = w2;
        ok[c] = true;
        sv[c] = 0.f;
    }
}

// One-sided (Hestenes) Jacobi: rotates column pairs of U = F V until mutually orthogonal; relative accuracy of the
// singular values is that of f32 even for ill-conditioned F.  Convergence is quadratic, so the loop stops after the
// first sweep whose largest pre-rotation cosine^2 was below 1e-9 (the rotations of that sweep bring it to round-off):
// 3-4 sweeps for any F met in practice, no extra verification sweep.
__device__ __forceinline__ Svd3 svd3(const Mat3& F) {
    Mat3 U = F, V = mat_identity();
#pragma unroll 1
    for (int sweep = 0; sweep < 12; ++sweep) {
        bool big = false;
        float nrm[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) nrm[c] = U.m[3 * c] * U.m[3 * c] + U.m[3 * c + 1] * U.m[3 * c + 1] + U.m[3 * c + 2] * U.m[3 * c + 2];
        SPK_JACOBI_ROT(0, 1)
        SPK_JACOBI_ROT(0, 2)
        SPK_JACOBI_ROT(1, 2)
        if (!big) break;
    }
    Svd3 out;
    bool ok[3];
    float smax = 0.f;
    float inv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float n2 = U.m[3 * c] * U.m[3 * c] + U.m[3 * c + 1] * U.m[3 * c + 1] + U.m[3 * c + 2] * U.m[3 * c + 2];
        float r = rsqrtf(fmaxf(n2, 1.0e-37f));
        r = r * fmaf(-0.5f * n2, r * r, 1.5f); // one Newton step: 1 / |u_c| to f32 round-off
        inv[c] = r;
        out.s[c] = n2 * r;
        smax = fmaxf(smax, out.s[c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        ok[c] = out.s[c] > 1.0e-30f && out.s[c] > 1.0e-7f * smax;
        if (ok[c]) { U.m[3 * c] *= inv[c]; U.m[3 * c + 1] *= inv[c]; U.m[3 * c + 2] *= inv[c]; }
    }
    if (!(ok[0] && ok[1] && ok[2])) svd3_complete_basis(U, out.s, ok);
    out.u = U;
    out.v = V;
    return out;
}
// U diag(s) V^T
__device__ __forceinline__ Mat3 svd_recompose(const Mat3& u, const float s[3], const Mat3& v) {
    Mat3 us;
#pragma unroll
    for (int c = 0; c < 3; ++c) { us.m[3 * c] = u.m[3 * c] * s[c]; us.m[3 * c + 1] = u.m[3 * c + 1] * s[c]; us.m[3 * c + 2] = u.m[3 * c + 2] * s[c]; }
    return mat_mul_bt(us, v);
}

// eigenvalue extrema of a symmetric 3x3 (lower triangle read, like nalgebra SymmetricEigen) by cyclic Jacobi
__device__ __forceinline__ void sym_eig_minmax(const Mat3& a, float& mn, float& mx) {
    float A00 = M_(a, 0, 0), A11 = M_(a, 1, 1), A22 = M_(a, 2, 2), A01 = M_(a, 1, 0), A02 = M_(a, 2, 0), A12 = M_(a, 2, 1);
#pragma unroll 1
    for (int sweep = 0; sweep < 10; ++sweep) {
        float off = A01 * A01 + A02 * A02 + A12 * A12;
        float diag = A00 * A00 + A11 * A11 + A22 * A22;
        if (off <= 1.0e-14f * diag || off == 0.f) break;
        // (0,1)
        if (A01 != 0.f) {
            float th = (A11 - A00) / (2.f * A01), t = copysignf(1.f, th) / (fabsf(th) + sqrtf(th * th + 1.f)), c = rsqrtf(t * t + 1.f), s = t * c;
            float a00 = A00 - t * A01, a11 = A11 + t * A01, a02 = c * A02 - s * A12, a12 = s * A02 + c * A12;
            A00 = a00; A11 = a11; A02 = a02; A12 = a12; A01 = 0.f;
        }
        // (0,2)
        if (A02 != 0.f) {
            float th = (A22 - A00) / (2.f * A02), t = copysignf(1.f, th) / (fabsf(th) + sqrtf(th * th + 1.f)), c = rsqrtf(t * t + 1.f), s = t * c;
            float a00 = A00 - t * A02, a22 = A22 + t * A02, a01 = c * A01 - s * A12, a12 = s * A01 + c * A12;
            A00 = a00; A22 = a22; A01 = a01; A12 = a12; A02 = 0.f;
        }
        // (1,2)
        if (A12 != 0.f) {
            float th = (A22 - A11) / (2.f * A12), t = copysignf(1.f, th) / (fabsf(th) + sqrtf(th * th + 1.f)), c = rsqrtf(t * t + 1.f), s = t * c;
            float a11 = A11 - t * A12, a22 = A22 + t * A12, a01 = c * A01 - s * A02, a02 = s * A01 + c * A02;
            A11 = a11; A22 = a22; A01 = a01; A02 = a02; A12 = 0.f;
        }
    }
    mn = fminf(A00, fminf(A11, A22));
    mx = fmaxf(A00, fmaxf(A11, A22));
}

// compiler-rt __powisf2: what Rust's f32::powi lowers to (eos_monaghan_sph.rs:36)
__device__ __forceinline__ float powi_f(float a, int b) {
    const bool recip = b < 0;
    float r = 1.f;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.f / r : r;
}

// ---------------------------------------------------------------------------------------------------- models
struct ParticleState { // everything g2p2g keeps in registers for one particle
    float x[3];
    uint32_t flags;
    Mat3 F;
    float jp, mass, volume0;
    float phase, psi_pos, elastic_hardening, plastic_hardening, log_vol_gain, nacc_alpha;
    uint32_t orig_id;
};

// Model mask: which constitutive / plastic / failure models a kernel instantiation is COMPILED for.  The library instantiates the
// fused kernel for a few common model sets (sand, fluid, elastic solids) next to the generic one and picks the narrowest
// instantiation that covers the uploaded model table (spk_g2p2g.cu): the dead models cost neither registers nor instruction cache.
#define SPK_MM_C(tag) (1u << (tag))
#define SPK_MM_P(tag) (1u << (4 + (tag)))
#define SPK_MM_F_MAX_STRESS (1u << 10)
#define SPK_MM_ALL 0xffffffffu
__host__ __device__ __forceinline__ uint32_t model_mask_of(uint32_t ctag, uint32_t ptag, uint32_t ftag) {
    return SPK_MM_C(ctag) | SPK_MM_P(ptag) | (ftag == SPK_FAILURE_MAXIMUM_STRESS ? SPK_MM_F_MAX_STRESS : 0u);
}

template <uint32_t MM = SPK_MM_ALL>
__device__ __forceinline__ bool model_is_fluid(const DevModel& m) { return (MM & SPK_MM_C(SPK_CONSTITUTIVE_MONAGHAN)) && m.ctag == SPK_CONSTITUTIVE_MONAGHAN; }
template <uint32_t MM = SPK_MM_ALL>
__device__ __forceinline__ bool model_needs_svd(const DevModel& m) {
    constexpr uint32_t plastic = SPK_MM_P(SPK_PLASTIC_DRUCKER_PRAGER) | SPK_MM_P(SPK_PLASTIC_NACC) | SPK_MM_P(SPK_PLASTIC_RANKINE) | SPK_MM_P(SPK_PLASTIC_SNOW);
    return ((MM & SPK_MM_C(SPK_CONSTITUTIVE_COROTATED)) && m.ctag == SPK_CONSTITUTIVE_COROTATED) || ((MM & plastic) && m.ptag <= SPK_PLASTIC_SNOW);
}

// eos_monaghan_sph.rs:28-38
__device__ __forceinline__ float monaghan_pressure(const DevModel& m, float mass, float volume0, float density_fluid) {
    float density0 = mass / volume0;
    return fmaxf(m.c[0] * (powi_f(density_fluid / density0, (int)__float_as_uint(m.c[1])) - 1.f), -m.c[3]);
}

// GpuPlasticModel::update_particle (gpu_plastic_model.rs:21-61) on a shared SVD.  Returns true when F changed
// (svd.s then holds the projected singular values and F has been recomposed).
template <uint32_t MM>
__device__ __forceinline__ bool plastic_update_other(const DevModel& m, ParticleState& p, Svd3& svd);
template <uint32_t MM = SPK_MM_ALL>
__device__ __forceinline__ bool plastic_update(const DevModel& m, ParticleState& p, Svd3& svd) {
    float* s = svd.s;
    if ((MM & SPK_MM_P(SPK_PLASTIC_DRUCKER_PRAGER)) && m.ptag == SPK_PLASTIC_DRUCKER_PRAGER) { // plasticity_drucker_prager.rs:36-104
        {
        if ((m.flags & 2u) && p.phase != 0.f) return false;
        const float h0 = m.p[0], h1 = m.p[1], h2 = m.p[2], h3 = m.p[3], lambda = m.p[4], mu = m.p[5], volume_correction = m.p[6];
        const float q = p.plastic_hardening;
        const float angle = h0 + (h1 * q - h3) * expf(-h2 * q);
        const float s_angle = sinf(angle);
        const float alpha = sqrtf(2.f / 3.f) * (2.f * s_angle) / (3.f - s_angle);
        const float d = 3.f;
        const float lg = p.log_vol_gain / d;
        float e0 = logf(s[0]) + lg, e1 = logf(s[1]) + lg, e2 = logf(s[2]) + lg;
        const float tr = e0 + e1 + e2;
        const float trd = tr / d;
        const float d0 = e0 - trd, d1 = e1 - trd, d2 = e2 - trd;
        float n0, n1, n2, dq;
        if ((d0 == 0.f && d1 == 0.f && d2 == 0.f) || tr > 0.f) {
            n0 = n1 = n2 = 1.f;
            dq = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
        } else {
            const float dn = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
            const float gamma = dn + (d * lambda + 2.f * mu) / (2.f * mu) * tr * alpha;
            if (gamma <= 0.f) return false;
            n0 = expf(e0 - gamma * (d0 / dn)); n1 = expf(e1 - gamma * (d1 / dn)); n2 = expf(e2 - gamma * (d2 / dn));
            dq = gamma;
        }
        const float prev_det = s[0] * s[1] * s[2];
        float new_det = n0 * n1 * n2;
        const float diff = new_det - prev_det;
        new_det = diff > 0.f ? new_det : prev_det + diff * volume_correction;
        p.jp *= prev_det / new_det;
        p.log_vol_gain += logf(prev_det) - logf(new_det);
        s[0] = n0; s[1] = n1; s[2] = n2;
        p.plastic_hardening += dq;
        p.F = svd_recompose(svd.u, s, svd.v);
        return true;
        }
    }
    constexpr uint32_t others = SPK_MM_P(SPK_PLASTIC_NACC) | SPK_MM_P(SPK_PLASTIC_RANKINE) | SPK_MM_P(SPK_PLASTIC_SNOW);
    if ((MM & others) && m.ptag >= SPK_PLASTIC_NACC && m.ptag <= SPK_PLASTIC_SNOW) return plastic_update_other<MM>(m, p, svd);
    return false; // Custom / None: no-op (gpu_plastic_model.rs:59)
}
template <uint32_t MM>
__device__ __forceinline__ bool plastic_update_other(const DevModel& m, ParticleState& p, Svd3& svd) {
    float* s = svd.s;
    if ((MM & SPK_MM_P(SPK_PLASTIC_NACC)) && m.ptag == SPK_PLASTIC_NACC) { // plasticity_nacc.rs:61-165
        {
        const float mu = m.p[0], kappa = m.p[1], xi = m.p[2], beta = m.p[3], mm = m.p[4];
        const bool hardening_enabled = (m.flags & 4u) != 0;
        const float d = 3.f;
        float alpha = p.nacc_alpha;
        const float q0 = s[0] * s[0], q1 = s[1] * s[1], q2 = s[2] * s[2];
        const float sq_tr = q0 + q1 + q2;
        const float p0 = kappa * (1.0e-5f + sinhf(xi * fmaxf(-alpha, 0.f)));
        const float j_e_tr = s[0] * s[1] * s[2];
        const float sc = mu * powf(j_e_tr, -2.f / d);
        const float t0 = sc * (q0 - sq_tr / d), t1 = sc * (q1 - sq_tr / d), t2 = sc * (q2 - sq_tr / d);
        const float psi_kappa = kappa / 2.f * (j_e_tr - 1.f / j_e_tr);
        const float p_tr = -psi_kappa * j_e_tr;
        if (p_tr > p0 || p_tr < -beta * p0) {
            const float j_e_n1 = p_tr > p0 ? sqrtf(-2.f * p0 / kappa + 1.f) : sqrtf(2.f * beta * p0 / kappa + 1.f);
            const float v = powf(j_e_n1, 1.f / d);
            s[0] = s[1] = s[2] = v;
            if (hardening_enabled) alpha += logf(j_e_tr / j_e_n1);
            p.nacc_alpha = alpha;
            p.F = svd_recompose(svd.u, s, svd.v);
            return true;
        }
        const float y0 = (1.f + 2.f * beta) * ((6.f - d) / 2.f);
        const float y1 = mm * mm * (p_tr + beta * p0) * (p_tr - p0);
        const float tn2 = t0 * t0 + t1 * t1 + t2 * t2;
        const float y = y0 * tn2 + y1;
        if (y < 1.0e-4f) return false;
        const float tn = sqrtf(tn2);
        if (hardening_enabled && p0 > 1.0e-4f && p_tr < p0 - 1.0e-4f && p_tr > -beta * p0 + 1.0e-4f) {
            const float p_c = (1.f - beta) * p0 / 2.f;
            const float q_tr = sqrtf((6.f - d) / 2.f) * tn;
            float dx = p_c - p_tr, dy = 0.f - q_tr;
            const float dn = sqrtf(dx * dx + dy * dy);
            dx /= dn; dy /= dn;
            const float cq = mm * mm * (p_c + beta * p0) * (p_c - p0);
            const float bq = mm * mm * dx * (2.f * p_c - p0 + beta * p0);
            const float aq = mm * mm * dx * dx + (1.f + 2.f * beta) * dy * dy;
            const float discr = sqrtf(bq * bq - 4.f * aq * cq);
            const float l1 = (-bq + discr) / (2.f * aq), l2 = (-bq - discr) / (2.f * aq);
            const float p1 = p_c + l1 * dx, p2 = p_c + l2 * dx;
            const float p_x = (p_tr - p_c) * (p1 - p_c) > 0.f ? p1 : p2;
            const float j_e_x = sqrtf(fabsf(-2.f * p_x / kappa + 1.f));
            if (j_e_x > 1.0e-4f) alpha += logf(j_e_tr / j_e_x);
        }
        const float k = sqrtf(-y1 / y0) * (powf(j_e_tr, 2.f / d) / mu);
        s[0] = sqrtf(k * (t0 / tn) + sq_tr / d); s[1] = sqrtf(k * (t1 / tn) + sq_tr / d); s[2] = sqrtf(k * (t2 / tn) + sq_tr / d);
        p.nacc_alpha = alpha;
        p.F = svd_recompose(svd.u, s, svd.v);
        return true;
        }
    }
    if ((MM & SPK_MM_P(SPK_PLASTIC_RANKINE)) && m.ptag == SPK_PLASTIC_RANKINE) { // plasticity_rankine.rs:35-78
        {
        const float mu = m.p[0], lambda = m.p[1], tensile = m.p[2], softening = m.p[3];
        float e[3] = {logf(s[0]), logf(s[1]), logf(s[2])};
        const float pr0 = e[0], pr1 = e[1], pr2 = e[2];
        int i0 = 0, i1 = 1, i2 = 2; // ascending order of e (insertion sort, ties keep index order)
        if (e[i0] > e[i1]) { int t = i0; i0 = i1; i1 = t; }
        if (e[i1] > e[i2]) { int t = i1; i1 = i2; i2 = t; if (e[i0] > e[i1]) { int u = i0; i0 = i1; i1 = u; } }
        const int e3 = i0, e2 = i1, e1 = i2;
        const float soft = tensile - (p.plastic_hardening - 1.f);
        const float sum = e[0] + e[1] + e[2];
        if (lambda * sum + 2.f * mu * e[e1] <= soft) return false;
        if ((2.f * mu + lambda) * e[e2] + lambda * (sum - e[e1]) <= soft) {
            e[e1] = (soft - lambda * (sum - e[e1])) / (2.f * mu + lambda);
        } else if ((2.f * mu + 3.f * lambda) * e[e3] <= soft) {
            const float ne = (soft - lambda * (sum - e[e1] - e[e2])) / (2.f * mu + 2.f * lambda);
            e[e1] = ne; e[e2] = ne;
        } else {
            const float ne = soft / (2.f * mu + 3.f * lambda);
            e[0] = e[1] = e[2] = ne;
        }
        const float dd0 = pr0 - e[0], dd1 = pr1 - e[1], dd2 = pr2 - e[2];
        p.plastic_hardening += softening * sqrtf(dd0 * dd0 + dd1 * dd1 + dd2 * dd2);
        p.plastic_hardening = fminf(p.plastic_hardening, tensile);
        s[0] = expf(e[0]); s[1] = expf(e[1]); s[2] = expf(e[2]);
        p.F = svd_recompose(svd.u, s, svd.v);
        return true;
        }
    }
    if ((MM & SPK_MM_P(SPK_PLASTIC_SNOW)) && m.ptag == SPK_PLASTIC_SNOW) { // plasticity_snow.rs:26-48
        {
        const float lo = 1.f - m.p[0], hi = 1.f + m.p[1];
        const float n0 = fminf(fmaxf(s[0], lo), hi), n1 = fminf(fmaxf(s[1], lo), hi), n2 = fminf(fmaxf(s[2], lo), hi);
        p.jp *= (s[0] * s[1] * s[2]) / (n0 * n1 * n2);
        p.elastic_hardening = expf(m.p[2] * (1.f - p.jp));
        s[0] = n0; s[1] = n1; s[2] = n2;
        p.F = svd_recompose(svd.u, s, svd.v);
        return true;
        }
    }
    return false; // Custom / None: no-op (gpu_plastic_model.rs:59)
}

// CorotatedLinearElasticity::kirchhoff_stress, elasticity_corotated_linear.rs:31-77, from the shared SVD of F.
__device__ __forceinline__ Mat3 corotated_stress_split(const DevModel& m, float phase, float hardening, const Mat3& F, const Svd3& svd, float j, float sph);
__device__ __forceinline__ Mat3 corotated_stress(const DevModel& m, float phase, float hardening, const Mat3& F, const Svd3& svd, float j) {
    const float lambda = m.c[1], mu = m.c[2];
    const float sph = lambda * hardening * (j - 1.f) * j;
    if (phase == 1.f) {
        // U (S - I) V^T F^T with U (S - I) V^T = F - U V^T
        Mat3 r = mat_mul_bt(svd.u, svd.v);
        Mat3 d;
#pragma unroll
        for (int i = 0; i < 9; ++i) d.m[i] = F.m[i] - r.m[i];
        Mat3 t = mat_mul_bt(d, F);
        const float k = 2.f * mu * hardening;
#pragma unroll
        for (int i = 0; i < 9; ++i) t.m[i] *= k;
        t.m[0] += sph; t.m[4] += sph; t.m[8] += sph;
        return t;
    }
    return corotated_stress_split(m, phase, hardening, F, svd, j, sph);
}
__device__ __forceinline__ Mat3 corotated_stress_split(const DevModel& m, float phase, float hardening, const Mat3& F, const Svd3& svd, float j, float sph) {
    const float mu = m.c[2];
    float sp[3], sn[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { sp[i] = fmaxf(svd.s[i] - 1.f, 0.f); sn[i] = fminf(svd.s[i] - 1.f, 0.f); }
    const float k = 2.f * mu * hardening;
    Mat3 pos = mat_mul_bt(svd_recompose(svd.u, sp, svd.v), F);
    Mat3 neg = mat_mul_bt(svd_recompose(svd.u, sn, svd.v), F);
#pragma unroll
    for (int i = 0; i < 9; ++i) { pos.m[i] *= k; neg.m[i] *= k; }
    if (j < 1.f) { neg.m[0] += sph; neg.m[4] += sph; neg.m[8] += sph; }
    else { pos.m[0] += sph; pos.m[4] += sph; pos.m[8] += sph; }
    const float phase_coeff = ((m.flags & 1u) && phase == 0.f) ? 0.f : 1.f;
    Mat3 out;
#pragma unroll
    for (int i = 0; i < 9; ++i) out.m[i] = pos.m[i] * phase_coeff + neg.m[i];
    return out;
}

// elasticity_neo_hookean.rs:43-67
__device__ __forceinline__ Mat3 neo_hookean_stress(const DevModel& m, float phase, float hardening, const Mat3& F, float j) {
    const float lambda = m.c[1], mu = m.c[2];
    const float pc = (1.f - 0.001f) * phase * phase + 0.001f;
    const float k = 2.f / 3.f * mu * hardening + lambda * hardening;
    Mat3 cg = mat_mul_bt(F, F);
    const float sphc = (cg.m[0] + cg.m[4] + cg.m[8]) / 3.f;
    cg.m[0] -= sphc; cg.m[4] -= sphc; cg.m[8] -= sphc;
    const float sc = (mu * hardening) * powf(j, -2.f / 3.f);
    const float vol = k / 2.f * (j * j - 1.f);
    Mat3 out;
    if (j >= 1.f) {
#pragma unroll
        for (int i = 0; i < 9; ++i) out.m[i] = (sc * cg.m[i]) * pc;
        out.m[0] = (sc * cg.m[0] + vol) * pc; out.m[4] = (sc * cg.m[4] + vol) * pc; out.m[8] = (sc * cg.m[8] + vol) * pc;
    } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) out.m[i] = (sc * cg.m[i]) * pc;
        out.m[0] += vol; out.m[4] += vol; out.m[8] += vol;
    }
    return out;
}

// timestep bounds: estimate_particle_timestep_length (particle_updater.rs:47-70) + timestep.rs:74-112.
// Returns the per-particle candidate BEFORE the max_dt/min_dt clamps (those are monotone and applied once to the
// reduced value, SURVEY §2.2 "fused-op opportunities").  +inf when the particle imposes no bound.
template <uint32_t MM = SPK_MM_ALL>
__device__ __forceinline__ float particle_dt_candidate(const DevModel& m, float h, float mass, float volume0, float f00, float elastic_hardening, float vx, float vy, float vz) {
    const float vn2 = vx * vx + vy * vy + vz * vz;
    float cand = FLT_MAX;
    bool velocity_bound = false;
    constexpr uint32_t solids = SPK_MM_C(SPK_CONSTITUTIVE_COROTATED) | SPK_MM_C(SPK_CONSTITUTIVE_NEO_HOOKEAN);
    if ((MM & solids) && (m.ctag == SPK_CONSTITUTIVE_COROTATED || m.ctag == SPK_CONSTITUTIVE_NEO_HOOKEAN)) {
        {
        // elasticity_sound_speed_timestep_bound.rs:22-34; NeoHookean is called with (cell_width, hardening) swapped
        // (gpu_constitutive_model.rs:129-137)
        const bool nh = m.ctag == SPK_CONSTITUTIVE_NEO_HOOKEAN;
        const float hard = nh ? h : elastic_hardening;
        const float cw = nh ? elastic_hardening : h;
        const float lambda = m.c[1], mu = m.c[2];
        const float bulk = (lambda + 2.f * mu / 3.f) * hard, shear = mu * hard;
        const float density0 = mass / volume0;
        // max(|v|, c_dir) = sqrt(max(v^2, c_dir^2)) bit for bit (sqrtf is monotone and correctly rounded; a NaN operand drops out of
        // fmaxf on either side of the identity): one square root instead of two
        const float denom = sqrtf(fmaxf(vn2, (bulk + 4.f / 3.f * shear) * 1.f / density0));
        cand = m.c[0] * cw / denom;
        // the velocity bound h / |v| below cannot be the smaller one when the numerator is at most h and the denominator at least |v|
        // (correctly rounded * and / are monotone): skipped, with its square root and division, for the usual CFL coefficient <= 1
        velocity_bound = !(!nh && m.c[0] <= 1.f);
        }
    } else if (model_is_fluid<MM>(m)) { // eos_monaghan_sph.rs:72-99
        {
        const float j = f00;
        const float density0 = mass / volume0;
        const float pressure = -monaghan_pressure(m, mass, volume0, density0 / j);
        const float single = (h / j) * sqrtf(density0 * (j - 1.f) / (6.f * pressure * 3.f));
        const float c_square = fmaxf(vn2, 1.f) / 0.1f;
        cand = fminf(single, h / sqrtf(c_square));
        velocity_bound = true;
        }
    } // Custom: ActiveTimestepBounds::NONE
    if (velocity_bound) cand = fminf(cand, h / sqrtf(vn2));
    return cand;
}
