// spk_math2d.cuh — dim2 counterpart of spk_math.cuh: register-resident 2x2 math, SVD and the constitutive / plasticity /
// failure models of the reference's `src_core/dynamics/models/*.rs` evaluated with DIM = 2, one thread per particle.
// As in 3D, ONE decomposition F = U diag(s) V^T per particle is shared by the plastic return mapping, the energy and the
// stress (the reference calls nalgebra's SVD at each consumer).  Constants written `DIM` in the Rust source are 2 here;
// literal 2/3, 4/3 and 3 of the Rust source stay literal (the Rust file:line of every formula is cited next to it).
#pragma once
#include <cfloat>

#include "spk_internal.cuh"

struct Mat2 { float m[4]; }; // column-major like nalgebra: m[c*2 + r]
#define M2_(a, r, c) ((a).m[(c) * 2 + (r)])

__device__ __forceinline__ Mat2 mat2_identity() { Mat2 a; a.m[0] = 1.f; a.m[1] = 0.f; a.m[2] = 0.f; a.m[3] = 1.f; return a; }
__device__ __forceinline__ Mat2 mat2_zero() { Mat2 a; a.m[0] = a.m[1] = a.m[2] = a.m[3] = 0.f; return a; }
__device__ __forceinline__ Mat2 mat2_mul(const Mat2& a, const Mat2& b) {
    Mat2 r;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) M2_(r, rr, c) = M2_(a, rr, 0) * M2_(b, 0, c) + M2_(a, rr, 1) * M2_(b, 1, c);
    return r;
}
__device__ __forceinline__ Mat2 mat2_mul_bt(const Mat2& a, const Mat2& b) { // a * b^T
    Mat2 r;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) M2_(r, rr, c) = M2_(a, rr, 0) * M2_(b, c, 0) + M2_(a, rr, 1) * M2_(b, c, 1);
    return r;
}
__device__ __forceinline__ float mat2_det(const Mat2& a) { return M2_(a, 0, 0) * M2_(a, 1, 1) - M2_(a, 1, 0) * M2_(a, 0, 1); }
__device__ __forceinline__ float inv_exact(float e) { return e == 0.f ? 0.f : 1.f / e; } // physics.rs:7-14
__device__ __forceinline__ float signum_f(float x) { return (x != x) ? x : copysignf(1.f, x); }

// ---------------------------------------------------------------------------------------------------- SVD
struct Svd2 { Mat2 u; float s[2]; Mat2 v; }; // F = u diag(s) v^T, s >= 0 (unordered)

// One-sided (Hestenes) Jacobi on the single column pair of a 2x2: one rotation orthogonalises the columns exactly, a second
// pass removes the round-off of the first.  IEEE division / sqrt (no MUFU shortcuts: the 2D kernel is not issue bound).
__device__ __forceinline__ Svd2 svd2(const Mat2& F) {
    Mat2 U = F, V = mat2_identity();
#pragma unroll
    for (int sweep = 0; sweep < 2; ++sweep) {
        const float al = U.m[0] * U.m[0] + U.m[1] * U.m[1], be = U.m[2] * U.m[2] + U.m[3] * U.m[3];
        const float ga = U.m[0] * U.m[2] + U.m[1] * U.m[3];
        if (ga * ga > 1.0e-16f * al * be) {
            const float zeta = (be - al) / (2.f * ga);
            const float t = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(1.f + zeta * zeta));
            const float cc = 1.f / sqrtf(1.f + t * t), ss = cc * t;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const float up = U.m[r], uq = U.m[2 + r];
                U.m[r] = cc * up - ss * uq; U.m[2 + r] = ss * up + cc * uq;
                const float vp = V.m[r], vq = V.m[2 + r];
                V.m[r] = cc * vp - ss * vq; V.m[2 + r] = ss * vp + cc * vq;
            }
        }
    }
    Svd2 out;
    out.s[0] = sqrtf(U.m[0] * U.m[0] + U.m[1] * U.m[1]);
    out.s[1] = sqrtf(U.m[2] * U.m[2] + U.m[3] * U.m[3]);
    const float smax = fmaxf(out.s[0], out.s[1]);
    const bool ok0 = out.s[0] > 1.0e-30f && out.s[0] > 1.0e-7f * smax, ok1 = out.s[1] > 1.0e-30f && out.s[1] > 1.0e-7f * smax;
    if (ok0) { U.m[0] /= out.s[0]; U.m[1] /= out.s[0]; }
    if (ok1) { U.m[2] /= out.s[1]; U.m[3] /= out.s[1]; }
    // rank-deficient F: complete U to an orthonormal basis (nalgebra's SVD returns a full U too)
    if (!ok0 && !ok1) { U = mat2_identity(); out.s[0] = out.s[1] = 0.f; }
    else if (!ok0) { U.m[0] = U.m[3]; U.m[1] = -U.m[2]; out.s[0] = 0.f; }
    else if (!ok1) { U.m[2] = -U.m[1]; U.m[3] = U.m[0]; out.s[1] = 0.f; }
    // nalgebra's 2x2 SVD (svd2.rs `svd_ordered2`) returns the singular values in DESCENDING order; Rankine's index sort
    // (plasticity_rankine.rs:47-49) depends on it, every other consumer is permutation invariant
    if (out.s[0] < out.s[1]) {
        float t;
        t = out.s[0]; out.s[0] = out.s[1]; out.s[1] = t;
        t = U.m[0]; U.m[0] = U.m[2]; U.m[2] = t; t = U.m[1]; U.m[1] = U.m[3]; U.m[3] = t;
        t = V.m[0]; V.m[0] = V.m[2]; V.m[2] = t; t = V.m[1]; V.m[1] = V.m[3]; V.m[3] = t;
    }
    out.u = U;
    out.v = V;
    return out;
}
__device__ __forceinline__ Mat2 svd2_recompose(const Mat2& u, const float s[2], const Mat2& v) { // U diag(s) V^T
    Mat2 us;
    us.m[0] = u.m[0] * s[0]; us.m[1] = u.m[1] * s[0]; us.m[2] = u.m[2] * s[1]; us.m[3] = u.m[3] * s[1];
    return mat2_mul_bt(us, v);
}
// eigenvalue extrema of a symmetric 2x2 (lower triangle read, like nalgebra SymmetricEigen)
__device__ __forceinline__ void sym2_eig_minmax(const Mat2& a, float& mn, float& mx) {
    const float p = M2_(a, 0, 0), q = M2_(a, 1, 1), o = M2_(a, 1, 0);
    const float mean = (p + q) * 0.5f, hd = (p - q) * 0.5f, rad = sqrtf(hd * hd + o * o);
    mn = mean - rad; mx = mean + rad;
}
__device__ __forceinline__ float powi_f(float a, int b) { // compiler-rt __powisf2 (eos_monaghan_sph.rs:36)
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
struct ParticleState2 {
    float x[2];
    uint32_t flags;
    Mat2 F;
    float jp, mass, volume0;
    float phase, psi_pos, elastic_hardening, plastic_hardening, log_vol_gain, nacc_alpha;
    uint32_t orig_id;
};

__device__ __forceinline__ bool model_is_fluid(const DevModel& m) { return m.ctag == SPK_CONSTITUTIVE_MONAGHAN; }
__device__ __forceinline__ bool model_needs_svd(const DevModel& m) { return m.ctag == SPK_CONSTITUTIVE_COROTATED || (m.ptag <= SPK_PLASTIC_SNOW); }

__device__ __forceinline__ float monaghan_pressure(const DevModel& m, float mass, float volume0, float density_fluid) { // eos_monaghan_sph.rs:28-38
    const float density0 = mass / volume0;
    return fmaxf(m.c[0] * (powi_f(density_fluid / density0, (int)__float_as_uint(m.c[1])) - 1.f), -m.c[3]);
}

// GpuPlasticModel::update_particle (gpu_plastic_model.rs:21-61) on the shared SVD; true when F changed
__device__ inline bool plastic_update2(const DevModel& m, ParticleState2& p, Svd2& svd) {
    float* s = svd.s;
    const float d = 2.f; // DIM
    switch (m.ptag) {
    case SPK_PLASTIC_DRUCKER_PRAGER: { // plasticity_drucker_prager.rs:36-104
        if ((m.flags & 2u) && p.phase != 0.f) return false;
        const float h0 = m.p[0], h1 = m.p[1], h2 = m.p[2], h3 = m.p[3], lambda = m.p[4], mu = m.p[5], volume_correction = m.p[6];
        const float q = p.plastic_hardening;
        const float angle = h0 + (h1 * q - h3) * expf(-h2 * q);
        const float s_angle = sinf(angle);
        const float alpha = sqrtf(2.f / 3.f) * (2.f * s_angle) / (3.f - s_angle);
        const float lg = p.log_vol_gain / d;
        const float e0 = logf(s[0]) + lg, e1 = logf(s[1]) + lg;
        const float tr = e0 + e1;
        const float trd = tr / d;
        const float d0 = e0 - trd, d1 = e1 - trd;
        float n0, n1, dq;
        if ((d0 == 0.f && d1 == 0.f) || tr > 0.f) {
            n0 = n1 = 1.f;
            dq = sqrtf(e0 * e0 + e1 * e1);
        } else {
            const float dn = sqrtf(d0 * d0 + d1 * d1);
            const float gamma = dn + (d * lambda + 2.f * mu) / (2.f * mu) * tr * alpha;
            if (gamma <= 0.f) return false;
            n0 = expf(e0 - gamma * (d0 / dn)); n1 = expf(e1 - gamma * (d1 / dn));
            dq = gamma;
        }
        const float prev_det = s[0] * s[1];
        float new_det = n0 * n1;
        const float diff = new_det - prev_det;
        new_det = diff > 0.f ? new_det : prev_det + diff * volume_correction;
        p.jp *= prev_det / new_det;
        p.log_vol_gain += logf(prev_det) - logf(new_det);
        s[0] = n0; s[1] = n1;
        p.plastic_hardening += dq;
        p.F = svd2_recompose(svd.u, s, svd.v);
        return true;
    }
    case SPK_PLASTIC_NACC: { // plasticity_nacc.rs:61-165
        const float mu = m.p[0], kappa = m.p[1], xi = m.p[2], beta = m.p[3], mm = m.p[4];
        const bool hardening_enabled = (m.flags & 4u) != 0;
        float alpha = p.nacc_alpha;
        const float q0 = s[0] * s[0], q1 = s[1] * s[1];
        const float sq_tr = q0 + q1;
        const float p0 = kappa * (1.0e-5f + sinhf(xi * fmaxf(-alpha, 0.f)));
        const float j_e_tr = s[0] * s[1];
        const float sc = mu * powf(j_e_tr, -2.f / d);
        const float t0 = sc * (q0 - sq_tr / d), t1 = sc * (q1 - sq_tr / d);
        const float psi_kappa = kappa / 2.f * (j_e_tr - 1.f / j_e_tr);
        const float p_tr = -psi_kappa * j_e_tr;
        if (p_tr > p0 || p_tr < -beta * p0) {
            const float j_e_n1 = p_tr > p0 ? sqrtf(-2.f * p0 / kappa + 1.f) : sqrtf(2.f * beta * p0 / kappa + 1.f);
            const float v = powf(j_e_n1, 1.f / d);
            s[0] = s[1] = v;
            if (hardening_enabled) alpha += logf(j_e_tr / j_e_n1);
            p.nacc_alpha = alpha;
            p.F = svd2_recompose(svd.u, s, svd.v);
            return true;
        }
        const float y0 = (1.f + 2.f * beta) * ((6.f - d) / 2.f);
        const float y1 = mm * mm * (p_tr + beta * p0) * (p_tr - p0);
        const float tn2 = t0 * t0 + t1 * t1;
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
        s[0] = sqrtf(k * (t0 / tn) + sq_tr / d); s[1] = sqrtf(k * (t1 / tn) + sq_tr / d);
        p.nacc_alpha = alpha;
        p.F = svd2_recompose(svd.u, s, svd.v);
        return true;
    }
    case SPK_PLASTIC_RANKINE: { // plasticity_rankine.rs:35-78 with idx = [0, 1, DIM - 1] and no `DIM == 3` branch
        const float mu = m.p[0], lambda = m.p[1], tensile = m.p[2], softening = m.p[3];
        const float pr0 = logf(s[0]), pr1 = logf(s[1]);
        float ev0 = pr0, ev1 = pr1;
        // sorted [0, 1, 1] by eigenvalue: e0 <= e1 -> (e3, e2, e1) = (0, 1, 1); e0 > e1 -> (1, 1, 0)
        const bool swapped = pr0 > pr1;
        const float e_e1 = swapped ? pr0 : pr1; // eigv[e1]
        const float e_e2 = pr1;                 // eigv[e2]: index 1 in both orders
        const float soft = tensile - (p.plastic_hardening - 1.f);
        const float sum = pr0 + pr1;
        if (lambda * sum + 2.f * mu * e_e1 <= soft) return false;
        if ((2.f * mu + lambda) * e_e2 + lambda * (sum - e_e1) <= soft) {
            const float ne = (soft - lambda * (sum - e_e1)) / (2.f * mu + lambda);
            if (swapped) ev0 = ne; else ev1 = ne;
        } else {
            const float ne = soft / (2.f * mu + 3.f * lambda);
            ev0 = ev1 = ne;
        }
        const float dd0 = pr0 - ev0, dd1 = pr1 - ev1;
        p.plastic_hardening += softening * sqrtf(dd0 * dd0 + dd1 * dd1);
        p.plastic_hardening = fminf(p.plastic_hardening, tensile);
        s[0] = expf(ev0); s[1] = expf(ev1);
        p.F = svd2_recompose(svd.u, s, svd.v);
        return true;
    }
    case SPK_PLASTIC_SNOW: { // plasticity_snow.rs:26-48
        const float lo = 1.f - m.p[0], hi = 1.f + m.p[1];
        const float n0 = fminf(fmaxf(s[0], lo), hi), n1 = fminf(fmaxf(s[1], lo), hi);
        p.jp *= (s[0] * s[1]) / (n0 * n1);
        p.elastic_hardening = expf(m.p[2] * (1.f - p.jp));
        s[0] = n0; s[1] = n1;
        p.F = svd2_recompose(svd.u, s, svd.v);
        return true;
    }
    default: return false;
    }
}

// CorotatedLinearElasticity::kirchhoff_stress, elasticity_corotated_linear.rs:31-77, from the shared SVD of F
__device__ inline Mat2 corotated_stress2(const DevModel& m, float phase, float hardening, const Mat2& F, const Svd2& svd, float j) {
    const float lambda = m.c[1], mu = m.c[2];
    const float sph = lambda * hardening * (j - 1.f) * j;
    const float k = 2.f * mu * hardening;
    if (phase == 1.f) {
        float sm[2] = {svd.s[0] - 1.f, svd.s[1] - 1.f};
        Mat2 t = mat2_mul_bt(svd2_recompose(svd.u, sm, svd.v), F);
#pragma unroll
        for (int i = 0; i < 4; ++i) t.m[i] *= k;
        t.m[0] += sph; t.m[3] += sph;
        return t;
    }
    float sp[2], sn[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) { sp[i] = fmaxf(svd.s[i] - 1.f, 0.f); sn[i] = fminf(svd.s[i] - 1.f, 0.f); }
    Mat2 pos = mat2_mul_bt(svd2_recompose(svd.u, sp, svd.v), F);
    Mat2 neg = mat2_mul_bt(svd2_recompose(svd.u, sn, svd.v), F);
#pragma unroll
    for (int i = 0; i < 4; ++i) { pos.m[i] *= k; neg.m[i] *= k; }
    if (j < 1.f) { neg.m[0] += sph; neg.m[3] += sph; }
    else { pos.m[0] += sph; pos.m[3] += sph; }
    const float phase_coeff = ((m.flags & 1u) && phase == 0.f) ? 0.f : 1.f;
    Mat2 out;
#pragma unroll
    for (int i = 0; i < 4; ++i) out.m[i] = pos.m[i] * phase_coeff + neg.m[i];
    return out;
}
// elasticity_neo_hookean.rs:43-67 (k: literal 2/3; J exponent and deviatoric trace: DIM)
__device__ inline Mat2 neo_hookean_stress2(const DevModel& m, float phase, float hardening, const Mat2& F, float j) {
    const float lambda = m.c[1], mu = m.c[2];
    const float pc = (1.f - 0.001f) * phase * phase + 0.001f;
    const float k = 2.f / 3.f * mu * hardening + lambda * hardening;
    Mat2 cg = mat2_mul_bt(F, F);
    const float sphc = (cg.m[0] + cg.m[3]) / 2.f;
    cg.m[0] -= sphc; cg.m[3] -= sphc;
    const float sc = (mu * hardening) * powf(j, -2.f / 2.f);
    const float vol = k / 2.f * (j * j - 1.f);
    Mat2 out;
#pragma unroll
    for (int i = 0; i < 4; ++i) out.m[i] = (sc * cg.m[i]) * pc;
    if (j >= 1.f) { out.m[0] = (sc * cg.m[0] + vol) * pc; out.m[3] = (sc * cg.m[3] + vol) * pc; }
    else { out.m[0] += vol; out.m[3] += vol; }
    return out;
}

// estimate_particle_timestep_length (particle_updater.rs:47-70) + timestep.rs:74-112: candidate BEFORE the clamps
__device__ inline float particle_dt_candidate2(const DevModel& m, float h, float mass, float volume0, float f00, float elastic_hardening, float vx, float vy) {
    const float vn2 = vx * vx + vy * vy;
    const float vn = sqrtf(vn2);
    float cand = FLT_MAX;
    bool velocity_bound = false;
    switch (m.ctag) {
    case SPK_CONSTITUTIVE_COROTATED:
    case SPK_CONSTITUTIVE_NEO_HOOKEAN: { // elasticity_sound_speed_timestep_bound.rs:22-34; NeoHookean: swapped arguments (quirk 4)
        const bool nh = m.ctag == SPK_CONSTITUTIVE_NEO_HOOKEAN;
        const float hard = nh ? h : elastic_hardening;
        const float cw = nh ? elastic_hardening : h;
        const float lambda = m.c[1], mu = m.c[2];
        const float bulk = (lambda + 2.f * mu / 3.f) * hard, shear = mu * hard;
        const float density0 = mass / volume0;
        const float c_dir = sqrtf((bulk + 4.f / 3.f * shear) * 1.f / density0);
        cand = m.c[0] * cw / fmaxf(vn, c_dir);
        velocity_bound = true;
        break;
    }
    case SPK_CONSTITUTIVE_MONAGHAN: { // eos_monaghan_sph.rs:72-99, d = DIM
        const float j = f00;
        const float density0 = mass / volume0;
        const float pressure = -monaghan_pressure(m, mass, volume0, density0 / j);
        const float single = (h / j) * sqrtf(density0 * (j - 1.f) / (6.f * pressure * 2.f));
        const float c_square = fmaxf(vn2, 1.f) / 0.1f;
        cand = fminf(single, h / sqrtf(c_square));
        velocity_bound = true;
        break;
    }
    default: break;
    }
    if (velocity_bound) cand = fminf(cand, h / vn);
    return cand;
}
