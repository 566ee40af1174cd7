// mpm_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs, never by the product path (sparkl_b200/ never links it).
//
// What it is: a restatement, in plain C++17/f32, of the algorithm the reference's GPU path executes for one
// `CudaMpmPipeline::step` (dimforge/sparkl 0.4.0; file:line of every restated function is cited inline; all
// paths are relative to the reference root).  It follows the reference's per-particle operation order
// (NBH_SHIFTS order, G2P -> update -> P2G, 3 SVD call sites, ...).  Two things are canonicalised because the
// reference leaves them to atomics races: active-block ids are the rank of the block key in
// ascending order, and particles inside a block are ordered by particle id.  Grid sums (float `red.add` in the
// reference, order unspecified) are accumulated in double and rounded once.
//
// Third-party arithmetic that is not under /root/reference: nalgebra 0.32.1 (`svd_unordered`,
// `try_symmetric_eigen`, `determinant`) and parry3d 0.13.0 (`Cuboid::project_point`).  Only U·f(Σ)·Vᵀ products
// and eigenvalue extrema are consumed by the path, so the SVD / eigen solvers here are Jacobi iterations run in
// double precision (the exact quantities nalgebra approximates in f32) and rounded to f32.
//
// PARITY STATUS: the reference holds no golden vectors for this path; this restatement is pinned
// against the reference's own shipped GPU kernels (oracle/_ref, see ref_harness.cpp) on the B200 box.
#include "../include/sparkl_abi.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

typedef float Real;

// ------------------------------------------------------------------------------------------------ small math
struct V3 {
    Real x, y, z;
    Real& operator[](int i) { return (&x)[i]; }
    Real operator[](int i) const { return (&x)[i]; }
};
static inline V3 v3(Real x, Real y, Real z) { return V3{x, y, z}; }
static inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
static inline V3 operator*(V3 a, Real s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator*(Real s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator/(V3 a, Real s) { return v3(a.x / s, a.y / s, a.z / s); }
static inline Real dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline Real norm_squared(V3 a) { return dot(a, a); }
static inline Real norm(V3 a) { return std::sqrt(norm_squared(a)); }
static inline bool is_zero(V3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }

// column-major 3x3, same storage as nalgebra Matrix3 / ParticleVolume::deformation_gradient
struct M3 {
    Real m[9];
    Real& operator()(int r, int c) { return m[c * 3 + r]; }
    Real operator()(int r, int c) const { return m[c * 3 + r]; }
};
static inline M3 m3_zero() { M3 a; for (int i = 0; i < 9; ++i) a.m[i] = 0.0f; return a; }
static inline M3 m3_identity() { M3 a = m3_zero(); a(0, 0) = a(1, 1) = a(2, 2) = 1.0f; return a; }
static inline M3 operator+(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] + b.m[i]; return r; }
static inline M3 operator-(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] - b.m[i]; return r; }
static inline M3 operator*(const M3& a, Real s) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] * s; return r; }
static inline M3 operator*(Real s, const M3& a) { return a * s; }
static inline M3 operator*(const M3& a, const M3& b) {
    M3 r;
    for (int c = 0; c < 3; ++c)
        for (int rr = 0; rr < 3; ++rr) {
            Real s = 0.0f;
            for (int k = 0; k < 3; ++k) s += a(rr, k) * b(k, c);
            r(rr, c) = s;
        }
    return r;
}
static inline V3 operator*(const M3& a, V3 v) {
    return v3(a(0, 0) * v.x + a(0, 1) * v.y + a(0, 2) * v.z, a(1, 0) * v.x + a(1, 1) * v.y + a(1, 2) * v.z,
              a(2, 0) * v.x + a(2, 1) * v.y + a(2, 2) * v.z);
}
static inline M3 transpose(const M3& a) { M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = a(j, i); return r; }
static inline Real trace(const M3& a) { return a(0, 0) + a(1, 1) + a(2, 2); }
// nalgebra Matrix3::determinant (cofactor expansion along the first row)
static inline Real determinant(const M3& a) {
    Real m12_23 = a(1, 1) * a(2, 2) - a(2, 1) * a(1, 2);
    Real m11_23 = a(1, 0) * a(2, 2) - a(2, 0) * a(1, 2);
    Real m11_22 = a(1, 0) * a(2, 1) - a(2, 0) * a(1, 1);
    return a(0, 0) * m12_23 - a(0, 1) * m11_23 + a(0, 2) * m11_22;
}
// v * w^T
static inline M3 outer(V3 v, V3 w) { M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = v[i] * w[j]; return r; }

// Rust semantics helpers
static inline Real rmin(Real a, Real b) { return std::fmin(a, b); } // f32::min ignores NaN
static inline Real rmax(Real a, Real b) { return std::fmax(a, b); }
static inline Real signum(Real x) { return std::isnan(x) ? x : (std::signbit(x) ? -1.0f : 1.0f); }
static inline int64_t sat_i64(Real x) {
    if (std::isnan(x)) return 0;
    if (x >= 9.2233720368547758e18f) return INT64_MAX;
    if (x <= -9.2233720368547758e18f) return INT64_MIN;
    return (int64_t)x;
}
static inline int32_t sat_i32(Real x) {
    if (std::isnan(x)) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
}
static inline uint64_t sat_u64(Real x) {
    if (std::isnan(x) || x <= 0.0f) return 0;
    if (x >= 1.8446744073709552e19f) return UINT64_MAX;
    return (uint64_t)x;
}
// compiler-rt __powisf2 (what `f32::powi` lowers to)
static inline Real powi(Real a, int b) {
    const bool recip = b < 0;
    Real r = 1.0f;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0f / r : r;
}
// src_core/utils/physics.rs:7-14
static inline Real inv_exact(Real e) { return e == 0.0f ? 0.0f : 1.0f / e; }

// ------------------------------------------------------------------------------------------------ SVD / eigen
// Stand-in for nalgebra 0.32.1 `svd_unordered(true,true)` (call sites: elasticity_corotated_linear.rs:38,88,107,
// plasticity_drucker_prager.rs:80, plasticity_nacc.rs:71, plasticity_rankine.rs:43, plasticity_snow.rs:36).
// One-sided Jacobi (Hestenes) in double; singular values are non-negative and unordered, U/V orthogonal.
// ORC_FAST (liboracle_fast.so, the CPU ARM of bench.py only — never the checker): the same algorithm compiled for speed: f32 Jacobi with
// an f32 stopping rule instead of the f64 iteration to round-off, f32 node sums, -O3 -march=x86-64-v3.  What a competently written CPU
// path of the reference does (its own SVDs are nalgebra's f32 ones); the parity tests always use the accurate build.
#ifdef ORC_FAST
typedef float SvdReal;
#define ORC_SVD_TOL 1e-7f
typedef float AccReal;
#else
typedef double SvdReal;
#define ORC_SVD_TOL 1e-17
typedef double AccReal;
#endif
struct Svd3 { M3 u; V3 s; M3 vt; };
static Svd3 svd_unordered(const M3& a) {
    SvdReal U[3][3], V[3][3];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { U[r][c] = a(r, c); V[r][c] = (r == c) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                SvdReal alpha = 0, beta = 0, gamma = 0;
                for (int r = 0; r < 3; ++r) { alpha += U[r][p] * U[r][p]; beta += U[r][q] * U[r][q]; gamma += U[r][p] * U[r][q]; }
                if (gamma == 0 || std::fabs(gamma) <= (SvdReal)ORC_SVD_TOL * std::sqrt(alpha * beta)) continue;
                rotated = true;
                SvdReal zeta = (beta - alpha) / (2 * gamma);
                SvdReal t = (zeta >= 0 ? (SvdReal)1 : (SvdReal)-1) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
                SvdReal c = 1 / std::sqrt(1 + t * t), s = c * t;
                for (int r = 0; r < 3; ++r) {
                    SvdReal up = U[r][p], uq = U[r][q];
                    U[r][p] = c * up - s * uq; U[r][q] = s * up + c * uq;
                    SvdReal vp = V[r][p], vq = V[r][q];
                    V[r][p] = c * vp - s * vq; V[r][q] = s * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    SvdReal sig[3];
    bool ok[3];
    SvdReal smax = 0;
    for (int c = 0; c < 3; ++c) { sig[c] = std::sqrt(U[0][c] * U[0][c] + U[1][c] * U[1][c] + U[2][c] * U[2][c]); smax = std::max(smax, sig[c]); }
    for (int c = 0; c < 3; ++c) {
        ok[c] = sig[c] > (SvdReal)(sizeof(SvdReal) == 8 ? 1e-300 : 1e-37) && sig[c] > (SvdReal)(sizeof(SvdReal) == 8 ? 1e-14 : 1e-7) * smax;
        if (ok[c]) for (int r = 0; r < 3; ++r) U[r][c] /= sig[c];
    }
    // complete U to an orthonormal basis where a singular value vanished
    for (int c = 0; c < 3; ++c) {
        if (ok[c]) continue;
        int a1 = (c + 1) % 3, a2 = (c + 2) % 3;
        SvdReal w[3] = {0, 0, 0};
        if (ok[a1] && ok[a2]) {
            w[0] = U[1][a1] * U[2][a2] - U[2][a1] * U[1][a2];
            w[1] = U[2][a1] * U[0][a2] - U[0][a1] * U[2][a2];
            w[2] = U[0][a1] * U[1][a2] - U[1][a1] * U[0][a2];
        } else {
            // pick any unit vector orthogonal to the valid columns (Gram-Schmidt over the canonical basis)
            SvdReal best = -1;
            for (int e = 0; e < 3; ++e) {
                SvdReal cand[3] = {0, 0, 0};
                cand[e] = 1;
                for (int k = 0; k < 3; ++k) {
                    if (k == c || !ok[k]) continue;
                    SvdReal d = cand[0] * U[0][k] + cand[1] * U[1][k] + cand[2] * U[2][k];
                    for (int r = 0; r < 3; ++r) cand[r] -= d * U[r][k];
                }
                SvdReal n = std::sqrt(cand[0] * cand[0] + cand[1] * cand[1] + cand[2] * cand[2]);
                if (n > best) { best = n; for (int r = 0; r < 3; ++r) w[r] = cand[r] / n; }
            }
        }
        for (int r = 0; r < 3; ++r) U[r][c] = w[r];
        ok[c] = true;
        sig[c] = 0.0;
    }
    Svd3 out;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { out.u(r, c) = (Real)U[r][c]; out.vt(c, r) = (Real)V[r][c]; }
    out.s = v3((Real)sig[0], (Real)sig[1], (Real)sig[2]);
    return out;
}
// SVD::recompose: U * diag(s) * V^T
static inline M3 recompose(const Svd3& svd) {
    M3 us = svd.u;
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) us(r, c) *= svd.s[c];
    return us * svd.vt;
}
// stand-in for nalgebra `try_symmetric_eigen(1e-6, 100)` eigenvalues (failure_maximum_stress.rs:21): cyclic Jacobi, double.
static void symmetric_eigenvalues(const M3& a, double ev[3]) {
    double A[3][3];
    // nalgebra's SymmetricEigen reads the lower triangle
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) A[r][c] = (r >= c) ? a(r, c) : a(c, r);
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
        if (off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) { double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
                for (int k = 0; k < 3; ++k) { double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
            }
    }
    ev[0] = A[0][0]; ev[1] = A[1][1]; ev[2] = A[2][2];
}

// ------------------------------------------------------------------------------------------------ grid keys
// src_kernels/gpu_grid.rs:73-120
static const uint64_t PACK_ORIGIN = 0x00100000ull;
static const uint64_t KEY_MASK = 0x1fffffull;
static inline uint64_t pack_key(uint64_t x, uint64_t y, uint64_t z) { return (x & KEY_MASK) | ((y & KEY_MASK) << 21) | ((z & KEY_MASK) << 42); }
static inline void unpack_key(uint64_t k, uint64_t out[3]) { out[0] = k & KEY_MASK; out[1] = (k >> 21) & KEY_MASK; out[2] = (k >> 42) & KEY_MASK; }
// src_kernels/gpu_grid.rs:271-279 (block_associated_to_point), main block coordinates before packing
static inline void main_block_of_point(const float p[3], Real h, uint64_t out[3]) {
    for (int a = 0; a < 3; ++a) {
        int64_t coord = sat_i64(std::round(p[a] / h));
        out[a] = (uint64_t)(coord - 2 + (int64_t)PACK_ORIGIN * 4) / 4; // `as usize / 4`
    }
}
// NBH_SHIFTS, src_kernels/gpu_grid.rs:32-60 (NBH_SHIFTS_SHARED[i] == x + 8y + 64z, :63-66)
static const int NBH[27][3] = {{2, 2, 2}, {2, 0, 2}, {2, 1, 2}, {0, 2, 2}, {0, 0, 2}, {0, 1, 2}, {1, 2, 2}, {1, 0, 2}, {1, 1, 2},
                               {2, 2, 0}, {2, 0, 0}, {2, 1, 0}, {0, 2, 0}, {0, 0, 0}, {0, 1, 0}, {1, 2, 0}, {1, 0, 0}, {1, 1, 0},
                               {2, 2, 1}, {2, 0, 1}, {2, 1, 1}, {0, 2, 1}, {0, 0, 1}, {0, 1, 1}, {1, 2, 1}, {1, 0, 1}, {1, 1, 1}};

struct Grid {
    std::vector<SpkActiveBlockHeader> headers;          // id == rank of virtual_id (canonical form)
    std::vector<SpkGpuGridNode> nodes;                  // 64 per block
    std::unordered_map<uint64_t, uint32_t> map;         // GridHashMap (hashmap.rs:37-88): only set semantics matter
    void clear() { headers.clear(); nodes.clear(); map.clear(); }
    int find(uint64_t key) const { auto it = map.find(key); return it == map.end() ? -1 : (int)it->second; }
};
static inline SpkGpuGridNode default_node() { // GpuGridNode::default, gpu_grid.rs:461-474
    SpkGpuGridNode n;
    std::memset(&n, 0, sizeof(n));
    n.projection_status_tag = SPK_PROJ_NOT_COMPUTED;
    return n;
}

} // namespace

struct orc_ctx {
    Real cell_width;
    std::vector<SpkParticleStatus> status;
    std::vector<SpkParticlePosition> pos;
    std::vector<SpkParticleVelocity> vel;
    std::vector<SpkParticleVolume> volume;
    std::vector<SpkParticlePhase> phase;
    std::vector<SpkParticleData> data;
    std::vector<uint32_t> sorted_ids;
    std::vector<SpkGpuParticleModel> models;
    std::vector<SpkGpuCollider> colliders;
    std::vector<std::vector<float>> collider_data; // owned copies of the heightfield matrices the colliders point to
    std::vector<std::vector<uint32_t>> collider_indices; // owned copies of the trimesh index lists
    Grid grids[2];
    int curr = 0; // grids[curr] = read-only previous grid, grids[1-curr] = next
    std::vector<SpkDispatchBlock2ActiveBlock> dispatch;
    uint64_t dt_acc = 0; // GpuTimestepLength
    uint32_t num_halo = 0;
    Grid& curr_grid() { return grids[curr]; }
    Grid& next_grid() { return grids[1 - curr]; }
};

namespace {

static inline M3 load_F(const SpkParticleVolume& v) { M3 f; std::memcpy(f.m, v.deformation_gradient, sizeof(f.m)); return f; }
static inline void store_F(SpkParticleVolume& v, const M3& f) { std::memcpy(v.deformation_gradient, f.m, sizeof(f.m)); }

// ------------------------------------------------------------------------------------------------ models
// elasticity_corotated_linear.rs:31-77
static M3 corotated_kirchhoff_stress(const SpkCorotatedLinearElasticity& m, Real phase, Real hardening, const M3& F) {
    Real j = determinant(F);
    Svd3 svd = svd_unordered(F);
    for (int i = 0; i < 3; ++i) svd.s[i] = svd.s[i] - 1.0f;
    if (phase == 1.0f) {
        return (2.0f * m.mu * hardening) * recompose(svd) * transpose(F) + (m.lambda * hardening * (j - 1.0f) * j) * m3_identity();
    }
    Svd3 pos_def = svd, neg_def = svd;
    for (int i = 0; i < 3; ++i) { pos_def.s[i] = rmax(pos_def.s[i], 0.0f); neg_def.s[i] = rmin(neg_def.s[i], 0.0f); }
    M3 pos_dev = 2.0f * m.mu * hardening * recompose(pos_def) * transpose(F);
    M3 neg_dev = 2.0f * m.mu * hardening * recompose(neg_def) * transpose(F);
    M3 spherical = (m.lambda * hardening * (j - 1.0f) * j) * m3_identity();
    M3 pos_part, neg_part;
    if (j < 1.0f) { pos_part = pos_dev; neg_part = neg_dev + spherical; }
    else { pos_part = pos_dev + spherical; neg_part = neg_dev; }
    Real phase_coeff = (m.split_stress_on_failure && phase == 0.0f) ? 0.0f : 1.0f;
    return pos_part * phase_coeff + neg_part;
}
// elasticity_corotated_linear.rs:104-118
static Real corotated_pos_energy(const SpkCorotatedLinearElasticity& m, const M3& F, Real hardening) {
    Real j = determinant(F);
    Svd3 svd = svd_unordered(F);
    V3 s;
    for (int i = 0; i < 3; ++i) s[i] = rmax(svd.s[i] - 1.0f, 0.0f);
    Real pos_dev_part = m.mu * hardening * norm_squared(s);
    Real spherical_part = m.lambda * hardening / 2.0f * (j - 1.0f) * (j - 1.0f);
    return j < 1.0f ? pos_dev_part : pos_dev_part + spherical_part;
}
// src_core/dynamics/timestep/elasticity_sound_speed_timestep_bound.rs:22-34
static Real sound_speed_bound(Real alpha, Real bulk, Real shear, Real density0, V3 v, Real cell_width) {
    Real c_dir = std::sqrt((bulk + 4.0f / 3.0f * shear) * 1.0f / density0);
    Real max_denom = rmax(norm(v), c_dir);
    return alpha * cell_width / max_denom;
}
// elasticity_corotated_linear.rs:126-142 and elasticity_neo_hookean.rs:149-165 (identical bodies)
static Real elastic_timestep_bound(Real cfl, Real lambda, Real mu, Real density0, V3 v, Real elastic_hardening, Real cell_width) {
    Real bulk = lambda + 2.0f * mu / 3.0f; // physics.rs:37-39
    Real shear = mu;
    return sound_speed_bound(cfl, bulk * elastic_hardening, shear * elastic_hardening, density0, v, cell_width);
}
static inline M3 deviatoric_part(const M3& t) { // DecomposedTensor::decompose
    Real sph = trace(t) / 3.0f;
    return t - sph * m3_identity();
}
static inline Real nh_phase_coeff(Real phase) { const Real R = 0.001f; return (1.0f - R) * phase * phase + R; } // elasticity_neo_hookean.rs:70-73
// elasticity_neo_hookean.rs:43-67
static M3 neo_hookean_kirchhoff_stress(const SpkNeoHookeanElasticity& m, Real phase, Real hardening, const M3& F) {
    Real pc = nh_phase_coeff(phase);
    Real j = determinant(F);
    Real k = 2.0f / 3.0f * m.mu * hardening + m.lambda * hardening;
    M3 cg = F * transpose(F);
    M3 dev = (m.mu * hardening) * std::pow(j, -2.0f / 3.0f) * deviatoric_part(cg);
    M3 vol = k / 2.0f * (j * j - 1.0f) * m3_identity();
    M3 pos_part, neg_part;
    if (j >= 1.0f) { pos_part = dev + vol; neg_part = m3_zero(); }
    else { pos_part = dev; neg_part = vol; }
    return pos_part * pc + neg_part;
}
// elasticity_neo_hookean.rs:117-141
static Real neo_hookean_pos_energy(const SpkNeoHookeanElasticity& m, Real phase, Real hardening, const M3& F) {
    Real pc = nh_phase_coeff(phase);
    Real j = determinant(F);
    Real k = 2.0f / 3.0f * m.mu * hardening + m.lambda * hardening;
    Real dev = hardening * m.mu / 2.0f * (trace(F * transpose(F)) * std::pow(j, -2.0f / 3.0f) - 3.0f);
    Real vol = k / 2.0f * ((j * j - 1.0f) / 2.0f - std::log(j));
    return j < 1.0f ? dev * pc : (dev + vol) * phase;
}
// eos_monaghan_sph.rs:28-38
static Real monaghan_pressure(const SpkMonaghanSphEos& m, Real mass, Real volume0, Real density_fluid) {
    Real density0 = mass / volume0;
    return rmax(m.pressure0 * (powi(density_fluid / density0, (int)m.gamma) - 1.0f), -m.max_neg_pressure);
}
// eos_monaghan_sph.rs:44-63
static M3 monaghan_kirchhoff_stress(const SpkMonaghanSphEos& m, Real mass, Real volume0, Real density_fluid, Real jdet, const M3& vel_grad) {
    M3 stress = (-monaghan_pressure(m, mass, volume0, density_fluid) * jdet) * m3_identity();
    if (m.viscosity != 0.0f) {
        M3 strain_rate = (vel_grad + transpose(vel_grad)) * 0.5f; // physics.rs:69-71
        stress = stress + (2.0f * m.viscosity * jdet) * deviatoric_part(strain_rate);
    }
    return stress;
}
// eos_monaghan_sph.rs:72-99
static Real monaghan_timestep_bound(const SpkMonaghanSphEos& m, Real j, Real mass, Real volume0, Real density_fluid, V3 v, Real cell_width) {
    Real density0 = mass / volume0;
    Real k = 6.0f, d = 3.0f;
    Real pressure = -monaghan_pressure(m, mass, volume0, density_fluid);
    Real single_particle_dt = (cell_width / j) * std::sqrt(density0 * (j - 1.0f) / (k * pressure * d));
    Real density_fluctuation = 0.1f;
    Real c_square = rmax(norm_squared(v), 1.0f) / density_fluctuation;
    Real cfl_dt = cell_width / std::sqrt(c_square);
    return rmin(single_particle_dt, cfl_dt);
}

// plasticity_drucker_prager.rs:36-104
static void drucker_prager_update(const SpkDruckerPragerPlasticity& m, Real phase, M3& F, Real& jp, Real& plastic_hardening, Real& log_vol_gain) {
    if (m.only_active_when_failed && phase != 0.0f) return;
    Svd3 svd = svd_unordered(F);
    // alpha(q), :64-69
    Real q = plastic_hardening;
    Real angle = m.h0 + (m.h1 * q - m.h3) * std::exp(-m.h2 * q);
    Real s_angle = std::sin(angle);
    Real alpha = std::sqrt(2.0f / 3.0f) * (2.0f * s_angle) / (3.0f - s_angle);
    // project_deformation_gradient, :36-62
    const Real d = 3.0f;
    V3 strain = v3(std::log(svd.s.x), std::log(svd.s.y), std::log(svd.s.z)) + v3(log_vol_gain / d, log_vol_gain / d, log_vol_gain / d);
    Real strain_trace = strain.x + strain.y + strain.z;
    V3 dev = strain - v3(strain_trace / d, strain_trace / d, strain_trace / d);
    V3 new_s;
    Real dq;
    if (is_zero(dev) || strain_trace > 0.0f) {
        new_s = v3(1.0f, 1.0f, 1.0f);
        dq = norm(strain);
    } else {
        Real gamma = norm(dev) + (d * m.lambda + 2.0f * m.mu) / (2.0f * m.mu) * strain_trace * alpha;
        if (gamma <= 0.0f) return; // inside the yield surface
        V3 hh = strain - gamma * (dev / norm(dev));
        new_s = v3(std::exp(hh.x), std::exp(hh.y), std::exp(hh.z));
        dq = gamma;
    }
    Real prev_det = svd.s.x * svd.s.y * svd.s.z;
    Real new_det = new_s.x * new_s.y * new_s.z;
    Real diff = new_det - prev_det;
    new_det = diff > 0.0f ? new_det : prev_det + diff * m.volume_correction;
    jp *= prev_det / new_det;
    log_vol_gain += std::log(prev_det) - std::log(new_det);
    svd.s = new_s;
    plastic_hardening += dq;
    F = recompose(svd);
}
// plasticity_nacc.rs:61-165
static void nacc_update(const SpkNaccPlasticity& m, M3& F, Real& alpha_io) {
    Real alpha = alpha_io;
    const Real xi = m.hardening_factor, beta = m.cohesion, mm = m.friction, d = 3.0f;
    Svd3 svd = svd_unordered(F);
    V3 sq = v3(svd.s.x * svd.s.x, svd.s.y * svd.s.y, svd.s.z * svd.s.z);
    Real sq_tr = sq.x + sq.y + sq.z;
    Real p0 = m.kappa * (1.0e-5f + std::sinh(xi * rmax(-alpha, 0.0f)));
    Real j_e_tr = svd.s.x * svd.s.y * svd.s.z;
    V3 s_tr = (m.mu * std::pow(j_e_tr, -2.0f / d)) * (sq - v3(sq_tr / d, sq_tr / d, sq_tr / d));
    Real psi_kappa = m.kappa / 2.0f * (j_e_tr - 1.0f / j_e_tr);
    Real p_tr = -psi_kappa * j_e_tr;
    if (p_tr > p0) {
        Real j_e_n1 = std::sqrt(-2.0f * p0 / m.kappa + 1.0f);
        Real v = std::pow(j_e_n1, 1.0f / d);
        svd.s = v3(v, v, v);
        if (m.hardening_enabled) alpha += std::log(j_e_tr / j_e_n1);
        F = recompose(svd); alpha_io = alpha; return;
    }
    if (p_tr < -beta * p0) {
        Real j_e_n1 = std::sqrt(2.0f * beta * p0 / m.kappa + 1.0f);
        Real v = std::pow(j_e_n1, 1.0f / d);
        svd.s = v3(v, v, v);
        if (m.hardening_enabled) alpha += std::log(j_e_tr / j_e_n1);
        F = recompose(svd); alpha_io = alpha; return;
    }
    Real y0 = (1.0f + 2.0f * beta) * ((6.0f - d) / 2.0f);
    Real y1 = mm * mm * (p_tr + beta * p0) * (p_tr - p0);
    Real y = y0 * norm_squared(s_tr) + y1;
    if (y < 1.0e-4f) return; // inside the yield surface: F, alpha unchanged
    if (m.hardening_enabled && p0 > 1.0e-4f && p_tr < p0 - 1.0e-4f && p_tr > -beta * p0 + 1.0e-4f) {
        Real p_c = (1.0f - beta) * p0 / 2.0f;
        Real q_tr = std::sqrt((6.0f - d) / 2.0f) * norm(s_tr);
        Real dx = p_c - p_tr, dy = 0.0f - q_tr;
        Real dn = std::sqrt(dx * dx + dy * dy);
        dx /= dn; dy /= dn;
        Real c = mm * mm * (p_c + beta * p0) * (p_c - p0);
        Real b = mm * mm * dx * (2.0f * p_c - p0 + beta * p0);
        Real a = mm * mm * dx * dx + (1.0f + 2.0f * beta) * dy * dy;
        Real discr = std::sqrt(b * b - 4.0f * a * c);
        Real l1 = (-b + discr) / (2.0f * a), l2 = (-b - discr) / (2.0f * a);
        Real p1 = p_c + l1 * dx, p2 = p_c + l2 * dx;
        Real p_x = (p_tr - p_c) * (p1 - p_c) > 0.0f ? p1 : p2;
        Real j_e_x = std::sqrt(std::fabs(-2.0f * p_x / m.kappa + 1.0f));
        if (j_e_x > 1.0e-4f) alpha += std::log(j_e_tr / j_e_x);
    }
    V3 b_e_n1 = (std::sqrt(-y1 / y0) * (std::pow(j_e_tr, 2.0f / d) / m.mu)) * (s_tr / norm(s_tr)) + v3(sq_tr / d, sq_tr / d, sq_tr / d);
    svd.s = v3(std::sqrt(b_e_n1.x), std::sqrt(b_e_n1.y), std::sqrt(b_e_n1.z));
    F = recompose(svd);
    alpha_io = alpha;
}
// plasticity_rankine.rs:35-78
static void rankine_update(const SpkRankinePlasticity& m, M3& F, Real& plastic_hardening) {
    const Real lambda = m.lambda, mu = m.mu;
    Svd3 svd = svd_unordered(F);
    V3 eigv = v3(std::log(svd.s.x), std::log(svd.s.y), std::log(svd.s.z));
    V3 prev = eigv;
    int idx[3] = {0, 1, 2};
    // sort_unstable_by ascending eigv (ties/NaN -> Equal); a 3-element insertion sort reproduces any stable/unstable tie-free order
    for (int i = 1; i < 3; ++i) { int k = idx[i]; int jx = i - 1; while (jx >= 0 && eigv[idx[jx]] > eigv[k]) { idx[jx + 1] = idx[jx]; --jx; } idx[jx + 1] = k; }
    int e3 = idx[0], e2 = idx[1], e1 = idx[2];
    Real soft = m.tensile_strength - (plastic_hardening - 1.0f);
    Real sum = eigv.x + eigv.y + eigv.z;
    if (lambda * sum + 2.0f * mu * eigv[e1] <= soft) {
        return;
    } else if ((2.0f * mu + lambda) * eigv[e2] + lambda * (sum - eigv[e1]) <= soft) {
        eigv[e1] = (soft - lambda * (sum - eigv[e1])) / (2.0f * mu + lambda);
    } else if ((2.0f * mu + 3.0f * lambda) * eigv[e3] <= soft) {
        Real ne = (soft - lambda * (sum - eigv[e1] - eigv[e2])) / (2.0f * mu + 2.0f * lambda);
        eigv[e1] = ne; eigv[e2] = ne;
    } else {
        Real ne = soft / (2.0f * mu + 3.0f * lambda);
        eigv = v3(ne, ne, ne);
    }
    plastic_hardening += m.softening_rate * norm(prev - eigv);
    plastic_hardening = rmin(plastic_hardening, m.tensile_strength);
    svd.s = v3(std::exp(eigv.x), std::exp(eigv.y), std::exp(eigv.z));
    F = recompose(svd);
}
// plasticity_snow.rs:26-48
static void snow_update(const SpkSnowPlasticity& m, M3& F, Real& elastic_hardening, Real& jp) {
    Svd3 svd = svd_unordered(F);
    V3 ns;
    for (int i = 0; i < 3; ++i) ns[i] = rmin(rmax(svd.s[i], 1.0f - m.min_epsilon), 1.0f + m.max_epsilon);
    jp *= (svd.s.x * svd.s.y * svd.s.z) / (ns.x * ns.y * ns.z);
    elastic_hardening = std::exp(m.hardening_coeff * (1.0f - jp));
    svd.s = ns;
    F = recompose(svd);
}
// failure_maximum_stress.rs:20-28
static bool maximum_stress_failed(const SpkMaximumStressFailure& m, const M3& stress) {
    for (int i = 0; i < 9; ++i) if (!std::isfinite(stress.m[i])) return false; // try_symmetric_eigen does not converge -> None
    double ev[3];
    symmetric_eigenvalues(stress, ev);
    Real mn = (Real)std::min(ev[0], std::min(ev[1], ev[2]));
    Real mx = (Real)std::max(ev[0], std::max(ev[1], ev[2]));
    return mx > m.max_principal_stress || (mx - mn) / 2.0f > m.max_shear_stress;
}

// GpuConstitutiveModel dispatch, src_kernels/gpu_constitutive_model.rs:19-149
static inline bool model_is_fluid(const SpkGpuParticleModel& m) { return m.constitutive_tag == SPK_CONSTITUTIVE_MONAGHAN; }
static Real model_pos_energy(const SpkGpuParticleModel& m, const SpkParticleData& d, const SpkParticleVolume& vol, Real phase) {
    switch (m.constitutive_tag) {
    case SPK_CONSTITUTIVE_COROTATED: return corotated_pos_energy(m.constitutive.corotated.m, load_F(vol), d.elastic_hardening);
    case SPK_CONSTITUTIVE_NEO_HOOKEAN: return neo_hookean_pos_energy(m.constitutive.neo_hookean.m, phase, d.elastic_hardening, load_F(vol));
    default: return 0.0f;
    }
}
static M3 model_kirchhoff_stress(const SpkGpuParticleModel& m, const SpkParticleData& d, const SpkParticleVolume& vol, Real phase, const M3& vel_grad) {
    switch (m.constitutive_tag) {
    case SPK_CONSTITUTIVE_COROTATED: return corotated_kirchhoff_stress(m.constitutive.corotated.m, phase, d.elastic_hardening, load_F(vol));
    case SPK_CONSTITUTIVE_NEO_HOOKEAN: return neo_hookean_kirchhoff_stress(m.constitutive.neo_hookean.m, phase, d.elastic_hardening, load_F(vol));
    case SPK_CONSTITUTIVE_MONAGHAN: {
        Real density_fluid = (vol.mass / vol.volume0) / vol.deformation_gradient[0]; // particle.rs:101-103
        return monaghan_kirchhoff_stress(m.constitutive.monaghan.m, vol.mass, vol.volume0, density_fluid, vol.plastic_deformation_gradient_det, vel_grad);
    }
    default: return m3_zero();
    }
}
// ActiveTimestepBounds: CONSTITUTIVE_MODEL_BOUND = 1, PARTICLE_VELOCITY_BOUND = 2 (constitutive_model.rs:3-12)
static inline unsigned model_active_bounds(const SpkGpuParticleModel& m) {
    switch (m.constitutive_tag) {
    case SPK_CONSTITUTIVE_COROTATED: case SPK_CONSTITUTIVE_NEO_HOOKEAN: return 1u | 2u | 8u;
    case SPK_CONSTITUTIVE_MONAGHAN: return 1u | 2u | 4u | 16u;
    default: return 0u;
    }
}
static Real model_timestep_bound(const SpkGpuParticleModel& m, const SpkParticleData& d, const SpkParticleVolume& vol, V3 v, Real cell_width) {
    Real density0 = vol.mass / vol.volume0;
    switch (m.constitutive_tag) {
    case SPK_CONSTITUTIVE_COROTATED: {
        const auto& c = m.constitutive.corotated.m;
        return elastic_timestep_bound(c.cfl_coeff, c.lambda, c.mu, density0, v, d.elastic_hardening, cell_width);
    }
    case SPK_CONSTITUTIVE_NEO_HOOKEAN: {
        // NOTE: the reference passes (cell_width, hardening) in swapped positions, gpu_constitutive_model.rs:129-137
        const auto& c = m.constitutive.neo_hookean.m;
        return elastic_timestep_bound(c.cfl_coeff, c.lambda, c.mu, density0, v, /*elastic_hardening=*/cell_width, /*cell_width=*/d.elastic_hardening);
    }
    case SPK_CONSTITUTIVE_MONAGHAN: {
        Real jf = vol.deformation_gradient[0];
        return monaghan_timestep_bound(m.constitutive.monaghan.m, jf, vol.mass, vol.volume0, density0 / jf, v, cell_width);
    }
    default: return 0.0f;
    }
}
// GpuPlasticModel::update_particle, src_kernels/gpu_plastic_model.rs:21-61
static void model_plastic_update(const SpkGpuParticleModel& m, SpkParticleData& d, SpkParticleVolume& vol, Real phase) {
    M3 F = load_F(vol);
    switch (m.plastic_tag) {
    case SPK_PLASTIC_DRUCKER_PRAGER: drucker_prager_update(m.plastic.drucker_prager.m, phase, F, vol.plastic_deformation_gradient_det, d.plastic_hardening, d.log_vol_gain); break;
    case SPK_PLASTIC_NACC: nacc_update(m.plastic.nacc.m, F, d.nacc_alpha); break;
    case SPK_PLASTIC_RANKINE: rankine_update(m.plastic.rankine.m, F, d.plastic_hardening); break;
    case SPK_PLASTIC_SNOW: snow_update(m.plastic.snow.m, F, d.elastic_hardening, vol.plastic_deformation_gradient_det); break;
    default: return;
    }
    store_F(vol, F);
}

// ------------------------------------------------------------------------------------------------ kernels
// QuadraticKernel::eval_all, src_core/dynamics/solver/kernel.rs:15-21
static inline void eval_all(Real x, Real w[3]) {
    w[0] = 0.5f * ((1.5f - x) * (1.5f - x));
    w[1] = 0.75f - ((x - 1.0f) * (x - 1.0f));
    w[2] = 0.5f * ((x - 0.5f) * (x - 0.5f));
}
// ParticlePosition::dir_to_associated_grid_node, particle.rs:37-47
static inline V3 dir_to_associated_grid_node(const float p[3], Real h) {
    V3 r;
    for (int a = 0; a < 3; ++a) r[a] = (std::round(p[a] / h) - 1.0f) * h - p[a];
    return r;
}

struct TileCell { // GridGatherData, g2p2g.rs:22-40 (fields the default updater consumes)
    V3 velocity; Real psi_velocity;
    AccReal mass, momentum[3], psi_momentum, psi_mass; // accumulators
};

// DefaultParticleUpdater::update_particle_and_compute_kirchhoff_stress, src_kernels/cuda/particle_updater.rs:73-236
// returns false when the reference returns None (no P2G)
static bool update_particle(const orc_ctx& ctx, Real dt, Real cell_width, uint32_t pid, SpkParticleStatus& st, SpkParticlePosition& pos,
                            SpkParticleVelocity& vel, SpkParticleVolume& vol, SpkParticlePhase& ph, SpkParticleData& data,
                            V3 interp_vel, M3& vel_grad, Real vel_grad_det, M3& stress_out) {
    (void)pid;
    const SpkGpuParticleModel& model = ctx.models[st.model_index];
    V3 v = interp_vel;
    const bool is_fluid = model_is_fluid(model);
    if (st.kinematic_vel_enabled) v = v3(st.kinematic_vel[0], st.kinematic_vel[1], st.kinematic_vel[2]);
    if (std::fabs(v.x) * dt >= cell_width || std::fabs(v.y) * dt >= cell_width || std::fabs(v.z) * dt >= cell_width) {
        for (int a = 0; a < 3; ++a) v[a] = signum(v[a]) * cell_width / dt;
    }
    for (int a = 0; a < 3; ++a) pos.point[a] += v[a] * dt;
    if (!is_fluid) {
        M3 F = load_F(vol);
        F = F + (vel_grad * dt) * F;
        store_F(vol, F);
    } else {
        vol.deformation_gradient[0] += (vel_grad_det * dt) * vol.deformation_gradient[0];
    }
    if (model.plastic_tag != SPK_PLASTIC_NONE) model_plastic_update(model, data, vol, ph.phase);
    if (st.is_static) { v = v3(0, 0, 0); vel_grad = m3_zero(); }
    for (int a = 0; a < 3; ++a) vel.vector[a] = v[a];
    Real density_def_grad = (vol.mass / vol.volume0) / determinant(load_F(vol));
    if (density_def_grad == 0.0f || st.failed || (!is_fluid && std::fabs(vol.deformation_gradient[0]) > 1.0e4f)) {
        st.failed = 1;
        store_F(vol, m3_identity());
        return false;
    }
    Real energy = model_pos_energy(model, data, vol, ph.phase);
    ph.psi_pos = rmax(ph.psi_pos, energy);
    if (model.failure_tag == SPK_FAILURE_MAXIMUM_STRESS) {
        M3 s = model_kirchhoff_stress(model, data, vol, ph.phase, vel_grad);
        if (maximum_stress_failed(model.failure.maximum_stress, s)) ph.phase = 0.0f;
    }
    stress_out = model_kirchhoff_stress(model, data, vol, ph.phase, vel_grad);
    return true;
}

// particle_g2p2g, src_kernels/cuda/g2p2g.rs:179-351
static void particle_g2p2g(const orc_ctx& ctx, Real dt, uint32_t pid, SpkParticleStatus& st, SpkParticlePosition& pos, SpkParticleVelocity& vel,
                           SpkParticleVolume& vol, SpkParticlePhase& ph, SpkParticleData& data, TileCell* tile) {
    const Real h = ctx.cell_width;
    const Real inv_d = 4.0f / (h * h);
    V3 ref = dir_to_associated_grid_node(pos.point, h);
    Real w[3][3];
    for (int a = 0; a < 3; ++a) eval_all(-ref[a] / h, w[a]);
    Real cell_before[3];
    int cib[3];
    for (int a = 0; a < 3; ++a) {
        Real e = pos.point[a] / h;
        cell_before[a] = std::round(e);
        int assoc_cell = sat_i32(std::round(e)) - 2;                       // particle.rs:28-35
        int assoc_block = sat_i32(std::floor((Real)assoc_cell / 4.0f)) * 4;
        cib[a] = assoc_cell - assoc_block;
    }
    int packed = (cib[0] + 1) + (cib[1] + 1) * 8 + (cib[2] + 1) * 64;
    V3 iv = v3(0, 0, 0);
    M3 grad = m3_zero();
    Real psi_pos_momentum = 0.0f, grad_det = 0.0f;
    for (int i = 0; i < 27; ++i) {
        const int* s = NBH[i];
        V3 dpt = ref + v3((Real)s[0], (Real)s[1], (Real)s[2]) * h;
        Real weight = w[0][s[0]] * w[1][s[1]] * w[2][s[2]];
        const TileCell& cell = tile[packed + s[0] + 8 * s[1] + 64 * s[2]];
        iv = iv + weight * cell.velocity;
        grad = grad + outer((weight * inv_d) * cell.velocity, dpt);
        psi_pos_momentum += weight * cell.psi_velocity;
        grad_det += weight * dot(cell.velocity, dpt) * inv_d;
    }
    (void)psi_pos_momentum; // interpolated but unused by the default updater
    M3 stress;
    if (!update_particle(ctx, dt, h, pid, st, pos, vel, vol, ph, data, iv, grad, grad_det, stress)) return;
    V3 ref2 = dir_to_associated_grid_node(pos.point, h);
    Real w2[3][3];
    for (int a = 0; a < 3; ++a) eval_all(-ref2[a] / h, w2[a]);
    M3 affine = vol.mass * grad - (vol.volume0 * inv_d * dt) * stress;
    V3 pv = v3(vel.vector[0], vel.vector[1], vel.vector[2]);
    V3 momentum = vol.mass * pv + v3(0, 0, 0) * dt;
    Real psi_mass = (ph.phase > 0.0f && !st.failed) ? vol.mass : 0.0f;
    Real psi_pos_mom = psi_mass * ph.psi_pos;
    int64_t mv[3];
    for (int a = 0; a < 3; ++a) mv[a] = sat_i64(std::round(pos.point[a] / h) - cell_before[a]);
    int64_t packed2 = (int64_t)packed + mv[0] + mv[1] * 8 + mv[2] * 64;
    for (int i = 0; i < 27; ++i) {
        const int* s = NBH[i];
        V3 dpt = ref2 + v3((Real)s[0], (Real)s[1], (Real)s[2]) * h;
        Real weight = w2[0][s[0]] * w2[1][s[1]] * w2[2][s[2]];
        int64_t ci = packed2 + s[0] + 8 * s[1] + 64 * s[2];
        if (ci < 0 || ci >= 512) continue; // cannot happen after the velocity clamp (|dx| < h); guards NaN inputs
        TileCell& cell = tile[ci];
        V3 am = weight * (affine * dpt + momentum);
        cell.mass += (AccReal)(weight * vol.mass);
        cell.momentum[0] += (AccReal)am.x; cell.momentum[1] += (AccReal)am.y; cell.momentum[2] += (AccReal)am.z;
        cell.psi_momentum += (AccReal)(weight * psi_pos_mom);
        cell.psi_mass += (AccReal)(weight * psi_mass);
    }
}

// parry3d 0.13 Aabb::project_local_point (aabb_point.rs), called through Cuboid::project_point with solid = false
// (src_kernels/gpu_collider.rs:53; grid_update.rs:59-64,114-119).  Returns the world-space projection.
struct Projection { V3 point; bool is_inside; };
static inline V3 quat_rotate(const float q[4], V3 v) { // nalgebra UnitQuaternion * Vector: v + w t + q x t, t = 2 q x v
    V3 u = v3(q[0], q[1], q[2]);
    V3 t = v3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x) * 2.0f;
    V3 c = v3(u.y * t.z - u.z * t.y, u.z * t.x - u.x * t.z, u.x * t.y - u.y * t.x);
    return v + t * q[3] + c;
}
static Projection cuboid_project_point(const SpkGpuCollider& c, V3 pt) {
    float qinv[4] = {-c.rotation[0], -c.rotation[1], -c.rotation[2], c.rotation[3]};
    V3 tr = v3(c.translation[0], c.translation[1], c.translation[2]);
    V3 lp = quat_rotate(qinv, pt - tr);
    V3 he = v3(c.shape.cuboid.half_extents[0], c.shape.cuboid.half_extents[1], c.shape.cuboid.half_extents[2]);
    V3 mins_pt = (-he) - lp, pt_maxs = lp - he;
    V3 shift;
    for (int a = 0; a < 3; ++a) shift[a] = rmax(mins_pt[a], 0.0f) - rmax(pt_maxs[a], 0.0f);
    Projection out;
    if (!is_zero(shift)) {
        out.is_inside = false;
        out.point = lp + shift;
    } else {
        Real best = -FLT_MAX;
        bool is_mins = false;
        int best_id = 0;
        for (int a = 0; a < 3; ++a) {
            if (mins_pt[a] < pt_maxs[a]) {
                if (pt_maxs[a] > best) { best_id = a; is_mins = false; best = pt_maxs[a]; }
            } else if (mins_pt[a] > best) { best_id = a; is_mins = true; best = mins_pt[a]; }
        }
        V3 sh = v3(0, 0, 0);
        sh[best_id] = is_mins ? best : -best;
        out.is_inside = true;
        out.point = lp + sh;
    }
    out.point = quat_rotate(c.rotation, out.point) + tr;
    return out;
}
// Triangle::project_local_point (parry3d 0.13 point_triangle.rs): closest point by Voronoi regions (Ericson 5.1.5)
static V3 triangle_closest_point(V3 p, V3 a, V3 b, V3 c) {
    V3 ab = b - a, ac = c - a, ap = p - a;
    Real d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0 && d2 <= 0) return a;
    V3 bp = p - b;
    Real d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0 && d4 <= d3) return b;
    Real vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) return a + ab * (d1 / (d1 - d3));
    V3 cp = p - c;
    Real d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0 && d5 <= d6) return c;
    Real vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) return a + ac * (d2 / (d2 - d6));
    Real va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) return b + (c - b) * ((d4 - d3) / ((d4 - d3) + (d5 - d6)));
    Real denom = 1.0f / (va + vb + vc);
    return a + ab * (vb * denom) + ac * (vc * denom);
}
// GenericHeightField::project_local_point_with_max_dist over map_elements_in_local_aabb (parry3d 0.13 point_heightfield.rs,
// heightfield3.rs; the crate is not under /root/reference: restated from its published algorithm), then the inside test
// of GpuColliderShape::project_point_with_max_dist (src_kernels/gpu_collider.rs:73-99)
static bool heightfield_project_point(const SpkGpuCollider& c, V3 pt, Real max_dist, Projection& out) {
    const auto& hf = c.shape.heightfield;
    float qinv[4] = {-c.rotation[0], -c.rotation[1], -c.rotation[2], c.rotation[3]};
    V3 tr = v3(c.translation[0], c.translation[1], c.translation[2]);
    V3 lp = quat_rotate(qinv, pt - tr);
    const int cells_x = (int)hf.ncols - 1, cells_z = (int)hf.nrows - 1;
    const Real cw = 1.0f / (Real)cells_x, ch = 1.0f / (Real)cells_z;
    V3 rmn = v3((lp.x - max_dist) / hf.scale[0], (lp.y - max_dist) / hf.scale[1], (lp.z - max_dist) / hf.scale[2]);
    V3 rmx = v3((lp.x + max_dist) / hf.scale[0], (lp.y + max_dist) / hf.scale[1], (lp.z + max_dist) / hf.scale[2]);
    if (rmx.x <= -0.5f || rmx.z <= -0.5f || rmn.x >= 0.5f || rmn.z >= 0.5f) return false;
    auto clampf = [](Real v, Real lo, Real hi) { return v < lo ? lo : (v > hi ? hi : v); };
    const int min_x = (int)clampf(std::floor((rmn.x + 0.5f) / cw), 0.0f, (Real)(cells_x - 1));
    const int min_z = (int)clampf(std::floor((rmn.z + 0.5f) / ch), 0.0f, (Real)(cells_z - 1));
    const int max_x = (int)clampf(std::ceil((rmx.x + 0.5f) / cw), 0.0f, (Real)cells_x);
    const int max_z = (int)clampf(std::ceil((rmx.z + 0.5f) / ch), 0.0f, (Real)cells_z);
    Real sq_smallest = FLT_MAX;
    bool found = false;
    V3 best = v3(0, 0, 0);
    for (int j = min_x; j < max_x; ++j)
        for (int i = min_z; i < max_z; ++i) {
            Real z0 = -0.5f + ch * (Real)i, z1 = -0.5f + ch * (Real)(i + 1), x0 = -0.5f + cw * (Real)j, x1 = -0.5f + cw * (Real)(j + 1);
            Real y00 = hf.heights[i + (size_t)j * hf.nrows], y10 = hf.heights[i + 1 + (size_t)j * hf.nrows];
            Real y01 = hf.heights[i + (size_t)(j + 1) * hf.nrows], y11 = hf.heights[i + 1 + (size_t)(j + 1) * hf.nrows];
            if ((y00 > rmx.y && y10 > rmx.y && y01 > rmx.y && y11 > rmx.y) || (y00 < rmn.y && y10 < rmn.y && y01 < rmn.y && y11 < rmn.y)) continue;
            V3 p00 = v3(x0 * hf.scale[0], y00 * hf.scale[1], z0 * hf.scale[2]), p10 = v3(x0 * hf.scale[0], y10 * hf.scale[1], z1 * hf.scale[2]);
            V3 p01 = v3(x1 * hf.scale[0], y01 * hf.scale[1], z0 * hf.scale[2]), p11 = v3(x1 * hf.scale[0], y11 * hf.scale[1], z1 * hf.scale[2]);
            for (int t = 0; t < 2; ++t) {
                V3 q = t == 0 ? triangle_closest_point(lp, p00, p10, p01) : triangle_closest_point(lp, p10, p11, p01);
                Real sq = norm_squared(lp - q);
                if (sq < sq_smallest) {
                    sq_smallest = sq;
                    if (std::sqrt(sq) <= max_dist) { best = q; found = true; }
                }
            }
        }
    if (!found) return false;
    if (hf.bellow_heightfield_is_solid) {
        bool in_footprint = lp.x >= -0.5f * hf.scale[0] && lp.x <= 0.5f * hf.scale[0] && lp.z >= -0.5f * hf.scale[2] && lp.z <= 0.5f * hf.scale[2];
        out.is_inside = in_footprint && (hf.flip_interior ? lp.y >= best.y : lp.y <= best.y);
    } else {
        out.is_inside = best.x == lp.x && best.y == lp.y && best.z == lp.z;
    }
    out.point = quat_rotate(c.rotation, best) + tr;
    return true;
}
// polyline_project_point, src_kernels/gpu_collider.rs:105-201 (dim3), with parry3d's Segment::project_local_point_and_get_location
// and Segment::planar_normal(2) restated
static bool polyline_project_point(const SpkGpuCollider& c, V3 point, Projection& out) {
    const auto& pl = c.shape.polyline;
    const int n = (int)pl.num_vertices;
    if (n == 0) return false;
    float qinv[4] = {-c.rotation[0], -c.rotation[1], -c.rotation[2], c.rotation[3]};
    V3 tr = v3(c.translation[0], c.translation[1], c.translation[2]);
    V3 local_point = quat_rotate(qinv, point - tr);
    auto vertex = [&](int i) { return v3(pl.vertices[3 * i], pl.vertices[3 * i + 1], pl.vertices[3 * i + 2]); };
    struct Best { V3 proj; int loc; Real dist; int id; V3 a, b; bool is_inside; } best{};
    bool have = false;
    for (int i = 0; i < n; ++i) { // 1) the closest segment (:118-138)
        V3 a = vertex(i), b = vertex((i + 1) % n);
        V3 ab = b - a, ap = local_point - a;
        Real ab_ap = dot(ab, ap), sqnab = dot(ab, ab);
        V3 q; int loc;
        if (ab_ap <= 0) { q = a; loc = 0; }
        else if (ab_ap >= sqnab) { q = b; loc = 1; }
        else { q = a + ab * (ab_ap / sqnab); loc = 2; }
        Real dist = norm(local_point - q);
        if (!have || dist < best.dist) { best = Best{q, loc, dist, i, a, b, q.x == local_point.x && q.y == local_point.y && q.z == local_point.z}; have = true; }
    }
    // 2) inside / outside (:140-186)
    V3 dir = best.b - best.a;
    Real nn = std::sqrt(dir.y * dir.y + dir.x * dir.x);
    if (nn > FLT_EPSILON) {
        V3 normal1 = v3(dir.y / nn, -dir.x / nn, 0);
        Real side = dot(point - best.proj, normal1); // WORLD point minus LOCAL projection, as in the reference
        if (best.loc == 2) best.is_inside = side <= 0;
        else {
            V3 dir2;
            if (best.loc == 0) { int adj = best.id == 0 ? n - 1 : best.id - 1; dir2 = -(vertex((adj + 1) % n) - vertex(adj)); }
            else { int adj = (best.id + 1) % n; dir2 = vertex((adj + 1) % n) - vertex(adj); }
            Real d = dot(normal1, dir2), threshold = 1.0e-3f * norm(dir2);
            best.is_inside = std::fabs(d) > threshold ? d >= 0 : side <= 0;
        }
    }
    best.proj.z = local_point.z; // :190-193
    if (pl.flip_interior) best.is_inside = !best.is_inside;
    out.is_inside = best.is_inside;
    out.point = quat_rotate(c.rotation, best.proj) + tr;
    return true;
}
// GpuColliderShape::TriMesh, src_kernels/gpu_collider.rs:58-68 with parry3d 0.13's TriMesh::project_point_with_max_dist(.., Real::MAX)
// restated (the crate is not under /root/reference): the closest point over all triangles — the QBVH only prunes — and, for a mesh
// without pseudo-normals (rapier's default trimesh flags), is_inside = the closest triangle's answer for solid = false, i.e. the
// query lies on it.  PARITY: unpinned against the reference's kernels (CudaTriMeshPtr layout not recovered), checked against an
// independent f64 brute force (tests/test_trimesh_cpu.py).
static bool trimesh_project_point(const SpkGpuCollider& c, V3 pt, Projection& out) {
    const auto& tm = c.shape.trimesh;
    if (tm.num_triangles == 0) return false;
    float qinv[4] = {-c.rotation[0], -c.rotation[1], -c.rotation[2], c.rotation[3]};
    V3 tr = v3(c.translation[0], c.translation[1], c.translation[2]);
    V3 lp = quat_rotate(qinv, pt - tr);
    auto vertex = [&](uint32_t i) { return v3(tm.vertices[3 * i], tm.vertices[3 * i + 1], tm.vertices[3 * i + 2]); };
    Real best_sq = FLT_MAX;
    V3 best = v3(0, 0, 0);
    for (uint32_t t = 0; t < tm.num_triangles; ++t) {
        V3 q = triangle_closest_point(lp, vertex(tm.indices[3 * t]), vertex(tm.indices[3 * t + 1]), vertex(tm.indices[3 * t + 2]));
        Real sq = norm_squared(lp - q);
        if (sq < best_sq) { best_sq = sq; best = q; }
    }
    out.is_inside = best.x == lp.x && best.y == lp.y && best.z == lp.z;
    if (tm.flip_interior) out.is_inside = !out.is_inside;
    out.point = quat_rotate(c.rotation, best) + tr;
    return true;
}
static bool collider_project(const SpkGpuCollider& c, V3 pt, Real max_dist, Projection& out) {
    if (c.shape_tag == SPK_SHAPE_TRIMESH) return trimesh_project_point(c, pt, out);
    if (c.shape_tag == SPK_SHAPE_HEIGHTFIELD) return heightfield_project_point(c, pt, max_dist, out);
    if (c.shape_tag == SPK_SHAPE_POLYLINE) return polyline_project_point(c, pt, out);
    if (c.shape_tag != SPK_SHAPE_CUBOID) return false;
    out = cuboid_project_point(c, pt);
    return true;
}
// sdf, grid_update.rs:54-71
static bool sdf(const orc_ctx& ctx, V3 q, Real max_dist, Real& out) {
    bool any = false;
    Real best = 0;
    for (const auto& c : ctx.colliders) {
        if (c.grid_boundary_handling == SPK_BOUNDARY_NONE) continue;
        Projection p;
        if (!collider_project(c, q, max_dist, p)) continue;
        Real d = norm(p.point - q) * (p.is_inside ? -1.0f : 1.0f);
        if (!any || d < best) { best = d; any = true; } // min_by keeps the first of equal elements
    }
    out = best;
    return any;
}
// sdf_gradient, grid_update.rs:74-94
static V3 sdf_gradient(const orc_ctx& ctx, Real h, V3 q) {
    const Real FACTOR = 0.1f;
    V3 g;
    for (int a = 0; a < 3; ++a) {
        V3 off = v3(0, 0, 0);
        off[a] = 1.0f * (h * FACTOR);
        Real sp, sn;
        bool hp = sdf(ctx, q + off, h * 2.0f, sp), hn = sdf(ctx, q + (-off), h * 2.0f, sn);
        g[a] = (hp && hn) ? (sp - sn) / h / FACTOR / 2.0f : 0.0f;
    }
    Real n = norm(g);
    if (n <= 1.0e-5f) return v3(0, 0, 0); // try_normalize(1e-5)
    return g / n;
}
// update_single_cell, grid_update.rs:96-198
static void update_single_cell(const orc_ctx& ctx, Real dt, SpkGpuGridNode& cell, V3 cell_pos, Real h, V3 gravity) {
    V3 mom = v3(cell.momentum_velocity[0], cell.momentum_velocity[1], cell.momentum_velocity[2]);
    V3 v = (mom + cell.mass * gravity * dt) * inv_exact(cell.mass);
    if (cell.projection_status_tag == SPK_PROJ_NOT_COMPUTED) {
        bool have = false;
        V3 best_dir = v3(0, 0, 0);
        Real best_dist = FLT_MAX;
        bool best_inside = false;
        size_t best_id = 0;
        for (size_t i = 0; i < ctx.colliders.size(); ++i) {
            const auto& c = ctx.colliders[i];
            if (c.grid_boundary_handling == SPK_BOUNDARY_NONE) continue;
            Projection p;
            if (!collider_project(c, cell_pos, h * 2.0f, p)) continue; // max_dist = cell_width * 2 (grid_update.rs:118)
            V3 dir = cell_pos - p.point;
            Real dist = norm(dir);
            if (dist < (have ? best_dist : FLT_MAX)) { have = true; best_dir = dir; best_dist = dist; best_inside = p.is_inside; best_id = i; }
        }
        if (have) {
            cell.projection_status_tag = best_inside ? SPK_PROJ_INSIDE : SPK_PROJ_OUTSIDE;
            cell.projection_status_payload = best_id;
            for (int a = 0; a < 3; ++a) cell.projection_scaled_dir[a] = best_dir[a];
        } else {
            cell.projection_status_tag = SPK_PROJ_TOO_FAR;
        }
        V3 n = sdf_gradient(ctx, h, cell_pos);
        for (int a = 0; a < 3; ++a) cell.collision_normal[a] = n[a];
    }
    if (cell.projection_status_tag == SPK_PROJ_INSIDE || cell.projection_status_tag == SPK_PROJ_OUTSIDE) {
        const bool is_inside = cell.projection_status_tag == SPK_PROJ_INSIDE;
        const SpkGpuCollider& col = ctx.colliders[cell.projection_status_payload];
        switch (col.grid_boundary_handling) {
        case SPK_BOUNDARY_STICK:
            if (is_inside) v = v3(0, 0, 0);
            break;
        case SPK_BOUNDARY_FRICTION:
        case SPK_BOUNDARY_FRICTION_Z_UP: {
            V3 normal = v3(cell.collision_normal[0], cell.collision_normal[1], cell.collision_normal[2]);
            if (!is_zero(normal)) {
                Real dist = norm(v3(cell.projection_scaled_dir[0], cell.projection_scaled_dir[1], cell.projection_scaled_dir[2]));
                bool apply = col.grid_boundary_handling == SPK_BOUNDARY_FRICTION || (col.grid_boundary_handling == SPK_BOUNDARY_FRICTION_Z_UP && normal.z >= 0.0f);
                if (apply) {
                    Real normal_vel = dot(v, normal);
                    if (normal_vel < 0.0f) {
                        Real dist_with_margin = dist - h;
                        if (is_inside || dist_with_margin <= 0.0f) {
                            V3 tangent = v - normal_vel * normal;
                            Real tn = norm(tangent);
                            v = tangent;
                            if (tn > 1.0e-10f) v = tangent / tn * rmax(tn + normal_vel * col.friction, 0.0f);
                        } else if (-normal_vel * dt > dist_with_margin) {
                            v = v - (dist_with_margin / dt + normal_vel) * normal;
                        }
                    }
                }
            }
            break;
        }
        default: break;
        }
    }
    for (int a = 0; a < 3; ++a) cell.momentum_velocity[a] = v[a];
    cell.psi_momentum_velocity *= inv_exact(cell.psi_mass);
}

// `red.global.add.f32` (src_kernels/cuda/atomic.rs, used by g2p2g.rs:542-550) flushes subnormal inputs and results
// to sign-preserving zero (PTX ISA, atom/red .f32 add) — this keeps 1/mass finite in grid_update.
static inline Real ftz(Real x) { return std::fabs(x) < FLT_MIN ? std::copysign(0.0f, x) : x; }
static inline Real red_add(Real dst, Real src) { return ftz(ftz(dst) + ftz(src)); }

static const Real DT_FACTOR = 1000000000000.0f; // GpuTimestepLength::FACTOR, src_kernels/gpu_timestep.rs:18
static inline uint64_t dt_from_sec(Real dt) { return sat_u64(dt * DT_FACTOR); }
static inline Real dt_into_sec(uint64_t v) { return (Real)v / DT_FACTOR; }

} // namespace

// ================================================================================================ C API
extern "C" {

orc_ctx* orc_create(float cell_width) {
    orc_ctx* c = new orc_ctx();
    c->cell_width = cell_width;
    return c;
}
void orc_destroy(orc_ctx* c) { delete c; }
int orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

int orc_particles_upload(orc_ctx* c, const SpkParticleStatus* status, const SpkParticlePosition* pos, const SpkParticleVelocity* vel,
                         const SpkParticleVolume* volume, const SpkParticlePhase* phase, const SpkParticleData* data, size_t n) {
    if (!c || (n && (!status || !pos || !vel || !volume || !phase))) return SPK_EINVAL;
    c->status.assign(status, status + n);
    c->pos.assign(pos, pos + n);
    c->vel.assign(vel, vel + n);
    c->volume.assign(volume, volume + n);
    c->phase.assign(phase, phase + n);
    if (data) c->data.assign(data, data + n);
    else c->data.assign(n, SpkParticleData{-0.01f, 1.0f, 1.0f, 0.0f}); // ParticleData::default, particle.rs:211-219
    c->sorted_ids.assign(n, 0);
    return SPK_OK;
}
int orc_particles_len(const orc_ctx* c, size_t* out) { *out = c->status.size(); return SPK_OK; }
int orc_particles_read(orc_ctx* c, int which, void* dst, size_t n) {
    if (!c || !dst || n > c->status.size()) return SPK_EINVAL;
    switch (which) {
    case SPK_ATTR_POSITION: std::memcpy(dst, c->pos.data(), n * sizeof(SpkParticlePosition)); break;
    case SPK_ATTR_VELOCITY: std::memcpy(dst, c->vel.data(), n * sizeof(SpkParticleVelocity)); break;
    case SPK_ATTR_STATUS: std::memcpy(dst, c->status.data(), n * sizeof(SpkParticleStatus)); break;
    case SPK_ATTR_VOLUME: std::memcpy(dst, c->volume.data(), n * sizeof(SpkParticleVolume)); break;
    case SPK_ATTR_PHASE: std::memcpy(dst, c->phase.data(), n * sizeof(SpkParticlePhase)); break;
    case SPK_ATTR_DATA: std::memcpy(dst, c->data.data(), n * sizeof(SpkParticleData)); break;
    case SPK_ATTR_SORTED_IDS: std::memcpy(dst, c->sorted_ids.data(), n * sizeof(uint32_t)); break;
    default: return SPK_EINVAL;
    }
    return SPK_OK;
}
int orc_models_upload(orc_ctx* c, const SpkGpuParticleModel* models, size_t n) { c->models.assign(models, models + n); return SPK_OK; }
int orc_colliders_upload(orc_ctx* c, const SpkGpuCollider* cols, size_t n) {
    for (size_t i = 0; i < n; ++i)
        if (cols[i].shape_tag > SPK_SHAPE_POLYLINE) return SPK_EUNSUPPORTED;
    c->colliders.assign(cols, cols + n);
    c->collider_data.assign(n, {});
    c->collider_indices.assign(n, {});
    for (size_t i = 0; i < n; ++i)
        if (cols[i].shape_tag == SPK_SHAPE_HEIGHTFIELD) { // keep an owned copy of the heights matrix
            const auto& hf = cols[i].shape.heightfield;
            if (!hf.heights || hf.nrows < 2 || hf.ncols < 2) return SPK_EINVAL;
            c->collider_data[i].assign(hf.heights, hf.heights + (size_t)hf.nrows * hf.ncols);
            c->colliders[i].shape.heightfield.heights = c->collider_data[i].data();
        } else if (cols[i].shape_tag == SPK_SHAPE_POLYLINE) {
            const auto& pl = cols[i].shape.polyline;
            if (pl.num_vertices && !pl.vertices) return SPK_EINVAL;
            c->collider_data[i].assign(pl.vertices, pl.vertices + (size_t)pl.num_vertices * 3);
            c->colliders[i].shape.polyline.vertices = c->collider_data[i].data();
        } else if (cols[i].shape_tag == SPK_SHAPE_TRIMESH) {
            const auto& tm = cols[i].shape.trimesh;
            if (!tm.vertices || !tm.indices || tm.num_vertices == 0 || tm.num_triangles == 0) return SPK_EINVAL;
            for (size_t k = 0; k < (size_t)tm.num_triangles * 3; ++k) if (tm.indices[k] >= tm.num_vertices) return SPK_EINVAL;
            c->collider_data[i].assign(tm.vertices, tm.vertices + (size_t)tm.num_vertices * 3);
            c->collider_indices[i].assign(tm.indices, tm.indices + (size_t)tm.num_triangles * 3);
            c->colliders[i].shape.trimesh.vertices = c->collider_data[i].data();
            c->colliders[i].shape.trimesh.indices = c->collider_indices[i].data();
        }
    return SPK_OK;
}

int orc_frame_begin(orc_ctx* c, const SpkSolverParameters* p) { c->dt_acc = dt_from_sec(p->dt); return SPK_OK; } // cuda_mpm_pipeline.rs:294-299

// CudaSparseGrid::launch_sort, src/cuda/cuda_sparse_grid.rs:147-489 with the kernels of src_kernels/cuda/sort.rs:58-280
int orc_stage_sort(orc_ctx* c, SpkSortInfo* info) {
    const size_t n = c->pos.size();
    Grid& g = c->next_grid();
    g.clear();
    std::vector<uint64_t> pkey(n);
    std::vector<uint64_t> mains;
    {
        uint64_t last = UINT64_MAX;
        for (size_t i = 0; i < n; ++i) {
            uint64_t b[3];
            main_block_of_point(c->pos[i].point, c->cell_width, b);
            pkey[i] = pack_key(b[0], b[1], b[2]);
            if (pkey[i] != last) { mains.push_back(pkey[i]); last = pkey[i]; }
        }
    }
    std::sort(mains.begin(), mains.end());
    mains.erase(std::unique(mains.begin(), mains.end()), mains.end());
    // touch_particle_blocks (sort.rs:58-72): the 8 blocks b + {0,1}^3 of every particle (gpu_grid.rs:307-321)
    std::vector<uint64_t> keys;
    keys.reserve(mains.size() * 8);
    for (uint64_t k : mains) {
        uint64_t b[3];
        unpack_key(k, b);
        for (int o = 0; o < 8; ++o) keys.push_back(pack_key(b[0] + ((o >> 2) & 1), b[1] + ((o >> 1) & 1), b[2] + (o & 1)));
    }
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    const size_t nab = keys.size();
    g.headers.resize(nab);
    g.map.reserve(nab * 2);
    for (size_t i = 0; i < nab; ++i) {
        g.headers[i] = SpkActiveBlockHeader{keys[i], 0, 0, 0, 0};
        g.map.emplace(keys[i], (uint32_t)i);
    }
    // update_block_particle_count (sort.rs:166-181)
    std::vector<uint32_t> pblock(n);
    for (size_t i = 0; i < n; ++i) {
        pblock[i] = g.map.find(pkey[i])->second;
        g.headers[pblock[i]].num_particles += 1;
    }
    // exclusive scan -> first_particle (sort.rs:183-199)
    uint32_t run = 0;
    for (size_t i = 0; i < nab; ++i) { g.headers[i].first_particle = run; run += g.headers[i].num_particles; }
    // finalize_particles_sort (sort.rs:201-221); ascending particle id inside a block (canonical form)
    std::vector<uint32_t> cursor(nab);
    for (size_t i = 0; i < nab; ++i) cursor[i] = g.headers[i].first_particle;
    c->sorted_ids.resize(n);
    for (size_t i = 0; i < n; ++i) c->sorted_ids[cursor[pblock[i]]++] = (uint32_t)i;
    // dispatch table (sort.rs:226-280), 128 = G2P2G_THREADS (src_kernels/cuda/mod.rs:15)
    c->dispatch.clear();
    for (size_t i = 0; i < nab; ++i) {
        uint32_t np = g.headers[i].num_particles;
        uint32_t mult = np / 128 + (np % 128 > 0);
        for (uint32_t k = 0; k < mult; ++k) c->dispatch.push_back(SpkDispatchBlock2ActiveBlock{(uint32_t)i, g.headers[i].first_particle + k * 128});
    }
    g.nodes.assign(nab * 64, default_node());
    if (info) {
        info->num_active_blocks = (uint32_t)nab;
        info->num_dispatch_blocks = (uint32_t)c->dispatch.size();
        info->num_dispatch_halo_blocks = 0;
        info->hashmap_capacity = 0;
    }
    return SPK_OK;
}

// reset_grid + copy_grid_projection_data, src_kernels/cuda/reset_grid.rs:8-60
int orc_stage_reset_grid(orc_ctx* c) {
    Grid& next = c->next_grid();
    const Grid& curr = c->curr_grid();
    for (auto& nd : next.nodes) nd = default_node();
    for (size_t b = 0; b < next.headers.size(); ++b) {
        int pb = curr.find(next.headers[b].virtual_id);
        if (pb < 0) continue;
        for (int k = 0; k < 64; ++k) {
            const SpkGpuGridNode& p = curr.nodes[(size_t)pb * 64 + k];
            SpkGpuGridNode& q = next.nodes[b * 64 + k];
            q.prev_mass = p.mass;
            q.projection_status_tag = p.projection_status_tag;
            q.projection_status_payload = p.projection_status_payload;
            std::memcpy(q.collision_normal, p.collision_normal, 12);
            std::memcpy(q.projection_scaled_dir, p.projection_scaled_dir, 12);
        }
    }
    return SPK_OK;
}

// estimate_timestep_length, src_kernels/cuda/timestep.rs:13-113
int orc_stage_estimate_dt(orc_ctx* c, float min_dt, float max_dt, float* dt_out) {
    const size_t n = c->pos.size();
    uint64_t acc = c->dt_acc;
#pragma omp parallel for reduction(min : acc) schedule(static)
    for (long long i = 0; i < (long long)n; ++i) {
        const SpkParticleStatus& st = c->status[i];
        if (st.failed) continue;
        Real dt = max_dt;
        const SpkGpuParticleModel& model = c->models[st.model_index];
        unsigned bounds = model_active_bounds(model);
        V3 v = v3(c->vel[i].vector[0], c->vel[i].vector[1], c->vel[i].vector[2]);
        Real cand = (bounds & 1u) ? model_timestep_bound(model, c->data[i], c->volume[i], v, c->cell_width) : FLT_MAX; // particle_updater.rs:57-69
        dt = rmin(dt, cand);
        if (bounds & 2u) dt = rmin(dt, c->cell_width / norm(v));
        if (dt < min_dt && max_dt > min_dt) dt = min_dt;
        uint64_t cnd = dt_from_sec(dt);
        if (cnd < acc) acc = cnd;
    }
    c->dt_acc = acc;
    if (dt_out) *dt_out = dt_into_sec(acc);
    return SPK_OK;
}
int orc_dt_acc_get(orc_ctx* c, uint64_t* out) { *out = c->dt_acc; return SPK_OK; }
int orc_dt_acc_set(orc_ctx* c, uint64_t v) { c->dt_acc = v; return SPK_OK; }

// g2p2g, src_kernels/cuda/g2p2g.rs:64-177 (+ transfer_* :353-553)
int orc_stage_g2p2g(orc_ctx* c, float dt, uint32_t damage_model) {
    (void)damage_model; // unused by the kernel, g2p2g.rs:190
    Grid& next = c->next_grid();
    const Grid& curr = c->curr_grid();
    const size_t nab = next.headers.size();
    std::vector<AccReal> acc(nab * 64 * 6, (AccReal)0);
    // 8 colours by block-coordinate parity: two blocks of one colour never write the same node (a block writes b + {0,1}^3)
    std::vector<std::vector<uint32_t>> colours(8);
    for (size_t b = 0; b < nab; ++b) {
        if (next.headers[b].num_particles == 0) continue;
        uint64_t k[3];
        unpack_key(next.headers[b].virtual_id, k);
        colours[(k[0] & 1) | ((k[1] & 1) << 1) | ((k[2] & 1) << 2)].push_back((uint32_t)b);
    }
    for (int col = 0; col < 8; ++col) {
        const std::vector<uint32_t>& blocks = colours[col];
#pragma omp parallel
        {
            std::vector<TileCell> tile(512);
#pragma omp for schedule(dynamic, 4)
            for (long long bi = 0; bi < (long long)blocks.size(); ++bi) {
                const uint32_t b = blocks[bi];
                const SpkActiveBlockHeader& hdr = next.headers[b];
                uint64_t base[3];
                unpack_key(hdr.virtual_id, base);
                int oct_next[8];
                // transfer_global_blocks_to_shared_memory, g2p2g.rs:353-474
                for (int o = 0; o < 8; ++o) {
                    int ox = o & 1, oy = (o >> 1) & 1, oz = (o >> 2) & 1;
                    uint64_t okey = pack_key(base[0] + ox, base[1] + oy, base[2] + oz);
                    int pb = curr.find(okey);
                    oct_next[o] = next.find(okey);
                    for (int id = 0; id < 64; ++id) {
                        int sx = id & 3, sy = (id >> 2) & 3, sz = id >> 4;
                        TileCell& t = tile[(ox * 4 + sx) + (oy * 4 + sy) * 8 + (oz * 4 + sz) * 64];
                        if (pb >= 0) {
                            const SpkGpuGridNode& gn = curr.nodes[(size_t)pb * 64 + id];
                            t.velocity = v3(gn.momentum_velocity[0], gn.momentum_velocity[1], gn.momentum_velocity[2]);
                            t.psi_velocity = gn.psi_momentum_velocity;
                        } else {
                            t.velocity = v3(0, 0, 0);
                            t.psi_velocity = 0.0f;
                        }
                        t.mass = 0; t.momentum[0] = t.momentum[1] = t.momentum[2] = 0; t.psi_momentum = 0; t.psi_mass = 0;
                    }
                }
                for (uint32_t k = 0; k < hdr.num_particles; ++k) {
                    uint32_t pid = c->sorted_ids[hdr.first_particle + k];
                    particle_g2p2g(*c, dt, pid, c->status[pid], c->pos[pid], c->vel[pid], c->volume[pid], c->phase[pid], c->data[pid], tile.data());
                }
                // transfer_shared_blocks_to_grid, g2p2g.rs:476-553
                for (int o = 0; o < 8; ++o) {
                    int ox = o & 1, oy = (o >> 1) & 1, oz = (o >> 2) & 1;
                    if (oct_next[o] < 0) continue; // the reference unwrap()s: cannot be missing after touch_particle_blocks
                    AccReal* a = &acc[(size_t)oct_next[o] * 64 * 6];
                    for (int id = 0; id < 64; ++id) {
                        int sx = id & 3, sy = (id >> 2) & 3, sz = id >> 4;
                        const TileCell& t = tile[(ox * 4 + sx) + (oy * 4 + sy) * 8 + (oz * 4 + sz) * 64];
                        a[id * 6 + 0] += (AccReal)ftz((Real)t.mass); a[id * 6 + 1] += (AccReal)ftz((Real)t.momentum[0]);
                        a[id * 6 + 2] += (AccReal)ftz((Real)t.momentum[1]); a[id * 6 + 3] += (AccReal)ftz((Real)t.momentum[2]);
                        a[id * 6 + 4] += (AccReal)ftz((Real)t.psi_momentum); a[id * 6 + 5] += (AccReal)ftz((Real)t.psi_mass);
                    }
                }
            }
        }
    }
    for (size_t i = 0; i < nab * 64; ++i) {
        SpkGpuGridNode& nd = next.nodes[i];
        nd.mass = red_add(nd.mass, (Real)acc[i * 6 + 0]);
        nd.momentum_velocity[0] = red_add(nd.momentum_velocity[0], (Real)acc[i * 6 + 1]);
        nd.momentum_velocity[1] = red_add(nd.momentum_velocity[1], (Real)acc[i * 6 + 2]);
        nd.momentum_velocity[2] = red_add(nd.momentum_velocity[2], (Real)acc[i * 6 + 3]);
        nd.psi_momentum_velocity = red_add(nd.psi_momentum_velocity, (Real)acc[i * 6 + 4]);
        nd.psi_mass = red_add(nd.psi_mass, (Real)acc[i * 6 + 5]);
    }
    return SPK_OK;
}

// Fluids-volume recomputation: MlsSolver::recompute_fluids_volumes, src/dynamics/solver/fluids_volume.rs:9-116 (the live CPU pass;
// the GPU kernels src_kernels/cuda/fluids_volume.rs:9-105 are the same arithmetic, commented out of the reference's build).  Runs after
// the sort on the zeroed next grid: P2G of the particle masses, G2P of the node masses into F[0][0] of the fluid particles, node masses
// back to zero.  Node sums in double then rounded once (the order of the reference's additions is unspecified on either path).
int orc_stage_recompute_fluids_volume(orc_ctx* c) {
    Grid& next = c->next_grid();
    const Real h = c->cell_width;
    const size_t n = c->pos.size();
    std::vector<double> acc(next.nodes.size(), 0.0);
    std::vector<long long> node((size_t)n * 27, -1);
    std::vector<Real> wgt((size_t)n * 27, 0.0f);
    for (size_t i = 0; i < n; ++i) {
        auto ref = dir_to_associated_grid_node(c->pos[i].point, h);
        Real w[3][3];
        uint64_t base[3];
        int cib[3];
        main_block_of_point(c->pos[i].point, h, base);
        for (int a = 0; a < 3; ++a) {
            eval_all(-ref[a] / h, w[a]);
            int assoc_cell = sat_i32(std::round(c->pos[i].point[a] / h)) - 2; // particle.rs:28-35
            int assoc_block = sat_i32(std::floor((Real)assoc_cell / 4.0f)) * 4;
            cib[a] = assoc_cell - assoc_block;
        }
        for (int k = 0; k < 27; ++k) {
            const int* s = NBH[k];
            int t[3];
            for (int a = 0; a < 3; ++a) t[a] = cib[a] + 1 + s[a];
            int b = next.find(pack_key(base[0] + (uint64_t)(t[0] >> 2), base[1] + (uint64_t)(t[1] >> 2), base[2] + (uint64_t)(t[2] >> 2)));
            if (b < 0) continue; // cannot happen after touch_particle_blocks
            node[i * 27 + k] = (long long)b * 64 + (t[0] & 3) * 1 + (t[1] & 3) * 4 + (t[2] & 3) * 16;
            wgt[i * 27 + k] = w[0][s[0]] * w[1][s[1]] * w[2][s[2]];
            acc[(size_t)node[i * 27 + k]] += (double)(wgt[i * 27 + k] * c->volume[i].mass); // fluids_volume.rs:40-47
        }
    }
    for (size_t i = 0; i < n; ++i) {
        if (!model_is_fluid(c->models[c->status[i].model_index])) continue; // :70-77
        Real new_mass = 0.0f;
        for (int k = 0; k < 27; ++k)
            if (node[i * 27 + k] >= 0) new_mass += wgt[i * 27 + k] * ftz((Real)acc[(size_t)node[i * 27 + k]]);
        Real new_density = new_mass / (h * h * h); // cell_width.powi(DIM)
        Real new_volume = c->volume[i].mass / new_density;
        c->volume[i].deformation_gradient[0] = new_volume / c->volume[i].volume0;
    }
    return SPK_OK; // the next grid's masses were never written: "Reset the modified cells" (:98-113) is implicit
}

// grid_update + swap_buffers, src_kernels/cuda/grid_update.rs:10-52, cuda_mpm_pipeline.rs:536-558
int orc_stage_grid_update(orc_ctx* c, float dt, const float gravity[3]) {
    Grid& next = c->next_grid();
    const V3 g = v3(gravity[0], gravity[1], gravity[2]);
    const long long nab = (long long)next.headers.size();
#pragma omp parallel for schedule(dynamic, 16)
    for (long long b = 0; b < nab; ++b) {
        uint64_t k[3];
        unpack_key(next.headers[b].virtual_id, k);
        for (int id = 0; id < 64; ++id) {
            int sx = id & 3, sy = (id >> 2) & 3, sz = id >> 4;
            int64_t ci[3] = {((int64_t)k[0] - (int64_t)PACK_ORIGIN) * 4 + sx, ((int64_t)k[1] - (int64_t)PACK_ORIGIN) * 4 + sy,
                             ((int64_t)k[2] - (int64_t)PACK_ORIGIN) * 4 + sz};
            V3 cell_pos = v3((Real)ci[0] * c->cell_width, (Real)ci[1] * c->cell_width, (Real)ci[2] * c->cell_width);
            update_single_cell(*c, dt, next.nodes[(size_t)b * 64 + id], cell_pos, c->cell_width, g);
        }
    }
    c->curr = 1 - c->curr;
    return SPK_OK;
}

// CudaMpmPipeline::step_generic, src/cuda/cuda_mpm_pipeline.rs:273-659 (single context)
int orc_step(orc_ctx* c, const SpkSolverParameters* params, const float gravity[3], SpkStepTimings* timings, float* simulated_time_out) {
    if (!c || !params || !gravity) return SPK_EINVAL;
    orc_frame_begin(c, params);
    if (timings) std::memset(timings, 0, sizeof(*timings));
    if (params->dt == 0.0f || c->pos.empty()) { if (simulated_time_out) *simulated_time_out = params->dt; return SPK_OK; }
    const Real min_dt = params->dt / (Real)params->max_num_substeps;
    Real remaining = params->dt;
    uint32_t niter = 0;
    while (remaining > 0.0f) {
        ++niter;
        SpkSortInfo info;
        orc_stage_sort(c, &info);
        orc_stage_reset_grid(c);
        Real dt;
        orc_stage_estimate_dt(c, min_dt, rmin(remaining, params->max_substep_dt), &dt);
        if (info.num_dispatch_blocks > 0) orc_stage_g2p2g(c, dt, params->damage_model);
        orc_stage_grid_update(c, dt, gravity);
        remaining -= dt;
        if (timings) { timings->num_substeps = niter; timings->last_dt = dt; timings->num_active_blocks = info.num_active_blocks; timings->num_dispatch_blocks = info.num_dispatch_blocks; }
        if (remaining <= 10.0f / DT_FACTOR || params->stop_after_one_substep) break;
    }
    if (simulated_time_out) *simulated_time_out = params->dt - remaining;
    return SPK_OK;
}

int orc_grid_num_active_blocks(orc_ctx* c, int which, uint32_t* out) { *out = (uint32_t)(which ? c->next_grid() : c->curr_grid()).headers.size(); return SPK_OK; }
int orc_grid_read_headers(orc_ctx* c, int which, SpkActiveBlockHeader* dst, uint32_t n) {
    Grid& g = which ? c->next_grid() : c->curr_grid();
    if (n > g.headers.size()) return SPK_EINVAL;
    std::memcpy(dst, g.headers.data(), n * sizeof(SpkActiveBlockHeader));
    return SPK_OK;
}
int orc_grid_read_nodes(orc_ctx* c, int which, SpkGpuGridNode* dst, uint32_t n_blocks) {
    Grid& g = which ? c->next_grid() : c->curr_grid();
    if (n_blocks > g.headers.size()) return SPK_EINVAL;
    std::memcpy(dst, g.nodes.data(), (size_t)n_blocks * 64 * sizeof(SpkGpuGridNode));
    return SPK_OK;
}
int orc_grid_read_dispatch(orc_ctx* c, SpkDispatchBlock2ActiveBlock* dst, uint32_t n) {
    if (n > c->dispatch.size()) return SPK_EINVAL;
    std::memcpy(dst, c->dispatch.data(), n * sizeof(SpkDispatchBlock2ActiveBlock));
    return SPK_OK;
}

// ---- halo exchange (multi-process path), same contract as spk_halo_* with HOST pointers; lets the world_size-2 gloo
// tests drive sparkl_b200.distributed on CPU.  tag_halo_blocks / copy_halo_to_staging / merge_halo_blocks,
// src_kernels/cuda/sort.rs:74-164.
int orc_halo_begin(orc_ctx* c) { c->num_halo = 0; return SPK_OK; }
int orc_halo_active_keys(orc_ctx* c, uint64_t* keys_out, uint32_t capacity) {
    Grid& g = c->next_grid();
    if (g.headers.size() > capacity) return SPK_EOVERFLOW;
    for (uint32_t i = 0; i < capacity; ++i) keys_out[i] = i < g.headers.size() ? g.headers[i].virtual_id : UINT64_MAX;
    return SPK_OK;
}
int orc_halo_tag(orc_ctx* c, const uint64_t* remote_keys, uint32_t n_remote) {
    Grid& g = c->next_grid();
    for (uint32_t i = 0; i < n_remote; ++i) {
        if (remote_keys[i] == UINT64_MAX) continue;
        int b = g.find(remote_keys[i]);
        if (b >= 0 && !(g.headers[b].halo_state & 1u)) { g.headers[b].halo_state |= 1u; ++c->num_halo; }
    }
    return SPK_OK;
}
int orc_halo_count(orc_ctx* c, uint32_t* out) { *out = c->num_halo; return SPK_OK; }
int orc_halo_pack(orc_ctx* c, uint64_t* keys_out, float* payload_out, uint32_t capacity) {
    Grid& g = c->next_grid();
    uint32_t k = 0;
    for (uint32_t i = 0; i < capacity; ++i) keys_out[i] = UINT64_MAX;
    for (size_t b = 0; b < g.headers.size(); ++b) {
        if (!(g.headers[b].halo_state & 1u)) continue;
        if (k >= capacity) return SPK_EOVERFLOW;
        keys_out[k] = g.headers[b].virtual_id;
        float* dst = payload_out + (size_t)k * 384;
        for (int cell = 0; cell < 64; ++cell) {
            const SpkGpuGridNode& nd = g.nodes[b * 64 + cell];
            dst[0 * 64 + cell] = nd.mass; dst[1 * 64 + cell] = nd.momentum_velocity[0]; dst[2 * 64 + cell] = nd.momentum_velocity[1];
            dst[3 * 64 + cell] = nd.momentum_velocity[2]; dst[4 * 64 + cell] = nd.psi_momentum_velocity; dst[5 * 64 + cell] = nd.psi_mass;
        }
        ++k;
    }
    return SPK_OK;
}
int orc_halo_merge(orc_ctx* c, const uint64_t* keys, const float* payload, uint32_t n_blocks) {
    Grid& g = c->next_grid();
    for (uint32_t k = 0; k < n_blocks; ++k) {
        if (keys[k] == UINT64_MAX) continue;
        int b = g.find(keys[k]);
        if (b < 0) continue;
        const float* src = payload + (size_t)k * 384;
        for (int cell = 0; cell < 64; ++cell) {
            SpkGpuGridNode& nd = g.nodes[(size_t)b * 64 + cell];
            nd.mass = red_add(nd.mass, src[0 * 64 + cell]);
            nd.momentum_velocity[0] = red_add(nd.momentum_velocity[0], src[1 * 64 + cell]);
            nd.momentum_velocity[1] = red_add(nd.momentum_velocity[1], src[2 * 64 + cell]);
            nd.momentum_velocity[2] = red_add(nd.momentum_velocity[2], src[3 * 64 + cell]);
            nd.psi_momentum_velocity = red_add(nd.psi_momentum_velocity, src[4 * 64 + cell]);
            nd.psi_mass = red_add(nd.psi_mass, src[5 * 64 + cell]);
        }
    }
    return SPK_OK;
}

// ---- unit hooks used by tests/ (known-answer checks of the building blocks)
void orc_svd3(const float* f_colmajor, float* u, float* s, float* vt) {
    M3 F; std::memcpy(F.m, f_colmajor, 36);
    Svd3 r = svd_unordered(F);
    std::memcpy(u, r.u.m, 36); std::memcpy(vt, r.vt.m, 36); s[0] = r.s.x; s[1] = r.s.y; s[2] = r.s.z;
}
uint64_t orc_block_key_of_point(const float* p, float cell_width) { uint64_t b[3]; main_block_of_point(p, cell_width, b); return pack_key(b[0], b[1], b[2]); }
void orc_quadratic_weights(float x, float* w3) { eval_all(x, w3); }
uint64_t orc_dt_from_sec(float dt) { return dt_from_sec(dt); }
float orc_dt_into_sec(uint64_t v) { return dt_into_sec(v); }
void orc_kirchhoff_stress(const SpkGpuParticleModel* m, const SpkParticleData* d, const SpkParticleVolume* vol, float phase, const float* vel_grad, float* out9) {
    M3 g; std::memcpy(g.m, vel_grad, 36);
    M3 s = model_kirchhoff_stress(*m, *d, *vol, phase, g);
    std::memcpy(out9, s.m, 36);
}
void orc_plastic_update(const SpkGpuParticleModel* m, SpkParticleData* d, SpkParticleVolume* vol, float phase) { model_plastic_update(*m, *d, *vol, phase); }
int orc_collider_project(const SpkGpuCollider* c, const float* pt, float max_dist, float* out_pt, int* inside) {
    Projection p;
    if (!collider_project(*c, v3(pt[0], pt[1], pt[2]), max_dist, p)) return 0;
    out_pt[0] = p.point.x; out_pt[1] = p.point.y; out_pt[2] = p.point.z;
    *inside = p.is_inside ? 1 : 0;
    return 1;
}
void orc_cuboid_project(const SpkGpuCollider* c, const float* pt, float* out_pt, int* inside) {
    Projection p = cuboid_project_point(*c, v3(pt[0], pt[1], pt[2]));
    out_pt[0] = p.point.x; out_pt[1] = p.point.y; out_pt[2] = p.point.z; *inside = p.is_inside;
}

} // extern "C"

// unit hook: DefaultParticleUpdater::update_particle_and_compute_kirchhoff_stress on ONE particle (tests/test_device_math3d_cpu.py
// compares the product's host-compiled updater with it).  Returns 1 when the reference returns Some(..) (P2G happens).
extern "C" int orc_update_particle(orc_ctx* c, float dt, SpkParticleStatus* st, SpkParticlePosition* pos, SpkParticleVelocity* vel, SpkParticleVolume* vol,
                                   SpkParticlePhase* ph, SpkParticleData* data, const float* interp_vel, float* vel_grad9, float* stress9) {
    M3 g, s = m3_zero();
    std::memcpy(g.m, vel_grad9, 36);
    const bool ok = update_particle(*c, dt, c->cell_width, 0, *st, *pos, *vel, *vol, *ph, *data, v3(interp_vel[0], interp_vel[1], interp_vel[2]), g, trace(g), s);
    std::memcpy(vel_grad9, g.m, 36);
    std::memcpy(stress9, s.m, 36);
    return ok ? 1 : 0;
}
