// mpm_oracle2d.cpp — CPU ORACLE of the 2D (`dim2`) build.  TEST INFRASTRUCTURE ONLY: imported by tests/ and
// __graft_entry__.smoke(), never by the product path (sparkl_b200/ never links it).
//
// What it is: the `dim2` counterpart of mpm_oracle.cpp — a restatement, in plain C++17/f32, of what the reference's GPU path
// executes for one `CudaMpmPipeline::step` when sparkl is built with the `dim2` feature (crate sparkl2d: 4x4-cell blocks,
// 4 associated blocks per particle, 9-node stencil, 32-bit-per-axis block keys).  The reference compiles
// ONE source tree for both dimensions (`#[cfg(feature = "dim2")]` / `DIM`), so every function below cites the same file:line as
// its 3D sibling and states what the dimension changes.  Same canonicalisation as the 3D oracle: block ids are the rank of the
// block key, particles inside a block are ordered by particle id, grid sums are accumulated in double and rounded once.
//
// Third-party arithmetic not under /root/reference: nalgebra 0.32.1 (`svd_unordered` of a 2x2, `try_symmetric_eigen`,
// `determinant`) and parry2d 0.13.0 (`Cuboid::project_point`, `Segment::normal`).  Only U·f(Σ)·Vᵀ products and eigenvalue
// extrema are consumed, so the 2x2 SVD / eigen-decomposition are computed in closed form in double and rounded to f32.
//
// PARITY STATUS: pinned against the reference itself on collider-free scenes.  The reference holds no golden vectors for this
// path; its dim2 kernels (resources/sparkl2d-kernels.ptx, assembled for sm_100a and driven by the -DSPK_DIM=2 build of
// oracle/ref_harness.cpp) were run on a B200 on ten seeded scenes (every constitutive / plastic / failure model) and their outputs
// are committed as tests/golden/*2.npz: this file reproduces them to ~1e-6 relative (tests/test_golden_cpu.py), and product, oracle
// and reference kernels run in lock-step in tests/test_dim2_gpu.py.  NOT YET pinned against the reference: the collider branch of
// grid_update in dim2 (the 280-byte dim2 GpuCollider image has been recovered from the PTX and the harness builds it, but the
// three-way collider tests have not run on a GPU yet: SPK_RUN_UNVERIFIED=1); until then the pins are the shared code path with the
// 3D oracle, brute-force geometry checks and the x <-> y symmetry tests of tests/test_dim2_cpu.py.
#define SPK_DIM 2
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
static const int CELLS = 16; // NUM_CELL_PER_BLOCK, gpu_grid.rs:9-10
static const int TILE = 64;  // NUM_SHARED_CELLS, g2p2g.rs:15-16

// ------------------------------------------------------------------------------------------------ small math
struct V2 {
    Real x, y;
    Real& operator[](int i) { return (&x)[i]; }
    Real operator[](int i) const { return (&x)[i]; }
};
static inline V2 v2(Real x, Real y) { return V2{x, y}; }
static inline V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
static inline V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
static inline V2 operator-(V2 a) { return v2(-a.x, -a.y); }
static inline V2 operator*(V2 a, Real s) { return v2(a.x * s, a.y * s); }
static inline V2 operator*(Real s, V2 a) { return v2(a.x * s, a.y * s); }
static inline V2 operator/(V2 a, Real s) { return v2(a.x / s, a.y / s); }
static inline Real dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
static inline Real norm_squared(V2 a) { return dot(a, a); }
static inline Real norm(V2 a) { return std::sqrt(norm_squared(a)); }
static inline bool is_zero(V2 a) { return a.x == 0.0f && a.y == 0.0f; }

// column-major 2x2, same storage as nalgebra Matrix2 / ParticleVolume::deformation_gradient
struct M2 {
    Real m[4];
    Real& operator()(int r, int c) { return m[c * 2 + r]; }
    Real operator()(int r, int c) const { return m[c * 2 + r]; }
};
static inline M2 m2_zero() { M2 a; for (int i = 0; i < 4; ++i) a.m[i] = 0.0f; return a; }
static inline M2 m2_identity() { M2 a = m2_zero(); a(0, 0) = a(1, 1) = 1.0f; return a; }
static inline M2 operator+(const M2& a, const M2& b) { M2 r; for (int i = 0; i < 4; ++i) r.m[i] = a.m[i] + b.m[i]; return r; }
static inline M2 operator-(const M2& a, const M2& b) { M2 r; for (int i = 0; i < 4; ++i) r.m[i] = a.m[i] - b.m[i]; return r; }
static inline M2 operator*(const M2& a, Real s) { M2 r; for (int i = 0; i < 4; ++i) r.m[i] = a.m[i] * s; return r; }
static inline M2 operator*(Real s, const M2& a) { return a * s; }
static inline M2 operator*(const M2& a, const M2& b) {
    M2 r;
    for (int c = 0; c < 2; ++c)
        for (int rr = 0; rr < 2; ++rr) r(rr, c) = a(rr, 0) * b(0, c) + a(rr, 1) * b(1, c);
    return r;
}
static inline V2 operator*(const M2& a, V2 v) { return v2(a(0, 0) * v.x + a(0, 1) * v.y, a(1, 0) * v.x + a(1, 1) * v.y); }
static inline M2 transpose(const M2& a) { M2 r; for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) r(i, j) = a(j, i); return r; }
static inline Real trace(const M2& a) { return a(0, 0) + a(1, 1); }
static inline Real determinant(const M2& a) { return a(0, 0) * a(1, 1) - a(1, 0) * a(0, 1); } // nalgebra Matrix2::determinant
static inline M2 outer(V2 v, V2 w) { M2 r; for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) r(i, j) = v[i] * w[j]; return r; }

// Rust semantics helpers (as in mpm_oracle.cpp)
static inline Real rmin(Real a, Real b) { return std::fmin(a, b); }
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
static inline Real powi(Real a, int b) { // compiler-rt __powisf2
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
static inline Real inv_exact(Real e) { return e == 0.0f ? 0.0f : 1.0f / e; } // src_core/utils/physics.rs:7-14

// ------------------------------------------------------------------------------------------------ SVD / eigen (2x2, closed form)
// Stand-in for nalgebra 0.32.1 `svd_unordered(true, true)` on a Matrix2.  Since 0.31 nalgebra special-cases 2x2 matrices
// (linalg/svd2.rs `svd_ordered2`): the closed form F = R(phi) diag(sx, sy) R(theta) from the E/F/G/H half-sums, singular values
// returned as (sx, |sy|) with sx >= |sy| — i.e. DESCENDING — and the sign of sy folded into the second row of V^T.  Restated
// here in double.  The order matters for exactly one consumer: Rankine's `idx = [0, 1, DIM - 1]` sort (plasticity_rankine.rs:47).
struct Svd2 { M2 u; V2 s; M2 vt; };
static Svd2 svd_unordered(const M2& a) {
    const double A = a(0, 0), B = a(0, 1), C = a(1, 0), D = a(1, 1);
    const double E = (A + D) * 0.5, F = (A - D) * 0.5, G = (C + B) * 0.5, H = (C - B) * 0.5;
    const double Q = std::hypot(E, H), R = std::hypot(F, G);
    double sx = Q + R, sy = Q - R;
    const double a1 = (G == 0.0 && F == 0.0) ? 0.0 : std::atan2(G, F), a2 = (H == 0.0 && E == 0.0) ? 0.0 : std::atan2(H, E);
    const double theta = (a2 - a1) * 0.5, phi = (a2 + a1) * 0.5;
    // a = R(phi) diag(sx, sy) R(theta)
    double u[2][2] = {{std::cos(phi), -std::sin(phi)}, {std::sin(phi), std::cos(phi)}};
    double vt[2][2] = {{std::cos(theta), -std::sin(theta)}, {std::sin(theta), std::cos(theta)}};
    if (sy < 0.0) { sy = -sy; vt[1][0] = -vt[1][0]; vt[1][1] = -vt[1][1]; }
    Svd2 out;
    for (int r = 0; r < 2; ++r) for (int c = 0; c < 2; ++c) { out.u(r, c) = (Real)u[r][c]; out.vt(r, c) = (Real)vt[r][c]; }
    out.s = v2((Real)sx, (Real)sy);
    return out;
}
static inline M2 recompose(const Svd2& svd) { // SVD::recompose: U * diag(s) * V^T
    M2 us = svd.u;
    for (int c = 0; c < 2; ++c) for (int r = 0; r < 2; ++r) us(r, c) *= svd.s[c];
    return us * svd.vt;
}
// eigenvalues of a symmetric 2x2 (lower triangle read, like nalgebra SymmetricEigen; failure_maximum_stress.rs:21)
static void symmetric_eigenvalues(const M2& a, double ev[2]) {
    const double p = a(0, 0), q = a(1, 1), o = a(1, 0);
    const double mean = (p + q) * 0.5, rad = std::hypot((p - q) * 0.5, o);
    ev[0] = mean - rad; ev[1] = mean + rad;
}

// ------------------------------------------------------------------------------------------------ grid keys
// src_kernels/gpu_grid.rs:73-120 (dim2): 32 bits per axis, x | y << 32, PACK_ORIGIN = 2^31
static const uint64_t PACK_ORIGIN = 0x80000000ull;
static const uint64_t KEY_MASK = 0xffffffffull;
static inline uint64_t pack_key(uint64_t x, uint64_t y) { return (x & KEY_MASK) | ((y & KEY_MASK) << 32); }
static inline void unpack_key(uint64_t k, uint64_t out[2]) { out[0] = k & KEY_MASK; out[1] = (k >> 32) & KEY_MASK; }
// gpu_grid.rs:271-279 (block_associated_to_point)
static inline void main_block_of_point(const float p[2], Real h, uint64_t out[2]) {
    for (int a = 0; a < 2; ++a) {
        int64_t coord = sat_i64(std::round(p[a] / h));
        out[a] = (uint64_t)(coord - 2 + (int64_t)PACK_ORIGIN * 4) / 4; // `as usize / 4`
    }
}
// NBH_SHIFTS (dim2), gpu_grid.rs:18-29 (NBH_SHIFTS_SHARED[i] == x + 8y, :67-68)
static const int NBH[9][2] = {{2, 2}, {2, 0}, {2, 1}, {0, 2}, {0, 0}, {0, 1}, {1, 2}, {1, 0}, {1, 1}};

struct Grid {
    std::vector<SpkActiveBlockHeader> headers;  // id == rank of virtual_id (canonical form)
    std::vector<SpkGpuGridNode> nodes;          // 16 per block, node = x + 4 y (gpu_grid.rs:145-148)
    std::unordered_map<uint64_t, uint32_t> map; // GridHashMap: only set semantics matter
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
    std::vector<std::vector<float>> collider_data; // owned copies of the polyline / trimesh vertices and the height samples
    std::vector<std::vector<uint32_t>> collider_idx; // owned copies of the trimesh indices
    Grid grids[2];
    int curr = 0;
    std::vector<SpkDispatchBlock2ActiveBlock> dispatch;
    uint64_t dt_acc = 0;
    Grid& curr_grid() { return grids[curr]; }
    Grid& next_grid() { return grids[1 - curr]; }
};

namespace {

static inline M2 load_F(const SpkParticleVolume& v) { M2 f; std::memcpy(f.m, v.deformation_gradient, sizeof(f.m)); return f; }
static inline void store_F(SpkParticleVolume& v, const M2& f) { std::memcpy(v.deformation_gradient, f.m, sizeof(f.m)); }

// ------------------------------------------------------------------------------------------------ models
// DIM-dependent constants are marked `DIM`; literal 2/3, 4/3 and 3 of the Rust source stay literal.
static const Real DIMF = 2.0f;

// elasticity_corotated_linear.rs:31-77 (dimension-generic source)
static M2 corotated_kirchhoff_stress(const SpkCorotatedLinearElasticity& m, Real phase, Real hardening, const M2& F) {
    Real j = determinant(F);
    Svd2 svd = svd_unordered(F);
    for (int i = 0; i < 2; ++i) svd.s[i] = svd.s[i] - 1.0f;
    if (phase == 1.0f) return (2.0f * m.mu * hardening) * recompose(svd) * transpose(F) + (m.lambda * hardening * (j - 1.0f) * j) * m2_identity();
    Svd2 pos_def = svd, neg_def = svd;
    for (int i = 0; i < 2; ++i) { pos_def.s[i] = rmax(pos_def.s[i], 0.0f); neg_def.s[i] = rmin(neg_def.s[i], 0.0f); }
    M2 pos_dev = 2.0f * m.mu * hardening * recompose(pos_def) * transpose(F);
    M2 neg_dev = 2.0f * m.mu * hardening * recompose(neg_def) * transpose(F);
    M2 spherical = (m.lambda * hardening * (j - 1.0f) * j) * m2_identity();
    M2 pos_part, neg_part;
    if (j < 1.0f) { pos_part = pos_dev; neg_part = neg_dev + spherical; }
    else { pos_part = pos_dev + spherical; neg_part = neg_dev; }
    Real phase_coeff = (m.split_stress_on_failure && phase == 0.0f) ? 0.0f : 1.0f;
    return pos_part * phase_coeff + neg_part;
}
// elasticity_corotated_linear.rs:104-118
static Real corotated_pos_energy(const SpkCorotatedLinearElasticity& m, const M2& F, Real hardening) {
    Real j = determinant(F);
    Svd2 svd = svd_unordered(F);
    V2 s;
    for (int i = 0; i < 2; ++i) s[i] = rmax(svd.s[i] - 1.0f, 0.0f);
    Real pos_dev_part = m.mu * hardening * norm_squared(s);
    Real spherical_part = m.lambda * hardening / 2.0f * (j - 1.0f) * (j - 1.0f);
    return j < 1.0f ? pos_dev_part : pos_dev_part + spherical_part;
}
// elasticity_sound_speed_timestep_bound.rs:22-34 (the 4/3 is a literal there)
static Real sound_speed_bound(Real alpha, Real bulk, Real shear, Real density0, V2 v, Real cell_width) {
    Real c_dir = std::sqrt((bulk + 4.0f / 3.0f * shear) * 1.0f / density0);
    Real max_denom = rmax(norm(v), c_dir);
    return alpha * cell_width / max_denom;
}
static Real elastic_timestep_bound(Real cfl, Real lambda, Real mu, Real density0, V2 v, Real elastic_hardening, Real cell_width) {
    Real bulk = lambda + 2.0f * mu / 3.0f; // physics.rs:37-39 (literal 3)
    return sound_speed_bound(cfl, bulk * elastic_hardening, mu * elastic_hardening, density0, v, cell_width);
}
static inline M2 deviatoric_part(const M2& t) { // DecomposedTensor::decompose: spherical part = trace / DIM (physics.rs:78-80)
    Real sph = trace(t) / DIMF;
    return t - sph * m2_identity();
}
static inline Real nh_phase_coeff(Real phase) { const Real R = 0.001f; return (1.0f - R) * phase * phase + R; }
// elasticity_neo_hookean.rs:43-67: k uses the literal 2/3, the J exponent uses DIM
static M2 neo_hookean_kirchhoff_stress(const SpkNeoHookeanElasticity& m, Real phase, Real hardening, const M2& F) {
    Real pc = nh_phase_coeff(phase);
    Real j = determinant(F);
    Real k = 2.0f / 3.0f * m.mu * hardening + m.lambda * hardening;
    M2 cg = F * transpose(F);
    M2 dev = (m.mu * hardening) * std::pow(j, -2.0f / DIMF) * deviatoric_part(cg);
    M2 vol = k / 2.0f * (j * j - 1.0f) * m2_identity();
    M2 pos_part, neg_part;
    if (j >= 1.0f) { pos_part = dev + vol; neg_part = m2_zero(); }
    else { pos_part = dev; neg_part = vol; }
    return pos_part * pc + neg_part;
}
// elasticity_neo_hookean.rs:117-141
static Real neo_hookean_pos_energy(const SpkNeoHookeanElasticity& m, Real phase, Real hardening, const M2& F) {
    Real pc = nh_phase_coeff(phase);
    Real j = determinant(F);
    Real k = 2.0f / 3.0f * m.mu * hardening + m.lambda * hardening;
    Real dev = hardening * m.mu / 2.0f * (trace(F * transpose(F)) * std::pow(j, -2.0f / DIMF) - DIMF);
    Real vol = k / 2.0f * ((j * j - 1.0f) / 2.0f - std::log(j));
    return j < 1.0f ? dev * pc : (dev + vol) * phase;
}
// eos_monaghan_sph.rs:28-38
static Real monaghan_pressure(const SpkMonaghanSphEos& m, Real mass, Real volume0, Real density_fluid) {
    Real density0 = mass / volume0;
    return rmax(m.pressure0 * (powi(density_fluid / density0, (int)m.gamma) - 1.0f), -m.max_neg_pressure);
}
// eos_monaghan_sph.rs:44-63
static M2 monaghan_kirchhoff_stress(const SpkMonaghanSphEos& m, Real mass, Real volume0, Real density_fluid, Real jdet, const M2& vel_grad) {
    M2 stress = (-monaghan_pressure(m, mass, volume0, density_fluid) * jdet) * m2_identity();
    if (m.viscosity != 0.0f) {
        M2 strain_rate = (vel_grad + transpose(vel_grad)) * 0.5f;
        stress = stress + (2.0f * m.viscosity * jdet) * deviatoric_part(strain_rate);
    }
    return stress;
}
// eos_monaghan_sph.rs:72-99: d = DIM (:85)
static Real monaghan_timestep_bound(const SpkMonaghanSphEos& m, Real j, Real mass, Real volume0, Real density_fluid, V2 v, Real cell_width) {
    Real density0 = mass / volume0;
    Real k = 6.0f, d = DIMF;
    Real pressure = -monaghan_pressure(m, mass, volume0, density_fluid);
    Real single_particle_dt = (cell_width / j) * std::sqrt(density0 * (j - 1.0f) / (k * pressure * d));
    Real density_fluctuation = 0.1f;
    Real c_square = rmax(norm_squared(v), 1.0f) / density_fluctuation;
    Real cfl_dt = cell_width / std::sqrt(c_square);
    return rmin(single_particle_dt, cfl_dt);
}

// plasticity_drucker_prager.rs:36-104: d = DIM (:42); alpha keeps its literal sqrt(2/3) and 3 (:64-69)
static void drucker_prager_update(const SpkDruckerPragerPlasticity& m, Real phase, M2& F, Real& jp, Real& plastic_hardening, Real& log_vol_gain) {
    if (m.only_active_when_failed && phase != 0.0f) return;
    Svd2 svd = svd_unordered(F);
    Real q = plastic_hardening;
    Real angle = m.h0 + (m.h1 * q - m.h3) * std::exp(-m.h2 * q);
    Real s_angle = std::sin(angle);
    Real alpha = std::sqrt(2.0f / 3.0f) * (2.0f * s_angle) / (3.0f - s_angle);
    const Real d = DIMF;
    V2 strain = v2(std::log(svd.s.x), std::log(svd.s.y)) + v2(log_vol_gain / d, log_vol_gain / d);
    Real strain_trace = strain.x + strain.y;
    V2 dev = strain - v2(strain_trace / d, strain_trace / d);
    V2 new_s;
    Real dq;
    if (is_zero(dev) || strain_trace > 0.0f) {
        new_s = v2(1.0f, 1.0f);
        dq = norm(strain);
    } else {
        Real gamma = norm(dev) + (d * m.lambda + 2.0f * m.mu) / (2.0f * m.mu) * strain_trace * alpha;
        if (gamma <= 0.0f) return;
        V2 hh = strain - gamma * (dev / norm(dev));
        new_s = v2(std::exp(hh.x), std::exp(hh.y));
        dq = gamma;
    }
    Real prev_det = svd.s.x * svd.s.y;
    Real new_det = new_s.x * new_s.y;
    Real diff = new_det - prev_det;
    new_det = diff > 0.0f ? new_det : prev_det + diff * m.volume_correction;
    jp *= prev_det / new_det;
    log_vol_gain += std::log(prev_det) - std::log(new_det);
    svd.s = new_s;
    plastic_hardening += dq;
    F = recompose(svd);
}
// plasticity_nacc.rs:61-165: d = DIM (:69), j_e_tr = s0 * s1 (:77-78)
static void nacc_update(const SpkNaccPlasticity& m, M2& F, Real& alpha_io) {
    Real alpha = alpha_io;
    const Real xi = m.hardening_factor, beta = m.cohesion, mm = m.friction, d = DIMF;
    Svd2 svd = svd_unordered(F);
    V2 sq = v2(svd.s.x * svd.s.x, svd.s.y * svd.s.y);
    Real sq_tr = sq.x + sq.y;
    Real p0 = m.kappa * (1.0e-5f + std::sinh(xi * rmax(-alpha, 0.0f)));
    Real j_e_tr = svd.s.x * svd.s.y;
    V2 s_tr = (m.mu * std::pow(j_e_tr, -2.0f / d)) * (sq - v2(sq_tr / d, sq_tr / d));
    Real psi_kappa = m.kappa / 2.0f * (j_e_tr - 1.0f / j_e_tr);
    Real p_tr = -psi_kappa * j_e_tr;
    if (p_tr > p0) {
        Real j_e_n1 = std::sqrt(-2.0f * p0 / m.kappa + 1.0f);
        Real v = std::pow(j_e_n1, 1.0f / d);
        svd.s = v2(v, v);
        if (m.hardening_enabled) alpha += std::log(j_e_tr / j_e_n1);
        F = recompose(svd); alpha_io = alpha; return;
    }
    if (p_tr < -beta * p0) {
        Real j_e_n1 = std::sqrt(2.0f * beta * p0 / m.kappa + 1.0f);
        Real v = std::pow(j_e_n1, 1.0f / d);
        svd.s = v2(v, v);
        if (m.hardening_enabled) alpha += std::log(j_e_tr / j_e_n1);
        F = recompose(svd); alpha_io = alpha; return;
    }
    Real y0 = (1.0f + 2.0f * beta) * ((6.0f - d) / 2.0f);
    Real y1 = mm * mm * (p_tr + beta * p0) * (p_tr - p0);
    Real y = y0 * norm_squared(s_tr) + y1;
    if (y < 1.0e-4f) return;
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
    V2 b_e_n1 = (std::sqrt(-y1 / y0) * (std::pow(j_e_tr, 2.0f / d) / m.mu)) * (s_tr / norm(s_tr)) + v2(sq_tr / d, sq_tr / d);
    svd.s = v2(std::sqrt(b_e_n1.x), std::sqrt(b_e_n1.y));
    F = recompose(svd);
    alpha_io = alpha;
}
// plasticity_rankine.rs:35-78 in 2D: idx = [0, 1, DIM - 1] = [0, 1, 1] sorted by eigv (:47-49), the `DIM == 3` branch (:64) is
// never taken, the last branch keeps its literal 3 * lambda
static void rankine_update(const SpkRankinePlasticity& m, M2& F, Real& plastic_hardening) {
    const Real lambda = m.lambda, mu = m.mu;
    Svd2 svd = svd_unordered(F);
    V2 eigv = v2(std::log(svd.s.x), std::log(svd.s.y));
    V2 prev = eigv;
    int idx[3] = {0, 1, 1};
    for (int i = 1; i < 3; ++i) { int k = idx[i]; int jx = i - 1; while (jx >= 0 && eigv[idx[jx]] > eigv[k]) { idx[jx + 1] = idx[jx]; --jx; } idx[jx + 1] = k; }
    int e2 = idx[1], e1 = idx[2];
    Real soft = m.tensile_strength - (plastic_hardening - 1.0f);
    Real sum = eigv.x + eigv.y;
    if (lambda * sum + 2.0f * mu * eigv[e1] <= soft) {
        return;
    } else if ((2.0f * mu + lambda) * eigv[e2] + lambda * (sum - eigv[e1]) <= soft) {
        eigv[e1] = (soft - lambda * (sum - eigv[e1])) / (2.0f * mu + lambda);
    } else {
        Real ne = soft / (2.0f * mu + 3.0f * lambda);
        eigv = v2(ne, ne);
    }
    plastic_hardening += m.softening_rate * norm(prev - eigv);
    plastic_hardening = rmin(plastic_hardening, m.tensile_strength);
    svd.s = v2(std::exp(eigv.x), std::exp(eigv.y));
    F = recompose(svd);
}
// plasticity_snow.rs:26-48
static void snow_update(const SpkSnowPlasticity& m, M2& F, Real& elastic_hardening, Real& jp) {
    Svd2 svd = svd_unordered(F);
    V2 ns;
    for (int i = 0; i < 2; ++i) ns[i] = rmin(rmax(svd.s[i], 1.0f - m.min_epsilon), 1.0f + m.max_epsilon);
    jp *= (svd.s.x * svd.s.y) / (ns.x * ns.y);
    elastic_hardening = std::exp(m.hardening_coeff * (1.0f - jp));
    svd.s = ns;
    F = recompose(svd);
}
// failure_maximum_stress.rs:20-28
static bool maximum_stress_failed(const SpkMaximumStressFailure& m, const M2& stress) {
    for (int i = 0; i < 4; ++i) if (!std::isfinite(stress.m[i])) return false;
    double ev[2];
    symmetric_eigenvalues(stress, ev);
    Real mn = (Real)ev[0], mx = (Real)ev[1];
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
static M2 model_kirchhoff_stress(const SpkGpuParticleModel& m, const SpkParticleData& d, const SpkParticleVolume& vol, Real phase, const M2& vel_grad) {
    switch (m.constitutive_tag) {
    case SPK_CONSTITUTIVE_COROTATED: return corotated_kirchhoff_stress(m.constitutive.corotated.m, phase, d.elastic_hardening, load_F(vol));
    case SPK_CONSTITUTIVE_NEO_HOOKEAN: return neo_hookean_kirchhoff_stress(m.constitutive.neo_hookean.m, phase, d.elastic_hardening, load_F(vol));
    case SPK_CONSTITUTIVE_MONAGHAN: {
        Real density_fluid = (vol.mass / vol.volume0) / vol.deformation_gradient[0];
        return monaghan_kirchhoff_stress(m.constitutive.monaghan.m, vol.mass, vol.volume0, density_fluid, vol.plastic_deformation_gradient_det, vel_grad);
    }
    default: return m2_zero();
    }
}
static inline unsigned model_active_bounds(const SpkGpuParticleModel& m) {
    switch (m.constitutive_tag) {
    case SPK_CONSTITUTIVE_COROTATED: case SPK_CONSTITUTIVE_NEO_HOOKEAN: return 1u | 2u | 8u;
    case SPK_CONSTITUTIVE_MONAGHAN: return 1u | 2u | 4u | 16u;
    default: return 0u;
    }
}
static Real model_timestep_bound(const SpkGpuParticleModel& m, const SpkParticleData& d, const SpkParticleVolume& vol, V2 v, Real cell_width) {
    Real density0 = vol.mass / vol.volume0;
    switch (m.constitutive_tag) {
    case SPK_CONSTITUTIVE_COROTATED: {
        const auto& c = m.constitutive.corotated.m;
        return elastic_timestep_bound(c.cfl_coeff, c.lambda, c.mu, density0, v, d.elastic_hardening, cell_width);
    }
    case SPK_CONSTITUTIVE_NEO_HOOKEAN: { // (cell_width, hardening) swapped, gpu_constitutive_model.rs:129-137
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
static void model_plastic_update(const SpkGpuParticleModel& m, SpkParticleData& d, SpkParticleVolume& vol, Real phase) {
    M2 F = load_F(vol);
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
static inline void eval_all(Real x, Real w[3]) { // QuadraticKernel::eval_all, kernel.rs:15-21
    w[0] = 0.5f * ((1.5f - x) * (1.5f - x));
    w[1] = 0.75f - ((x - 1.0f) * (x - 1.0f));
    w[2] = 0.5f * ((x - 0.5f) * (x - 0.5f));
}
static inline V2 dir_to_associated_grid_node(const float p[2], Real h) { // particle.rs:37-47
    V2 r;
    for (int a = 0; a < 2; ++a) r[a] = (std::round(p[a] / h) - 1.0f) * h - p[a];
    return r;
}

struct TileCell { // GridGatherData, g2p2g.rs:22-40
    V2 velocity; Real psi_velocity;
    double mass, momentum[2], psi_momentum, psi_mass;
};

// DefaultParticleUpdater::update_particle_and_compute_kirchhoff_stress, particle_updater.rs:73-236
static bool update_particle(const orc_ctx& ctx, Real dt, Real cell_width, SpkParticleStatus& st, SpkParticlePosition& pos, SpkParticleVelocity& vel,
                            SpkParticleVolume& vol, SpkParticlePhase& ph, SpkParticleData& data, V2 interp_vel, M2& vel_grad, Real vel_grad_det,
                            M2& stress_out) {
    const SpkGpuParticleModel& model = ctx.models[st.model_index];
    V2 v = interp_vel;
    const bool is_fluid = model_is_fluid(model);
    if (st.kinematic_vel_enabled) v = v2(st.kinematic_vel[0], st.kinematic_vel[1]);
    if (std::fabs(v.x) * dt >= cell_width || std::fabs(v.y) * dt >= cell_width) {
        for (int a = 0; a < 2; ++a) v[a] = signum(v[a]) * cell_width / dt;
    }
    for (int a = 0; a < 2; ++a) pos.point[a] += v[a] * dt;
    if (!is_fluid) {
        M2 F = load_F(vol);
        F = F + (vel_grad * dt) * F;
        store_F(vol, F);
    } else {
        vol.deformation_gradient[0] += (vel_grad_det * dt) * vol.deformation_gradient[0];
    }
    if (model.plastic_tag != SPK_PLASTIC_NONE) model_plastic_update(model, data, vol, ph.phase);
    if (st.is_static) { v = v2(0, 0); vel_grad = m2_zero(); }
    for (int a = 0; a < 2; ++a) vel.vector[a] = v[a];
    Real density_def_grad = (vol.mass / vol.volume0) / determinant(load_F(vol));
    if (density_def_grad == 0.0f || st.failed || (!is_fluid && std::fabs(vol.deformation_gradient[0]) > 1.0e4f)) {
        st.failed = 1;
        store_F(vol, m2_identity());
        return false;
    }
    Real energy = model_pos_energy(model, data, vol, ph.phase);
    ph.psi_pos = rmax(ph.psi_pos, energy);
    if (model.failure_tag == SPK_FAILURE_MAXIMUM_STRESS) {
        M2 s = model_kirchhoff_stress(model, data, vol, ph.phase, vel_grad);
        if (maximum_stress_failed(model.failure.maximum_stress, s)) ph.phase = 0.0f;
    }
    stress_out = model_kirchhoff_stress(model, data, vol, ph.phase, vel_grad);
    return true;
}

// particle_g2p2g, g2p2g.rs:179-351 (dim2: packed = (x + 1) + (y + 1) * 8, 9 neighbours)
static void particle_g2p2g(const orc_ctx& ctx, Real dt, SpkParticleStatus& st, SpkParticlePosition& pos, SpkParticleVelocity& vel, SpkParticleVolume& vol,
                           SpkParticlePhase& ph, SpkParticleData& data, TileCell* tile) {
    const Real h = ctx.cell_width;
    const Real inv_d = 4.0f / (h * h); // Kernel::inv_d
    V2 ref = dir_to_associated_grid_node(pos.point, h);
    Real w[2][3];
    for (int a = 0; a < 2; ++a) eval_all(-ref[a] / h, w[a]);
    Real cell_before[2];
    int cib[2];
    for (int a = 0; a < 2; ++a) {
        Real e = pos.point[a] / h;
        cell_before[a] = std::round(e);
        int assoc_cell = sat_i32(std::round(e)) - 2; // particle.rs:28-35
        int assoc_block = sat_i32(std::floor((Real)assoc_cell / 4.0f)) * 4;
        cib[a] = assoc_cell - assoc_block;
    }
    int packed = (cib[0] + 1) + (cib[1] + 1) * 8;
    V2 iv = v2(0, 0);
    M2 grad = m2_zero();
    Real grad_det = 0.0f;
    for (int i = 0; i < 9; ++i) {
        const int* s = NBH[i];
        V2 dpt = ref + v2((Real)s[0], (Real)s[1]) * h;
        Real weight = w[0][s[0]] * w[1][s[1]];
        const TileCell& cell = tile[packed + s[0] + 8 * s[1]];
        iv = iv + weight * cell.velocity;
        grad = grad + outer((weight * inv_d) * cell.velocity, dpt);
        grad_det += weight * dot(cell.velocity, dpt) * inv_d;
    }
    M2 stress;
    if (!update_particle(ctx, dt, h, st, pos, vel, vol, ph, data, iv, grad, grad_det, stress)) return;
    V2 ref2 = dir_to_associated_grid_node(pos.point, h);
    Real w2[2][3];
    for (int a = 0; a < 2; ++a) eval_all(-ref2[a] / h, w2[a]);
    M2 affine = vol.mass * grad - (vol.volume0 * inv_d * dt) * stress;
    V2 momentum = vol.mass * v2(vel.vector[0], vel.vector[1]) + v2(0, 0) * dt;
    Real psi_mass = (ph.phase > 0.0f && !st.failed) ? vol.mass : 0.0f;
    Real psi_pos_mom = psi_mass * ph.psi_pos;
    int64_t mv[2];
    for (int a = 0; a < 2; ++a) mv[a] = sat_i64(std::round(pos.point[a] / h) - cell_before[a]);
    int64_t packed2 = (int64_t)packed + mv[0] + mv[1] * 8;
    for (int i = 0; i < 9; ++i) {
        const int* s = NBH[i];
        V2 dpt = ref2 + v2((Real)s[0], (Real)s[1]) * h;
        Real weight = w2[0][s[0]] * w2[1][s[1]];
        int64_t ci = packed2 + s[0] + 8 * s[1];
        if (ci < 0 || ci >= TILE) continue; // cannot happen after the velocity clamp; guards NaN inputs
        TileCell& cell = tile[ci];
        V2 am = weight * (affine * dpt + momentum);
        cell.mass += (double)(weight * vol.mass);
        cell.momentum[0] += (double)am.x; cell.momentum[1] += (double)am.y;
        cell.psi_momentum += (double)(weight * psi_pos_mom);
        cell.psi_mass += (double)(weight * psi_mass);
    }
}

// ------------------------------------------------------------------------------------------------ colliders (parry2d)
struct Projection { V2 point; bool is_inside; };
static inline V2 rot(const float c[2], V2 v) { return v2(c[0] * v.x - c[1] * v.y, c[1] * v.x + c[0] * v.y); }     // UnitComplex * Vector
static inline V2 rot_inv(const float c[2], V2 v) { return v2(c[0] * v.x + c[1] * v.y, -c[1] * v.x + c[0] * v.y); } // inverse
// parry2d 0.13 Aabb::project_local_point through Cuboid::project_point with solid = false (gpu_collider.rs:53)
static Projection cuboid_project_point(const SpkGpuCollider& c, V2 pt) {
    V2 tr = v2(c.translation[0], c.translation[1]);
    V2 lp = rot_inv(c.rotation, pt - tr);
    V2 he = v2(c.shape.cuboid.half_extents[0], c.shape.cuboid.half_extents[1]);
    V2 mins_pt = (-he) - lp, pt_maxs = lp - he;
    V2 shift;
    for (int a = 0; a < 2; ++a) shift[a] = rmax(mins_pt[a], 0.0f) - rmax(pt_maxs[a], 0.0f);
    Projection out;
    if (!is_zero(shift)) {
        out.is_inside = false;
        out.point = lp + shift;
    } else {
        Real best = -FLT_MAX;
        bool is_mins = false;
        int best_id = 0;
        for (int a = 0; a < 2; ++a) {
            if (mins_pt[a] < pt_maxs[a]) {
                if (pt_maxs[a] > best) { best_id = a; is_mins = false; best = pt_maxs[a]; }
            } else if (mins_pt[a] > best) { best_id = a; is_mins = true; best = mins_pt[a]; }
        }
        V2 sh = v2(0, 0);
        sh[best_id] = is_mins ? best : -best;
        out.is_inside = true;
        out.point = lp + sh;
    }
    out.point = rot(c.rotation, out.point) + tr;
    return out;
}
// polyline_project_point, src_kernels/gpu_collider.rs:105-201 (dim2: normal1 = Segment::normal() = normalize((dir.y, -dir.x)))
static bool polyline_project_point(const SpkGpuCollider& c, V2 point, Projection& out) {
    const auto& pl = c.shape.polyline;
    const int n = (int)pl.num_vertices;
    if (n == 0) return false;
    V2 tr = v2(c.translation[0], c.translation[1]);
    V2 local_point = rot_inv(c.rotation, point - tr);
    auto vertex = [&](int i) { return v2(pl.vertices[2 * i], pl.vertices[2 * i + 1]); };
    struct Best { V2 proj; int loc; Real dist; int id; V2 a, b; bool is_inside; } best{};
    bool have = false;
    for (int i = 0; i < n; ++i) {
        V2 a = vertex(i), b = vertex((i + 1) % n);
        V2 ab = b - a, ap = local_point - a;
        Real ab_ap = dot(ab, ap), sqnab = dot(ab, ab);
        V2 q; int loc;
        if (ab_ap <= 0) { q = a; loc = 0; }
        else if (ab_ap >= sqnab) { q = b; loc = 1; }
        else { q = a + ab * (ab_ap / sqnab); loc = 2; }
        Real dist = norm(local_point - q);
        if (!have || dist < best.dist) { best = Best{q, loc, dist, i, a, b, q.x == local_point.x && q.y == local_point.y}; have = true; }
    }
    V2 dir = best.b - best.a;
    Real nn = std::sqrt(dir.y * dir.y + dir.x * dir.x);
    if (nn > FLT_EPSILON) {
        V2 normal1 = v2(dir.y / nn, -dir.x / nn);
        Real side = dot(point - best.proj, normal1); // WORLD point minus LOCAL projection, as in the reference (:176,181)
        if (best.loc == 2) best.is_inside = side <= 0;
        else {
            V2 dir2;
            if (best.loc == 0) { int adj = best.id == 0 ? n - 1 : best.id - 1; dir2 = -(vertex((adj + 1) % n) - vertex(adj)); }
            else { int adj = (best.id + 1) % n; dir2 = vertex((adj + 1) % n) - vertex(adj); }
            Real d = dot(normal1, dir2), threshold = 1.0e-3f * norm(dir2);
            best.is_inside = std::fabs(d) > threshold ? d >= 0 : side <= 0;
        }
    }
    if (pl.flip_interior) best.is_inside = !best.is_inside;
    out.is_inside = best.is_inside;
    out.point = rot(c.rotation, best.proj) + tr;
    return true;
}
// GenericHeightField::project_local_point_with_max_dist over map_elements_in_local_aabb (parry2d 0.13 point_heightfield.rs,
// heightfield2.rs; restated from the published algorithm, like its 3D sibling in mpm_oracle.cpp), then the inside test of
// GpuColliderShape::project_point_with_max_dist (src_kernels/gpu_collider.rs:73-99).  Unpinned against the reference's kernels.
static bool heightfield_project_point(const SpkGpuCollider& c, V2 pt, Real max_dist, Projection& out) {
    const auto& hf = c.shape.heightfield;
    V2 tr = v2(c.translation[0], c.translation[1]);
    V2 lp = rot_inv(c.rotation, pt - tr);
    const int cells = (int)hf.num_heights - 1;
    const Real seg = 1.0f / (Real)cells;
    const Real rminx = (lp.x - max_dist) / hf.scale[0], rmaxx = (lp.x + max_dist) / hf.scale[0];
    if (rmaxx <= -0.5f || rminx >= 0.5f) return false;
    auto clampf = [](Real v, Real lo, Real hi) { return v < lo ? lo : (v > hi ? hi : v); };
    const int min_x = (int)clampf(std::floor((rminx + 0.5f) / seg), 0.0f, (Real)(cells - 1));
    const int max_x = (int)clampf(std::ceil((rmaxx + 0.5f) / seg), 0.0f, (Real)cells);
    Real sq_smallest = FLT_MAX;
    bool found = false;
    V2 best = v2(0, 0);
    for (int i = min_x; i < max_x; ++i) {
        V2 a = v2((-0.5f + seg * (Real)i) * hf.scale[0], hf.heights[i] * hf.scale[1]);
        V2 b = v2((-0.5f + seg * (Real)(i + 1)) * hf.scale[0], hf.heights[i + 1] * hf.scale[1]);
        V2 ab = b - a, ap = lp - a;
        Real ab_ap = dot(ab, ap), sqnab = dot(ab, ab);
        V2 q = ab_ap <= 0 ? a : (ab_ap >= sqnab ? b : a + ab * (ab_ap / sqnab));
        Real sq = norm_squared(lp - q);
        if (sq < sq_smallest) {
            sq_smallest = sq;
            if (std::sqrt(sq) <= max_dist) { best = q; found = true; }
        }
    }
    if (!found) return false;
    if (hf.bellow_heightfield_is_solid) {
        bool in_footprint = lp.x >= -0.5f * hf.scale[0] && lp.x <= 0.5f * hf.scale[0];
        out.is_inside = in_footprint && (hf.flip_interior ? lp.y >= best.y : lp.y <= best.y);
    } else {
        out.is_inside = best.x == lp.x && best.y == lp.y;
    }
    out.point = rot(c.rotation, best) + tr;
    return true;
}
// parry2d 0.13 Triangle::project_local_point_and_get_location (query/point/point_triangle.rs with the `dim2` feature), solid = false.
// parry2d is an un-vendored dependency (crates-ptx/Cargo.lock pins 0.13.0): restated from its published algorithm — vertex regions,
// edge regions through perp products, and for a point inside the triangle the nearest of the three edge lines, flagged inside.
// PARITY: unpinned against the reference's kernels (its TriMesh payload holds a device QBVH that is not reproduced); checked against an
// independent f64 brute force (tests/test_dim2_cpu.py).
static void triangle_project2(V2 a, V2 b, V2 c, V2 p, V2& q, bool& inside) {
    V2 ab = b - a, ac = c - a, ap = p - a;
    Real ab_ap = dot(ab, ap), ac_ap = dot(ac, ap);
    inside = false;
    if (ab_ap <= 0.0f && ac_ap <= 0.0f) { q = a; inside = p.x == a.x && p.y == a.y; return; }
    V2 bp = p - b;
    Real ab_bp = dot(ab, bp), ac_bp = dot(ac, bp);
    if (ab_bp >= 0.0f && ac_bp <= ab_bp) { q = b; inside = p.x == b.x && p.y == b.y; return; }
    V2 cp = p - c;
    Real ab_cp = dot(ab, cp), ac_cp = dot(ac, cp);
    if (ac_cp >= 0.0f && ab_cp <= ac_cp) { q = c; inside = p.x == c.x && p.y == c.y; return; }
    V2 bc = c - b;
    Real n = ab.x * ac.y - ab.y * ac.x;
    Real sq_ab = dot(ab, ab), sq_ac = dot(ac, ac), sq_bc = dot(bc, bc);
    Real vc = n * (ab.x * ap.y - ab.y * ap.x);
    if (vc < 0.0f && ab_ap >= 0.0f && ab_bp <= 0.0f) { Real v = ab_ap / sq_ab; q = a + ab * v; inside = q.x == p.x && q.y == p.y; return; }
    Real vb = -n * (ac.x * cp.y - ac.y * cp.x);
    if (vb < 0.0f && ac_ap >= 0.0f && ac_cp <= 0.0f) { Real w = ac_ap / sq_ac; q = a + ac * w; inside = q.x == p.x && q.y == p.y; return; }
    Real va = n * (bc.x * bp.y - bc.y * bp.x);
    if (va < 0.0f && ac_bp - ab_bp >= 0.0f && ab_cp - ac_cp >= 0.0f) { Real w = dot(bc, bp) / sq_bc; q = b + bc * w; inside = q.x == p.x && q.y == p.y; return; }
    Real v = ab_ap / (ab_ap - ab_bp), w = ac_ap / (ac_ap - ac_cp), u = (ac_bp - ab_bp) / (ac_bp - ab_bp + ab_cp - ac_cp);
    Real sq_ap = dot(ap, ap), sq_bp = dot(bp, bp);
    Real d_ab = sq_ap - sq_ab * v * v, d_ac = sq_ap - sq_ac * w * w, d_bc = sq_bp - sq_bc * u * u;
    inside = true;
    if (d_ab < d_ac) q = d_ab < d_bc ? a + ab * v : b + bc * u;
    else q = d_ac < d_bc ? a + ac * w : b + bc * u;
}
// GpuColliderShape::TriMesh, src_kernels/gpu_collider.rs:58-68: project_point_with_max_dist(.., Real::MAX) = the closest triangle of the
// whole mesh (ties: lowest index), then flip_interior
static bool trimesh_project_point(const SpkGpuCollider& c, V2 pt, Projection& out) {
    const auto& tm = c.shape.trimesh;
    V2 tr = v2(c.translation[0], c.translation[1]);
    V2 lp = rot_inv(c.rotation, pt - tr);
    Real best_sq = FLT_MAX;
    bool found = false, best_in = false;
    V2 best = v2(0, 0);
    for (uint32_t t = 0; t < tm.num_triangles; ++t) {
        const uint32_t* id = tm.indices + 3 * (size_t)t;
        V2 q; bool in;
        triangle_project2(v2(tm.vertices[2 * id[0]], tm.vertices[2 * id[0] + 1]), v2(tm.vertices[2 * id[1]], tm.vertices[2 * id[1] + 1]),
                          v2(tm.vertices[2 * id[2]], tm.vertices[2 * id[2] + 1]), lp, q, in);
        Real sq = norm_squared(lp - q);
        if (sq < best_sq) { best_sq = sq; best = q; best_in = in; found = true; }
    }
    if (!found) return false;
    out.is_inside = tm.flip_interior ? !best_in : best_in;
    out.point = rot(c.rotation, best) + tr;
    return true;
}
static bool collider_project(const SpkGpuCollider& c, V2 pt, Real max_dist, Projection& out) {
    if (c.shape_tag == SPK_SHAPE_TRIMESH) return trimesh_project_point(c, pt, out);
    if (c.shape_tag == SPK_SHAPE_HEIGHTFIELD) return heightfield_project_point(c, pt, max_dist, out);
    if (c.shape_tag == SPK_SHAPE_POLYLINE) return polyline_project_point(c, pt, out);
    if (c.shape_tag != SPK_SHAPE_CUBOID) return false;
    out = cuboid_project_point(c, pt);
    return true;
}
// sdf, grid_update.rs:54-71
static bool sdf(const orc_ctx& ctx, V2 q, Real max_dist, Real& out) {
    bool any = false;
    Real best = 0;
    for (const auto& c : ctx.colliders) {
        if (c.grid_boundary_handling == SPK_BOUNDARY_NONE) continue;
        Projection p;
        if (!collider_project(c, q, max_dist, p)) continue;
        Real d = norm(p.point - q) * (p.is_inside ? -1.0f : 1.0f);
        if (!any || d < best) { best = d; any = true; }
    }
    out = best;
    return any;
}
// sdf_gradient, grid_update.rs:74-94 (x and y only in dim2, :89)
static V2 sdf_gradient(const orc_ctx& ctx, Real h, V2 q) {
    const Real FACTOR = 0.1f;
    V2 g;
    for (int a = 0; a < 2; ++a) {
        V2 off = v2(0, 0);
        off[a] = 1.0f * (h * FACTOR);
        Real sp, sn;
        bool hp = sdf(ctx, q + off, h * 2.0f, sp), hn = sdf(ctx, q + (-off), h * 2.0f, sn);
        g[a] = (hp && hn) ? (sp - sn) / h / FACTOR / 2.0f : 0.0f;
    }
    Real n = norm(g);
    if (n <= 1.0e-5f) return v2(0, 0);
    return g / n;
}
// update_single_cell, grid_update.rs:96-198 (dim2: apply_friction = true for Friction and FrictionZUp, :159-160)
static void update_single_cell(const orc_ctx& ctx, Real dt, SpkGpuGridNode& cell, V2 cell_pos, Real h, V2 gravity) {
    V2 mom = v2(cell.momentum_velocity[0], cell.momentum_velocity[1]);
    V2 v = (mom + cell.mass * gravity * dt) * inv_exact(cell.mass);
    if (cell.projection_status_tag == SPK_PROJ_NOT_COMPUTED) {
        bool have = false;
        V2 best_dir = v2(0, 0);
        Real best_dist = FLT_MAX;
        bool best_inside = false;
        size_t best_id = 0;
        for (size_t i = 0; i < ctx.colliders.size(); ++i) {
            const auto& c = ctx.colliders[i];
            if (c.grid_boundary_handling == SPK_BOUNDARY_NONE) continue;
            Projection p;
            if (!collider_project(c, cell_pos, h * 2.0f, p)) continue;
            V2 dir = cell_pos - p.point;
            Real dist = norm(dir);
            if (dist < (have ? best_dist : FLT_MAX)) { have = true; best_dir = dir; best_dist = dist; best_inside = p.is_inside; best_id = i; }
        }
        if (have) {
            cell.projection_status_tag = best_inside ? SPK_PROJ_INSIDE : SPK_PROJ_OUTSIDE;
            cell.projection_status_payload = best_id;
            for (int a = 0; a < 2; ++a) cell.projection_scaled_dir[a] = best_dir[a];
        } else {
            cell.projection_status_tag = SPK_PROJ_TOO_FAR;
        }
        V2 n = sdf_gradient(ctx, h, cell_pos);
        for (int a = 0; a < 2; ++a) cell.collision_normal[a] = n[a];
    }
    if (cell.projection_status_tag == SPK_PROJ_INSIDE || cell.projection_status_tag == SPK_PROJ_OUTSIDE) {
        const bool is_inside = cell.projection_status_tag == SPK_PROJ_INSIDE;
        const SpkGpuCollider& col = ctx.colliders[cell.projection_status_payload];
        switch (col.grid_boundary_handling) {
        case SPK_BOUNDARY_STICK:
            if (is_inside) v = v2(0, 0);
            break;
        case SPK_BOUNDARY_FRICTION:
        case SPK_BOUNDARY_FRICTION_Z_UP: {
            V2 normal = v2(cell.collision_normal[0], cell.collision_normal[1]);
            if (!is_zero(normal)) {
                Real dist = norm(v2(cell.projection_scaled_dir[0], cell.projection_scaled_dir[1]));
                Real normal_vel = dot(v, normal);
                if (normal_vel < 0.0f) {
                    Real dist_with_margin = dist - h;
                    if (is_inside || dist_with_margin <= 0.0f) {
                        V2 tangent = v - normal_vel * normal;
                        Real tn = norm(tangent);
                        v = tangent;
                        if (tn > 1.0e-10f) v = tangent / tn * rmax(tn + normal_vel * col.friction, 0.0f);
                    } else if (-normal_vel * dt > dist_with_margin) {
                        v = v - (dist_with_margin / dt + normal_vel) * normal;
                    }
                }
            }
            break;
        }
        default: break;
        }
    }
    for (int a = 0; a < 2; ++a) cell.momentum_velocity[a] = v[a];
    cell.psi_momentum_velocity *= inv_exact(cell.psi_mass);
}

static inline Real ftz(Real x) { return std::fabs(x) < FLT_MIN ? std::copysign(0.0f, x) : x; } // red.global.add.f32 flushes subnormals
static inline Real red_add(Real dst, Real src) { return ftz(ftz(dst) + ftz(src)); }

static const Real DT_FACTOR = 1000000000000.0f; // GpuTimestepLength::FACTOR, gpu_timestep.rs:18
static inline uint64_t dt_from_sec(Real dt) { return sat_u64(dt * DT_FACTOR); }
static inline Real dt_into_sec(uint64_t v) { return (Real)v / DT_FACTOR; }

} // namespace

// ================================================================================================ C API (same names as the 3D oracle)
extern "C" {

int orc_dim(void) { return 2; }
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
    else c->data.assign(n, SpkParticleData{-0.01f, 1.0f, 1.0f, 0.0f});
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
    c->collider_idx.assign(n, {});
    for (size_t i = 0; i < n; ++i)
        if (cols[i].shape_tag == SPK_SHAPE_HEIGHTFIELD) {
            const auto& hf = cols[i].shape.heightfield;
            if (!hf.heights || hf.num_heights < 2) return SPK_EINVAL;
            c->collider_data[i].assign(hf.heights, hf.heights + hf.num_heights);
            c->colliders[i].shape.heightfield.heights = c->collider_data[i].data();
        } else if (cols[i].shape_tag == SPK_SHAPE_POLYLINE) {
            const auto& pl = cols[i].shape.polyline;
            if (pl.num_vertices && !pl.vertices) return SPK_EINVAL;
            c->collider_data[i].assign(pl.vertices, pl.vertices + (size_t)pl.num_vertices * 2);
            c->colliders[i].shape.polyline.vertices = c->collider_data[i].data();
        } else if (cols[i].shape_tag == SPK_SHAPE_TRIMESH) {
            const auto& tm = cols[i].shape.trimesh;
            if (!tm.vertices || !tm.indices || tm.num_vertices == 0 || tm.num_triangles == 0) return SPK_EINVAL;
            for (size_t k = 0; k < (size_t)tm.num_triangles * 3; ++k)
                if (tm.indices[k] >= tm.num_vertices) return SPK_EINVAL;
            c->collider_data[i].assign(tm.vertices, tm.vertices + (size_t)tm.num_vertices * 2);
            c->collider_idx[i].assign(tm.indices, tm.indices + (size_t)tm.num_triangles * 3);
            c->colliders[i].shape.trimesh.vertices = c->collider_data[i].data();
            c->colliders[i].shape.trimesh.indices = c->collider_idx[i].data();
        }
    return SPK_OK;
}

int orc_frame_begin(orc_ctx* c, const SpkSolverParameters* p) { c->dt_acc = dt_from_sec(p->dt); return SPK_OK; }

// CudaSparseGrid::launch_sort, src/cuda/cuda_sparse_grid.rs:147-489 with the kernels of src_kernels/cuda/sort.rs:58-280
int orc_stage_sort(orc_ctx* c, SpkSortInfo* info) {
    const size_t n = c->pos.size();
    Grid& g = c->next_grid();
    g.clear();
    std::vector<uint64_t> pkey(n), mains;
    for (size_t i = 0; i < n; ++i) {
        uint64_t b[2];
        main_block_of_point(c->pos[i].point, c->cell_width, b);
        pkey[i] = pack_key(b[0], b[1]);
        mains.push_back(pkey[i]);
    }
    std::sort(mains.begin(), mains.end());
    mains.erase(std::unique(mains.begin(), mains.end()), mains.end());
    // touch_particle_blocks (sort.rs:58-72): the 4 blocks b + {0,1}^2 of every particle (gpu_grid.rs:293-304)
    std::vector<uint64_t> keys;
    for (uint64_t k : mains) {
        uint64_t b[2];
        unpack_key(k, b);
        for (int o = 0; o < 4; ++o) keys.push_back(pack_key(b[0] + ((o >> 1) & 1), b[1] + (o & 1)));
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
    std::vector<uint32_t> pblock(n);
    for (size_t i = 0; i < n; ++i) {
        pblock[i] = g.map.find(pkey[i])->second;
        g.headers[pblock[i]].num_particles += 1;
    }
    uint32_t run = 0;
    for (size_t i = 0; i < nab; ++i) { g.headers[i].first_particle = run; run += g.headers[i].num_particles; }
    std::vector<uint32_t> cursor(nab);
    for (size_t i = 0; i < nab; ++i) cursor[i] = g.headers[i].first_particle;
    c->sorted_ids.resize(n);
    for (size_t i = 0; i < n; ++i) c->sorted_ids[cursor[pblock[i]]++] = (uint32_t)i;
    // dispatch table (sort.rs:226-280): G2P2G_THREADS = 64 in dim2 (src_kernels/cuda/mod.rs:12-13)
    c->dispatch.clear();
    for (size_t i = 0; i < nab; ++i) {
        uint32_t np = g.headers[i].num_particles;
        uint32_t mult = np / 64 + (np % 64 > 0);
        for (uint32_t k = 0; k < mult; ++k) c->dispatch.push_back(SpkDispatchBlock2ActiveBlock{(uint32_t)i, g.headers[i].first_particle + k * 64});
    }
    g.nodes.assign(nab * CELLS, default_node());
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
        for (int k = 0; k < CELLS; ++k) {
            const SpkGpuGridNode& p = curr.nodes[(size_t)pb * CELLS + k];
            SpkGpuGridNode& q = next.nodes[b * CELLS + k];
            q.prev_mass = p.mass;
            q.projection_status_tag = p.projection_status_tag;
            q.projection_status_payload = p.projection_status_payload;
            std::memcpy(q.collision_normal, p.collision_normal, 8);
            std::memcpy(q.projection_scaled_dir, p.projection_scaled_dir, 8);
        }
    }
    return SPK_OK;
}

// estimate_timestep_length, src_kernels/cuda/timestep.rs:13-113
int orc_stage_estimate_dt(orc_ctx* c, float min_dt, float max_dt, float* dt_out) {
    const size_t n = c->pos.size();
    uint64_t acc = c->dt_acc;
    for (size_t i = 0; i < n; ++i) {
        const SpkParticleStatus& st = c->status[i];
        if (st.failed) continue;
        Real dt = max_dt;
        const SpkGpuParticleModel& model = c->models[st.model_index];
        unsigned bounds = model_active_bounds(model);
        V2 v = v2(c->vel[i].vector[0], c->vel[i].vector[1]);
        Real cand = (bounds & 1u) ? model_timestep_bound(model, c->data[i], c->volume[i], v, c->cell_width) : FLT_MAX;
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
    (void)damage_model;
    Grid& next = c->next_grid();
    const Grid& curr = c->curr_grid();
    const size_t nab = next.headers.size();
    std::vector<double> acc(nab * CELLS * 6, 0.0);
    std::vector<TileCell> tile(TILE);
    for (size_t b = 0; b < nab; ++b) {
        const SpkActiveBlockHeader& hdr = next.headers[b];
        if (hdr.num_particles == 0) continue;
        uint64_t base[2];
        unpack_key(hdr.virtual_id, base);
        int oct_next[4];
        // transfer_global_blocks_to_shared_memory, g2p2g.rs:353-474 (dim2: 4 quadrants, shared index x + 8 y)
        for (int o = 0; o < 4; ++o) {
            int ox = o & 1, oy = (o >> 1) & 1;
            uint64_t okey = pack_key(base[0] + ox, base[1] + oy);
            int pb = curr.find(okey);
            oct_next[o] = next.find(okey);
            for (int id = 0; id < CELLS; ++id) {
                int sx = id & 3, sy = id >> 2;
                TileCell& t = tile[(ox * 4 + sx) + (oy * 4 + sy) * 8];
                if (pb >= 0) {
                    const SpkGpuGridNode& gn = curr.nodes[(size_t)pb * CELLS + id];
                    t.velocity = v2(gn.momentum_velocity[0], gn.momentum_velocity[1]);
                    t.psi_velocity = gn.psi_momentum_velocity;
                } else {
                    t.velocity = v2(0, 0);
                    t.psi_velocity = 0.0f;
                }
                t.mass = 0; t.momentum[0] = t.momentum[1] = 0; t.psi_momentum = 0; t.psi_mass = 0;
            }
        }
        for (uint32_t k = 0; k < hdr.num_particles; ++k) {
            uint32_t pid = c->sorted_ids[hdr.first_particle + k];
            particle_g2p2g(*c, dt, c->status[pid], c->pos[pid], c->vel[pid], c->volume[pid], c->phase[pid], c->data[pid], tile.data());
        }
        // transfer_shared_blocks_to_grid, g2p2g.rs:476-553
        for (int o = 0; o < 4; ++o) {
            int ox = o & 1, oy = (o >> 1) & 1;
            if (oct_next[o] < 0) continue;
            double* a = &acc[(size_t)oct_next[o] * CELLS * 6];
            for (int id = 0; id < CELLS; ++id) {
                int sx = id & 3, sy = id >> 2;
                const TileCell& t = tile[(ox * 4 + sx) + (oy * 4 + sy) * 8];
                a[id * 6 + 0] += (double)ftz((Real)t.mass); a[id * 6 + 1] += (double)ftz((Real)t.momentum[0]);
                a[id * 6 + 2] += (double)ftz((Real)t.momentum[1]);
                a[id * 6 + 4] += (double)ftz((Real)t.psi_momentum); a[id * 6 + 5] += (double)ftz((Real)t.psi_mass);
            }
        }
    }
    for (size_t i = 0; i < nab * CELLS; ++i) {
        SpkGpuGridNode& nd = next.nodes[i];
        nd.mass = red_add(nd.mass, (Real)acc[i * 6 + 0]);
        nd.momentum_velocity[0] = red_add(nd.momentum_velocity[0], (Real)acc[i * 6 + 1]);
        nd.momentum_velocity[1] = red_add(nd.momentum_velocity[1], (Real)acc[i * 6 + 2]);
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
    std::vector<long long> node((size_t)n * 9, -1);
    std::vector<Real> wgt((size_t)n * 9, 0.0f);
    for (size_t i = 0; i < n; ++i) {
        auto ref = dir_to_associated_grid_node(c->pos[i].point, h);
        Real w[2][3];
        uint64_t base[2];
        int cib[2];
        main_block_of_point(c->pos[i].point, h, base);
        for (int a = 0; a < 2; ++a) {
            eval_all(-ref[a] / h, w[a]);
            int assoc_cell = sat_i32(std::round(c->pos[i].point[a] / h)) - 2; // particle.rs:28-35
            int assoc_block = sat_i32(std::floor((Real)assoc_cell / 4.0f)) * 4;
            cib[a] = assoc_cell - assoc_block;
        }
        for (int k = 0; k < 9; ++k) {
            const int* s = NBH[k];
            int t[2];
            for (int a = 0; a < 2; ++a) t[a] = cib[a] + 1 + s[a];
            int b = next.find(pack_key(base[0] + (uint64_t)(t[0] >> 2), base[1] + (uint64_t)(t[1] >> 2)));
            if (b < 0) continue; // cannot happen after touch_particle_blocks
            node[i * 9 + k] = (long long)b * 16 + (t[0] & 3) * 1 + (t[1] & 3) * 4;
            wgt[i * 9 + k] = w[0][s[0]] * w[1][s[1]];
            acc[(size_t)node[i * 9 + k]] += (double)(wgt[i * 9 + k] * c->volume[i].mass); // fluids_volume.rs:40-47
        }
    }
    for (size_t i = 0; i < n; ++i) {
        if (!model_is_fluid(c->models[c->status[i].model_index])) continue; // :70-77
        Real new_mass = 0.0f;
        for (int k = 0; k < 9; ++k)
            if (node[i * 9 + k] >= 0) new_mass += wgt[i * 9 + k] * ftz((Real)acc[(size_t)node[i * 9 + k]]);
        Real new_density = new_mass / (h * h); // cell_width.powi(DIM)
        Real new_volume = c->volume[i].mass / new_density;
        c->volume[i].deformation_gradient[0] = new_volume / c->volume[i].volume0;
    }
    return SPK_OK; // the next grid's masses were never written: "Reset the modified cells" (:98-113) is implicit
}

// grid_update + swap_buffers, src_kernels/cuda/grid_update.rs:10-52, cuda_mpm_pipeline.rs:536-558
int orc_stage_grid_update(orc_ctx* c, float dt, const float gravity[2]) {
    Grid& next = c->next_grid();
    const V2 g = v2(gravity[0], gravity[1]);
    for (size_t b = 0; b < next.headers.size(); ++b) {
        uint64_t k[2];
        unpack_key(next.headers[b].virtual_id, k);
        for (int id = 0; id < CELLS; ++id) {
            int sx = id & 3, sy = id >> 2;
            int64_t ci[2] = {((int64_t)k[0] - (int64_t)PACK_ORIGIN) * 4 + sx, ((int64_t)k[1] - (int64_t)PACK_ORIGIN) * 4 + sy};
            V2 cell_pos = v2((Real)ci[0] * c->cell_width, (Real)ci[1] * c->cell_width);
            update_single_cell(*c, dt, next.nodes[b * CELLS + id], cell_pos, c->cell_width, g);
        }
    }
    c->curr = 1 - c->curr;
    return SPK_OK;
}

// CudaMpmPipeline::step_generic, src/cuda/cuda_mpm_pipeline.rs:273-659 (single context)
int orc_step(orc_ctx* c, const SpkSolverParameters* params, const float gravity[2], SpkStepTimings* timings, float* simulated_time_out) {
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
    std::memcpy(dst, g.nodes.data(), (size_t)n_blocks * CELLS * sizeof(SpkGpuGridNode));
    return SPK_OK;
}
int orc_grid_read_dispatch(orc_ctx* c, SpkDispatchBlock2ActiveBlock* dst, uint32_t n) {
    if (n > c->dispatch.size()) return SPK_EINVAL;
    std::memcpy(dst, c->dispatch.data(), n * sizeof(SpkDispatchBlock2ActiveBlock));
    return SPK_OK;
}

// ---- unit hooks used by tests/
void orc_svd2(const float* f_colmajor, float* u, float* s, float* vt) {
    M2 F; std::memcpy(F.m, f_colmajor, 16);
    Svd2 r = svd_unordered(F);
    std::memcpy(u, r.u.m, 16); std::memcpy(vt, r.vt.m, 16); s[0] = r.s.x; s[1] = r.s.y;
}
uint64_t orc_block_key_of_point(const float* p, float cell_width) { uint64_t b[2]; main_block_of_point(p, cell_width, b); return pack_key(b[0], b[1]); }
void orc_quadratic_weights(float x, float* w3) { eval_all(x, w3); }
uint64_t orc_dt_from_sec(float dt) { return dt_from_sec(dt); }
float orc_dt_into_sec(uint64_t v) { return dt_into_sec(v); }
void orc_kirchhoff_stress(const SpkGpuParticleModel* m, const SpkParticleData* d, const SpkParticleVolume* vol, float phase, const float* vel_grad, float* out4) {
    M2 g; std::memcpy(g.m, vel_grad, 16);
    M2 s = model_kirchhoff_stress(*m, *d, *vol, phase, g);
    std::memcpy(out4, s.m, 16);
}
void orc_plastic_update(const SpkGpuParticleModel* m, SpkParticleData* d, SpkParticleVolume* vol, float phase) { model_plastic_update(*m, *d, *vol, phase); }
int orc_collider_project(const SpkGpuCollider* c, const float* pt, float max_dist, float* out_pt, int* inside) {
    Projection p;
    if (!collider_project(*c, v2(pt[0], pt[1]), max_dist, p)) return 0;
    out_pt[0] = p.point.x; out_pt[1] = p.point.y;
    *inside = p.is_inside ? 1 : 0;
    return 1;
}

} // extern "C"
