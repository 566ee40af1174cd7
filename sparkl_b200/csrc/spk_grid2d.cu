// spk_grid2d.cu — dim2 build of spk_grid.cu: grid reset + collider-cache carry (K13/K14, src_kernels/cuda/reset_grid.rs:8-60),
// grid momentum update with parry2d cuboid / polyline SDF colliders and boundary conditions (K19,
// src_kernels/cuda/grid_update.rs:10-198 with the `dim2` feature), grid export in the reference's 56-byte dim2 node layout, and
// the halo kernels of the portable multi-process path (K3/K4/K17/K18, sort.rs:74-164).  Compiled only into libsparkl_b200_2d.so.
#include <cfloat>

#include "spk_device.cuh"
#include "spk_math2d.cuh"

#if SPK_DIM != 2
#error "spk_grid2d.cu is the dim2 build (compile with -DSPK_DIM=2)"
#endif

namespace {

constexpr int CELLS = SPK_CELLS_PER_BLOCK;               // 16
constexpr int ACC_BLOCK = SPK_ACC_CHANNELS * CELLS;      // floats per block in `acc`
constexpr int BLOCKS_PER_CTA = 256 / CELLS;              // 16 blocks per 256-thread CTA, one thread per cell

// the reference's `red.global.add.f32` flushes subnormals; the TMA bulk reduction does not: every consumer flushes (see spk_grid.cu)
__device__ __forceinline__ float spk_ftz(float x) { return fabsf(x) < FLT_MIN ? copysignf(0.f, x) : x; }

// ---- reset + carry -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reset_grid_kernel(GridBuffers next, GridBuffers prev, uint32_t nab) {
    const uint32_t b = blockIdx.x * BLOCKS_PER_CTA + (threadIdx.x >> SPK_CELL_SHIFT);
    const uint32_t cell = threadIdx.x & (CELLS - 1);
    if (b >= nab) return;
    float* acc = next.acc + (size_t)b * ACC_BLOCK;
#pragma unroll
    for (int ch = 0; ch < SPK_ACC_CHANNELS; ++ch) acc[ch * CELLS + cell] = 0.f; // GpuGridNode::default (reset_grid.rs:24-26)
    const int32_t pb = next.nbr_prev[b * SPK_NBR + 0];                          // same key in the previous grid (reset_grid.rs:46)
    const size_t node = (size_t)b * CELLS + cell;
    if (pb >= 0) {
        const size_t pnode = (size_t)pb * CELLS + cell;
        float4 n = prev.cacheN[pnode];
        n.w = spk_ftz(prev.acc[(size_t)pb * ACC_BLOCK + cell]); // prev_mass = prev.mass (reset_grid.rs:54)
        next.cacheN[node] = n;
        next.cacheD[node] = prev.cacheD[pnode];
        next.proj[node] = prev.proj[pnode];
    } else {
        next.cacheN[node] = make_float4(0.f, 0.f, 0.f, 0.f);
        next.cacheD[node] = make_float4(0.f, 0.f, 0.f, 0.f);
        next.proj[node] = SPK_PROJ_NOT_COMPUTED;
    }
}

// ---- colliders (parry2d 0.13) ---------------------------------------------------------------------------------------
struct Vec2 { float x, y; };
__device__ __forceinline__ Vec2 rot(const float c[4], Vec2 v) { return Vec2{c[0] * v.x - c[1] * v.y, c[1] * v.x + c[0] * v.y}; }      // UnitComplex * v
__device__ __forceinline__ Vec2 rot_inv(const float c[4], Vec2 v) { return Vec2{c[0] * v.x + c[1] * v.y, -c[1] * v.x + c[0] * v.y}; } // inverse
// Aabb::project_local_point with solid = false through Cuboid::project_point (gpu_collider.rs:53)
__device__ inline void cuboid_project(const DevCollider& c, Vec2 pt, Vec2& out, bool& inside) {
    const Vec2 lp = rot_inv(c.rot, Vec2{pt.x - c.tr[0], pt.y - c.tr[1]});
    const float l[2] = {lp.x, lp.y};
    float mins_pt[2], pt_maxs[2], shift[2];
    bool zero = true;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        mins_pt[a] = -c.half_extents[a] - l[a];
        pt_maxs[a] = l[a] - c.half_extents[a];
        shift[a] = fmaxf(mins_pt[a], 0.f) - fmaxf(pt_maxs[a], 0.f);
        zero = zero && shift[a] == 0.f;
    }
    float r[2] = {l[0], l[1]};
    if (!zero) {
        inside = false;
        r[0] += shift[0]; r[1] += shift[1];
    } else {
        float best = -FLT_MAX;
        bool is_mins = false;
        int best_id = 0;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            if (mins_pt[a] < pt_maxs[a]) {
                if (pt_maxs[a] > best) { best_id = a; is_mins = false; best = pt_maxs[a]; }
            } else if (mins_pt[a] > best) { best_id = a; is_mins = true; best = mins_pt[a]; }
        }
        inside = true;
#pragma unroll
        for (int a = 0; a < 2; ++a) if (a == best_id) r[a] += is_mins ? best : -best;
    }
    const Vec2 w = rot(c.rot, Vec2{r[0], r[1]});
    out = Vec2{w.x + c.tr[0], w.y + c.tr[1]};
}
// polyline_project_point, src_kernels/gpu_collider.rs:105-201 (dim2: normal1 = Segment::normal()); the inside tests subtract the
// LOCAL projection from the WORLD point as the reference does (:176,181)
__device__ inline bool polyline_project(const DevCollider& c, Vec2 pt, Vec2& out, bool& inside) {
    const int n = (int)c.nrows;
    if (n == 0) return false;
    const Vec2 lp = rot_inv(c.rot, Vec2{pt.x - c.tr[0], pt.y - c.tr[1]});
    const float* v = c.heights;
    float best_dist = FLT_MAX;
    Vec2 best{0.f, 0.f}, sa{0.f, 0.f}, sb{0.f, 0.f};
    int best_id = 0, best_loc = 2; // 0 / 1: on vertex a / b, 2: on the edge
    bool best_eq = false;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1) % n;
        const Vec2 a{v[2 * i], v[2 * i + 1]}, b{v[2 * j], v[2 * j + 1]};
        const Vec2 ab{b.x - a.x, b.y - a.y}, ap{lp.x - a.x, lp.y - a.y};
        const float ab_ap = ab.x * ap.x + ab.y * ap.y, sqnab = ab.x * ab.x + ab.y * ab.y;
        Vec2 q; int loc;
        if (ab_ap <= 0.f) { q = a; loc = 0; }
        else if (ab_ap >= sqnab) { q = b; loc = 1; }
        else { const float u = ab_ap / sqnab; q = Vec2{a.x + ab.x * u, a.y + ab.y * u}; loc = 2; }
        const float dx = lp.x - q.x, dy = lp.y - q.y;
        const float dist = sqrtf(dx * dx + dy * dy);
        if (i == 0 || dist < best_dist) { best_dist = dist; best = q; best_id = i; best_loc = loc; sa = a; sb = b; best_eq = q.x == lp.x && q.y == lp.y; }
    }
    inside = best_eq;
    const Vec2 dir{sb.x - sa.x, sb.y - sa.y};
    const float nn = sqrtf(dir.y * dir.y + dir.x * dir.x);
    if (nn > FLT_EPSILON) {
        const Vec2 nrm{dir.y / nn, -dir.x / nn};
        const float side = (pt.x - best.x) * nrm.x + (pt.y - best.y) * nrm.y;
        if (best_loc == 2) {
            inside = side <= 0.f;
        } else {
            Vec2 dir2;
            if (best_loc == 0) {
                const int adj = best_id == 0 ? n - 1 : best_id - 1, aj = (adj + 1) % n;
                dir2 = Vec2{-(v[2 * aj] - v[2 * adj]), -(v[2 * aj + 1] - v[2 * adj + 1])};
            } else {
                const int adj = (best_id + 1) % n, aj = (adj + 1) % n;
                dir2 = Vec2{v[2 * aj] - v[2 * adj], v[2 * aj + 1] - v[2 * adj + 1]};
            }
            const float dot = nrm.x * dir2.x + nrm.y * dir2.y;
            const float threshold = 1.0e-3f * sqrtf(dir2.x * dir2.x + dir2.y * dir2.y);
            inside = fabsf(dot) > threshold ? dot >= 0.f : side <= 0.f;
        }
    }
    if (c.hf_flags & 2u) inside = !inside;
    const Vec2 w = rot(c.rot, best);
    out = Vec2{w.x + c.tr[0], w.y + c.tr[1]};
    return true;
}
// heightfield projection: parry2d 0.13 `GenericHeightField::project_local_point_with_max_dist` over `map_elements_in_local_aabb`
// (heightfield2.rs: the segments whose x range meets [p.x - max_dist, p.x + max_dist]), restated; then the inside test of
// GpuColliderShape::project_point_with_max_dist (src_kernels/gpu_collider.rs:73-99)
__device__ inline bool heightfield_project(const DevCollider& c, Vec2 pt, float max_dist, Vec2& out, bool& inside) {
    const Vec2 lp = rot_inv(c.rot, Vec2{pt.x - c.tr[0], pt.y - c.tr[1]});
    const int cells = (int)c.nrows - 1;
    const float seg = 1.f / (float)cells;
    const float rminx = (lp.x - max_dist) / c.scale[0], rmaxx = (lp.x + max_dist) / c.scale[0];
    if (rmaxx <= -0.5f || rminx >= 0.5f) return false; // outside of the heightfield bounds
    const int min_x = (int)fminf(fmaxf(floorf((rminx + 0.5f) / seg), 0.f), (float)(cells - 1));
    const int max_x = (int)fminf(fmaxf(ceilf((rmaxx + 0.5f) / seg), 0.f), (float)cells);
    float sq_smallest = FLT_MAX;
    bool found = false;
    Vec2 best{0.f, 0.f};
    for (int i = min_x; i < max_x; ++i) {
        const Vec2 a{(-0.5f + seg * (float)i) * c.scale[0], c.heights[i] * c.scale[1]};
        const Vec2 b{(-0.5f + seg * (float)(i + 1)) * c.scale[0], c.heights[i + 1] * c.scale[1]};
        const Vec2 ab{b.x - a.x, b.y - a.y}, ap{lp.x - a.x, lp.y - a.y};
        const float ab_ap = ab.x * ap.x + ab.y * ap.y, sqnab = ab.x * ab.x + ab.y * ab.y;
        Vec2 q; // Segment::project_local_point
        if (ab_ap <= 0.f) q = a;
        else if (ab_ap >= sqnab) q = b;
        else { const float u = ab_ap / sqnab; q = Vec2{a.x + ab.x * u, a.y + ab.y * u}; }
        const float dx = lp.x - q.x, dy = lp.y - q.y;
        const float sq = dx * dx + dy * dy;
        if (sq < sq_smallest) {
            sq_smallest = sq;
            if (sqrtf(sq) <= max_dist) { best = q; found = true; }
        }
    }
    if (!found) return false;
    if (c.hf_flags & 1u) { // bellow_heightfield_is_solid: within the x footprint and on the solid side of the surface
        const bool in_footprint = lp.x >= -0.5f * c.scale[0] && lp.x <= 0.5f * c.scale[0];
        inside = in_footprint && ((c.hf_flags & 2u) ? lp.y >= best.y : lp.y <= best.y);
    } else {
        inside = best.x == lp.x && best.y == lp.y;
    }
    const Vec2 w = rot(c.rot, best);
    out = Vec2{w.x + c.tr[0], w.y + c.tr[1]};
    return true;
}
// parry2d 0.13 Triangle::project_local_point_and_get_location (query/point/point_triangle.rs, `dim2` paths) with solid = false:
// vertex / edge Voronoi regions as in 3D (perp products instead of cross products); a point inside the triangle is projected onto
// the nearest of the three edge LINES (squared distances by Pythagoras) and flagged inside.  Restated from the published algorithm.
__device__ inline void triangle_project2(Vec2 a, Vec2 b, Vec2 c, Vec2 p, Vec2& q, bool& inside) {
    const Vec2 ab{b.x - a.x, b.y - a.y}, ac{c.x - a.x, c.y - a.y}, ap{p.x - a.x, p.y - a.y};
    const float ab_ap = ab.x * ap.x + ab.y * ap.y, ac_ap = ac.x * ap.x + ac.y * ap.y;
    inside = false;
    if (ab_ap <= 0.f && ac_ap <= 0.f) { q = a; inside = p.x == a.x && p.y == a.y; return; }
    const Vec2 bp{p.x - b.x, p.y - b.y};
    const float ab_bp = ab.x * bp.x + ab.y * bp.y, ac_bp = ac.x * bp.x + ac.y * bp.y;
    if (ab_bp >= 0.f && ac_bp <= ab_bp) { q = b; inside = p.x == b.x && p.y == b.y; return; }
    const Vec2 cp{p.x - c.x, p.y - c.y};
    const float ab_cp = ab.x * cp.x + ab.y * cp.y, ac_cp = ac.x * cp.x + ac.y * cp.y;
    if (ac_cp >= 0.f && ab_cp <= ac_cp) { q = c; inside = p.x == c.x && p.y == c.y; return; }
    const Vec2 bc{c.x - b.x, c.y - b.y};
    const float n = ab.x * ac.y - ab.y * ac.x;
    const float sq_ab = ab.x * ab.x + ab.y * ab.y, sq_ac = ac.x * ac.x + ac.y * ac.y, sq_bc = bc.x * bc.x + bc.y * bc.y;
    const float vc = n * (ab.x * ap.y - ab.y * ap.x);
    if (vc < 0.f && ab_ap >= 0.f && ab_bp <= 0.f) { const float v = ab_ap / sq_ab; q = Vec2{a.x + ab.x * v, a.y + ab.y * v}; inside = q.x == p.x && q.y == p.y; return; }
    const float vb = -n * (ac.x * cp.y - ac.y * cp.x);
    if (vb < 0.f && ac_ap >= 0.f && ac_cp <= 0.f) { const float w = ac_ap / sq_ac; q = Vec2{a.x + ac.x * w, a.y + ac.y * w}; inside = q.x == p.x && q.y == p.y; return; }
    const float va = n * (bc.x * bp.y - bc.y * bp.x);
    if (va < 0.f && ac_bp - ab_bp >= 0.f && ab_cp - ac_cp >= 0.f) {
        const float w = (bc.x * bp.x + bc.y * bp.y) / sq_bc;
        q = Vec2{b.x + bc.x * w, b.y + bc.y * w}; inside = q.x == p.x && q.y == p.y; return;
    }
    // inside the (solid) triangle, solid = false: the closest edge
    const float v = ab_ap / (ab_ap - ab_bp), w = ac_ap / (ac_ap - ac_cp), u = (ac_bp - ab_bp) / (ac_bp - ab_bp + ab_cp - ac_cp);
    const float sq_ap = ap.x * ap.x + ap.y * ap.y, sq_bp = bp.x * bp.x + bp.y * bp.y;
    const float d_ab = sq_ap - sq_ab * v * v, d_ac = sq_ap - sq_ac * w * w, d_bc = sq_bp - sq_bc * u * u;
    inside = true;
    if (d_ab < d_ac) {
        if (d_ab < d_bc) q = Vec2{a.x + ab.x * v, a.y + ab.y * v};
        else q = Vec2{b.x + bc.x * u, b.y + bc.y * u};
    } else {
        if (d_ac < d_bc) q = Vec2{a.x + ac.x * w, a.y + ac.y * w};
        else q = Vec2{b.x + bc.x * u, b.y + bc.y * u};
    }
}
// TriMesh::project_point_with_max_dist(.., Real::MAX) (gpu_collider.rs:58-68): the closest triangle over the whole mesh (parry's QBVH
// only prunes; ties go to the lowest triangle index here), flip_interior applied afterwards
__device__ inline bool trimesh_project(const DevCollider& c, Vec2 pt, Vec2& out, bool& inside) {
    const Vec2 lp = rot_inv(c.rot, Vec2{pt.x - c.tr[0], pt.y - c.tr[1]});
    const float* verts = c.heights;
    const float* idx = c.heights + (size_t)c.nrows * 2;
    float best_sq = FLT_MAX;
    Vec2 best{0.f, 0.f};
    bool best_in = false, found = false;
    for (uint32_t t = 0; t < c.ncols; ++t) {
        const uint32_t i0 = __float_as_uint(idx[3 * t]), i1 = __float_as_uint(idx[3 * t + 1]), i2 = __float_as_uint(idx[3 * t + 2]);
        Vec2 q; bool in;
        triangle_project2(Vec2{verts[2 * i0], verts[2 * i0 + 1]}, Vec2{verts[2 * i1], verts[2 * i1 + 1]}, Vec2{verts[2 * i2], verts[2 * i2 + 1]}, lp, q, in);
        const float dx = lp.x - q.x, dy = lp.y - q.y;
        const float sq = dx * dx + dy * dy;
        if (sq < best_sq) { best_sq = sq; best = q; best_in = in; found = true; }
    }
    if (!found) return false;
    inside = (c.hf_flags & 2u) ? !best_in : best_in;
    const Vec2 w = rot(c.rot, best);
    out = Vec2{w.x + c.tr[0], w.y + c.tr[1]};
    return true;
}
__device__ __forceinline__ bool collider_project(const DevCollider& c, Vec2 pt, float max_dist, Vec2& out, bool& inside) {
    if (c.shape == SPK_SHAPE_TRIMESH) return trimesh_project(c, pt, out, inside);
    if (c.shape == SPK_SHAPE_HEIGHTFIELD) return heightfield_project(c, pt, max_dist, out, inside);
    if (c.shape == SPK_SHAPE_POLYLINE) return polyline_project(c, pt, out, inside);
    cuboid_project(c, pt, out, inside);
    return true;
}
// sdf, grid_update.rs:54-71
__device__ inline bool sdf(const DevCollider* cols, uint32_t ncol, Vec2 q, float max_dist, float& out) {
    bool any = false;
    float best = 0.f;
    for (uint32_t i = 0; i < ncol; ++i) {
        if (cols[i].handling == SPK_BOUNDARY_NONE) continue;
        Vec2 p; bool inside;
        if (!collider_project(cols[i], q, max_dist, p, inside)) continue;
        const float dx = p.x - q.x, dy = p.y - q.y;
        const float d = sqrtf(dx * dx + dy * dy) * (inside ? -1.f : 1.f);
        if (!any || d < best) { best = d; any = true; }
    }
    out = best;
    return any;
}

__global__ void __launch_bounds__(256) grid_update_kernel(GridBuffers g, uint32_t nab, const DevCollider* __restrict__ cols, uint32_t ncol, float h, float dt,
                                                           float gx, float gy) {
    const uint32_t b = blockIdx.x * BLOCKS_PER_CTA + (threadIdx.x >> SPK_CELL_SHIFT);
    const uint32_t cell = threadIdx.x & (CELLS - 1);
    if (b >= nab) return;
    const float* acc = g.acc + (size_t)b * ACC_BLOCK;
    const size_t node = (size_t)b * CELLS + cell;
    const float mass = spk_ftz(acc[cell]);
    const float psi_mass = spk_ftz(acc[5 * CELLS + cell]);
    const float inv_m = inv_exact(mass);
    // (momentum + mass * gravity * dt) * inv_exact(mass), grid_update.rs:104-105
    float vx = (spk_ftz(acc[1 * CELLS + cell]) + mass * gx * dt) * inv_m;
    float vy = (spk_ftz(acc[2 * CELLS + cell]) + mass * gy * dt) * inv_m;
    uint32_t proj = g.proj[node];
    float4 cn = g.cacheN[node];
    float4 cd = g.cacheD[node];
    if (ncol == 0) {
        if ((proj & 3u) == SPK_PROJ_NOT_COMPUTED) { proj = SPK_PROJ_TOO_FAR; g.proj[node] = proj; }
    } else if ((proj & 3u) == SPK_PROJ_NOT_COMPUTED) {
        unsigned long long kx, ky;
        spk_unpack_key(g.blk_key[b], kx, ky);
        const long long cx = ((long long)kx - SPK_PACK_ORIGIN) * 4 + (cell & 3), cy = ((long long)ky - SPK_PACK_ORIGIN) * 4 + (cell >> 2);
        // grid_update.rs:34-38.  __fmul_rn: the product must be ROUNDED before the collider code subtracts from it, as in the
        // reference's PTX (the node position is a plain mul.f32 reused by every subtraction); a fused cy * h - y_surface decides
        // the `dist - h <= 0` test of the friction branch differently for nodes exactly one cell above an axis-aligned surface
        const Vec2 pos{__fmul_rn((float)cx, h), __fmul_rn((float)cy, h)};
        bool have = false, best_inside = false;
        float best_dist = FLT_MAX;
        Vec2 best_dir{0.f, 0.f};
        uint32_t best_id = 0;
        for (uint32_t i = 0; i < ncol; ++i) {
            if (cols[i].handling == SPK_BOUNDARY_NONE) continue;
            Vec2 p; bool inside;
            if (!collider_project(cols[i], pos, h * 2.f, p, inside)) continue; // max_dist = cell_width * 2, grid_update.rs:118
            const Vec2 dir{pos.x - p.x, pos.y - p.y};
            const float dist = sqrtf(dir.x * dir.x + dir.y * dir.y);
            if (dist < (have ? best_dist : FLT_MAX)) { have = true; best_dist = dist; best_dir = dir; best_inside = inside; best_id = i; }
        }
        if (have) {
            proj = (best_inside ? SPK_PROJ_INSIDE : SPK_PROJ_OUTSIDE) | (best_id << 2);
            cd = make_float4(best_dir.x, best_dir.y, 0.f, 0.f);
        } else {
            proj = SPK_PROJ_TOO_FAR;
        }
        // sdf_gradient, grid_update.rs:74-94: central differences at +-0.1 h along x and y
        const float off = 1.f * (h * 0.1f);
        float grad[2];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            Vec2 qp = pos, qn = pos;
            if (a == 0) { qp.x += off; qn.x += -off; } else { qp.y += off; qn.y += -off; }
            float sp, sn;
            const bool hp = sdf(cols, ncol, qp, h * 2.f, sp), hn = sdf(cols, ncol, qn, h * 2.f, sn);
            grad[a] = (hp && hn) ? (sp - sn) / h / 0.1f / 2.f : 0.f;
        }
        const float gn = sqrtf(grad[0] * grad[0] + grad[1] * grad[1]);
        if (gn <= 1.0e-5f) cn.x = cn.y = 0.f; // try_normalize(1e-5).unwrap_or(zeros)
        else { cn.x = grad[0] / gn; cn.y = grad[1] / gn; }
        cn.z = 0.f;
        g.proj[node] = proj;
        g.cacheN[node] = cn;
        g.cacheD[node] = cd;
    }
    const uint32_t tag = proj & 3u;
    if (tag == SPK_PROJ_INSIDE || tag == SPK_PROJ_OUTSIDE) { // grid_update.rs:142-194
        const bool is_inside = tag == SPK_PROJ_INSIDE;
        const DevCollider& col = cols[proj >> 2];
        if (col.handling == SPK_BOUNDARY_STICK) {
            if (is_inside) vx = vy = 0.f;
        } else if (col.handling == SPK_BOUNDARY_FRICTION || col.handling == SPK_BOUNDARY_FRICTION_Z_UP) { // same in 2D (:159-160)
            if (!(cn.x == 0.f && cn.y == 0.f)) {
                const float dist = sqrtf(cd.x * cd.x + cd.y * cd.y);
                const float normal_vel = vx * cn.x + vy * cn.y;
                if (normal_vel < 0.f) {
                    const float dist_with_margin = dist - h;
                    if (is_inside || dist_with_margin <= 0.f) {
                        const float tx = vx - normal_vel * cn.x, ty = vy - normal_vel * cn.y;
                        const float tn = sqrtf(tx * tx + ty * ty);
                        vx = tx; vy = ty;
                        if (tn > 1.0e-10f) {
                            const float k = fmaxf(tn + normal_vel * col.friction, 0.f);
                            vx = tx / tn * k; vy = ty / tn * k;
                        }
                    } else if (-normal_vel * dt > dist_with_margin) {
                        const float k = dist_with_margin / dt + normal_vel;
                        vx -= k * cn.x; vy -= k * cn.y;
                    }
                }
            }
        }
    }
    const float psi_v = spk_ftz(acc[4 * CELLS + cell]) * inv_exact(psi_mass); // grid_update.rs:197
    g.vel[node] = make_float4(vx, vy, 0.f, psi_v);
}

// ---- export to the reference's dim2 GpuGridNode layout ------------------------------------------------------------
__global__ void grid_export_kernel(GridBuffers g, uint32_t nblocks, SpkGpuGridNode* out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nblocks * CELLS) return;
    const uint32_t b = t >> SPK_CELL_SHIFT, cell = t & (CELLS - 1);
    const float* acc = g.acc + (size_t)b * ACC_BLOCK;
    SpkGpuGridNode n;
    n.mass = spk_ftz(acc[cell]);
    n.psi_mass = spk_ftz(acc[5 * CELLS + cell]);
    if (g.has_velocity) {
        const float4 v = g.vel[t];
        n.momentum_velocity[0] = v.x; n.momentum_velocity[1] = v.y;
        n.psi_momentum_velocity = v.w;
    } else {
        n.momentum_velocity[0] = spk_ftz(acc[1 * CELLS + cell]); n.momentum_velocity[1] = spk_ftz(acc[2 * CELLS + cell]);
        n.psi_momentum_velocity = spk_ftz(acc[4 * CELLS + cell]);
    }
    const float4 cn = g.cacheN[t], cd = g.cacheD[t];
    const uint32_t proj = g.proj[t];
    n.prev_mass = cn.w;
    n.projection_status_tag = proj & 3u;
    n.projection_status_payload = ((proj & 3u) == SPK_PROJ_INSIDE || (proj & 3u) == SPK_PROJ_OUTSIDE) ? (proj >> 2) : 0;
    n.collision_normal[0] = cn.x; n.collision_normal[1] = cn.y;
    n.projection_scaled_dir[0] = cd.x; n.projection_scaled_dir[1] = cd.y;
    out[t] = n;
}

#if defined(__CUDACC__) || defined(SPK_EMU) // everything below needs the device hash map / launch syntax; the per-node kernels above are also host-compiled by tests/host/
// ---- halo path (portable multi-process protocol, same contract as the 3D build with 96 floats per block) ---------------
__global__ void halo_keys_kernel(GridBuffers g, uint32_t nab, unsigned long long* out, uint32_t capacity) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < capacity) out[b] = b < nab ? g.blk_key[b] : SPK_EMPTY_KEY;
}
__global__ void halo_tag_kernel(GridBuffers g, const unsigned long long* __restrict__ remote, uint32_t n, DeviceCounters* counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (remote[i] == SPK_EMPTY_KEY) return;
    const int32_t b = spk_hash_get(g, remote[i]);
    if (b >= 0 && atomicExch(&g.blk_halo[b], 1u) == 0u) atomicAdd(&counters->num_halo, 1u);
}
__global__ void __launch_bounds__(CELLS) halo_pack_kernel(GridBuffers g, uint32_t nab, unsigned long long* keys, float* payload, uint32_t capacity, uint32_t* cursor) {
    const uint32_t b = blockIdx.x;
    if (b >= nab || !(g.blk_halo[b] & 1u)) return;
    __shared__ uint32_t s_slot;
    if (threadIdx.x == 0) s_slot = atomicAdd(cursor, 1u);
    __syncthreads();
    const uint32_t slot = s_slot;
    if (slot >= capacity) return;
    if (threadIdx.x == 0) keys[slot] = g.blk_key[b];
    const float* acc = g.acc + (size_t)b * ACC_BLOCK;
    float* dst = payload + (size_t)slot * ACC_BLOCK;
#pragma unroll
    for (int ch = 0; ch < SPK_ACC_CHANNELS; ++ch) dst[ch * CELLS + threadIdx.x] = acc[ch * CELLS + threadIdx.x];
}
__global__ void __launch_bounds__(CELLS) halo_merge_kernel(GridBuffers g, const unsigned long long* __restrict__ keys, const float* __restrict__ payload) {
    if (keys[blockIdx.x] == SPK_EMPTY_KEY) return;
    const int32_t b = spk_hash_get(g, keys[blockIdx.x]);
    if (b < 0) return;
    float* acc = g.acc + (size_t)b * ACC_BLOCK;
    const float* src = payload + (size_t)blockIdx.x * ACC_BLOCK;
#pragma unroll
    for (int ch = 0; ch < SPK_ACC_CHANNELS; ++ch) atomicAdd(&acc[ch * CELLS + threadIdx.x], src[ch * CELLS + threadIdx.x]);
}

#endif // __CUDACC__

} // namespace

#if defined(__CUDACC__) || defined(SPK_EMU)
cudaError_t spk_launch_reset_grid(spk_ctx* c) {
    GridBuffers& next = c->grid[1 - c->gcur];
    const uint32_t nab = next.nab_host;
    next.has_velocity = 0;
    if (!nab) return cudaSuccess;
    reset_grid_kernel<<<(nab + BLOCKS_PER_CTA - 1) / BLOCKS_PER_CTA, 256, 0, c->stream>>>(next, c->grid[c->gcur], nab);
    c->launches += 1;
    return cudaGetLastError();
}

cudaError_t spk_launch_grid_update(spk_ctx* c, float dt, const float gravity[SPK_DIM]) {
    GridBuffers& next = c->grid[1 - c->gcur];
    const uint32_t nab = next.nab_host;
    if (nab) {
        grid_update_kernel<<<(nab + BLOCKS_PER_CTA - 1) / BLOCKS_PER_CTA, 256, 0, c->stream>>>(next, nab, c->colliders, (uint32_t)c->ncolliders, c->cell_width, dt,
                                                                                              gravity[0], gravity[1]);
        c->launches += 1;
    }
    next.has_velocity = 1;
    c->gcur = 1 - c->gcur; // CudaSparseGrid::swap_buffers, cuda_sparse_grid.rs:108-110
    return cudaGetLastError();
}

cudaError_t spk_launch_grid_export(spk_ctx* c, int which_grid, SpkGpuGridNode* d_out, uint32_t nblocks) {
    GridBuffers& g = which_grid ? c->grid[1 - c->gcur] : c->grid[c->gcur];
    if (!nblocks) return cudaSuccess;
    grid_export_kernel<<<(nblocks * CELLS + 255) / 256, 256, 0, c->stream>>>(g, nblocks, d_out);
    return cudaGetLastError();
}

cudaError_t spk_launch_halo_keys(spk_ctx* c, unsigned long long* out, uint32_t capacity) {
    GridBuffers& g = c->grid[1 - c->gcur];
    if (!capacity) return cudaSuccess;
    halo_keys_kernel<<<(capacity + 255) / 256, 256, 0, c->stream>>>(g, g.nab_host, out, capacity);
    c->launches += 1;
    return cudaGetLastError();
}
cudaError_t spk_launch_halo_tag(spk_ctx* c, const unsigned long long* remote, uint32_t n) {
    GridBuffers& g = c->grid[1 - c->gcur];
    if (!n) return cudaSuccess;
    c->counters_fresh = false;
    halo_tag_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(g, remote, n, c->d_counters);
    c->launches += 1;
    return cudaGetLastError();
}
cudaError_t spk_launch_halo_pack(spk_ctx* c, unsigned long long* keys, float* payload, uint32_t capacity) {
    GridBuffers& g = c->grid[1 - c->gcur];
    cudaError_t e = cudaMemsetAsync(&c->d_counters->halo_cursor, 0, 4, c->stream);
    if (e == cudaSuccess && capacity) e = cudaMemsetAsync(keys, 0xff, (size_t)capacity * 8, c->stream);
    if (e != cudaSuccess || !g.nab_host) return e;
    c->counters_fresh = false;
    halo_pack_kernel<<<g.nab_host, CELLS, 0, c->stream>>>(g, g.nab_host, keys, payload, capacity, &c->d_counters->halo_cursor);
    c->launches += 1;
    return cudaGetLastError();
}
cudaError_t spk_launch_halo_merge(spk_ctx* c, const unsigned long long* keys, const float* payload, uint32_t n) {
    GridBuffers& g = c->grid[1 - c->gcur];
    if (!n) return cudaSuccess;
    halo_merge_kernel<<<n, CELLS, 0, c->stream>>>(g, keys, payload);
    c->launches += 1;
    return cudaGetLastError();
}
#endif // __CUDACC__
