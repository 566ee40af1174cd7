// spk_grid.cu — grid reset + collider-cache carry (K13/K14, src_kernels/cuda/reset_grid.rs:8-60), grid momentum update
// with cuboid SDF colliders and boundary conditions (K19, src_kernels/cuda/grid_update.rs:10-198), grid export in the
// reference's 72-byte node layout, and the halo kernels of the multi-process path (K3/K4/K17/K18, sort.rs:74-164).
#include <cfloat>

#include "spk_device.cuh"
#include "spk_math.cuh"

namespace {

// The reference accumulates the grid with `red.global.add.f32`, which flushes subnormal inputs and results to zero
// (PTX ISA); the TMA bulk reduction used here does not, so every consumer of the accumulators flushes explicitly.
// This keeps 1/mass finite exactly where the reference's is.
__device__ __forceinline__ float spk_ftz(float x) { return fabsf(x) < FLT_MIN ? copysignf(0.f, x) : x; }

// ---- reset + carry: one warp-pair (64 threads) per block, one thread per cell ---------------------------------
__global__ void __launch_bounds__(256) reset_grid_kernel(GridBuffers next, GridBuffers prev, uint32_t nab) {
    const uint32_t b = blockIdx.x * 4 + (threadIdx.x >> 6);
    const uint32_t cell = threadIdx.x & 63;
    if (b >= nab) return;
    float* acc = next.acc + (size_t)b * (SPK_ACC_CHANNELS * 64);
#pragma unroll
    for (int ch = 0; ch < SPK_ACC_CHANNELS; ++ch) acc[ch * 64 + cell] = 0.f; // GpuGridNode::default (reset_grid.rs:24-26)
    const int32_t pb = next.nbr_prev[b * 8 + 0];                              // same key in the previous grid (reset_grid.rs:46)
    const size_t node = (size_t)b * 64 + cell;
    if (pb >= 0) {
        const size_t pnode = (size_t)pb * 64 + cell;
        float4 n = prev.cacheN[pnode];
        n.w = spk_ftz(prev.acc[(size_t)pb * (SPK_ACC_CHANNELS * 64) + SPK_ACC_SWZ(cell)]); // prev_mass = prev.mass (reset_grid.rs:54)
        next.cacheN[node] = n;
        next.cacheD[node] = prev.cacheD[pnode];
        next.proj[node] = prev.proj[pnode];
    } else {
        next.cacheN[node] = make_float4(0.f, 0.f, 0.f, 0.f);
        next.cacheD[node] = make_float4(0.f, 0.f, 0.f, 0.f);
        next.proj[node] = SPK_PROJ_NOT_COMPUTED;
    }
}

// ---- cuboid projection: parry3d 0.13 Aabb::project_local_point with solid = false (gpu_collider.rs:53) -----------
struct Vec3 { float x, y, z; };
__device__ __forceinline__ Vec3 quat_rotate(const float q[4], Vec3 v) {
    Vec3 t{2.f * (q[1] * v.z - q[2] * v.y), 2.f * (q[2] * v.x - q[0] * v.z), 2.f * (q[0] * v.y - q[1] * v.x)};
    Vec3 c{q[1] * t.z - q[2] * t.y, q[2] * t.x - q[0] * t.z, q[0] * t.y - q[1] * t.x};
    return Vec3{v.x + q[3] * t.x + c.x, v.y + q[3] * t.y + c.y, v.z + q[3] * t.z + c.z};
}
__device__ inline void cuboid_project(const DevCollider& c, Vec3 pt, Vec3& out, bool& inside) {
    const float qinv[4] = {-c.rot[0], -c.rot[1], -c.rot[2], c.rot[3]};
    const Vec3 lp = quat_rotate(qinv, Vec3{pt.x - c.tr[0], pt.y - c.tr[1], pt.z - c.tr[2]});
    const float l[3] = {lp.x, lp.y, lp.z};
    float mins_pt[3], pt_maxs[3], shift[3];
    bool zero = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        mins_pt[a] = -c.half_extents[a] - l[a];
        pt_maxs[a] = l[a] - c.half_extents[a];
        shift[a] = fmaxf(mins_pt[a], 0.f) - fmaxf(pt_maxs[a], 0.f);
        zero = zero && shift[a] == 0.f;
    }
    float r[3] = {l[0], l[1], l[2]};
    if (!zero) {
        inside = false;
#pragma unroll
        for (int a = 0; a < 3; ++a) r[a] += shift[a];
    } else {
        float best = -FLT_MAX;
        bool is_mins = false;
        int best_id = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (mins_pt[a] < pt_maxs[a]) {
                if (pt_maxs[a] > best) { best_id = a; is_mins = false; best = pt_maxs[a]; }
            } else if (mins_pt[a] > best) { best_id = a; is_mins = true; best = mins_pt[a]; }
        }
        inside = true;
#pragma unroll
        for (int a = 0; a < 3; ++a) if (a == best_id) r[a] += is_mins ? best : -best;
    }
    const Vec3 w = quat_rotate(c.rot, Vec3{r[0], r[1], r[2]});
    out = Vec3{w.x + c.tr[0], w.y + c.tr[1], w.z + c.tr[2]};
}
// ---- heightfield projection: parry3d 0.13 `GenericHeightField::project_local_point_with_max_dist` (point_heightfield.rs)
// over `map_elements_in_local_aabb` (heightfield3.rs), restated; then the inside test of
// GpuColliderShape::project_point_with_max_dist (src_kernels/gpu_collider.rs:73-99).
// Closest point on a triangle: Ericson, Real-Time Collision Detection 5.1.5 (the Voronoi-region walk parry uses).
__device__ __forceinline__ Vec3 v_sub(Vec3 a, Vec3 b) { return Vec3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ float v_dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ inline Vec3 closest_point_on_triangle(Vec3 p, Vec3 a, Vec3 b, Vec3 c) {
    const Vec3 ab = v_sub(b, a), ac = v_sub(c, a), ap = v_sub(p, a);
    const float d1 = v_dot(ab, ap), d2 = v_dot(ac, ap);
    if (d1 <= 0.f && d2 <= 0.f) return a;
    const Vec3 bp = v_sub(p, b);
    const float d3 = v_dot(ab, bp), d4 = v_dot(ac, bp);
    if (d3 >= 0.f && d4 <= d3) return b;
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
        const float v = d1 / (d1 - d3);
        return Vec3{a.x + v * ab.x, a.y + v * ab.y, a.z + v * ab.z};
    }
    const Vec3 cp = v_sub(p, c);
    const float d5 = v_dot(ab, cp), d6 = v_dot(ac, cp);
    if (d6 >= 0.f && d5 <= d6) return c;
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
        const float w = d2 / (d2 - d6);
        return Vec3{a.x + w * ac.x, a.y + w * ac.y, a.z + w * ac.z};
    }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
        const float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        return Vec3{b.x + w * (c.x - b.x), b.y + w * (c.y - b.y), b.z + w * (c.z - b.z)};
    }
    const float denom = 1.f / (va + vb + vc);
    const float v = vb * denom, w = vc * denom;
    return Vec3{a.x + ab.x * v + ac.x * w, a.y + ab.y * v + ac.y * w, a.z + ab.z * v + ac.z * w};
}
__device__ inline bool heightfield_project(const DevCollider& c, Vec3 pt, float max_dist, Vec3& out, bool& inside) {
    const float qinv[4] = {-c.rot[0], -c.rot[1], -c.rot[2], c.rot[3]};
    const Vec3 lp = quat_rotate(qinv, Vec3{pt.x - c.tr[0], pt.y - c.tr[1], pt.z - c.tr[2]});
    const int cells_x = (int)c.ncols - 1, cells_z = (int)c.nrows - 1; // HeightField::ncols() / nrows(): number of CELLS
    const float cw = 1.f / (float)cells_x, ch = 1.f / (float)cells_z; // unit_cell_width / unit_cell_height
    const float rminx = (lp.x - max_dist) / c.scale[0], rminy = (lp.y - max_dist) / c.scale[1], rminz = (lp.z - max_dist) / c.scale[2];
    const float rmaxx = (lp.x + max_dist) / c.scale[0], rmaxy = (lp.y + max_dist) / c.scale[1], rmaxz = (lp.z + max_dist) / c.scale[2];
    if (rmaxx <= -0.5f || rmaxz <= -0.5f || rminx >= 0.5f || rminz >= 0.5f) return false; // outside of the heightfield bounds
    // quantize_floor / quantize_ceil (heightfield3.rs)
    const int min_x = (int)fminf(fmaxf(floorf((rminx + 0.5f) / cw), 0.f), (float)(cells_x - 1));
    const int min_z = (int)fminf(fmaxf(floorf((rminz + 0.5f) / ch), 0.f), (float)(cells_z - 1));
    const int max_x = (int)fminf(fmaxf(ceilf((rmaxx + 0.5f) / cw), 0.f), (float)cells_x);
    const int max_z = (int)fminf(fmaxf(ceilf((rmaxz + 0.5f) / ch), 0.f), (float)cells_z);
    float sq_smallest = FLT_MAX;
    bool found = false;
    Vec3 best{0.f, 0.f, 0.f};
    for (int j = min_x; j < max_x; ++j) {
        for (int i = min_z; i < max_z; ++i) {
            const float z0 = -0.5f + ch * (float)i, z1 = -0.5f + ch * (float)(i + 1);
            const float x0 = -0.5f + cw * (float)j, x1 = -0.5f + cw * (float)(j + 1);
            const float y00 = c.heights[i + (size_t)j * c.nrows], y10 = c.heights[i + 1 + (size_t)j * c.nrows];
            const float y01 = c.heights[i + (size_t)(j + 1) * c.nrows], y11 = c.heights[i + 1 + (size_t)(j + 1) * c.nrows];
            if ((y00 > rmaxy && y10 > rmaxy && y01 > rmaxy && y11 > rmaxy) || (y00 < rminy && y10 < rminy && y01 < rminy && y11 < rminy)) continue;
            const Vec3 p00{x0 * c.scale[0], y00 * c.scale[1], z0 * c.scale[2]}, p10{x0 * c.scale[0], y10 * c.scale[1], z1 * c.scale[2]};
            const Vec3 p01{x1 * c.scale[0], y01 * c.scale[1], z0 * c.scale[2]}, p11{x1 * c.scale[0], y11 * c.scale[1], z1 * c.scale[2]};
#pragma unroll
            for (int t = 0; t < 2; ++t) { // default cell status: triangles (p00, p10, p01) and (p10, p11, p01)
                const Vec3 q = t == 0 ? closest_point_on_triangle(lp, p00, p10, p01) : closest_point_on_triangle(lp, p10, p11, p01);
                const Vec3 d = v_sub(lp, q);
                const float sq = v_dot(d, d);
                if (sq < sq_smallest) {
                    sq_smallest = sq;
                    if (sqrtf(sq) <= max_dist) { best = q; found = true; }
                }
            }
        }
    }
    if (!found) return false;
    if (c.hf_flags & 1u) { // bellow_heightfield_is_solid: inside = within the x/z footprint and on the solid side of the surface
        const bool in_footprint = lp.x >= -0.5f * c.scale[0] && lp.x <= 0.5f * c.scale[0] && lp.z >= -0.5f * c.scale[2] && lp.z <= 0.5f * c.scale[2];
        inside = in_footprint && ((c.hf_flags & 2u) ? lp.y >= best.y : lp.y <= best.y);
    } else {
        inside = best.x == lp.x && best.y == lp.y && best.z == lp.z;
    }
    const Vec3 w = quat_rotate(c.rot, best);
    out = Vec3{w.x + c.tr[0], w.y + c.tr[1], w.z + c.tr[2]};
    return true;
}
// polyline_project_point, src_kernels/gpu_collider.rs:105-201 (dim3): closest segment of the closed polygon (full 3D distance,
// first of equal distances wins), inside / outside from the segment's normal in the xy plane (`planar_normal(2)`), with the
// vertex rule of parry's Polyline::project_local_point_assuming_solid_interior_ccw; the projection keeps the query's z.
// Quirk kept: the inside tests subtract the LOCAL projection from the WORLD point (gpu_collider.rs:176,181).
__device__ inline bool polyline_project(const DevCollider& c, Vec3 pt, Vec3& out, bool& inside) {
    const int n = (int)c.nrows;
    if (n == 0) return false;
    const float qinv[4] = {-c.rot[0], -c.rot[1], -c.rot[2], c.rot[3]};
    const Vec3 lp = quat_rotate(qinv, Vec3{pt.x - c.tr[0], pt.y - c.tr[1], pt.z - c.tr[2]});
    const float* v = c.heights;
    float best_dist = FLT_MAX;
    Vec3 best{0.f, 0.f, 0.f}, sa{0.f, 0.f, 0.f}, sb{0.f, 0.f, 0.f};
    int best_id = 0, best_loc = 2; // 0 / 1: on vertex a / b, 2: on the edge
    bool best_eq = false;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1) % n;
        const Vec3 a{v[3 * i], v[3 * i + 1], v[3 * i + 2]}, b{v[3 * j], v[3 * j + 1], v[3 * j + 2]};
        const Vec3 ab = v_sub(b, a), ap = v_sub(lp, a);
        const float ab_ap = v_dot(ab, ap), sqnab = v_dot(ab, ab);
        Vec3 q; int loc;
        if (ab_ap <= 0.f) { q = a; loc = 0; }
        else if (ab_ap >= sqnab) { q = b; loc = 1; }
        else { const float u = ab_ap / sqnab; q = Vec3{a.x + ab.x * u, a.y + ab.y * u, a.z + ab.z * u}; loc = 2; }
        const Vec3 d = v_sub(lp, q);
        const float dist = sqrtf(v_dot(d, d));
        if (i == 0 || dist < best_dist) { best_dist = dist; best = q; best_id = i; best_loc = loc; sa = a; sb = b; best_eq = q.x == lp.x && q.y == lp.y && q.z == lp.z; }
    }
    inside = best_eq;
    const Vec3 dir = v_sub(sb, sa);
    const float nn = sqrtf(dir.y * dir.y + dir.x * dir.x);
    if (nn > FLT_EPSILON) { // planar_normal(2) = normalize((dir.y, -dir.x, 0))
        const Vec3 nrm{dir.y / nn, -dir.x / nn, 0.f};
        const float side = (pt.x - best.x) * nrm.x + (pt.y - best.y) * nrm.y + (pt.z - best.z) * nrm.z;
        if (best_loc == 2) {
            inside = side <= 0.f;
        } else {
            Vec3 dir2;
            if (best_loc == 0) { // previous segment, reversed
                const int adj = best_id == 0 ? n - 1 : best_id - 1, aj = (adj + 1) % n;
                dir2 = Vec3{-(v[3 * aj] - v[3 * adj]), -(v[3 * aj + 1] - v[3 * adj + 1]), -(v[3 * aj + 2] - v[3 * adj + 2])};
            } else {
                const int adj = (best_id + 1) % n, aj = (adj + 1) % n;
                dir2 = Vec3{v[3 * aj] - v[3 * adj], v[3 * aj + 1] - v[3 * adj + 1], v[3 * aj + 2] - v[3 * adj + 2]};
            }
            const float dot = v_dot(nrm, dir2);
            const float threshold = 1.0e-3f * sqrtf(v_dot(dir2, dir2));
            inside = fabsf(dot) > threshold ? dot >= 0.f : side <= 0.f;
        }
    }
    best.z = lp.z;
    if (c.hf_flags & 2u) inside = !inside;
    const Vec3 w = quat_rotate(c.rot, best);
    out = Vec3{w.x + c.tr[0], w.y + c.tr[1], w.z + c.tr[2]};
    return true;
}
// GpuColliderShape::TriMesh, src_kernels/gpu_collider.rs:58-68: `trimesh.project_point_with_max_dist(position, point, solid, Real::MAX)`
// (no distance limit), then flip_interior.  parry3d 0.13's TriMesh projection returns the closest point over all triangles (its QBVH
// only prunes the search); without pseudo-normals (rapier's default trimesh flags) is_inside is the triangle's own answer for
// solid = false: true only when the query lies on the triangle.  Exhaustive search here: the projection is computed once per node
// and then carried in the cache (reset_grid_kernel), and collider meshes are small next to the grid.
__device__ inline bool trimesh_project(const DevCollider& c, Vec3 pt, Vec3& out, bool& inside) {
    const int nv = (int)c.nrows, nt = (int)c.ncols;
    if (nt == 0) return false;
    const float qinv[4] = {-c.rot[0], -c.rot[1], -c.rot[2], c.rot[3]};
    const Vec3 lp = quat_rotate(qinv, Vec3{pt.x - c.tr[0], pt.y - c.tr[1], pt.z - c.tr[2]});
    const float* v = c.heights;
    const uint32_t* idx = reinterpret_cast<const uint32_t*>(c.heights + (size_t)nv * 3);
    float best_sq = FLT_MAX;
    Vec3 best{0.f, 0.f, 0.f};
    for (int t = 0; t < nt; ++t) {
        const uint32_t i0 = idx[3 * t], i1 = idx[3 * t + 1], i2 = idx[3 * t + 2];
        const Vec3 a{v[3 * i0], v[3 * i0 + 1], v[3 * i0 + 2]}, b{v[3 * i1], v[3 * i1 + 1], v[3 * i1 + 2]}, cc{v[3 * i2], v[3 * i2 + 1], v[3 * i2 + 2]};
        const Vec3 q = closest_point_on_triangle(lp, a, b, cc);
        const Vec3 d = v_sub(lp, q);
        const float sq = v_dot(d, d);
        if (sq < best_sq) { best_sq = sq; best = q; }
    }
    inside = best.x == lp.x && best.y == lp.y && best.z == lp.z;
    if (c.hf_flags & 2u) inside = !inside;
    const Vec3 w = quat_rotate(c.rot, best);
    out = Vec3{w.x + c.tr[0], w.y + c.tr[1], w.z + c.tr[2]};
    return true;
}
// GpuColliderShape::project_point_with_max_dist, gpu_collider.rs:42-103 (solid = false)
__device__ __forceinline__ bool collider_project(const DevCollider& c, Vec3 pt, float max_dist, Vec3& out, bool& inside) {
    if (c.shape == SPK_SHAPE_TRIMESH) return trimesh_project(c, pt, out, inside);
    if (c.shape == SPK_SHAPE_HEIGHTFIELD) return heightfield_project(c, pt, max_dist, out, inside);
    if (c.shape == SPK_SHAPE_POLYLINE) return polyline_project(c, pt, out, inside);
    cuboid_project(c, pt, out, inside);
    return true;
}
// sdf, grid_update.rs:54-71
__device__ inline bool sdf(const DevCollider* cols, uint32_t ncol, Vec3 q, float max_dist, float& out) {
    bool any = false;
    float best = 0.f;
    for (uint32_t i = 0; i < ncol; ++i) {
        if (cols[i].handling == SPK_BOUNDARY_NONE) continue;
        Vec3 p; bool inside;
        if (!collider_project(cols[i], q, max_dist, p, inside)) continue;
        const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
        const float d = sqrtf(dx * dx + dy * dy + dz * dz) * (inside ? -1.f : 1.f);
        if (!any || d < best) { best = d; any = true; }
    }
    out = best;
    return any;
}

__global__ void __launch_bounds__(256) grid_update_kernel(GridBuffers g, uint32_t nab, const DevCollider* __restrict__ cols, uint32_t ncol, float h, float dt,
                                                           float gx, float gy, float gz, float* hacc) {
    const uint32_t b = blockIdx.x * 4 + (threadIdx.x >> 6);
    const uint32_t cell = threadIdx.x & 63;
    if (b >= nab) return;
    if (hacc && (g.blk_halo[b] & 1u)) { // multi-GPU, merge_halo_blocks (sort.rs:142-164): fold what the peers pushed, clear it for the next use
        float* a = g.acc + (size_t)b * (SPK_ACC_CHANNELS * 64);
        float* hh = hacc + (size_t)b * (SPK_ACC_CHANNELS * 64);
#pragma unroll
        for (int ch = 0; ch < SPK_ACC_CHANNELS; ++ch) {
            const uint32_t i = ch * 64 + SPK_ACC_SWZ(cell); // exactly the entries this thread reads below, through `acc`
            a[i] += __ldcg(&hh[i]);
            hh[i] = 0.f;
        }
    }
    const float* acc = g.acc + (size_t)b * (SPK_ACC_CHANNELS * 64) + SPK_ACC_SWZ(cell) - cell; // acc[ch * 64 + cell] below: the swizzled slot
    const size_t node = (size_t)b * 64 + cell;
    // the reference's `red.global.add.f32` flushes subnormals (PTX ISA); keep 1/mass finite the same way
    const float mass = spk_ftz(acc[cell]);
    const float psi_mass = spk_ftz(acc[5 * 64 + cell]);
    const float inv_m = inv_exact(mass);
    // (momentum + mass * gravity * dt) * inv_exact(mass), grid_update.rs:104-105
    float vx = (spk_ftz(acc[64 + cell]) + mass * gx * dt) * inv_m;
    float vy = (spk_ftz(acc[128 + cell]) + mass * gy * dt) * inv_m;
    float vz = (spk_ftz(acc[192 + cell]) + mass * gz * dt) * inv_m;
    uint32_t proj = g.proj[node];
    float4 cn = g.cacheN[node];
    float4 cd = g.cacheD[node];
    if (ncol == 0) {
        if ((proj & 3u) == SPK_PROJ_NOT_COMPUTED) { proj = SPK_PROJ_TOO_FAR; g.proj[node] = proj; } // no collider: best_proj = None (grid_update.rs:138-140)
    } else if ((proj & 3u) == SPK_PROJ_NOT_COMPUTED) {
        unsigned long long kx, ky, kz;
        spk_unpack_key(g.blk_key[b], kx, ky, kz);
        const long long cx = ((long long)kx - SPK_PACK_ORIGIN) * 4 + (cell & 3), cy = ((long long)ky - SPK_PACK_ORIGIN) * 4 + ((cell >> 2) & 3),
                        cz = ((long long)kz - SPK_PACK_ORIGIN) * 4 + (cell >> 4);
        // grid_update.rs:34-38; __fmul_rn keeps the products rounded (no fma contraction with the collider code's subtractions),
        // as in the reference: it decides ties such as `dist - h <= 0` for nodes exactly one cell above an axis-aligned surface
        const Vec3 pos{__fmul_rn((float)cx, h), __fmul_rn((float)cy, h), __fmul_rn((float)cz, h)};
        bool have = false, best_inside = false;
        float best_dist = FLT_MAX;
        Vec3 best_dir{0.f, 0.f, 0.f};
        uint32_t best_id = 0;
        for (uint32_t i = 0; i < ncol; ++i) {
            if (cols[i].handling == SPK_BOUNDARY_NONE) continue;
            Vec3 p; bool inside;
            if (!collider_project(cols[i], pos, h * 2.f, p, inside)) continue; // max_dist = cell_width * 2, grid_update.rs:118
            const Vec3 dir{pos.x - p.x, pos.y - p.y, pos.z - p.z};
            const float dist = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
            if (dist < (have ? best_dist : FLT_MAX)) { have = true; best_dist = dist; best_dir = dir; best_inside = inside; best_id = i; }
        }
        if (have) {
            proj = (best_inside ? SPK_PROJ_INSIDE : SPK_PROJ_OUTSIDE) | (best_id << 2);
            cd = make_float4(best_dir.x, best_dir.y, best_dir.z, 0.f);
        } else {
            proj = SPK_PROJ_TOO_FAR;
        }
        // sdf_gradient, grid_update.rs:74-94: central differences at +-0.1 h
        const float off = 1.f * (h * 0.1f);
        float grad[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            Vec3 qp = pos, qn = pos;
            if (a == 0) { qp.x += off; qn.x += -off; } else if (a == 1) { qp.y += off; qn.y += -off; } else { qp.z += off; qn.z += -off; }
            float sp, sn;
            const bool hp = sdf(cols, ncol, qp, h * 2.f, sp), hn = sdf(cols, ncol, qn, h * 2.f, sn);
            grad[a] = (hp && hn) ? (sp - sn) / h / 0.1f / 2.f : 0.f;
        }
        const float gn = sqrtf(grad[0] * grad[0] + grad[1] * grad[1] + grad[2] * grad[2]);
        if (gn <= 1.0e-5f) cn.x = cn.y = cn.z = 0.f; // try_normalize(1e-5).unwrap_or(zeros)
        else { cn.x = grad[0] / gn; cn.y = grad[1] / gn; cn.z = grad[2] / gn; }
        g.proj[node] = proj;
        g.cacheN[node] = cn;
        g.cacheD[node] = cd;
    }
    const uint32_t tag = proj & 3u;
    if (tag == SPK_PROJ_INSIDE || tag == SPK_PROJ_OUTSIDE) { // grid_update.rs:142-194
        const bool is_inside = tag == SPK_PROJ_INSIDE;
        const DevCollider& col = cols[proj >> 2];
        if (col.handling == SPK_BOUNDARY_STICK) {
            if (is_inside) vx = vy = vz = 0.f;
        } else if (col.handling == SPK_BOUNDARY_FRICTION || col.handling == SPK_BOUNDARY_FRICTION_Z_UP) {
            if (!(cn.x == 0.f && cn.y == 0.f && cn.z == 0.f)) {
                const float dist = sqrtf(cd.x * cd.x + cd.y * cd.y + cd.z * cd.z);
                const bool apply = col.handling == SPK_BOUNDARY_FRICTION || cn.z >= 0.f;
                if (apply) {
                    const float normal_vel = vx * cn.x + vy * cn.y + vz * cn.z;
                    if (normal_vel < 0.f) {
                        const float dist_with_margin = dist - h;
                        if (is_inside || dist_with_margin <= 0.f) {
                            const float tx = vx - normal_vel * cn.x, ty = vy - normal_vel * cn.y, tz = vz - normal_vel * cn.z;
                            const float tn = sqrtf(tx * tx + ty * ty + tz * tz);
                            vx = tx; vy = ty; vz = tz;
                            if (tn > 1.0e-10f) {
                                const float k = fmaxf(tn + normal_vel * col.friction, 0.f);
                                vx = tx / tn * k; vy = ty / tn * k; vz = tz / tn * k;
                            }
                        } else if (-normal_vel * dt > dist_with_margin) {
                            const float k = dist_with_margin / dt + normal_vel;
                            vx -= k * cn.x; vy -= k * cn.y; vz -= k * cn.z;
                        }
                    }
                }
            }
        }
    }
    const float psi_v = spk_ftz(acc[4 * 64 + cell]) * inv_exact(psi_mass); // grid_update.rs:197
    g.vel[node] = make_float4(vx, vy, vz, psi_v);
}

// ---- export to the reference's GpuGridNode layout --------------------------------------------------------------
__global__ void grid_export_kernel(GridBuffers g, uint32_t nblocks, SpkGpuGridNode* out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nblocks * 64) return;
    const uint32_t b = t >> 6, cell = t & 63;
    const float* acc = g.acc + (size_t)b * (SPK_ACC_CHANNELS * 64) + SPK_ACC_SWZ(cell) - cell;
    SpkGpuGridNode n;
    n.mass = spk_ftz(acc[cell]);
    n.psi_mass = spk_ftz(acc[5 * 64 + cell]);
    if (g.has_velocity) { // after grid_update the reference's node holds velocities in the momentum fields
        const float4 v = g.vel[t];
        n.momentum_velocity[0] = v.x; n.momentum_velocity[1] = v.y; n.momentum_velocity[2] = v.z;
        n.psi_momentum_velocity = v.w;
    } else {
        n.momentum_velocity[0] = spk_ftz(acc[64 + cell]); n.momentum_velocity[1] = spk_ftz(acc[128 + cell]); n.momentum_velocity[2] = spk_ftz(acc[192 + cell]);
        n.psi_momentum_velocity = spk_ftz(acc[256 + cell]);
    }
    const float4 cn = g.cacheN[t], cd = g.cacheD[t];
    const uint32_t proj = g.proj[t];
    n.prev_mass = cn.w;
    n._pad0 = 0;
    n.projection_status_tag = proj & 3u;
    n.projection_status_payload = ((proj & 3u) == SPK_PROJ_INSIDE || (proj & 3u) == SPK_PROJ_OUTSIDE) ? (proj >> 2) : 0;
    n.collision_normal[0] = cn.x; n.collision_normal[1] = cn.y; n.collision_normal[2] = cn.z;
    n.projection_scaled_dir[0] = cd.x; n.projection_scaled_dir[1] = cd.y; n.projection_scaled_dir[2] = cd.z;
    out[t] = n;
}

#if defined(__CUDACC__) || defined(SPK_EMU) // everything below needs the device hash map / launch syntax; the per-node kernels above are also host-compiled by tests/host/
// ---- halo path (one process per GPU) -----------------------------------------------------------------------------
__global__ void halo_keys_kernel(GridBuffers g, uint32_t nab, unsigned long long* out, uint32_t capacity) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < capacity) out[b] = b < nab ? g.blk_key[b] : SPK_EMPTY_KEY;
}
// tag_halo_blocks (sort.rs:74-96): a remote active block that is also active here becomes IS_HALO
__global__ void halo_tag_kernel(GridBuffers g, const unsigned long long* __restrict__ remote, uint32_t n, DeviceCounters* counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (remote[i] == SPK_EMPTY_KEY) return; // padding
    const int32_t b = spk_hash_get(g, remote[i]);
    if (b >= 0 && atomicExch(&g.blk_halo[b], 1u) == 0u) atomicAdd(&counters->num_halo, 1u);
}
// copy_halo_to_staging (sort.rs:116-140), payload reduced to the 6 summed channels
__global__ void __launch_bounds__(64) halo_pack_kernel(GridBuffers g, uint32_t nab, unsigned long long* keys, float* payload, uint32_t capacity, uint32_t* cursor) {
    const uint32_t b = blockIdx.x;
    if (b >= nab || !(g.blk_halo[b] & 1u)) return;
    __shared__ uint32_t s_slot;
    if (threadIdx.x == 0) s_slot = atomicAdd(cursor, 1u);
    __syncthreads();
    const uint32_t slot = s_slot;
    if (slot >= capacity) return;
    if (threadIdx.x == 0) keys[slot] = g.blk_key[b];
    const float* acc = g.acc + (size_t)b * (SPK_ACC_CHANNELS * 64);
    float* dst = payload + (size_t)slot * (SPK_ACC_CHANNELS * 64);
#pragma unroll
    for (int ch = 0; ch < SPK_ACC_CHANNELS; ++ch) dst[ch * 64 + threadIdx.x] = acc[ch * 64 + threadIdx.x];
}
// merge_halo_blocks (sort.rs:142-164)
__global__ void __launch_bounds__(64) halo_merge_kernel(GridBuffers g, const unsigned long long* __restrict__ keys, const float* __restrict__ payload) {
    if (keys[blockIdx.x] == SPK_EMPTY_KEY) return; // padding
    const int32_t b = spk_hash_get(g, keys[blockIdx.x]);
    if (b < 0) return;
    float* acc = g.acc + (size_t)b * (SPK_ACC_CHANNELS * 64);
    const float* src = payload + (size_t)blockIdx.x * (SPK_ACC_CHANNELS * 64);
#pragma unroll
    for (int ch = 0; ch < SPK_ACC_CHANNELS; ++ch) atomicAdd(&acc[ch * 64 + threadIdx.x], src[ch * 64 + threadIdx.x]);
}

#endif // __CUDACC__

} // namespace

#if defined(__CUDACC__) || defined(SPK_EMU)
cudaError_t spk_launch_reset_grid(spk_ctx* c) {
    GridBuffers& next = c->grid[1 - c->gcur];
    const uint32_t nab = next.nab_host;
    next.has_velocity = 0;
    if (!nab) return cudaSuccess;
    reset_grid_kernel<<<(nab + 3) / 4, 256, 0, c->stream>>>(next, c->grid[c->gcur], nab);
    c->launches += 1;
    return cudaGetLastError();
}

cudaError_t spk_launch_grid_update(spk_ctx* c, float dt, const float gravity[SPK_DIM]) {
    GridBuffers& next = c->grid[1 - c->gcur];
    const uint32_t nab = next.nab_host;
    if (nab) {
        grid_update_kernel<<<(nab + 3) / 4, 256, 0, c->stream>>>(next, nab, c->colliders, (uint32_t)c->ncolliders, c->cell_width, dt, gravity[0], gravity[1], gravity[2],
                                                                 c->dist_merge_in_update ? spk_dist_hacc_of_next(c) : nullptr);
        c->launches += 1;
    }
    next.has_velocity = 1;
    c->gcur = 1 - c->gcur; // CudaSparseGrid::swap_buffers, cuda_sparse_grid.rs:108-110
    return cudaGetLastError();
}

cudaError_t spk_launch_grid_export(spk_ctx* c, int which_grid, SpkGpuGridNode* d_out, uint32_t nblocks) {
    GridBuffers& g = which_grid ? c->grid[1 - c->gcur] : c->grid[c->gcur];
    if (!nblocks) return cudaSuccess;
    grid_export_kernel<<<(nblocks * 64 + 255) / 256, 256, 0, c->stream>>>(g, nblocks, d_out);
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
    if (e == cudaSuccess && capacity) e = cudaMemsetAsync(keys, 0xff, (size_t)capacity * 8, c->stream); // EMPTY padding
    if (e != cudaSuccess || !g.nab_host) return e;
    c->counters_fresh = false;
    halo_pack_kernel<<<g.nab_host, 64, 0, c->stream>>>(g, g.nab_host, keys, payload, capacity, &c->d_counters->halo_cursor);
    c->launches += 1;
    return cudaGetLastError();
}
cudaError_t spk_launch_halo_merge(spk_ctx* c, const unsigned long long* keys, const float* payload, uint32_t n) {
    GridBuffers& g = c->grid[1 - c->gcur];
    if (!n) return cudaSuccess;
    halo_merge_kernel<<<n, 64, 0, c->stream>>>(g, keys, payload);
    c->launches += 1;
    return cudaGetLastError();
}
#endif // __CUDACC__
