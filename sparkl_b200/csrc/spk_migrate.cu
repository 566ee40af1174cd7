// spk_migrate.cu — device-side particle migration between slabs ("pack + send/recv of 128-B particle records
// to neighbours").  The reference never re-homes particles (static ownership, src/cuda/cuda_mpm_pipeline.rs:177-194); this is the
// fast path of sparkl_b200/distributed.py::migrate: nothing but the per-destination counts crosses PCIe.
//
//   spk_particles_migrate_out: owner of a particle = slab r with bounds[r] <= block_x < bounds[r + 1] (block_x of
//     GpuGrid::block_associated_to_point, gpu_grid.rs:271-279; the outer two bounds are open ended).  Particles owned by other
//     ranks leave as SpkParticleRecord rows in DEVICE memory, grouped by destination rank and ascending particle id inside a
//     group; the others keep their relative order and are renumbered densely (a stable partition, so results do not depend on
//     scheduling).  Passing dev_records = NULL only counts.
//   spk_particles_migrate_in: appends records from DEVICE memory (ids continue after the current ones, in record order).
//
// Both are built from the streams <-> reference-layout conversions the upload / read-back paths already use (pack_kernel /
// unpack_kernel, spk_api.cu) plus a three-kernel stable multi-way partition (per-CTA bucket histogram with match.any, one-CTA
// exclusive scan of the bucket-major table, scatter).  A slow path by design: it runs every few frames, moves ~2 x 128 B per
// resident particle through HBM when anything leaves, and nothing when nobody does (the sort state stays valid then).
#include <algorithm>
#include <vector>

#include "spk_device.cuh"

#define CK(expr)                               \
    do {                                       \
        cudaError_t _e = (expr);               \
        if (_e != cudaSuccess) return (int)_e; \
    } while (0)

namespace {

constexpr int MG_THREADS = 256;
constexpr int MG_WARPS = MG_THREADS / 32;
constexpr int MG_MAX_BUCKETS = SPK_MAX_RANKS + 2; // bucket 0 = stays, 1 + r = leaves for rank r, last = padding lanes
constexpr int MG_SCAN_THREADS = 1024;

struct MgBounds { long long b[SPK_MAX_RANKS + 1]; uint32_t world, rank; };
struct MgAos { // the reference's per-attribute arrays (indexed by particle id) + global ids, all in device memory
    SpkParticleStatus* status; SpkParticlePosition* pos; SpkParticleVelocity* vel; SpkParticleVolume* volume; SpkParticlePhase* phase;
    SpkParticleData* data; uint32_t* gid;
};

size_t mg_align(size_t b) { return (b + 255) & ~size_t(255); }
size_t mg_aos_bytes(size_t n) {
    return mg_align(n * sizeof(SpkParticleStatus)) + mg_align(n * sizeof(SpkParticlePosition)) + mg_align(n * sizeof(SpkParticleVelocity)) +
           mg_align(n * sizeof(SpkParticleVolume)) + mg_align(n * sizeof(SpkParticlePhase)) + mg_align(n * sizeof(SpkParticleData)) + mg_align(n * 4);
}
MgAos mg_carve(char*& cur, size_t n) {
    MgAos a;
    a.status = (SpkParticleStatus*)cur; cur += mg_align(n * sizeof(SpkParticleStatus));
    a.pos = (SpkParticlePosition*)cur; cur += mg_align(n * sizeof(SpkParticlePosition));
    a.vel = (SpkParticleVelocity*)cur; cur += mg_align(n * sizeof(SpkParticleVelocity));
    a.volume = (SpkParticleVolume*)cur; cur += mg_align(n * sizeof(SpkParticleVolume));
    a.phase = (SpkParticlePhase*)cur; cur += mg_align(n * sizeof(SpkParticlePhase));
    a.data = (SpkParticleData*)cur; cur += mg_align(n * sizeof(SpkParticleData));
    a.gid = (uint32_t*)cur; cur += mg_align(n * 4);
    return a;
}

// bucket of a particle: 0 when its block column belongs to this rank's slab, 1 + owner otherwise
__device__ __forceinline__ uint32_t mg_bucket(const MgBounds& B, float x, float h) {
    const long long bx = (long long)spk_block_coord(spk_assoc_cell(x, h)) - (long long)SPK_PACK_ORIGIN;
    uint32_t owner = 0;
    for (uint32_t r = 1; r < B.world; ++r) owner += (B.b[r] <= bx) ? 1u : 0u; // interior bounds <= bx (searchsorted side = right)
    return owner == B.rank ? 0u : owner + 1u;
}

// stable rank of the calling thread among the CTA's threads with the same bucket; s_cnt[w][b] = members of bucket b in warp w
__device__ __forceinline__ uint32_t mg_rank_in_cta(uint32_t bucket, uint32_t (*s_cnt)[MG_MAX_BUCKETS]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < MG_WARPS * MG_MAX_BUCKETS; k += MG_THREADS) (&s_cnt[0][0])[k] = 0;
    __syncthreads();
    const unsigned mask = __match_any_sync(0xffffffffu, bucket);
    if (lane == __ffs(mask) - 1) s_cnt[warp][bucket] = __popc(mask);
    __syncthreads();
    uint32_t r = __popc(mask & spk_lanemask_lt());
    for (int w = 0; w < warp; ++w) r += s_cnt[w][bucket];
    return r;
}

__global__ void mg_iota_kernel(uint32_t* out, uint32_t first, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = first + i;
}

// pass 1: bucket per particle id + bucket-major histogram table[bucket][cta]
__global__ void __launch_bounds__(MG_THREADS) mg_classify_kernel(const SpkParticlePosition* __restrict__ pos, uint32_t n, float h, MgBounds B,
                                                                  uint8_t* __restrict__ bucket_of, uint32_t* __restrict__ table, uint32_t ncta) {
    __shared__ uint32_t s_cnt[MG_WARPS][MG_MAX_BUCKETS];
    const uint32_t j = blockIdx.x * MG_THREADS + threadIdx.x;
    const uint32_t pad = B.world + 1;
    const uint32_t bucket = j < n ? mg_bucket(B, pos[j].point[0], h) : pad;
    mg_rank_in_cta(bucket, s_cnt);
    if (j < n) bucket_of[j] = (uint8_t)bucket;
    if (threadIdx.x < pad) {
        uint32_t tot = 0;
        for (int w = 0; w < MG_WARPS; ++w) tot += s_cnt[w][threadIdx.x];
        table[threadIdx.x * ncta + blockIdx.x] = tot;
    }
}

// pass 2 (one CTA): exclusive scan of the table in place; starts[b] = first output slot of bucket b, starts[nb] = n
__global__ void __launch_bounds__(MG_SCAN_THREADS) mg_scan_kernel(uint32_t* table, uint32_t entries, uint32_t ncta, uint32_t nbuckets, uint32_t* starts) {
    __shared__ uint32_t s_sum[MG_SCAN_THREADS];
    const uint32_t per = (entries + MG_SCAN_THREADS - 1) / MG_SCAN_THREADS;
    const uint32_t lo = min(threadIdx.x * per, entries), hi = min(lo + per, entries);
    uint32_t local = 0;
    for (uint32_t k = lo; k < hi; ++k) local += table[k];
    s_sum[threadIdx.x] = local;
    __syncthreads();
    for (uint32_t d = 1; d < MG_SCAN_THREADS; d <<= 1) { // Hillis-Steele inclusive scan of the per-thread sums
        const uint32_t v = threadIdx.x >= d ? s_sum[threadIdx.x - d] : 0u;
        __syncthreads();
        s_sum[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = s_sum[threadIdx.x] - local;
    for (uint32_t k = lo; k < hi; ++k) { const uint32_t v = table[k]; table[k] = run; run += v; }
    __syncthreads();
    if (threadIdx.x < nbuckets) starts[threadIdx.x] = table[threadIdx.x * ncta];
    if (threadIdx.x == 0) starts[nbuckets] = s_sum[MG_SCAN_THREADS - 1];
}

// pass 3: stable scatter — bucket 0 to the compacted arrays at the new dense id, the others to the record buffer
__global__ void __launch_bounds__(MG_THREADS) mg_scatter_kernel(MgAos in, uint32_t n, const uint8_t* __restrict__ bucket_of, const uint32_t* __restrict__ table,
                                                                 uint32_t ncta, uint32_t pad, uint32_t n_stay, MgAos out, SpkParticleRecord* __restrict__ records) {
    __shared__ uint32_t s_cnt[MG_WARPS][MG_MAX_BUCKETS];
    const uint32_t j = blockIdx.x * MG_THREADS + threadIdx.x;
    const uint32_t bucket = j < n ? bucket_of[j] : pad;
    const uint32_t r = mg_rank_in_cta(bucket, s_cnt);
    if (j >= n) return;
    const uint32_t dst = table[bucket * ncta + blockIdx.x] + r;
    if (bucket == 0) {
        out.status[dst] = in.status[j]; out.pos[dst] = in.pos[j]; out.vel[dst] = in.vel[j]; out.volume[dst] = in.volume[j];
        out.phase[dst] = in.phase[j]; out.data[dst] = in.data[j]; out.gid[dst] = in.gid[j];
    } else {
        SpkParticleRecord rec;
        rec.status = in.status[j]; rec.position = in.pos[j]; rec.velocity = in.vel[j]; rec.volume = in.volume[j];
        rec.phase = in.phase[j]; rec.data = in.data[j]; rec.gid = in.gid[j];
        records[dst - n_stay] = rec;
    }
}

__global__ void mg_split_kernel(const SpkParticleRecord* __restrict__ records, uint32_t n, MgAos out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const SpkParticleRecord rec = records[k];
    out.status[k] = rec.status; out.pos[k] = rec.position; out.vel[k] = rec.velocity; out.volume[k] = rec.volume;
    out.phase[k] = rec.phase; out.data[k] = rec.data; out.gid[k] = rec.gid;
}

// Scratch of the migration passes: one grow-only allocation kept in the context.  (A cudaMalloc + cudaFree pair per call — 0.5 GB at 2 M
// particles, twice per migration — cost ~30 ms per call at 8 x 2 M particles, more than forty substeps.)
struct DevTemp { char* p = nullptr; };
cudaError_t mg_scratch(spk_ctx* c, DevTemp& t, size_t bytes) {
    if (c->mg_scratch_bytes < bytes) {
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) return e;
        if (c->mg_scratch) cudaFree(c->mg_scratch);
        c->mg_scratch = nullptr;
        c->mg_scratch_bytes = 0;
        const size_t want = bytes + bytes / 8;
        e = cudaMalloc((void**)&c->mg_scratch, want);
        if (e != cudaSuccess) return e;
        c->mg_scratch_bytes = want;
    }
    t.p = c->mg_scratch;
    return cudaSuccess;
}

// all particle attributes -> reference-layout arrays in id order, on the device
int mg_export_all(spk_ctx* c, const MgAos& a, size_t n) {
    CK(spk_launch_unpack(c, SPK_ATTR_STATUS, a.status, n));
    CK(spk_launch_unpack(c, SPK_ATTR_POSITION, a.pos, n));
    CK(spk_launch_unpack(c, SPK_ATTR_VELOCITY, a.vel, n));
    CK(spk_launch_unpack(c, SPK_ATTR_VOLUME, a.volume, n));
    CK(spk_launch_unpack(c, SPK_ATTR_PHASE, a.phase, n));
    CK(spk_launch_unpack(c, SPK_ATTR_DATA, a.data, n));
    if (c->gid) CK(cudaMemcpyAsync(a.gid, c->gid, n * 4, cudaMemcpyDeviceToDevice, c->stream));
    else if (n) mg_iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(a.gid, 0u, (uint32_t)n);
    return (int)cudaGetLastError();
}
int mg_ensure_gids(spk_ctx* c, size_t valid) { // allocate the id -> global id table (identity over the first `valid` ids)
    if (c->gid) return SPK_OK;
    CK(cudaMalloc((void**)&c->gid, std::max<size_t>(c->cap, 1) * 4));
    if (valid) mg_iota_kernel<<<(unsigned)((valid + 255) / 256), 256, 0, c->stream>>>(c->gid, 0u, (uint32_t)valid);
    return (int)cudaGetLastError();
}

} // namespace

int spk_particles_set_gids(spk_ctx* c, const uint32_t* gids, size_t n) {
    if (!c || (n && !gids) || n != c->n) return SPK_EINVAL;
    CK(cudaSetDevice(c->device));
    int rc = mg_ensure_gids(c, 0);
    if (rc) return rc;
    if (n) CK(cudaMemcpyAsync(c->gid, gids, n * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SPK_OK;
}

int spk_particles_read_gids(spk_ctx* c, uint32_t* dst, size_t n) {
    if (!c || (n && !dst) || n > c->n) return SPK_EINVAL;
    CK(cudaSetDevice(c->device));
    if (c->gid) {
        if (n) CK(cudaMemcpyAsync(dst, c->gid, n * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    } else {
        for (size_t i = 0; i < n; ++i) dst[i] = (uint32_t)i;
    }
    return SPK_OK;
}

int spk_particles_migrate_out(spk_ctx* c, const int64_t* bounds, uint32_t world, uint32_t rank, SpkParticleRecord* dev_records, size_t capacity,
                              uint32_t* counts_out) {
    if (!c || !bounds || !counts_out || world == 0 || world > SPK_MAX_RANKS || rank >= world) return SPK_EINVAL;
    for (uint32_t r = 0; r < world; ++r) if (bounds[r] >= bounds[r + 1]) return SPK_EINVAL;
    CK(cudaSetDevice(c->device));
    for (uint32_t r = 0; r < world; ++r) counts_out[r] = 0;
    const size_t n = c->n;
    if (n == 0) return SPK_OK;
    MgBounds B;
    for (uint32_t r = 0; r <= world; ++r) B.b[r] = bounds[r];
    B.world = world; B.rank = rank;
    const uint32_t ncta = (uint32_t)((n + MG_THREADS - 1) / MG_THREADS), nbuckets = world + 1, entries = nbuckets * ncta;
    const bool count_only = dev_records == nullptr;
    // count-only needs the positions alone; the full pass needs every attribute in id order, twice (source and compacted)
    const size_t bytes = (count_only ? mg_align(n * sizeof(SpkParticlePosition)) : 2 * mg_aos_bytes(n)) + mg_align(n) + mg_align((size_t)(entries + ncta) * 4) +
                         mg_align((MG_MAX_BUCKETS + 1) * 4);
    DevTemp tmp;
    CK(mg_scratch(c, tmp, bytes));
    char* cur = tmp.p;
    MgAos in{}, out{};
    if (count_only) { in.pos = (SpkParticlePosition*)cur; cur += mg_align(n * sizeof(SpkParticlePosition)); }
    else { in = mg_carve(cur, n); out = mg_carve(cur, n); }
    uint8_t* bucket_of = (uint8_t*)cur; cur += mg_align(n);
    uint32_t* table = (uint32_t*)cur; cur += mg_align((size_t)(entries + ncta) * 4);
    uint32_t* starts = (uint32_t*)cur;
    if (count_only) CK(spk_launch_unpack(c, SPK_ATTR_POSITION, in.pos, n));
    else { int rc = mg_export_all(c, in, n); if (rc) return rc; }
    mg_classify_kernel<<<ncta, MG_THREADS, 0, c->stream>>>(in.pos, (uint32_t)n, c->cell_width, B, bucket_of, table, ncta);
    mg_scan_kernel<<<1, MG_SCAN_THREADS, 0, c->stream>>>(table, entries, ncta, nbuckets, starts);
    c->launches += 2;
    CK(cudaGetLastError());
    uint32_t h_starts[MG_MAX_BUCKETS + 1];
    CK(cudaMemcpyAsync(h_starts, starts, (nbuckets + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (uint32_t r = 0; r < world; ++r) counts_out[r] = h_starts[r + 2] - h_starts[r + 1];
    const size_t n_stay = h_starts[1], n_leave = n - n_stay;
    if (count_only || n_leave == 0) return SPK_OK; // nothing changes (the sort state stays valid)
    if (n_leave > capacity) return SPK_EOVERFLOW;
    mg_scatter_kernel<<<ncta, MG_THREADS, 0, c->stream>>>(in, (uint32_t)n, bucket_of, table, ncta, world + 1, (uint32_t)n_stay, out, dev_records);
    c->launches += 1;
    CK(cudaGetLastError());
    int rc = mg_ensure_gids(c, 0);
    if (rc) return rc;
    if ((rc = spk_particles_resize(c, n_stay, true))) return rc; // n_stay <= capacity: nothing is reallocated
    CK(spk_launch_pack_upload(c, out.status, out.pos, out.vel, out.volume, out.phase, out.data, 0, n_stay));
    if (n_stay) CK(cudaMemcpyAsync(c->gid, out.gid, n_stay * 4, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream)); // the records are complete and the temporaries can go
    return SPK_OK;
}

int spk_particles_migrate_in(spk_ctx* c, const SpkParticleRecord* dev_records, size_t m) {
    if (!c || (m && !dev_records)) return SPK_EINVAL;
    if (m == 0) { // a rank that receives nothing in a collective migration still takes the path of an edited set in the next step
        c->sorted_valid = false; // (the stand-alone dt kernel of the peer-memory pipeline ends in a barrier every rank must reach)
        c->dt_cache_valid = false;
        return SPK_OK;
    }
    CK(cudaSetDevice(c->device));
    const size_t old_n = c->n, new_n = old_n + m;
    if (new_n >= 0xffffffffull) return SPK_EOVERFLOW;
    const bool grow = new_n > c->cap;
    DevTemp tmp;
    CK(mg_scratch(c, tmp, mg_aos_bytes(m) + (grow ? mg_aos_bytes(old_n) : 0)));
    char* cur = tmp.p;
    MgAos inc = mg_carve(cur, m);
    mg_split_kernel<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(dev_records, (uint32_t)m, inc);
    c->launches += 1;
    CK(cudaGetLastError());
    int rc;
    if (grow) { // device-side growth: park the residents in reference layout, reallocate the streams, pack them back
        MgAos old = mg_carve(cur, old_n);
        if ((rc = mg_export_all(c, old, old_n))) return rc;
        if ((rc = spk_particles_resize(c, new_n, true))) return rc; // reallocates (frees the id table too)
        if ((rc = mg_ensure_gids(c, 0))) return rc;
        CK(spk_launch_pack_upload(c, old.status, old.pos, old.vel, old.volume, old.phase, old.data, 0, old_n));
        if (old_n) CK(cudaMemcpyAsync(c->gid, old.gid, old_n * 4, cudaMemcpyDeviceToDevice, c->stream));
    } else {
        if ((rc = mg_ensure_gids(c, old_n))) return rc;
        if ((rc = spk_particles_resize(c, new_n, false))) return rc;
    }
    CK(spk_launch_pack_upload(c, inc.status, inc.pos, inc.vel, inc.volume, inc.phase, inc.data, old_n, m));
    CK(cudaMemcpyAsync(c->gid + old_n, inc.gid, m * 4, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SPK_OK;
}
