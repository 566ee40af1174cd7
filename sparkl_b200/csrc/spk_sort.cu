// spk_sort.cu — particle binning, sparse-grid block activation, block ids / scans, neighbour tables, permutation.
//
// Replaces K1-K12 of the reference (src_kernels/cuda/sort.rs:58-280, hashmap.rs:37-96,
// prefix_sum.rs, sequenced by src/cuda/cuda_sparse_grid.rs:147-489).  Redesign:
//  * ONE pass over positions (the reference reads them three times: K2, K5, K10): a warp first de-duplicates its
//    block keys with match.any (particles are kept physically block-sorted, so a warp usually holds 1-2 keys), the
//    group leader inserts the key and reserves ranks with one atomicAdd per (warp, block).
//  * the 8 associated blocks are activated per occupied block, not per particle (8 CAS per particle in K2).
//  * block ids are the rank of the hash slot (reduce-then-scan over the table), so ids need no atomic counter and the
//    three scans of the reference (first_particle, dispatch base, id) collapse into one.
//  * neighbour-block tables (ids of key + {0,1}^3 in the next and in the previous grid) are built once per block so
//    g2p2g / reset never probe the hash map.
#include "spk_device.cuh"

#ifdef SPK_EMU // the CUDA emulation of the CPU test suite pays per emulated thread, not per unit of latency: coarse work items there
#ifndef SPK_CELL_SCAN_SLOTS
#define SPK_CELL_SCAN_SLOTS 16
#endif
#ifndef SPK_SCAN_ITEMS
#define SPK_SCAN_ITEMS 16
#endif
#endif
#ifndef SPK_CELL_SCAN_SLOTS
#define SPK_CELL_SCAN_SLOTS 2
#endif
#ifndef SPK_MORTON_FACTOR
#define SPK_MORTON_FACTOR 4 /* Morton-table entries per hash slot (power of two) */
#endif
static_assert(SPK_MORTON_FACTOR == 1 || SPK_MORTON_FACTOR == 2 || SPK_MORTON_FACTOR == 4,
              "ensure_slot_scratch / ensure_staging (spk_api.cu) size the Morton table and the scan partials for at most 4 entries per hash slot");
#ifndef SPK_PERM_ITEMS
#define SPK_PERM_ITEMS 4
#endif
namespace {

constexpr int BIN_THREADS = 256;
constexpr int SCAN_THREADS = 256;
#ifndef SPK_SCAN_ITEMS
#define SPK_SCAN_ITEMS 2
#endif
constexpr int SCAN_ITEMS = SPK_SCAN_ITEMS; // slots per thread (few: the table has only tens of thousands of slots, the scan wants CTAs, not long threads)
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr int SPK_RANK_BITS = 32 - SPK_CELL_SHIFT; // rank_of = cell << SPK_RANK_BITS | rank inside the cell
constexpr uint32_t SPK_RANK_MASK = (1u << SPK_RANK_BITS) - 1u;

// ---- Morton-ordered block ids.  Block ids are ranks in a scan (no atomic counter); the scan runs over a dense table indexed by the
// Morton code of the block coordinate relative to the bounding box of the active blocks (`mslot`: code -> hash slot + 1), so ids —
// and with them the particle order, the per-block grid chunks and the dispatch order of g2p2g — follow a Z-curve through space
// (the interleave of src_core/utils/morton3.rs:3-15, generalised to per-axis bit counts so a flat or elongated domain still fits).
// The box is the one of the PREVIOUS sort, grown by the one block per substep a particle can travel (+1 for the {0,1}^3 activation);
// a block outside it (first sort, upload in between, ...) raises `fail` and this sort ranks the hash slots instead: ids are then valid
// but unordered, and the next sort has a fresh box.  mstate: [0..2] origin, [3..5] bits per axis, [6] valid, [7] fail.
__device__ __forceinline__ uint32_t spk_morton_index(unsigned long long key, const uint32_t* ms) {
#if SPK_DIM == 3
    unsigned long long k[3];
    spk_unpack_key(key, k[0], k[1], k[2]);
#else
    unsigned long long k[2];
    spk_unpack_key(key, k[0], k[1]);
#endif
    uint32_t d[SPK_DIM], bits[SPK_DIM];
    uint32_t maxbits = 0;
#pragma unroll
    for (int a = 0; a < SPK_DIM; ++a) {
        d[a] = (uint32_t)k[a] - ms[a];
        bits[a] = ms[3 + a];
        if (k[a] < ms[a] || (k[a] - ms[a]) >> bits[a]) return 0xffffffffu;
        maxbits = max(maxbits, bits[a]);
    }
    uint32_t out = 0, pos = 0;
    for (uint32_t l = 0; l < maxbits; ++l)
#pragma unroll
        for (int a = 0; a < SPK_DIM; ++a)
            if (l < bits[a]) out |= ((d[a] >> l) & 1u) << pos++;
    return out;
}
__device__ __forceinline__ void spk_morton_register(uint32_t* mslot, uint32_t mcap, uint32_t* ms, unsigned long long key, uint32_t slot) {
    if (!ms[6] || slot == SPK_NO_BLOCK) return;
    const uint32_t m = spk_morton_index(key, ms);
    if (m >= mcap) ms[7] = 1u; // outside the box (0xffffffff), or a box sized for the other grid buffer's larger table
    else mslot[m] = slot + 1u;
}
// the hash slot behind entry i of the scan domain (Morton table, or the hash table itself when there is no usable box)
__device__ __forceinline__ uint32_t spk_scan_slot(uint32_t i, bool morton, const uint32_t* __restrict__ mslot, const unsigned long long* __restrict__ keys,
                                                  uint32_t cap_hash) {
    if (morton) return mslot[i] - 1u; // 0 -> SPK_NO_BLOCK
    return (i < cap_hash && keys[i] != SPK_EMPTY_KEY) ? i : SPK_NO_BLOCK;
}

// K1: one pass over the positions.  Per particle: block key + cell inside the block; per (warp, block) one hash insert; per
// (warp, block, cell) one atomicAdd on the cell counter of the block's hash slot, which also hands out the particle's rank inside
// its cell.  Sorting by (block, cell) instead of by block alone costs nothing extra here and makes every g2p2g warp read a handful
// of cells (shared-memory broadcasts in G2P, few P2G bins per CTA).
__global__ void __launch_bounds__(BIN_THREADS) bin_particles_kernel(const float4* __restrict__ a0, uint32_t n, float h, float inv_h, unsigned long long* keys,
                                                                     uint32_t cap_hash, uint32_t* cell_cnt, uint32_t* __restrict__ slot_of,
                                                                     uint32_t* __restrict__ rank_of, DeviceCounters* counters) {
    const uint32_t i = blockIdx.x * BIN_THREADS + threadIdx.x;
    const bool valid = i < n;
    const unsigned active = __ballot_sync(0xffffffffu, valid);
    if (!valid) return;
    const float4 p = a0[i];
    // associated cells once (correctly rounded p / h through the host's RN(1 / h): bit-identical to div.rn, spk_div_h), then block key and
    // cell inside the block from the same integers
    const long long acx = spk_assoc_cell_m(p.x, h, inv_h), acy = spk_assoc_cell_m(p.y, h, inv_h);
#if SPK_DIM == 3
    const long long acz = spk_assoc_cell_m(p.z, h, inv_h);
    const unsigned long long key = spk_pack_key(spk_block_coord(acx), spk_block_coord(acy), spk_block_coord(acz));
    const uint32_t cell = ((uint32_t)(acx - 2) & 3u) | (((uint32_t)(acy - 2) & 3u) << 2) | (((uint32_t)(acz - 2) & 3u) << 4);
#else
    const unsigned long long key = spk_pack_key(spk_block_coord(acx), spk_block_coord(acy));
    const uint32_t cell = ((uint32_t)(acx - 2) & 3u) | (((uint32_t)(acy - 2) & 3u) << 2);
#endif
    const unsigned group = __match_any_sync(active, key);
    const unsigned cgroup = group & __match_any_sync(active, (unsigned long long)cell);
    const int leader = __ffs(group) - 1, cleader = __ffs(cgroup) - 1;
    const int lane = threadIdx.x & 31;
    uint32_t slot = SPK_NO_BLOCK, base = 0;
    if (lane == leader) {
        slot = spk_hash_insert(keys, cap_hash, key);
        if (slot == SPK_NO_BLOCK) counters->overflow = 1;
    }
    slot = __shfl_sync(group, slot, leader);
    if (lane == cleader && slot != SPK_NO_BLOCK) base = atomicAdd(&cell_cnt[(size_t)slot * SPK_CELLS_PER_BLOCK + cell], (uint32_t)__popc(cgroup));
    base = __shfl_sync(cgroup, base, cleader);
    const uint32_t rank = base + __popc(cgroup & spk_lanemask_lt());
    if (rank >> SPK_RANK_BITS) counters->overflow = 1; // > 2^26 particles in one cell
    slot_of[i] = slot;
    rank_of[i] = (cell << SPK_RANK_BITS) | (rank & SPK_RANK_MASK);
}

// per hash slot (one warp each): exclusive scan of the slot's cell counters (in place: they become the cell offsets inside the
// block), particle count of the block, and touch_particle_blocks (sort.rs:58-72) done per occupied block: activate key + {0,1}^3
// (gpu_grid.rs:307-321).  Slots created here own no particles: their counters are all zero whether or not their warp sees the key.
constexpr int CELL_SCAN_SLOTS = SPK_CELL_SCAN_SLOTS; // hash slots per warp (most are empty: the table is at most half full)
__global__ void __launch_bounds__(256) cell_scan_kernel(unsigned long long* keys, uint32_t cap_hash, uint32_t* cell_cnt, uint32_t* __restrict__ slot_cnt,
                                                          DeviceCounters* counters, uint32_t* bbox, uint32_t* mslot, uint32_t mcap) {
    const int lane = threadIdx.x & 31;
    uint32_t* ms = bbox + 8;
    if (blockIdx.x == 0 && threadIdx.x < 6) bbox[threadIdx.x] = threadIdx.x < 3 ? 0xffffffffu : 0u; // empty box for assign_ids_kernel's min / max
    const uint32_t s0 = (blockIdx.x * 8u + (threadIdx.x >> 5)) * CELL_SCAN_SLOTS;
    for (uint32_t s = s0; s < s0 + CELL_SCAN_SLOTS && s < cap_hash; ++s) {
        // lane 0's view of the slot decides for the warp: the lanes that activated neighbours in the previous iteration arrive here
        // later than the others, and a key inserted in between (by any warp) must not split the warp at the `continue`s below
        const unsigned long long key = __shfl_sync(0xffffffffu, keys[s], 0);
        if (key == SPK_EMPTY_KEY) continue;
        uint32_t* cc = cell_cnt + (size_t)s * SPK_CELLS_PER_BLOCK;
#if SPK_DIM == 3
        const uint32_t c0 = cc[2 * lane], c1 = cc[2 * lane + 1]; // lane owns cells 2 lane, 2 lane + 1
        const uint32_t mine = c0 + c1;
#else
        const uint32_t c0 = lane < SPK_CELLS_PER_BLOCK ? cc[lane] : 0u;
        const uint32_t mine = c0;
#endif
        uint32_t inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
        if (total == 0) continue; // a block activated as a neighbour only
#if SPK_DIM == 3
        cc[2 * lane] = inc - mine;
        cc[2 * lane + 1] = inc - mine + c0;
#else
        if (lane < SPK_CELLS_PER_BLOCK) cc[lane] = inc - mine;
#endif
        if (lane == 0) { slot_cnt[s] = total; spk_morton_register(mslot, mcap, ms, key, s); }
        if (lane >= 1 && lane < SPK_NBR) {
            const unsigned long long nkey = spk_key_offset(key, lane);
            const uint32_t ns = spk_hash_insert(keys, cap_hash, nkey);
            if (ns == SPK_NO_BLOCK) counters->overflow = 1;
            spk_morton_register(mslot, mcap, ms, nkey, ns);
        }
    }
}

struct Triple { uint32_t occ, cnt, mult; };
__device__ __forceinline__ Triple operator+(Triple a, Triple b) { return Triple{a.occ + b.occ, a.cnt + b.cnt, a.mult + b.mult}; }

__device__ __forceinline__ Triple block_exclusive_scan(Triple v, Triple& total) {
    __shared__ Triple warp_tot[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    Triple inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Triple o{__shfl_up_sync(0xffffffffu, inc.occ, d), __shfl_up_sync(0xffffffffu, inc.cnt, d), __shfl_up_sync(0xffffffffu, inc.mult, d)};
        if (lane >= d) inc = inc + o;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    Triple base{0, 0, 0}, tot{0, 0, 0};
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        if (w < wid) base = base + warp_tot[w];
        tot = tot + warp_tot[w];
    }
    __syncthreads();
    total = tot;
    return Triple{base.occ + inc.occ - v.occ, base.cnt + inc.cnt - v.cnt, base.mult + inc.mult - v.mult};
}

__device__ __forceinline__ Triple slot_triple(const uint32_t* slot_cnt, uint32_t s, uint32_t chunk) { // s: hash slot or SPK_NO_BLOCK
    Triple t{0, 0, 0};
    if (s != SPK_NO_BLOCK) {
        t.occ = 1;
        t.cnt = slot_cnt[s];
        t.mult = (t.cnt + chunk - 1) / chunk;
    }
    return t;
}

// pass 1 of reduce-then-scan: per-CTA totals of (occupied slots, particles, dispatch entries)
__global__ void __launch_bounds__(SCAN_THREADS) slot_totals_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ slot_cnt,
                                                                    uint32_t cap_hash, uint32_t chunk, Triple* partials, const uint32_t* __restrict__ mslot,
                                                                    const uint32_t* __restrict__ ms) {
    Triple acc{0, 0, 0};
    const bool morton = ms[6] && !ms[7];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) acc = acc + slot_triple(slot_cnt, spk_scan_slot(base + k, morton, mslot, keys, cap_hash), chunk);
    Triple total;
    block_exclusive_scan(acc, total);
    if (threadIdx.x == 0) partials[blockIdx.x] = total;
}

// pass 2: ids = rank of the slot, first_particle / dispatch base = exclusive prefix sums (replaces K6-K9, K11, K12)
__global__ void __launch_bounds__(SCAN_THREADS) assign_ids_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ slot_cnt,
                                                                   uint32_t cap_hash, uint32_t chunk, const Triple* __restrict__ partials, uint32_t* vals,
                                                                   GridBuffers g, DispatchEntry* dispatch, uint32_t dispatch_cap, DeviceCounters* counters, uint32_t* bbox,
                                                                   const uint32_t* __restrict__ mslot) {
    __shared__ Triple s_base;
    const uint32_t* ms = bbox + 8;
    const bool morton = ms[6] && !ms[7];
    // redundant reduction of the partials of all previous CTAs (at most a few hundred values)
    Triple pre{0, 0, 0};
    for (uint32_t j = threadIdx.x; j < blockIdx.x; j += SCAN_THREADS) pre = pre + partials[j];
    Triple pre_tot;
    block_exclusive_scan(pre, pre_tot);
    if (threadIdx.x == 0) s_base = pre_tot;
    __syncthreads();
    const Triple cta_base = s_base;

    Triple item[SCAN_ITEMS];
    uint32_t slot_of_item[SCAN_ITEMS];
    Triple acc{0, 0, 0};
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        slot_of_item[k] = spk_scan_slot(base + k, morton, mslot, keys, cap_hash);
        item[k] = slot_triple(slot_cnt, slot_of_item[k], chunk);
        acc = acc + item[k];
    }
    Triple total;
    Triple run = block_exclusive_scan(acc, total) + cta_base;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (item[k].occ) {
            const uint32_t s = slot_of_item[k], id = run.occ;
            vals[s] = id;
            if (id < g.cap_blocks) {
                g.blk_key[id] = keys[s];
                g.blk_first[id] = run.cnt;
                g.blk_count[id] = item[k].cnt;
                g.blk_halo[id] = 0; // HaloState::EMPTY, gpu_grid.rs:353-368
                g.blk_dfirst[id] = run.mult;
            } else {
                counters->overflow = 1;
            }
            for (uint32_t m = 0; m < item[k].mult; ++m) {
                const uint32_t d = run.mult + m;
                if (d < dispatch_cap) dispatch[d] = DispatchEntry{id, run.cnt + m * chunk, min(chunk, item[k].cnt - m * chunk), 0};
                else counters->overflow = 1;
            }
        }
        run = run + item[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) {
        counters->nab = run.occ;
        counters->ndispatch = run.mult;
        counters->nparticles_binned = run.cnt;
    }
    { // bounding box of the active blocks (raw key coordinates; the next sort's Morton box, and what the peers of a multi-GPU run are told):
      // thread -> warp -> CTA, then six atomics per CTA
        __shared__ uint32_t s_box[SCAN_THREADS / 32][6];
        uint32_t lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k)
            if (item[k].occ) {
                unsigned long long kx, ky, kz = 0;
#if SPK_DIM == 3
                spk_unpack_key(keys[slot_of_item[k]], kx, ky, kz);
#else
                spk_unpack_key(keys[slot_of_item[k]], kx, ky);
#endif
                lo[0] = min(lo[0], (uint32_t)kx); lo[1] = min(lo[1], (uint32_t)ky); lo[2] = min(lo[2], (uint32_t)kz);
                hi[0] = max(hi[0], (uint32_t)kx); hi[1] = max(hi[1], (uint32_t)ky); hi[2] = max(hi[2], (uint32_t)kz);
            }
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = __reduce_min_sync(0xffffffffu, lo[a]); hi[a] = __reduce_max_sync(0xffffffffu, hi[a]); }
        if ((threadIdx.x & 31) == 0)
            for (int a = 0; a < 3; ++a) { s_box[threadIdx.x >> 5][a] = lo[a]; s_box[threadIdx.x >> 5][3 + a] = hi[a]; }
        __syncthreads();
        if (threadIdx.x < 6) {
            uint32_t v = s_box[0][threadIdx.x];
            for (int w = 1; w < SCAN_THREADS / 32; ++w) v = threadIdx.x < 3 ? min(v, s_box[w][threadIdx.x]) : max(v, s_box[w][threadIdx.x]);
            // (a CTA without blocks holds the neutral elements; the plain read first skips the atomic once the box already covers this CTA's)
            const uint32_t cur = *reinterpret_cast<volatile uint32_t*>(&bbox[threadIdx.x]);
            if (threadIdx.x < 3) { if (v < cur) atomicMin(&bbox[threadIdx.x], v); }
            else if (v > cur) atomicMax(&bbox[threadIdx.x], v);
        }
    }
}

// neighbour-block tables: for block b, ids of key + {0,1}^3 in the next grid (always present) and in the previous grid
// (launched for the grid's block CAPACITY before the host has read the block count back, so that the read-back overlaps this
// kernel and scatter_perm_kernel; the true count comes from the device-side record)
__global__ void build_tables_kernel(GridBuffers next, GridBuffers prev, DeviceCounters* counters, int32_t* __restrict__ dispatch_nbr,
                                    uint32_t dispatch_cap, uint32_t chunk, uint32_t* bbox, uint32_t mcap) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) { // the counter record of this sort has been published: reset the words the rest of the substep accumulates into (no memsets)
        counters->dt_raw_bits = SPK_DT_RAW_NONE; // g2p2g's fused CFL candidate (atomicMin)
        counters->num_halo = 0;                  // multi-GPU halo detection (atomicAdd)
        // the Morton box of the NEXT sort: this sort's bounding box grown by one block towards the origin and two away from it
        uint32_t* ms = bbox + 8;
        for (int k = 0; k < 8; ++k) bbox[16 + k] = ms[k]; // the box THIS sort used, for spk_grid_morton_box
        uint32_t total_bits = 0;
        bool ok = counters->nab > 0 && !counters->overflow;
        for (int a = 0; a < SPK_DIM; ++a) {
            ok = ok && bbox[a] >= 1u && bbox[3 + a] >= bbox[a];
            const uint32_t ext = bbox[3 + a] - bbox[a] + 3u;
            uint32_t b = 0;
            while (b < 31u && (1u << b) < ext) ++b;
            ms[a] = bbox[a] - 1u;
            ms[3 + a] = b;
            total_bits += b;
        }
        ms[6] = (ok && total_bits < 32u && (1u << total_bits) <= mcap) ? 1u : 0u;
        ms[7] = 0u;
    }
    const uint32_t b = t >> SPK_NBR_SHIFT, o = t & (SPK_NBR - 1);
    const bool live = b < counters->nab && b < next.cap_blocks;
    if (!live) return;
    const unsigned long long key = spk_key_offset(next.blk_key[b], (int)o);
    const int32_t ns = spk_hash_get(next, key), np = spk_hash_get(prev, key);
    next.nbr_self[b * SPK_NBR + o] = ns;
    next.nbr_prev[b * SPK_NBR + o] = np;
    // the same ids once more per g2p2g dispatch entry of the block: a CTA then reads its entry and its tile's block ids in one round trip
    const uint32_t d0 = next.blk_dfirst[b], nd = (next.blk_count[b] + chunk - 1) / chunk;
    for (uint32_t d = d0; d < d0 + nd && d < dispatch_cap; ++d) {
        dispatch_nbr[(size_t)d * (2 * SPK_NBR) + o] = ns;
        dispatch_nbr[(size_t)d * (2 * SPK_NBR) + SPK_NBR + o] = np;
    }
}

// finalize_particles_sort (sort.rs:201-221): sorted slot -> index in the current physical order; inside a block particles are ordered
// by cell
__global__ void __launch_bounds__(BIN_THREADS) scatter_perm_kernel(const uint32_t* __restrict__ slot_of, const uint32_t* __restrict__ rank_of, uint32_t n,
                                                                    const uint32_t* __restrict__ vals, const uint32_t* __restrict__ blk_first,
                                                                    const uint32_t* __restrict__ cell_off, uint32_t cap_blocks, uint32_t* __restrict__ perm) {
    // SPK_PERM_ITEMS particles per thread, every level of the dependent chain (slot -> id -> first particle, cell offset) loaded for all
    // of them before the next level is touched: the kernel is pure memory latency
    const uint32_t i0 = blockIdx.x * (BIN_THREADS * SPK_PERM_ITEMS) + threadIdx.x;
    uint32_t slot[SPK_PERM_ITEMS], cr[SPK_PERM_ITEMS], id[SPK_PERM_ITEMS], first[SPK_PERM_ITEMS], off[SPK_PERM_ITEMS];
#pragma unroll
    for (int k = 0; k < SPK_PERM_ITEMS; ++k) {
        const uint32_t i = i0 + k * BIN_THREADS;
        slot[k] = i < n ? slot_of[i] : SPK_NO_BLOCK;
        cr[k] = i < n ? rank_of[i] : 0u;
    }
#pragma unroll
    for (int k = 0; k < SPK_PERM_ITEMS; ++k) {
        id[k] = slot[k] != SPK_NO_BLOCK ? vals[slot[k]] : 0xffffffffu;
        off[k] = slot[k] != SPK_NO_BLOCK ? cell_off[(size_t)slot[k] * SPK_CELLS_PER_BLOCK + (cr[k] >> SPK_RANK_BITS)] : 0u;
    }
#pragma unroll
    for (int k = 0; k < SPK_PERM_ITEMS; ++k) first[k] = id[k] < cap_blocks ? blk_first[id[k]] : 0u;
#pragma unroll
    for (int k = 0; k < SPK_PERM_ITEMS; ++k)
        if (id[k] < cap_blocks) perm[first[k] + off[k] + (cr[k] & SPK_RANK_MASK)] = i0 + k * BIN_THREADS;
}

} // namespace

cudaError_t spk_launch_sort_bin(spk_ctx* c) {
    GridBuffers& g = c->grid[1 - c->gcur];
    cudaStream_t s = c->stream;
    cudaError_t e;
    if ((e = cudaMemsetAsync(g.keys, 0xff, (size_t)g.cap_hash * 8, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->slot_cnt, 0, (size_t)g.cap_hash * 4, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->cell_cnt, 0, (size_t)g.cap_hash * SPK_CELLS_PER_BLOCK * 4, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d_counters, 0, 16, s)) != cudaSuccess) return e; // nab, ndispatch, overflow, nparticles_binned
    const uint32_t mcap = g.cap_hash * SPK_MORTON_FACTOR; // entries of the Morton table (= of the scan domain)
    if ((e = cudaMemsetAsync(c->mslot, 0, (size_t)mcap * 4, s)) != cudaSuccess) return e;
    c->counters_fresh = false;
    const uint32_t n = (uint32_t)c->n;
    if (n) {
        bin_particles_kernel<<<(n + BIN_THREADS - 1) / BIN_THREADS, BIN_THREADS, 0, s>>>(c->pb[c->cur].a0, n, c->cell_width, 1.0f / c->cell_width, g.keys, g.cap_hash, c->cell_cnt,
                                                                                        c->slot_of, c->rank_of, c->d_counters);
        cell_scan_kernel<<<(g.cap_hash + 8 * CELL_SCAN_SLOTS - 1) / (8 * CELL_SCAN_SLOTS), 256, 0, s>>>(g.keys, g.cap_hash, c->cell_cnt, c->slot_cnt, c->d_counters, c->dist_bbox, c->mslot, mcap);
        c->launches += 2;
    } else { // no particles (a rank of a multi-GPU run may own none): the empty bounding box cell_scan_kernel would have written
        if ((e = cudaMemsetAsync(c->dist_bbox, 0xff, 12, s)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(c->dist_bbox + 3, 0, 12, s)) != cudaSuccess) return e;
    }
    const uint32_t ntiles = (mcap + SCAN_TILE - 1) / SCAN_TILE;
    if ((size_t)ntiles * sizeof(Triple) > c->staging_bytes) return cudaErrorInvalidValue; // sized by spk_stage_sort (ensure_staging) before the launch
    Triple* partials = (Triple*)c->staging;
    slot_totals_kernel<<<ntiles, SCAN_THREADS, 0, s>>>(g.keys, c->slot_cnt, g.cap_hash, c->chunk, partials, c->mslot, c->dist_bbox + 8);
    assign_ids_kernel<<<ntiles, SCAN_THREADS, 0, s>>>(g.keys, c->slot_cnt, g.cap_hash, c->chunk, partials, g.vals, g, c->dispatch, c->dispatch_cap, c->d_counters, c->dist_bbox,
                                                      c->mslot);
    c->launches += 2;
    return cudaGetLastError();
}

cudaError_t spk_launch_sort_finish(spk_ctx* c) {
    GridBuffers& g = c->grid[1 - c->gcur];
    GridBuffers& prev = c->grid[c->gcur];
    cudaStream_t s = c->stream;
    const uint32_t n = (uint32_t)c->n;
    {
        build_tables_kernel<<<(g.cap_blocks * SPK_NBR + 255) / 256, 256, 0, s>>>(g, prev, c->d_counters, c->dispatch_nbr, c->dispatch_cap, c->chunk, c->dist_bbox,
                                                                                 g.cap_hash * SPK_MORTON_FACTOR);
        c->launches += 1;
        { cudaError_t e = cudaEventRecord(c->ev_tables, s); if (e != cudaSuccess) return e; } // spk_step: the grid reset may start (side stream)
        // multi-GPU: this rank's hash table and bounding box are final -> tell the peers now (the arrive half of the barrier that precedes the
        // halo detection); its wait half runs ~60 us of independent work later (permutation, grid reset), when the peers' flags have landed
        if (spk_dist_active(c) && c->dist_split_barriers) {
            cudaError_t e = spk_dist_launch_barrier(c, 0, 1);
            if (e != cudaSuccess) return e;
            c->dist_s1_arrived = 1;
        }
    }
    if (n) {
        scatter_perm_kernel<<<(n + BIN_THREADS * SPK_PERM_ITEMS - 1) / (BIN_THREADS * SPK_PERM_ITEMS), BIN_THREADS, 0, s>>>(c->slot_of, c->rank_of, n, g.vals, g.blk_first, c->cell_cnt, g.cap_blocks, c->perm);
        c->launches += 1;
    }
    return cudaGetLastError();
}
