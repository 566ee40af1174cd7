// spk_dist.cu — multi-GPU halo accumulation over NVLink peer memory, one process per GPU.
//
// Replaces the reference's multi-GPU machinery: the host-side gather of every device's active block
// headers (src/cuda/cuda_sparse_grid.rs:256-322), K3/K4 halo tagging, K17 `copy_halo_to_staging`, one
// cuMemcpyDtoDAsync per device pair (src/cuda/cuda_mpm_pipeline.rs:458-472), K18 `merge_halo_blocks` and the host loop
// that takes the dt minimum (:393-398).  Semantics are unchanged: a block active on several ranks ends the substep with
// the full sum on every one of them, dt is the minimum over ranks.
//
// B200 design: every rank shares ONE device arena with its peers through CUDA IPC (mapped over NVLink / NVSwitch):
//     [mailbox | hash keys, hash vals, halo accumulators (hacc) of grid buffer 0 | the same for grid buffer 1]
// and all cross-GPU traffic is device-initiated:
//   * detect : every rank probes its peers' hash tables directly (P2P loads) for its own active block keys — no key
//              lists travel, no host round trip; the result is a per-block table of the peer's block ids.
//   * push   : after G2P2G a rank adds the accumulators of its shared blocks straight into the peers' `hacc` arrays with
//              NVLink atomics (red.add.f32), addressed by the peer's own block ids — no pack / staging / unpack.
//   * merge  : each rank folds its `hacc` into its grid and clears it.
//   * barrier: a one-CTA kernel writes an epoch flag (+ the dt candidate) into every peer's mailbox and spins on its own
//              mailbox; two per substep.  The dt minimum rides on the second one.
// The host never blocks for the exchange: the only read-back per substep is the 64-byte counter record of the sort.
#include "spk_internal.cuh"
#if SPK_DIM == 2 || defined(SPK_EMU)
// dim2 build (and the host emulation of tests/host/emu/): the NVLink peer-memory path is not built (the 2D scenes of the reference are single-GPU sized); the entry points
// exist so both libraries export the same symbols, and fail loudly.
bool spk_dist_active(const spk_ctx*) { return false; }
bool spk_dist_grid_storage(spk_ctx*, int, uint32_t, unsigned long long**, uint32_t**) { return false; }
cudaError_t spk_dist_launch_barrier(spk_ctx*, int, int) { return cudaErrorNotSupported; }
cudaError_t spk_dist_launch_detect(spk_ctx*) { return cudaErrorNotSupported; }
cudaError_t spk_dist_launch_push(spk_ctx*) { return cudaErrorNotSupported; }
cudaError_t spk_dist_launch_merge(spk_ctx*) { return cudaErrorNotSupported; }
float* spk_dist_hacc_of_next(spk_ctx*) { return nullptr; }
extern "C" {
int spk_dist_export(spk_ctx*, SpkDistDescriptor*) { return SPK_EUNSUPPORTED; }
int spk_dist_attach(spk_ctx*, uint32_t, uint32_t, const SpkDistDescriptor*) { return SPK_EUNSUPPORTED; }
int spk_dist_detach(spk_ctx* c) { return c ? SPK_OK : SPK_EINVAL; }
int spk_dist_info(spk_ctx*, uint32_t*, uint32_t*) { return SPK_EUNSUPPORTED; }
}
#else
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "spk_device.cuh"

#define CK(expr)                               \
    do {                                       \
        cudaError_t _e = (expr);               \
        if (_e != cudaSuccess) return (int)_e; \
    } while (0)

namespace {

constexpr size_t MAILBOX_BYTES = 4096;
constexpr size_t ALIGN = 1024;
__host__ __device__ inline size_t align_up(size_t v) { return (v + ALIGN - 1) / ALIGN * ALIGN; }
// arena layout as a function of the hash capacity (identical arithmetic on the owner and on every peer)
__host__ __device__ inline size_t arena_grid_bytes(uint32_t cap_hash) {
    return align_up((size_t)cap_hash * 8) + align_up((size_t)cap_hash * 4) + align_up((size_t)(cap_hash / 2) * SPK_ACC_CHANNELS * 64 * 4);
}
__host__ __device__ inline size_t arena_bytes(uint32_t cap_hash) { return MAILBOX_BYTES + 2 * arena_grid_bytes(cap_hash); }
__host__ __device__ inline DistGridView arena_grid(char* base, uint32_t cap_hash, int g) {
    char* p = base + MAILBOX_BYTES + (size_t)g * arena_grid_bytes(cap_hash);
    DistGridView v;
    v.keys = reinterpret_cast<unsigned long long*>(p);
    p += align_up((size_t)cap_hash * 8);
    v.vals = reinterpret_cast<uint32_t*>(p);
    p += align_up((size_t)cap_hash * 4);
    v.hacc = reinterpret_cast<float*>(p);
    return v;
}

struct DescriptorWire { // what travels in SpkDistDescriptor (opaque to the caller)
    cudaIpcMemHandle_t handle;
    uint32_t cap_hash;
    uint32_t device;
    unsigned long long bytes;
    uint32_t magic;
};
static_assert(sizeof(DescriptorWire) <= sizeof(SpkDistDescriptor), "descriptor too small");
constexpr uint32_t MAGIC = 0x53504b44u; // "SPKD"

__device__ __forceinline__ void st_sys_u32(volatile uint32_t* p, uint32_t v) { asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_sys_u32(const volatile uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Cross-GPU barrier + small all-to-all of one 64-bit value.  One CTA; thread p talks to peer p.
//   mode 0: plain barrier, publishes `gnext` (which grid buffer is this rank's next grid)
//   mode 1: additionally min-reduces counters->dt_raw_bits over the ranks (fused CFL reduction of g2p2g)
//   mode 2: additionally min-reduces counters->dt_acc over the ranks (stand-alone estimate_dt kernel)
//   phase bit 0: ARRIVE (payload + flag into every peer's mailbox), bit 1: WAIT (spin on this rank's mailbox, reduce).  spk_step issues
//   the two halves apart, with independent work in between, so the flags usually have landed when the wait half runs
__global__ void dist_barrier_kernel(DistState d, uint32_t epoch, uint32_t gnext, int mode, int phase, DeviceCounters* counters, long long timeout_cycles,
                                    DeviceCounters* host_record, const uint32_t* __restrict__ bbox) {
    const uint32_t p = threadIdx.x;
    const uint32_t slot = epoch & 1u;
    unsigned long long mine = 0;
    if (mode == 1) mine = counters->dt_raw_bits;
    else if (mode == 2) mine = counters->dt_acc;
    __shared__ unsigned long long s_val[SPK_MAX_RANKS];
    if (p < d.world) {
        s_val[p] = mine;
        if (p != d.rank && (phase & 1)) {
            DistMailbox* remote = d.peer_mailbox[p];
            remote->val[slot][d.rank] = mine;
            remote->gnext[slot][d.rank] = gnext;
            if (mode == 0) // the barrier before the detection carries the bounding box of this rank's active blocks
                for (int k = 0; k < 6; ++k) remote->bbox[slot][d.rank][k] = bbox[k];
            __threadfence_system(); // payload (and everything this rank wrote before, incl. NVLink atomics) before the flag
            st_sys_u32(&remote->flag[d.rank], epoch);
        }
        if (p != d.rank && (phase & 2)) {
            const long long t0 = clock64();
            while ((int)(ld_sys_u32(&d.mailbox->flag[p]) - epoch) < 0) {
                // a peer that died must not hang the GPU: give up, and leave a STICKY mark (pad[1] is outside the words the sort clears
                // every substep) that the host turns into SPK_EPEER at its next look at the counters and at the end of spk_step
                if (timeout_cycles > 0 && clock64() - t0 > timeout_cycles) {
                    counters->pad[1] = 0x100u + p;
                    host_record->pad[1] = 0x100u + p; // mapped pinned memory: the host sees it without another read-back
                    __threadfence_system();
                    break;
                }
                __nanosleep(64);
            }
            __threadfence_system();
            s_val[p] = __ldcv(&d.mailbox->val[slot][p]);
        }
    }
    __syncthreads();
    if (p == 0 && mode != 0 && (phase & 2)) {
        unsigned long long m = s_val[0];
        for (uint32_t r = 1; r < d.world; ++r) m = s_val[r] < m ? s_val[r] : m;
        if (mode == 1) counters->dt_raw_bits = (uint32_t)m;
        else counters->dt_acc = m;
    }
}

// For every active block of this rank and every peer: the peer's id of the same block key (probe of the peer's hash
// table over NVLink), -1 when the peer does not hold it.  Shared blocks are tagged IS_HALO (tag_halo_blocks, sort.rs:74-96).
__global__ void __launch_bounds__(256) dist_detect_kernel(DistState d, GridBuffers next, uint32_t nab, uint32_t epoch, int32_t* peer_ids, float** push_ptrs,
                                                          DeviceCounters* counters) {
    const uint32_t t = blockIdx.x * 256 + threadIdx.x;
    if (t < d.world) // where each rank's halo accumulator of THIS substep lives (for the push fused into g2p2g's flush)
        push_ptrs[t] = t == d.rank ? nullptr : d.peer_grid[t][__ldcv(&d.mailbox->gnext[epoch & 1u][t]) & 1u].hacc;
    const uint32_t b = t / d.world, p = t % d.world;
    if (b >= nab) return;
    int32_t id = -1;
    if (p != d.rank) {
        const uint32_t g = __ldcv(&d.mailbox->gnext[epoch & 1u][p]) & 1u;
        const unsigned long long* keys = d.peer_grid[p][g].keys;
        const uint32_t cap = d.peer_cap_hash[p], mask = cap - 1;
        const unsigned long long key = next.blk_key[b];
        // geometric pre-filter: only a block inside the peer's bounding box can be active there too — for slabs that is the
        // layer of blocks at the common face; every other block skips the probe of the peer's hash table over NVLink
        unsigned long long kx, ky, kz;
        spk_unpack_key(key, kx, ky, kz);
        const uint32_t* bb = const_cast<const uint32_t*>(d.mailbox->bbox[epoch & 1u][p]);
        const bool maybe = (uint32_t)kx >= __ldcv(&bb[0]) && (uint32_t)kx <= __ldcv(&bb[3]) && (uint32_t)ky >= __ldcv(&bb[1]) && (uint32_t)ky <= __ldcv(&bb[4]) &&
                           (uint32_t)kz >= __ldcv(&bb[2]) && (uint32_t)kz <= __ldcv(&bb[5]);
        if (maybe) {
        // GridHashMap::get (hashmap.rs:72-87) on the PEER's table: volatile loads, never served from a stale L1 line
        uint32_t slot = spk_hash(key, mask);
        for (uint32_t probe = 0; probe < cap; ++probe) {
            const unsigned long long k = __ldcv(&keys[slot]);
            if (k == key) { id = (int32_t)__ldcv(&d.peer_grid[p][g].vals[slot]); break; }
            if (k == SPK_EMPTY_KEY) break;
            slot = (slot + 1) & mask;
        }
        if ((uint32_t)id >= cap / 2) id = -1;
        if (id >= 0 && atomicExch(&next.blk_halo[b], 1u) == 0u) atomicAdd(&counters->num_halo, 1u);
        }
    }
    peer_ids[(size_t)b * d.world + p] = id;
}

// One CTA per active block; shared blocks add their six accumulated channels into each holder's `hacc` (NVLink red.add).
__global__ void __launch_bounds__(128) dist_push_kernel(DistState d, GridBuffers next, uint32_t nab, uint32_t epoch, const int32_t* __restrict__ peer_ids) {
    const uint32_t b = blockIdx.x;
    if (b >= nab || !(next.blk_halo[b] & 1u)) return;
    const float* acc = next.acc + (size_t)b * (SPK_ACC_CHANNELS * 64);
    float v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] = acc[threadIdx.x + k * 128];
    for (uint32_t p = 0; p < d.world; ++p) {
        const int32_t id = peer_ids[(size_t)b * d.world + p];
        if (id < 0) continue;
        const uint32_t g = __ldcv(&d.mailbox->gnext[epoch & 1u][p]) & 1u;
        float* dst = d.peer_grid[p][g].hacc + (size_t)id * (SPK_ACC_CHANNELS * 64);
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (v[k] != 0.f) atomicAdd(dst + threadIdx.x + k * 128, v[k]);
    }
    __threadfence_system();
}

// merge_halo_blocks (sort.rs:142-164): fold what the peers pushed into this rank's grid, clear the staging for the next use
__global__ void __launch_bounds__(128) dist_merge_kernel(GridBuffers next, uint32_t nab, float* hacc) {
    const uint32_t b = blockIdx.x;
    if (b >= nab || !(next.blk_halo[b] & 1u)) return;
    float* acc = next.acc + (size_t)b * (SPK_ACC_CHANNELS * 64);
    float* h = hacc + (size_t)b * (SPK_ACC_CHANNELS * 64);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t i = threadIdx.x + k * 128;
        acc[i] += __ldcg(&h[i]);
        h[i] = 0.f;
    }
}

} // namespace

bool spk_dist_active(const spk_ctx* c) { return c->dist.world > 1; }

// hash keys / vals of grid buffer `g` live in the shared arena once spk_dist_export has run
bool spk_dist_grid_storage(spk_ctx* c, int g, uint32_t cap_hash, unsigned long long** keys, uint32_t** vals) {
    if (!c->dist_arena || cap_hash != c->dist_cap_hash) return false;
    DistGridView v = arena_grid(c->dist_arena, c->dist_cap_hash, g);
    *keys = v.keys;
    *vals = v.vals;
    return true;
}

// phase 3: the whole barrier; 1: arrive only (takes a new epoch, remembered per mode); 2: the wait half of the last arrive of this mode.
// Runs on the context's current stream (spk_step points it at the grid-update stream for the wait half of the second barrier).
cudaError_t spk_dist_launch_barrier(spk_ctx* c, int mode, int phase) {
    DistState& d = c->dist;
    if (phase & 1) {
        d.epoch += 1;
        c->dist_epoch_of[mode] = d.epoch;
        c->dist_gnext_of[mode] = (uint32_t)(1 - c->gcur);
    }
    dist_barrier_kernel<<<1, SPK_MAX_RANKS, 0, c->stream>>>(d, c->dist_epoch_of[mode], c->dist_gnext_of[mode], mode, phase, c->d_counters, c->dist_timeout_cycles,
                                                            c->h_counters_dev, c->dist_bbox);
    c->launches += 1;
    c->counters_fresh = false;
    return cudaGetLastError();
}
cudaError_t spk_dist_launch_detect(spk_ctx* c) {
    DistState& d = c->dist;
    GridBuffers& next = c->grid[1 - c->gcur];
    const uint32_t nab = next.nab_host;
    if (!nab) return cudaSuccess; // (num_halo was zeroed by this substep's sort: build_tables_kernel)
    dist_detect_kernel<<<(nab * d.world + 255) / 256, 256, 0, c->stream>>>(d, next, nab, c->dist_epoch_of[0], c->dist_peer_ids, c->dist_push_ptrs, c->d_counters);
    c->launches += 1;
    return cudaGetLastError();
}
cudaError_t spk_dist_launch_push(spk_ctx* c) {
    DistState& d = c->dist;
    GridBuffers& next = c->grid[1 - c->gcur];
    if (!next.nab_host) return cudaSuccess;
    dist_push_kernel<<<next.nab_host, 128, 0, c->stream>>>(d, next, next.nab_host, c->dist_epoch_of[0], c->dist_peer_ids);
    c->launches += 1;
    return cudaGetLastError();
}
float* spk_dist_hacc_of_next(spk_ctx* c) { return c->dist_arena && c->dist.world > 1 ? arena_grid(c->dist_arena, c->dist_cap_hash, 1 - c->gcur).hacc : nullptr; }
cudaError_t spk_dist_launch_merge(spk_ctx* c) {
    GridBuffers& next = c->grid[1 - c->gcur];
    if (!next.nab_host || c->dist_merge_in_update) return cudaSuccess; // folded into grid_update_kernel
    dist_merge_kernel<<<next.nab_host, 128, 0, c->stream>>>(next, next.nab_host, arena_grid(c->dist_arena, c->dist_cap_hash, 1 - c->gcur).hacc);
    c->launches += 1;
    return cudaGetLastError();
}

extern "C" {

int spk_dist_export(spk_ctx* c, SpkDistDescriptor* out) {
    if (!c || !out) return SPK_EINVAL;
    { int jrc = spk_join_grid_update(c); if (jrc) return jrc; }
    if (c->dist_arena) return SPK_ESTATE;
    CK(cudaSetDevice(c->device));
    // the shared arena has a fixed capacity: size it generously from what the particle count suggested (growth of the hash
    // map would move it and invalidate the peers' mappings, so in distributed runs an overflow is an error instead)
    const uint32_t cap = std::max(c->grid[0].cap_hash, c->grid[1].cap_hash) * 2u;
    const size_t bytes = arena_bytes(cap);
    char* arena = nullptr;
    CK(cudaMalloc((void**)&arena, bytes));
    CK(cudaMemsetAsync(arena, 0, bytes, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->dist_arena = arena;
    c->dist_cap_hash = cap;
    c->dist_arena_bytes = bytes;
    // both grid buffers move their hash tables into the arena (their contents are rebuilt by the next sort)
    int rc = spk_dist_realloc_grids(c, cap);
    if (rc) return rc;
    CK(cudaMalloc((void**)&c->dist_peer_ids, (size_t)(cap / 2) * SPK_MAX_RANKS * sizeof(int32_t)));
    CK(cudaMalloc((void**)&c->dist_push_ptrs, SPK_MAX_RANKS * sizeof(float*)));
    CK(cudaMemset(c->dist_push_ptrs, 0, SPK_MAX_RANKS * sizeof(float*)));
    c->dist_push_wait = getenv("SPK_DIST_PUSH_WAIT") ? 1 : 0;
    c->dist_merge_in_update = 1;
    if (const char* m = getenv("SPK_DIST_MERGE")) c->dist_merge_in_update = std::strcmp(m, "kernel") != 0;
    c->dist_split_barriers = getenv("SPK_DIST_SPLIT") ? atoi(getenv("SPK_DIST_SPLIT")) : 1;
    c->dist_s1_arrived = 0;
    // halo push: a kernel of its own after g2p2g (default; measured 12 us per substep faster at 4 M particles per GPU: the pushes of the few
    // shared blocks stretch the whole fused kernel) or fused into g2p2g's flush (A/B: SPK_DIST_PUSH=fused)
    c->dist_fused_push = 0;
    if (const char* m = getenv("SPK_DIST_PUSH")) c->dist_fused_push = std::strcmp(m, "fused") == 0;
    {   // barrier timeout: SPK_DIST_TIMEOUT_MS (default 3000, 0 = unbounded), converted with the device's SM clock
        long long ms = 3000;
        if (const char* t = getenv("SPK_DIST_TIMEOUT_MS")) ms = atoll(t);
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
        c->dist_timeout_cycles = ms <= 0 ? 0 : ms * (long long)(khz > 0 ? khz : 2000000);
    }
    CK(cudaMemsetAsync(&c->d_counters->pad[1], 0, 4, c->stream));
    DescriptorWire w;
    std::memset(&w, 0, sizeof(w));
    CK(cudaIpcGetMemHandle(&w.handle, arena));
    w.cap_hash = cap;
    w.device = (uint32_t)c->device;
    w.bytes = bytes;
    w.magic = MAGIC;
    std::memset(out, 0, sizeof(*out));
    std::memcpy(out, &w, sizeof(w));
    return SPK_OK;
}

int spk_dist_attach(spk_ctx* c, uint32_t rank, uint32_t world, const SpkDistDescriptor* all) {
    if (!c || !all || world < 1 || world > SPK_MAX_RANKS || rank >= world) return SPK_EINVAL;
    if (!c->dist_arena || c->dist.world) return SPK_ESTATE;
    CK(cudaSetDevice(c->device));
    DistState d;
    std::memset(&d, 0, sizeof(d));
    d.rank = rank;
    d.world = world;
    d.mailbox = reinterpret_cast<DistMailbox*>(c->dist_arena);
    for (uint32_t p = 0; p < world; ++p) {
        DescriptorWire w;
        std::memcpy(&w, &all[p], sizeof(w));
        if (w.magic != MAGIC) return SPK_EINVAL;
        char* base = c->dist_arena;
        if (p != rank) {
            void* mapped = nullptr;
            CK(cudaIpcOpenMemHandle(&mapped, w.handle, cudaIpcMemLazyEnablePeerAccess));
            base = (char*)mapped;
            c->dist_mapped[p] = mapped;
        }
        d.peer_mailbox[p] = reinterpret_cast<DistMailbox*>(base);
        d.peer_cap_hash[p] = w.cap_hash;
        for (int g = 0; g < 2; ++g) d.peer_grid[p][g] = arena_grid(base, w.cap_hash, g);
    }
    c->dist = d;
    return SPK_OK;
}

int spk_dist_detach(spk_ctx* c) {
    if (!c) return SPK_EINVAL;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (uint32_t p = 0; p < SPK_MAX_RANKS; ++p)
        if (c->dist_mapped[p]) { cudaIpcCloseMemHandle(c->dist_mapped[p]); c->dist_mapped[p] = nullptr; }
    std::memset(&c->dist, 0, sizeof(c->dist));
    cudaMemset(&c->d_counters->pad[1], 0, 4); // the sticky peer-timeout mark dies with the attachment
    c->h_counters->pad[1] = 0;
    return SPK_OK;
}

int spk_dist_info(spk_ctx* c, uint32_t* num_halo_blocks, uint32_t* epoch) {
    if (!c) return SPK_EINVAL;
    if (num_halo_blocks) *num_halo_blocks = c->h_counters->num_halo;
    if (epoch) *epoch = c->dist.epoch;
    return SPK_OK;
}

} // extern "C"

#endif // SPK_DIM
