// spk_internal.cuh — device data layout and context of libsparkl_b200 (B200 / sm_100a).
//
// HBM layout:
//  * particles: "array of float4 streams" (AoSoA-free SoA of 16-byte records), double buffered.  g2p2g reads the
//    current buffer through the sort permutation and writes the next buffer in sorted order, so particles are
//    physically re-sorted by block every substep for free and all stores are fully coalesced.
//  * sparse grid: per active block 64 cells; `vel` (float4: v.xyz, psi_velocity) is one contiguous 1 KB chunk per
//    block so a g2p2g CTA stages its 2x2x2 block tile with 8 TMA bulk copies; `acc` ([6][64] floats: mass,
//    momentum xyz, psi_momentum, psi_mass) is one contiguous 1.5 KB chunk per block so the tile is flushed with 8
//    TMA bulk reduce-adds.  Collider cache (`proj`, `cacheN`, `cacheD`) is touched only by reset/grid_update.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sparkl_abi.h"

// SPK_DIM (include/sparkl_abi.h) selects the build: 3 -> libsparkl_b200.so, 2 -> libsparkl_b200_2d.so.  Host sequencing, hash
// map, scans and particle streams are shared; the block shape (4^DIM cells, 2^DIM associated blocks) and the kernels that walk
// the stencil (spk_g2p2g*.cu, spk_grid*.cu) are per dimension, like the reference's `dim2` / `dim3` features.
#if SPK_DIM == 3
#define SPK_CELLS_PER_BLOCK 64
#define SPK_TILE_CELLS 512
#define SPK_NBR 8 /* NUM_ASSOC_BLOCKS, src_kernels/gpu_grid.rs:13-16 */
#define SPK_NBR_SHIFT 3
#define SPK_CELL_SHIFT 6
#define SPK_REF_G2P2G_THREADS 128u /* G2P2G_THREADS of the reference, src_kernels/cuda/mod.rs:15 (dispatch read-back) */
#else
#define SPK_CELLS_PER_BLOCK 16
#define SPK_TILE_CELLS 64
#define SPK_NBR 4
#define SPK_NBR_SHIFT 2
#define SPK_CELL_SHIFT 4
#define SPK_REF_G2P2G_THREADS 64u /* dim2: src_kernels/cuda/mod.rs:12-13 */
#endif
#ifndef SPK_CELL_THREADS
#define SPK_CELL_THREADS 256 /* threads per CTA == particles per dispatch entry of the cell-coherent g2p2g kernel (mode 2) */
#endif
#define SPK_ACC_CHANNELS 6
// Position of a cell inside the 64-float channel chunk of `acc` (NOT of `vel` / the collider cache).  dim3: bit 1 of the cell's z flips
// bit 1 of its x, so that the 3 x 3 (y, z) columns of a P2G stencil fall into 9 different shared-memory banks when a g2p2g CTA adds a
// bin's sums into its tile (without it the columns sz = 0 and sz = 2 share a bank).  The tile chunk is copied to / from global memory
// as a whole (TMA), so every kernel that addresses one node of `acc` uses the same mapping.
#if SPK_DIM == 3
#define SPK_ACC_SWZ(cell) ((cell) ^ (((cell) >> 4) & 2u))
#else
#define SPK_ACC_SWZ(cell) (cell)
#endif
#define SPK_EMPTY_KEY 0xffffffffffffffffull
#define SPK_NO_BLOCK 0xffffffffu
#ifndef SPK_G2P2G_CHUNK
#if SPK_DIM == 3
#define SPK_G2P2G_CHUNK 256 /* particles per g2p2g dispatch entry == threads per CTA */
#else
#define SPK_G2P2G_CHUNK 128 /* dim2: a 4x4-cell block holds ~64 particles at rest */
#endif
#endif

// particle flag bits (record A0.w)
#define SPK_F_FAILED 1u
#define SPK_F_STATIC 2u
#define SPK_F_KINEMATIC 4u
#define SPK_F_MODEL_SHIFT 8

struct ParticleBuffers { // one of the two ping-pong buffers
    float4* a0;          // x, y, z, flags (bits)
    float4* a1;          // F col-major 0..3
    float4* a2;          // F 4..7
    float4* a3;          // F8, Jp, mass, volume0
    float4* a4;          // phase, psi_pos, elastic_hardening, plastic_hardening
    float4* a5;          // log_vol_gain, nacc_alpha, orig_id (bits), unused
};

struct DevModel { // compact device image of GpuParticleModel (include/sparkl_abi.h)
    uint32_t ctag, ptag, ftag, flags; // flags: bit0 split_stress_on_failure, bit1 only_active_when_failed, bit2 hardening_enabled
    float c[4];                       // corotated: cfl, lambda, mu | neo-hookean: cfl, lambda, mu | monaghan: pressure0, gamma(bits), viscosity, max_neg_pressure
    float p[8];                       // DP: h0,h1,h2,h3,lambda,mu,volume_correction | NACC: mu,kappa,hardening_factor,cohesion,friction | Rankine: mu,lambda,tensile,softening | Snow: min_eps,max_eps,hardening_coeff
    float f[2];                       // max_principal_stress, max_shear_stress
};

struct DevCollider {
    float half_extents[3];
    float rot[4]; // i, j, k, w (dim2: re, im of the unit complex, rest unused)
    float tr[3];
    float friction;
    uint32_t handling;
    uint32_t shape;       // SPK_SHAPE_CUBOID / SPK_SHAPE_HEIGHTFIELD
    uint32_t nrows, ncols; // heightfield: height samples along z / x; polyline: nrows = number of vertices
    const float* heights; // heightfield: device copy, column-major nrows x ncols; polyline: vertices (x, y, z); trimesh: nrows vertices (x, y, z)
                          // followed by ncols x 3 vertex indices stored as raw u32 bits
    float scale[3];
    uint32_t hf_flags;    // bit0 bellow_heightfield_is_solid, bit1 flip_interior
};

struct GridBuffers { // one of the two sparse-grid buffers (curr / next)
    // hash map: key -> block id (open addressing, linear probing, capacity power of two)
    unsigned long long* keys;
    uint32_t* vals;
    uint32_t cap_hash;
    uint32_t cap_blocks; // capacity of all per-block arrays (== cap_hash / 2)
    // per block
    unsigned long long* blk_key;
    uint32_t* blk_first;
    uint32_t* blk_count;
    uint32_t* blk_halo;
    uint32_t* blk_dfirst; // first g2p2g dispatch entry of the block (it owns ceil(count / chunk) consecutive entries)
    int32_t* nbr_self; // [cap_blocks][8] ids of key + {0,1}^3 in THIS grid (octant bit0 = x, bit1 = y, bit2 = z)
    int32_t* nbr_prev; // [cap_blocks][8] ids of the same keys in the OTHER (previous) grid, -1 when absent
    // per node
    float4* vel;     // [cap_blocks*64] velocity.xyz, psi_velocity (valid after grid_update)
    float* acc;      // [cap_blocks][6][64]
    uint32_t* proj;  // [cap_blocks*64] projection status: tag | collider << 2
    float4* cacheN;  // collision_normal.xyz, prev_mass
    float4* cacheD;  // projection_scaled_dir.xyz, unused
    uint32_t external_hash; // keys / vals live in the shared multi-GPU arena (not owned by this struct)
    uint32_t nab_host;      // host copy of the number of active blocks
    uint32_t has_velocity;  // 1 after grid_update (momentum_velocity holds velocities)
};

struct DeviceCounters { // one 64-byte record read back once per substep (pinned mirror on the host)
    uint32_t nab;            // number of active blocks
    uint32_t ndispatch;      // number of g2p2g dispatch entries
    uint32_t overflow;       // hash map full / block capacity exceeded
    uint32_t nparticles_binned;
    unsigned long long dt_acc; // GpuTimestepLength
    uint32_t num_halo;
    uint32_t halo_cursor;
    uint32_t dt_raw_bits;    // min over non-failed particles of the un-clamped dt candidate (float bits), written by g2p2g's
                             // epilogue for the NEXT substep; >= SPK_DT_RAW_NONE when no particle contributed
    uint32_t pad[5];         // pad[0]: invalid model index seen by the asynchronous upload; pad[1]: STICKY 0x100 + peer once a device-side barrier of
                             // the multi-GPU path timed out on that peer (never cleared by the per-substep memsets; spk_dist.cu)
};
#define SPK_DT_RAW_NONE 0x7f7f7f7fu /* memset pattern; valid candidates are capped at 1e30f */

struct DispatchEntry { uint32_t block; uint32_t first; uint32_t count; uint32_t pad; };

// ---- multi-GPU over NVLink peer memory (spk_dist.cu) ---------------------------------------------------------------
#define SPK_MAX_RANKS 16
struct DistMailbox { // first 4 KB of every rank's shared arena; written by the peers
    volatile uint32_t flag[SPK_MAX_RANKS];        // flag[p]: last barrier epoch signalled by rank p
    uint32_t gnext[2][SPK_MAX_RANKS];             // [epoch & 1][p]: which grid buffer is rank p's next grid
    unsigned long long val[2][SPK_MAX_RANKS];     // [epoch & 1][p]: rank p's dt candidate
    uint32_t bbox[2][SPK_MAX_RANKS][6];           // [epoch & 1][p]: bounding box of rank p's active blocks (min x, y, z, max x, y, z of the raw
                                                  // key coordinates): a block outside it cannot be shared with rank p and is not probed
};
struct DistGridView { unsigned long long* keys; uint32_t* vals; float* hacc; }; // one grid buffer's shared arrays
struct DistState {   // passed by value to the kernels
    uint32_t rank, world, epoch, pad;
    DistMailbox* mailbox;                          // this rank's
    DistMailbox* peer_mailbox[SPK_MAX_RANKS];      // peer-mapped
    DistGridView peer_grid[SPK_MAX_RANKS][2];      // peer-mapped (entry [rank] = this rank's own)
    uint32_t peer_cap_hash[SPK_MAX_RANKS];
};

struct spk_ctx {
    int device;
    float cell_width;
    cudaStream_t stream, halo_stream;
    // spk_step runs the grid update of substep n on its own stream, next to the particle sort of substep n + 1 (both only depend on
    // g2p2g of substep n); `gu_pending`: the main stream has not yet been ordered behind the last grid update
    cudaStream_t gu_stream; cudaEvent_t ev_transfer_done, ev_gu_done, ev_tables; bool gu_pending; int overlap_grid_update;
    // particles
    size_t n, cap;
    int cur; // index of the current particle buffer
    ParticleBuffers pb[2];
    float4* vel;     // vx, vy, vz, unused — aligned with pb[cur]
    float4* cold;    // kinematic_vel.xyz, radius0 — indexed by ORIGINAL particle id
    uint32_t* slot_of; // hash slot of the particle's block (current order)
    uint32_t* rank_of; // cell inside its block << (32 - SPK_CELL_SHIFT) | rank inside that cell
    uint32_t* perm;    // sorted index -> index in current order
    uint32_t* gid;     // optional global particle ids (multi-GPU migration), indexed by particle id; NULL = identity
    bool sorted_valid; // perm/slot tables describe pb[cur]
    uint32_t dt_raw_host; bool dt_raw_host_valid; // the candidate as the last sort's counter record carried it (the device word is reset by that sort)
    bool dt_cache_valid; // d_counters->dt_raw_bits describes pb[cur] (set by g2p2g, cleared by any particle/model edit)
    unsigned long long dt_acc_host; // GpuTimestepLength accumulator (host-authoritative)
    // models / colliders
    DevModel* models; size_t nmodels; uint32_t model_flags_any; // bit0: any model needs an SVD
    uint32_t model_mask;  // union of model_mask_of(...) over the uploaded model table (SPK_MM_*, spk_math.cuh): picks the kernel instantiation
    int g2p2g_stream;     // mode 2: 1 (default) persistent CTAs that prefetch their next dispatch entry, 0: one CTA per entry (SPK_G2P2G_STREAM=0)
    int g2p2g_prefetch;   // mode 2: 1: every CTA warms the L2 with the particles of a later dispatch entry (SPK_G2P2G_PREFETCH=1; off by default: no gain measured)
    int g2p2g_specialize; // 1 (default): use the model-set specialisations of the fused kernel; 0: always the generic one (SPK_G2P2G_GENERIC=1)
    DevCollider* colliders; size_t ncolliders;
    float* collider_data; // device copies of the heightfield matrices, one allocation
    // grid
    GridBuffers grid[2];
    int gcur;             // grid[gcur] = curr (read only), grid[1-gcur] = next
    uint32_t* slot_cnt;   // [cap_hash of next] particles per hash slot
    uint32_t* cell_cnt;   // [cap_hash of next][cells per block] particles per (hash slot, cell); after cell_scan_kernel: exclusive offsets inside the block
    char* mg_scratch; size_t mg_scratch_bytes; // grow-only scratch of the particle-migration passes (spk_migrate.cu)
    uint32_t* mslot;      // [4 x cap_hash of next] Morton table of the current sort: Morton code of a block -> hash slot + 1 (spk_sort.cu)
    uint32_t slot_cnt_cap;
    DispatchEntry* dispatch; uint32_t dispatch_cap;
    int32_t* dispatch_nbr; // [dispatch_cap][2 * SPK_NBR]: nbr_self | nbr_prev of the entry's block, so that a g2p2g CTA finds its tile with ONE load
    uint32_t chunk;       // particles per g2p2g dispatch entry
    int g2p2g_mode;       // 0: per-particle shared atomics, 1: cell-binned register accumulation (default)
    bool g2p2g_configured; // the >48 KB dynamic shared memory opt-in was set on this context's device
    // counters
    DeviceCounters* d_counters;
    DeviceCounters* h_counters; // pinned, mapped
    DeviceCounters* h_counters_dev; // device address of h_counters
    unsigned long long* h_upload; // pinned 8-byte staging for dt_acc -> device
    bool counters_fresh;         // h_counters was read after the last kernel that writes d_counters
    // scratch for read-back / staging
    void* staging; size_t staging_bytes;
    // asynchronous upload: own staging buffer, host->device copies on h2d_stream, pack kernel on the main stream
    cudaStream_t h2d_stream; void* up_staging; size_t up_staging_bytes; cudaEvent_t up_ready, up_consumed; bool up_inflight;
    bool up_staged, up_staged_data; size_t up_staged_n; size_t up_off[6]; // a staged (not yet committed) upload
    // asynchronous read-back: one device staging buffer + two events per attribute; copies run on halo_stream
    void* async_buf[8]; size_t async_bytes[8];
    cudaEvent_t async_ready[8], async_done[8]; bool async_inflight[8];
    // multi-GPU (all zero / null in single-GPU runs)
    DistState dist;
    char* dist_arena; size_t dist_arena_bytes; uint32_t dist_cap_hash;
    uint32_t* dist_bbox;                 // [24] device: [0..5] bounding box of this context's active blocks of the current sort (assign_ids_kernel; sent to the
                                         // peers of a multi-GPU run), [8..15] the Morton box the current sort ranks its blocks in (spk_sort.cu: mstate),
                                         // [16..23] the box the last completed sort used (spk_grid_morton_box)
    int dist_merge_in_update;            // 1 (default): grid_update folds the peers' halo sums itself; 0: separate merge kernel (A/B: SPK_DIST_MERGE=kernel)
    int dist_push_wait;                  // 1: a CTA that pushed halo sums waits for their remote completion (A/B: SPK_DIST_PUSH_WAIT=1); 0 (default): the
                                         // kernel boundary + the barrier's system fence order them before the flag
    int32_t* dist_peer_ids;              // [cap_blocks][world]: the peer's id of each of this rank's active blocks, -1 if absent
    void* dist_mapped[SPK_MAX_RANKS];    // cudaIpcOpenMemHandle results (to close)
    float** dist_push_ptrs;              // [SPK_MAX_RANKS] device array: peer-mapped hacc of every rank's next grid (written by detect)
    long long dist_timeout_cycles;       // device-side barrier: give up on a peer after this many SM cycles (0 = wait for ever)
    uint32_t dist_epoch_of[3], dist_gnext_of[3]; // epoch / next-grid index of the last ARRIVE of each barrier mode (the wait half may be issued later)
    int dist_split_barriers;             // 1 (default): arrive / wait halves of the two barriers are issued apart (A/B: SPK_DIST_SPLIT=0)
    int dist_s1_arrived;                 // 1: the arrive half of the pre-detection barrier was issued by the sort (spk_launch_sort_finish)
    int dist_fused_push;                 // 0 (default): a push kernel after g2p2g; 1: g2p2g's flush pushes shared blocks itself (A/B: SPK_DIST_PUSH=fused)
    // custom particle kernels (spk_set_particle_kernels_launcher); all NULL = DefaultCudaParticleKernelsLauncher
    SpkParticleKernelsLauncher launcher;
    // timing
    cudaEvent_t ev[12];
    cudaEvent_t ev_counters;
    // ring of event pairs around every g2p2g launch (recorded without synchronising; read by spk_g2p2g_time)
    cudaEvent_t kev[2 * 256]; uint32_t kev_count; bool kev_on; // recorded right after the counter record has been published to the host
    uint32_t launches;
    uint32_t g2p2g_launches;
    int sm_count;
};

// ---- kernels launch wrappers (defined in the .cu files) ---------------------------------------------------------
cudaError_t spk_launch_sort_bin(spk_ctx* c);                 // K1: hash insert + count + rank per particle
cudaError_t spk_launch_sort_finish(spk_ctx* c);              // K2..K5: neighbours, ids/scan, tables, scatter
cudaError_t spk_launch_reset_grid(spk_ctx* c);
cudaError_t spk_launch_estimate_dt(spk_ctx* c, float min_dt, float max_dt);
cudaError_t spk_launch_g2p2g(spk_ctx* c, float dt);
cudaError_t spk_launch_grid_update(spk_ctx* c, float dt, const float gravity[SPK_DIM]);
cudaError_t spk_launch_fluids_volume(spk_ctx* c);             // spk_fluids.cu
cudaError_t spk_launch_pack_upload(spk_ctx* c, const void* d_status, const void* d_pos, const void* d_vel, const void* d_volume,
                                   const void* d_phase, const void* d_data, size_t first, size_t n);
cudaError_t spk_launch_unpack(spk_ctx* c, int which, void* d_out, size_t n);
int spk_particles_resize(spk_ctx* c, size_t n, bool fresh);   // spk_api.cu: capacity / counts / grid sizing of an upload, without the copy
cudaError_t spk_launch_grid_export(spk_ctx* c, int which_grid, SpkGpuGridNode* d_out, uint32_t nblocks);
cudaError_t spk_launch_halo_keys(spk_ctx* c, unsigned long long* out, uint32_t capacity);
cudaError_t spk_launch_halo_tag(spk_ctx* c, const unsigned long long* remote, uint32_t n);
cudaError_t spk_launch_halo_pack(spk_ctx* c, unsigned long long* keys, float* payload, uint32_t capacity);
cudaError_t spk_launch_halo_merge(spk_ctx* c, const unsigned long long* keys, const float* payload, uint32_t n);
// spk_dist.cu
bool spk_dist_active(const spk_ctx* c);
bool spk_dist_grid_storage(spk_ctx* c, int g, uint32_t cap_hash, unsigned long long** keys, uint32_t** vals);
int spk_dist_realloc_grids(spk_ctx* c, uint32_t cap_hash); // defined in spk_api.cu
int spk_join_grid_update(spk_ctx* c);                        // spk_api.cu: main stream waits for a grid update left on its own stream
cudaError_t spk_dist_launch_barrier(spk_ctx* c, int mode, int phase = 3); // mode 0 barrier, 1 + min of dt_raw_bits, 2 + min of dt_acc; phase 1 arrive, 2 wait, 3 both
cudaError_t spk_dist_launch_detect(spk_ctx* c);
cudaError_t spk_dist_launch_push(spk_ctx* c);
cudaError_t spk_dist_launch_merge(spk_ctx* c);
float* spk_dist_hacc_of_next(spk_ctx* c); // this rank's halo accumulator of the next grid (NULL without an attachment)
