// spk_device.cuh — device helpers shared by the kernels: block keys, hash map, small reductions.
#pragma once
#include "spk_internal.cuh"

#if SPK_DIM == 3
#define SPK_PACK_ORIGIN 0x00100000ll
#define SPK_KEY_MASK 0x1fffffull

// BlockVirtualId::pack / unpack — src_kernels/gpu_grid.rs:96-120 (21 bits per axis, x | y<<21 | z<<42)
__host__ __device__ __forceinline__ unsigned long long spk_pack_key(unsigned long long x, unsigned long long y, unsigned long long z) {
    return (x & SPK_KEY_MASK) | ((y & SPK_KEY_MASK) << 21) | ((z & SPK_KEY_MASK) << 42);
}
__host__ __device__ __forceinline__ void spk_unpack_key(unsigned long long k, unsigned long long& x, unsigned long long& y, unsigned long long& z) {
    x = k & SPK_KEY_MASK; y = (k >> 21) & SPK_KEY_MASK; z = (k >> 42) & SPK_KEY_MASK;
}
// key of block (b + o), o in {0,1}^3 encoded bit0 = x, bit1 = y, bit2 = z (g2p2g.rs:376-387 octant convention)
__device__ __forceinline__ unsigned long long spk_key_offset(unsigned long long k, int o) {
    unsigned long long x, y, z;
    spk_unpack_key(k, x, y, z);
    return spk_pack_key(x + (o & 1), y + ((o >> 1) & 1), z + ((o >> 2) & 1));
}
#else
#define SPK_PACK_ORIGIN 0x80000000ll
#define SPK_KEY_MASK 0xffffffffull

// BlockVirtualId::pack / unpack (dim2) — src_kernels/gpu_grid.rs:90-94,106-109 (32 bits per axis, x | y<<32, origin 2^31)
__host__ __device__ __forceinline__ unsigned long long spk_pack_key(unsigned long long x, unsigned long long y) {
    return (x & SPK_KEY_MASK) | ((y & SPK_KEY_MASK) << 32);
}
__host__ __device__ __forceinline__ void spk_unpack_key(unsigned long long k, unsigned long long& x, unsigned long long& y) {
    x = k & SPK_KEY_MASK; y = (k >> 32) & SPK_KEY_MASK;
}
// key of block (b + o), o in {0,1}^2 encoded bit0 = x, bit1 = y
__device__ __forceinline__ unsigned long long spk_key_offset(unsigned long long k, int o) {
    unsigned long long x, y;
    spk_unpack_key(k, x, y);
    return spk_pack_key(x + (o & 1), y + ((o >> 1) & 1));
}
#endif

// Associated cell of a coordinate: round(p / h) with IEEE division and round-half-away (Rust f32::round), saturating
// cast to i64 with NaN -> 0 (gpu_grid.rs:273, PTX: div.rn.f32 + trunc(x + copysign(0.5, x)) + cvt.rzi.s64.f32).
__device__ __forceinline__ long long spk_assoc_cell(float p, float h) {
    float q = roundf(__fdiv_rn(p, h));
    if (q != q) return 0;
    return __float2ll_rz(q); // saturates on overflow
}
// x / h, correctly rounded, from the correctly rounded reciprocal y = RN(1 / h) (Markstein: q0 = x y, r = x - q0 h exactly
// by FMA, q = q0 + r y).  Bit-identical to div.rn.f32 over 4e9 random operands and the cell widths of every scene
// (scripts/check_div.c); operands outside the safe exponent range take the IEEE division.  3 instructions instead of ~12.
__device__ __forceinline__ float spk_div_h(float x, float h, float inv_h) {
    const float q0 = x * inv_h;
    const float q = fmaf(fmaf(-q0, h, x), inv_h, q0);
    const float ax = fabsf(x);
    return (ax < 1.0e30f && ax > 1.0e-30f) ? q : __fdiv_rn(x, h);
}
// the same cell through spk_div_h (bit-identical quotient, ~10 instructions fewer per axis)
__device__ __forceinline__ long long spk_assoc_cell_m(float p, float h, float inv_h) {
    float q = roundf(spk_div_h(p, h, inv_h));
    if (q != q) return 0;
    return __float2ll_rz(q);
}
// block_associated_to_point, gpu_grid.rs:271-279: ((cell - 2 + 2^22) as usize) / 4 per axis
__device__ __forceinline__ unsigned long long spk_block_coord(long long cell) {
    return (unsigned long long)(cell - 2 + SPK_PACK_ORIGIN * 4) / 4ull;
}
#if SPK_DIM == 3
__device__ __forceinline__ unsigned long long spk_block_key_of_point(float x, float y, float z, float h) {
    return spk_pack_key(spk_block_coord(spk_assoc_cell(x, h)), spk_block_coord(spk_assoc_cell(y, h)), spk_block_coord(spk_assoc_cell(z, h)));
}
#else
__device__ __forceinline__ unsigned long long spk_block_key_of_point(float x, float y, float /*unused*/, float h) {
    return spk_pack_key(spk_block_coord(spk_assoc_cell(x, h)), spk_block_coord(spk_assoc_cell(y, h)));
}
#endif

#if defined(__CUDACC__) || defined(SPK_EMU) // hash map + warp helpers: device only (the arithmetic helpers above are also host-compiled by tests/host/;
// SPK_EMU = the whole library compiled for the host under the CUDA emulation shim of tests/host/emu/)
// 64-bit finaliser hash (same family as hashmap.rs:7-14; only set semantics of the map are observable)
__device__ __forceinline__ uint32_t spk_hash(unsigned long long key, uint32_t mask) {
    key ^= key >> 33;
    key *= 0xff51afd7ed558ccdull;
    key ^= key >> 33;
    key *= 0xc4ceb9fe1a85ec53ull;
    key ^= key >> 33;
    return (uint32_t)key & mask;
}
// insert-if-absent; returns the slot or SPK_NO_BLOCK when the table is full (GridHashMap::insert_nonexistant_with, hashmap.rs:48-70)
__device__ __forceinline__ uint32_t spk_hash_insert(unsigned long long* keys, uint32_t cap, unsigned long long key) {
    const uint32_t mask = cap - 1;
    uint32_t slot = spk_hash(key, mask);
    // a probe sequence this long only happens when the table is (nearly) full: report overflow, the host grows the
    // table and re-runs the sort (cuda_sparse_grid.rs:162-253) instead of crawling through a saturated table
    const uint32_t max_probes = cap < 512u ? cap : 512u;
    for (uint32_t probe = 0; probe < max_probes; ++probe) {
        unsigned long long prev = keys[slot];
        if (prev == key) return slot;
        if (prev == SPK_EMPTY_KEY) {
            prev = atomicCAS(&keys[slot], SPK_EMPTY_KEY, key);
            if (prev == SPK_EMPTY_KEY || prev == key) return slot;
        }
        slot = (slot + 1) & mask;
    }
    return SPK_NO_BLOCK;
}
// GridHashMap::get, hashmap.rs:72-87; returns the slot or SPK_NO_BLOCK
__device__ __forceinline__ uint32_t spk_hash_find_slot(const unsigned long long* __restrict__ keys, uint32_t cap, unsigned long long key) {
    const uint32_t mask = cap - 1;
    uint32_t slot = spk_hash(key, mask);
    for (uint32_t probe = 0; probe < cap; ++probe) {
        unsigned long long k = keys[slot];
        if (k == key) return slot;
        if (k == SPK_EMPTY_KEY) return SPK_NO_BLOCK;
        slot = (slot + 1) & mask;
    }
    return SPK_NO_BLOCK;
}
__device__ __forceinline__ int32_t spk_hash_get(const GridBuffers& g, unsigned long long key) {
    if (g.cap_hash == 0) return -1;
    uint32_t slot = spk_hash_find_slot(g.keys, g.cap_hash, key);
    return slot == SPK_NO_BLOCK ? -1 : (int32_t)g.vals[slot];
}

__device__ __forceinline__ uint32_t spk_lanemask_lt() {
#ifdef SPK_EMU
    return (1u << (threadIdx.x & 31u)) - 1u;
#else
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
#endif
}
#endif // __CUDACC__
