// cuda_emu.h — TEST INFRASTRUCTURE: a minimal CUDA execution model for the host, so that the PRODUCT'S OWN .cu sources (kernels and
// host sequencing) can be compiled with g++ and run on a GPU-less machine by the test suite (tests/test_emulated_cuda_cpu.py).
//
// It is not a product path: libsparkl_b200_emu*.so are built into tests/host/ only, loaded only by tests, and the product loader
// (sparkl_b200/_lib.py) never looks for them — without a GPU the product still fails loudly.
//
// Model: a kernel launch runs its CTAs one after the other; inside a CTA every CUDA thread is a fiber (ucontext) scheduled round
// robin on one OS thread.  A fiber runs until it reaches a synchronisation point (__syncthreads or a *_sync warp collective), where
// it parks until every live thread it synchronises with has arrived; exited threads count as arrived, as on Volta+.  Because only
// one fiber runs at a time, atomics are plain read-modify-writes and results are deterministic (fixed summation order) — numerics
// otherwise follow the source: no fast-math, IEEE stand-ins for the MUFU approximations (rsqrtf, __fdividef).
// TMA bulk copies / reductions and mbarriers are replaced in the sources themselves (#ifdef SPK_EMU: plain loops).
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>

namespace emu {
struct Idx { unsigned x, y, z; };
extern Idx tid, bid, bdim, gdim;
void launch(unsigned grid, unsigned block, const std::function<void()>& body);
void block_sync();
void warp_sync(unsigned mask);
unsigned long long* warp_scratch(); // 32 slots of the calling thread's warp
unsigned char* dynamic_smem();
inline unsigned lane() { return tid.x & 31u; }
} // namespace emu

#define threadIdx emu::tid
#define blockIdx emu::bid
#define blockDim emu::bdim
#define gridDim emu::gdim
#undef __shared__
#define __shared__ static
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#define emu_launch emu::launch

using std::isfinite;
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }

// ---- synchronisation and warp collectives ---------------------------------------------------------------------------------
static inline void __syncthreads() { emu::block_sync(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_sync(mask); }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline void __nanosleep(unsigned) {}
static inline long long clock64() { return 0; }
template <typename T>
static inline T emu_exchange(unsigned mask, T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    unsigned long long* s = emu::warp_scratch();
    unsigned long long bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    s[emu::lane()] = bits;
    emu::warp_sync(mask);
    T out = v;
    if (src_lane >= 0 && src_lane < 32) std::memcpy(&out, &s[src_lane], sizeof(T));
    emu::warp_sync(mask);
    return out;
}
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src) { return emu_exchange(mask, v, src & 31); }
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d) { const int l = (int)emu::lane(); return emu_exchange(mask, v, l >= (int)d ? l - (int)d : l); }
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int x) { return emu_exchange(mask, v, (int)emu::lane() ^ x); }
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    unsigned long long* s = emu::warp_scratch();
    s[emu::lane()] = pred ? 1ull : 0ull;
    emu::warp_sync(mask);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) if ((mask >> l & 1u) && s[l]) r |= 1u << l;
    emu::warp_sync(mask);
    return r;
}
static inline unsigned __match_any_sync(unsigned mask, unsigned long long v) {
    unsigned long long* s = emu::warp_scratch();
    s[emu::lane()] = v;
    emu::warp_sync(mask);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) if ((mask >> l & 1u) && s[l] == v) r |= 1u << l;
    emu::warp_sync(mask);
    return r;
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
    unsigned long long* s = emu::warp_scratch();
    s[emu::lane()] = v;
    emu::warp_sync(mask);
    unsigned r = 0xffffffffu;
    for (int l = 0; l < 32; ++l) if ((mask >> l & 1u) && (unsigned)s[l] < r) r = (unsigned)s[l];
    emu::warp_sync(mask);
    return r;
}

static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    unsigned long long* s = emu::warp_scratch();
    s[emu::lane()] = v;
    emu::warp_sync(mask);
    unsigned r = 0u;
    for (int l = 0; l < 32; ++l) if ((mask >> l & 1u) && (unsigned)s[l] > r) r = (unsigned)s[l];
    emu::warp_sync(mask);
    return r;
}

// ---- atomics (one fiber runs at a time) -----------------------------------------------------------------------------------------
template <typename T> static inline T atomicAdd(T* p, T v) { const T old = *p; *p = old + v; return old; }
static inline unsigned atomicAdd(unsigned* p, int v) { const unsigned old = *p; *p = old + (unsigned)v; return old; }
static inline int atomicAdd(int* p, int v) { const int old = *p; *p = old + v; return old; }
template <typename T> static inline T atomicMin(T* p, T v) { const T old = *p; if (v < old) *p = v; return old; }
template <typename T> static inline T atomicMax(T* p, T v) { const T old = *p; if (v > old) *p = v; return old; }
template <typename T> static inline T atomicExch(T* p, T v) { const T old = *p; *p = v; return old; }
template <typename T> static inline T atomicCAS(T* p, T cmp, T v) { const T old = *p; if (old == cmp) *p = v; return old; }

// ---- scalar intrinsics ------------------------------------------------------------------------------------------------------------
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline long long __float2ll_rz(float x) { // cvt.rzi.s64.f32: saturating, NaN -> 0
    if (x != x) return 0;
    if (x >= 9.2233720368547758e18f) return INT64_MAX;
    if (x <= -9.2233720368547758e18f) return INT64_MIN;
    return (long long)x;
}
static inline unsigned long long __float2ull_rz(float x) {
    if (x != x || x <= 0.f) return 0ull;
    if (x >= 1.8446744073709552e19f) return 0xffffffffffffffffull;
    return (unsigned long long)x;
}
template <typename T> static inline T __ldcv(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }

// kernel function attributes do not exist here
template <typename K> static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
