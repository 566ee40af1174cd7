// grid3d_host.cpp — TEST INFRASTRUCTURE: compiles the per-node dim3 grid kernels of the product (sparkl_b200/csrc/spk_grid.cu: reset +
// collider-cache carry, grid update with Cuboid / HeightField / Polyline colliders and boundary conditions, export) for the host
// with g++ and runs them thread by thread; tests/test_device_grid3d_cpu.py compares them with the CPU restatement of the reference
// on GPU-less CI.  Nothing in the product links this file.
#define SPK_DIM 3
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline float __fdividef(float a, float b) { return a / b; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
static inline long long __float2ll_rz(float x) {
    if (x != x) return 0;
    if (x >= 9.2233720368547758e18f) return INT64_MAX;
    if (x <= -9.2233720368547758e18f) return INT64_MIN;
    return (long long)x;
}
using std::isfinite;
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
struct HkIdx { unsigned x, y, z; };
static HkIdx hk_threadIdx, hk_blockIdx, hk_blockDim;
#define threadIdx hk_threadIdx
#define blockIdx hk_blockIdx
#define blockDim hk_blockDim
#include "../../sparkl_b200/csrc/spk_grid.cu"

namespace {
template <typename K, typename... Args>
void launch(K kernel, unsigned grid, unsigned block, Args... args) {
    hk_blockDim = HkIdx{block, 1, 1};
    for (unsigned b = 0; b < grid; ++b)
        for (unsigned t = 0; t < block; ++t) {
            hk_blockIdx = HkIdx{b, 0, 0};
            hk_threadIdx = HkIdx{t, 0, 0};
            kernel(args...);
        }
}
} // namespace

extern "C" {
void hk3_reset_grid(GridBuffers next, GridBuffers prev, uint32_t nab) { if (nab) launch(reset_grid_kernel, (nab + 3) / 4, 256, next, prev, nab); }
void hk3_grid_update(GridBuffers g, uint32_t nab, const DevCollider* cols, uint32_t ncol, float h, float dt, float gx, float gy, float gz) {
    if (nab) launch(grid_update_kernel, (nab + 3) / 4, 256, g, nab, cols, ncol, h, dt, gx, gy, gz, (float*)nullptr);
}
void hk3_grid_export(GridBuffers g, uint32_t nblocks, SpkGpuGridNode* out) { if (nblocks) launch(grid_export_kernel, (nblocks * 64 + 255) / 256, 256, g, nblocks, out); }
int hk3_sizeof_gridbuffers(void) { return (int)sizeof(GridBuffers); }
int hk3_sizeof_devcollider(void) { return (int)sizeof(DevCollider); }
}
