// kernels2d_host.cpp — TEST INFRASTRUCTURE: compiles the per-particle body of the dim2 G2P2G kernel
// (sparkl_b200/csrc/spk_g2p2g2d_body.cuh) and the per-node dim2 grid kernels (sparkl_b200/csrc/spk_grid2d.cu: reset, grid update
// with colliders, export) for the HOST with g++, and emulates the launch glue (one CTA per dispatch entry, TMA tile staging and
// bulk-reduce flush replaced by plain loops).  tests/test_dim2_cpu.py drives it in lock-step with the dim2 CPU restatement, so the
// arithmetic and index logic of the CUDA sources is checked on GPU-less CI.  Nothing in the product links this file.
#define SPK_DIM 2
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline long long __float2ll_rz(float x) { // cvt.rzi.s64.f32: saturating, NaN -> 0
    if (x != x) return 0;
    if (x >= 9.2233720368547758e18f) return INT64_MAX;
    if (x <= -9.2233720368547758e18f) return INT64_MIN;
    return (long long)x;
}
static inline float atomicAdd(float* p, float v) { const float old = *p; *p = old + v; return old; }
using std::isfinite;
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
struct HkIdx { unsigned x, y, z; };
static HkIdx hk_threadIdx, hk_blockIdx, hk_blockDim;
#define threadIdx hk_threadIdx
#define blockIdx hk_blockIdx
#define blockDim hk_blockDim
#include "../../sparkl_b200/csrc/spk_g2p2g2d_body.cuh"
#include "../../sparkl_b200/csrc/spk_grid2d.cu"

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
int hk_chunk(void) { return SPK_G2P2G_CHUNK; }
// the glue of g2p2g2d_kernel (spk_g2p2g2d.cu) with the tile staged / flushed by plain loops
void hk_g2p2g(ParticleBuffers src, ParticleBuffers dst, float4* vel_out, const float4* cold, const uint32_t* perm, const DispatchEntry* dispatch, uint32_t ndispatch,
              const DevModel* models, const int32_t* nbr_self, const int32_t* nbr_prev, float* acc, const float4* prev_vel, float h, float dt, uint32_t* dt_raw_bits) {
    const float inv_h = 1.0f / h;
    std::vector<float4> s_vel(SPK_TILE_CELLS);
    std::vector<float> s_acc(SPK_NBR * G2P2G2D_ACC_BLOCK);
    for (uint32_t b = 0; b < ndispatch; ++b) {
        const DispatchEntry ent = dispatch[b];
        std::fill(s_acc.begin(), s_acc.end(), 0.f);
        for (int o = 0; o < SPK_NBR; ++o) {
            const int pb = nbr_prev[ent.block * SPK_NBR + o];
            for (int cidx = 0; cidx < SPK_CELLS_PER_BLOCK; ++cidx)
                s_vel[o * SPK_CELLS_PER_BLOCK + cidx] = pb >= 0 ? prev_vel[(size_t)pb * SPK_CELLS_PER_BLOCK + cidx] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (uint32_t tid = 0; tid < ent.count; ++tid) {
            Particle2dWork k;
            g2p2g2d_load(src, perm[ent.first + tid], h, inv_h, k);
            g2p2g2d_update<DefaultParticleUpdater2>(s_vel.data(), models, cold, h, dt, k, nullptr);
            const unsigned bits = g2p2g2d_dt_bits<DefaultParticleUpdater2>(models, h, k, nullptr);
            if (bits < *dt_raw_bits) *dt_raw_bits = bits;
            g2p2g2d_store(dst, vel_out, ent.first + tid, k);
            g2p2g2d_scatter(s_acc.data(), h, inv_h, dt, k);
        }
        for (int o = 0; o < SPK_NBR; ++o) {
            const int nid = nbr_self[ent.block * SPK_NBR + o];
            if (nid < 0) continue;
            for (int i = 0; i < G2P2G2D_ACC_BLOCK; ++i) acc[(size_t)nid * G2P2G2D_ACC_BLOCK + i] += s_acc[o * G2P2G2D_ACC_BLOCK + i];
        }
    }
}
void hk_reset_grid(GridBuffers next, GridBuffers prev, uint32_t nab) {
    if (nab) launch(reset_grid_kernel, (nab + BLOCKS_PER_CTA - 1) / BLOCKS_PER_CTA, 256, next, prev, nab);
}
void hk_grid_update(GridBuffers g, uint32_t nab, const DevCollider* cols, uint32_t ncol, float h, float dt, float gx, float gy) {
    if (nab) launch(grid_update_kernel, (nab + BLOCKS_PER_CTA - 1) / BLOCKS_PER_CTA, 256, g, nab, cols, ncol, h, dt, gx, gy);
}
void hk_grid_export(GridBuffers g, uint32_t nblocks, SpkGpuGridNode* out) {
    if (nblocks) launch(grid_export_kernel, (nblocks * CELLS + 255) / 256, 256, g, nblocks, out);
}
int hk_sizeof_gridbuffers(void) { return (int)sizeof(GridBuffers); }
int hk_sizeof_devcollider(void) { return (int)sizeof(DevCollider); }
int hk_sizeof_particlebuffers(void) { return (int)sizeof(ParticleBuffers); }
}
