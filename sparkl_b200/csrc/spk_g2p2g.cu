// spk_g2p2g.cu — the fused G2P -> particle update -> P2G kernel (replaces K16 `g2p2g`, src_kernels/cuda/g2p2g.rs:64-553
// and the DefaultParticleUpdater, src_kernels/cuda/particle_updater.rs:73-236) and the CFL reduction (K15,
// src_kernels/cuda/timestep.rs:13-113).
//
// One CTA processes one dispatch entry = up to `chunk` particles of ONE particle block (4x4x4 cells).
//   stage 0  the 2x2x2-block velocity tile (8 x 1 KB, contiguous per block) is staged in shared memory with TMA bulk
//            copies through the neighbour table (no hash probes); blocks absent from the previous grid read as zero
//            (g2p2g.rs:441-472).
//   stage 1  one thread per particle: gather through the sort permutation, 27-node G2P from the tile, constitutive
//            update in registers (one shared SVD), coalesced store into the NEXT particle buffer at the sorted index.
//   stage 2  P2G into a shared-memory accumulation tile laid out [octant][channel][cell] (see p2g_* below).
//   stage 3  the tile is flushed to the next grid with 8 TMA bulk reduce-adds (1.5 KB each, f32 add at L2).
#include "spk_g2p2g_kernel.cuh"

// Model-set specialisations of the default kernel (see SPK_MM_* in spk_math.cuh): same source, dead models compiled out.
namespace {
constexpr uint32_t MM_NO_PLASTICITY = SPK_MM_P(SPK_PLASTIC_NONE) | SPK_MM_P(SPK_PLASTIC_CUSTOM); // both are no-ops (gpu_plastic_model.rs:59)
constexpr uint32_t MM_SAND = SPK_MM_C(SPK_CONSTITUTIVE_COROTATED) | SPK_MM_P(SPK_PLASTIC_DRUCKER_PRAGER) | MM_NO_PLASTICITY;   // sand3, cube_through_sand3
constexpr uint32_t MM_FLUID = SPK_MM_C(SPK_CONSTITUTIVE_MONAGHAN) | MM_NO_PLASTICITY;                                              // fluids3
constexpr uint32_t MM_ELASTIC = SPK_MM_C(SPK_CONSTITUTIVE_COROTATED) | SPK_MM_C(SPK_CONSTITUTIVE_NEO_HOOKEAN) | MM_NO_PLASTICITY; // elasticity3, the N256M block
} // namespace

cudaError_t spk_launch_g2p2g(spk_ctx* c, float dt) {
    const uint32_t mm = c->model_mask;
    if (c->g2p2g_mode == 2 && c->g2p2g_specialize && mm) {
        if (!(mm & ~MM_SAND)) return spk_launch_g2p2g_with<DefaultParticleUpdaterT<MM_SAND>>(c, dt);
        if (!(mm & ~MM_FLUID)) return spk_launch_g2p2g_with<DefaultParticleUpdaterT<MM_FLUID>>(c, dt);
        if (!(mm & ~MM_ELASTIC)) return spk_launch_g2p2g_with<DefaultParticleUpdaterT<MM_ELASTIC>>(c, dt);
    }
    return spk_launch_g2p2g_with<DefaultParticleUpdater>(c, dt);
}

cudaError_t spk_launch_estimate_dt(spk_ctx* c, float min_dt, float max_dt) { return spk_launch_estimate_dt_with<DefaultParticleUpdater>(c, min_dt, max_dt); }
