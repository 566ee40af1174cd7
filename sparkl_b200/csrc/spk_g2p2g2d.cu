// spk_g2p2g2d.cu — dim2 build of the fused G2P -> particle update -> P2G kernel (K16 `g2p2g`, src_kernels/cuda/g2p2g.rs:64-553
// with the `dim2` feature: 4x4-cell blocks, a 2x2-block / 8x8-node shared tile, 9-node stencil) and of the CFL reduction (K15,
// src_kernels/cuda/timestep.rs:13-113).  Compiled only into libsparkl_b200_2d.so (-DSPK_DIM=2); spk_g2p2g.cu is its 3D sibling.
//
// One CTA = one dispatch entry = up to 128 particles of ONE particle block.  Same data flow as the 3D kernel, sized for 2D:
//   stage 0  the 2x2-block velocity tile (4 x 256 B, contiguous per block) is staged with TMA bulk copies through the neighbour
//            table; blocks absent from the previous grid read as zero (g2p2g.rs:441-472)
//   stage 1  one thread per particle: gather through the sort permutation, 9-node G2P, constitutive update in registers (one
//            shared 2x2 SVD), coalesced store into the NEXT particle buffer at the sorted index
//   stage 2  P2G with shared-memory float atomics into a [quadrant][channel][cell] tile (9 nodes x 5 channels per particle:
//            2D blocks hold ~64 particles, the cell-binned register accumulation of the 3D kernel does not pay here)
//   stage 3  the tile is flushed with 4 TMA bulk reduce-adds (384 B each); the CFL candidate of the next substep rides along
#include "spk_g2p2g2d_kernel.cuh"

cudaError_t spk_launch_g2p2g(spk_ctx* c, float dt) { return spk_launch_g2p2g_with<DefaultParticleUpdater2>(c, dt); }

cudaError_t spk_launch_estimate_dt(spk_ctx* c, float min_dt, float max_dt) { return spk_launch_estimate_dt_with<DefaultParticleUpdater2>(c, min_dt, max_dt); }
