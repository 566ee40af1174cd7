// spk_fluids.cu — fluids-volume recomputation: the density re-estimation pass the reference keeps in
// src_kernels/cuda/fluids_volume.rs:9-105 (two kernels, commented out of its build: src_kernels/cuda/mod.rs:2,18) and runs on the
// CPU in src/dynamics/solver/fluids_volume.rs:9-116 after the sort when `force_fluids_volume_recomputation` is set.
//
//   P2G:  node.mass += w(node, p) * p.mass            for every particle            (fluids_volume.rs:26-49 of the kernels)
//   G2P:  m = sum_nodes w * node.mass;  density = m / h^DIM;  F[0][0] = (p.mass / density) / p.volume0   for FLUID particles (:80-104)
//   then the touched node masses return to zero                                     (src/dynamics/solver/fluids_volume.rs:98-113)
//
// `spk_step` ignores the flag exactly like the reference's CUDA pipeline; the pass is the explicit stage
// `spk_stage_recompute_fluids_volume`, valid between `spk_stage_reset_grid` and `spk_stage_g2p2g` (the next grid is zeroed, the
// particle -> block map of the sort is current).  One thread per particle, both dimensions from this one source; the masses go
// through the `mass` channel of the next grid's accumulator chunks ([block][6][cells], the layout g2p2g flushes into) with
// red.global.add, the same instruction the reference kernel uses.
#include "spk_device.cuh"

namespace {

constexpr int FV_THREADS = 256;
constexpr int FV_NODES = SPK_DIM == 3 ? 27 : 9;

__device__ __forceinline__ void fv_weights(float x, float w[3]) { // QuadraticKernel::eval_all, src_core/dynamics/solver/kernel.rs:15-21
    w[0] = 0.5f * ((1.5f - x) * (1.5f - x));
    w[1] = 0.75f - ((x - 1.f) * (x - 1.f));
    w[2] = 0.5f * ((x - 0.5f) * (x - 0.5f));
}

struct FvStencil {
    float w[SPK_DIM][3];
    int t[SPK_DIM]; // stencil-base tile coordinate in [1, 4] (tile = the 2^DIM blocks b + {0,1}^DIM, 8 cells per axis)
};

__device__ __forceinline__ FvStencil fv_stencil(const float4 r0, float h, float inv_h) {
    FvStencil s;
    const float x[3] = {r0.x, r0.y, r0.z};
#pragma unroll
    for (int a = 0; a < SPK_DIM; ++a) {
        const float q = roundf(spk_div_h(x[a], h, inv_h));
        const float ref = (q - 1.f) * h - x[a];              // dir_to_associated_grid_node, particle.rs:37-47
        fv_weights(spk_div_h(-ref, h, inv_h), s.w[a]);
        const int assoc_cell = (q != q) ? -2 : (int)__float2ll_rz(q) - 2;
        s.t[a] = (assoc_cell & 3) + 1;                       // associated_cell_index_in_block_off_by_two, particle.rs:28-35
    }
    return s;
}

// address of the mass accumulator of stencil node k (x fastest, like NBH_SHIFTS, gpu_grid.rs:31-68), or nullptr when the block is
// not active (cannot happen after the sort: every particle's 2^DIM associated blocks are activated)
__device__ __forceinline__ float* fv_node_mass(const GridBuffers& g, const int32_t* __restrict__ nbr, const FvStencil& s, int k, float& weight) {
    int cell = 0, oct = 0;
    weight = 1.f;
    int kk = k;
#pragma unroll
    for (int a = 0; a < SPK_DIM; ++a) {
        const int d = kk % 3;
        kk /= 3;
        const int t = s.t[a] + d;
        cell |= (t & 3) << (2 * a);
        oct |= (t >> 2) << a;
        weight = a == 0 ? s.w[0][d] : weight * s.w[a][d];    // w[0][x] * w[1][y] (* w[2][z]), fluids_volume.rs:40-43
    }
    const int32_t nid = nbr[oct];
    if (nid < 0) return nullptr;
    return g.acc + ((size_t)nid * SPK_ACC_CHANNELS + 0) * SPK_CELLS_PER_BLOCK + SPK_ACC_SWZ((uint32_t)cell);
}

__device__ __forceinline__ const int32_t* fv_block_of(const GridBuffers& next, const uint32_t* __restrict__ slot_of, uint32_t i) {
    const uint32_t slot = slot_of[i];
    if (slot == SPK_NO_BLOCK) return nullptr;
    const uint32_t b = next.vals[slot];
    if (b >= next.cap_blocks) return nullptr;
    return next.nbr_self + (size_t)b * SPK_NBR;
}

// recompute_fluids_volume_p2g, src_kernels/cuda/fluids_volume.rs:9-49
__global__ void __launch_bounds__(FV_THREADS) fluids_volume_p2g_kernel(ParticleBuffers pb, const uint32_t* __restrict__ slot_of, GridBuffers next, uint32_t n,
                                                                        float h, float inv_h) {
    const uint32_t i = blockIdx.x * FV_THREADS + threadIdx.x;
    if (i >= n) return;
    const int32_t* nbr = fv_block_of(next, slot_of, i);
    if (!nbr) return;
    const FvStencil s = fv_stencil(pb.a0[i], h, inv_h);
    const float mass = pb.a3[i].z;
    for (int k = 0; k < FV_NODES; ++k) {
        float weight;
        float* m = fv_node_mass(next, nbr, s, k, weight);
        if (m) atomicAdd(m, weight * mass); // result unused: compiles to REDG.E.ADD.F32.FTZ.RN
    }
}

// recompute_fluids_volume_g2p, src_kernels/cuda/fluids_volume.rs:51-105
__global__ void __launch_bounds__(FV_THREADS) fluids_volume_g2p_kernel(ParticleBuffers pb, const uint32_t* __restrict__ slot_of, GridBuffers next,
                                                                        const DevModel* __restrict__ models, uint32_t n, float h, float inv_h) {
    const uint32_t i = blockIdx.x * FV_THREADS + threadIdx.x;
    if (i >= n) return;
    const float4 r0 = pb.a0[i];
    if (models[__float_as_uint(r0.w) >> SPK_F_MODEL_SHIFT].ctag != SPK_CONSTITUTIVE_MONAGHAN) return; // only the fluids (:75-77)
    const int32_t* nbr = fv_block_of(next, slot_of, i);
    if (!nbr) return;
    const FvStencil s = fv_stencil(r0, h, inv_h);
    float new_mass = 0.f;
    for (int k = 0; k < FV_NODES; ++k) {
        float weight;
        const float* m = fv_node_mass(next, nbr, s, k, weight);
        if (m) new_mass += weight * *m;
    }
    float cell_volume = h * h; // cell_width.powi(DIM)
    if (SPK_DIM == 3) cell_volume *= h;
    const float4 r3 = pb.a3[i];
    const float new_density = __fdiv_rn(new_mass, cell_volume);
    const float new_volume = __fdiv_rn(r3.z, new_density);
    pb.a1[i].x = __fdiv_rn(new_volume, r3.w); // deformation_gradient[(0, 0)]
}

// "Reset the modified cells", src/dynamics/solver/fluids_volume.rs:98-113
__global__ void fluids_volume_clear_kernel(GridBuffers next, const DeviceCounters* __restrict__ counters) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b = t / SPK_CELLS_PER_BLOCK, cell = t % SPK_CELLS_PER_BLOCK;
    if (b >= counters->nab || b >= next.cap_blocks) return;
    next.acc[((size_t)b * SPK_ACC_CHANNELS + 0) * SPK_CELLS_PER_BLOCK + cell] = 0.f;
}

} // namespace

cudaError_t spk_launch_fluids_volume(spk_ctx* c) {
    const uint32_t n = (uint32_t)c->n;
    if (!n) return cudaSuccess;
    GridBuffers& next = c->grid[1 - c->gcur];
    cudaStream_t s = c->stream;
    const float h = c->cell_width, inv_h = 1.0f / h;
    const uint32_t grid = (n + FV_THREADS - 1) / FV_THREADS;
    fluids_volume_p2g_kernel<<<grid, FV_THREADS, 0, s>>>(c->pb[c->cur], c->slot_of, next, n, h, inv_h);
    fluids_volume_g2p_kernel<<<grid, FV_THREADS, 0, s>>>(c->pb[c->cur], c->slot_of, next, c->models, n, h, inv_h);
    if (next.nab_host) fluids_volume_clear_kernel<<<(next.nab_host * SPK_CELLS_PER_BLOCK + 255) / 256, 256, 0, s>>>(next, c->d_counters);
    c->launches += 3;
    return cudaGetLastError();
}
