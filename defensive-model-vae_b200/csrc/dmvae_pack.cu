// dmvae_pack.cu - torch-layout parameters -> kernel-layout ("packed") weight arena.
//
// Replaces nothing in the reference by itself: it is the layout change that lets the
// fused kernels stream every GEMM operand with contiguous TMA bulk copies.  Forward
// operands are stored transposed and zero-padded, Wt[k][Np]; data-gradient operands
// are aligned copies W[n][k] (see Layout in dmvae_common.cuh).
#include "dmvae_common.cuh"
#include "dmvae_pack.cuh"

namespace dmvae {

// One thread per packed element: every segment of the arena owns a run of blocks, so that a full repack is a
// single short wave.
__global__ void __launch_bounds__(PACK_THREADS) pack_kernel(const __grid_constant__ Layout lo, const __grid_constant__ PackPlan plan,
                                                            const float* __restrict__ p, float* __restrict__ q,
                                                            long long* step_inc) {
  if (step_inc != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *step_inc += 1;   // this launch completes update t + 1
  int sgm = 0;
  while (sgm + 1 < plan.n && (int)blockIdx.x >= plan.block0[sgm + 1]) ++sgm;
  const int idx = ((int)blockIdx.x - plan.block0[sgm]) * PACK_THREADS + (int)threadIdx.x;
  if (idx >= plan.count[sgm]) return;
  pack_element(lo, plan.type[sgm], plan.id[sgm], idx, p, q);
}

cudaError_t launch_pack(const Layout& lo, const float* params, float* packed, cudaStream_t stream, long long* step_inc) {
  const PackPlan plan = make_pack_plan(lo);
  pack_kernel<<<plan.block0[plan.n], PACK_THREADS, 0, stream>>>(lo, plan, params, packed, step_inc);
  return cudaGetLastError();
}

}  // namespace dmvae
