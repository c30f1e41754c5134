// dmvae_probe_tc.cu - roofline probe for the tensor-core kernels: how fast does this GPU issue dense
// tcgen05.mma kind::tf32 products?  Nothing in the reference corresponds to it; bench.py times it to MEASURE the
// TF32 rate that the 3xTF32 ceiling of the training / generation kernels is derived from (one third of it),
// instead of assuming half the bf16 rate.
//
// One CTA per SM; one elected thread issues `iters` rounds of 16 dense MMAs (M = 128, K = 8 per instruction) on
// operands that never change, alternating between two accumulators so that no product waits for the one before it,
// and commits once at the end.  mode 0: both operands from shared memory, N = 256; mode 1: A from tensor memory,
// N = 128 - the shape the training chain issues.
#include "dmvae_common.cuh"
#include "dmvae_launch.h"
#include "dmvae_prof.h"
#include "dmvae_tc.cuh"

namespace dmvae {

// mbarrier wait that ends the kernel with an error after ~4 s instead of hanging the GPU (probe kernels only)
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  long long t0 = 0;
  for (int spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
    if ((spin & 255) == 255) {
      long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) __trap();
    }
  }
}

__global__ void __launch_bounds__(128, 1) tf32_probe_kernel(long long iters, int mode, float* sink) {
  __shared__ __align__(1024) float a_img[128 * 8];    // one K step: [k-chunk of 4][16 row groups][8 rows][4 k]
  __shared__ __align__(1024) float b_img[256 * 8];
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 8; i += 128) a_img[i] = 1.0f + 1e-3f * (float)(i & 63);
  for (int i = tid; i < 256 * 8; i += 128) b_img[i] = 0.5f - 1e-3f * (float)(i & 31);
  if (tid == 0) {
    mbar_init(&done, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (mode == 1) {   // an A operand in tensor memory: columns 256.. (its values do not matter)
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    tmem_st4(lane_base + 256, __float_as_uint(1.0f), __float_as_uint(0.5f), __float_as_uint(0.25f), __float_as_uint(2.0f));
    tmem_st4(lane_base + 260, __float_as_uint(1.0f), __float_as_uint(0.5f), __float_as_uint(0.25f), __float_as_uint(2.0f));
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) {
    const int N = mode == 0 ? 256 : 128;
    const uint32_t idesc = umma_idesc_tf32(128, N);
    const uint64_t a_desc = umma_desc(smem_u32(a_img), 128u * 16u, 128u);
    const uint64_t b_desc = umma_desc(smem_u32(b_img), (uint32_t)N * 16u, 128u);
    if (elect_one()) {
      for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const uint32_t d = tmem + (uint32_t)((r & 1) * N);          // two accumulators, alternating
          if (mode == 0) umma_tf32_ss(d, a_desc, b_desc, idesc, it > 0 || r > 1 ? 1u : 0u);
          else umma_tf32_ts(d, tmem + 256u, b_desc, idesc, it > 0 || r > 1 ? 1u : 0u);
        }
      }
      umma_commit(&done);
    }
    __syncwarp();
  }
  mbar_wait(&done, 0);
  tc_fence_after();
  if (warp == 0) {
    uint32_t v[4];
    tmem_ld4(tmem + ((uint32_t)(0) << 16), v);
    tmem_ld_wait();
    if (__uint_as_float(v[0]) == 123.456f) sink[0] = __uint_as_float(v[1]);   // keeps the products observable
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// mode 2 / 3: the same question for a PAIR of CTAs on the two SMs of one TPC (cta_group::2): one instruction computes
// M = 128 rows, 64 in each CTA, against N = 128 columns whose operand image is split between the two CTAs' shared
// memories (each holds 64 columns).  mode 2: A in tensor memory (64 rows per CTA, duplicated over the two lane halves
// as that instruction shape requires); mode 3: A in shared memory.  If the pair needs about half the time a single CTA
// needs for M = 128, N = 128, a 128-row tile split over two SMs halves the latency of a layer; if it needs the same
// time (as M = 64 does in a single CTA), the split buys nothing.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) tf32_pair_probe_kernel(long long iters, int mode, float* sink) {
  __shared__ __align__(1024) float a_img[64 * 8];     // this CTA's 64 rows of A (one K step)
  __shared__ __align__(1024) float b_img[64 * 8];     // this CTA's 64 columns of B
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = tid; i < 64 * 8; i += 128) {
    a_img[i] = 1.0f + 1e-3f * (float)(i & 63);
    b_img[i] = 0.5f - 1e-3f * (float)(i & 31);
  }
  if (tid == 0) {
    mbar_init(&done, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  {   // an A operand in tensor memory, all 128 lanes (the two lane halves hold the same 64 rows)
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    tmem_st4(lane_base + 256, __float_as_uint(1.0f), __float_as_uint(0.5f), __float_as_uint(0.25f), __float_as_uint(2.0f));
    tmem_st4(lane_base + 260, __float_as_uint(1.0f), __float_as_uint(0.5f), __float_as_uint(0.25f), __float_as_uint(2.0f));
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    tc_fence_after();
  }
  if (rank == 0 && warp == 0) {
    const uint32_t idesc = umma_idesc_tf32(128, 128);
    const uint64_t a_desc = umma_desc(smem_u32(a_img), 64u * 16u, 128u);
    const uint64_t b_desc = umma_desc(smem_u32(b_img), 64u * 16u, 128u);
    if (elect_one()) {
      for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const uint32_t d = tmem + (uint32_t)((r & 1) * 64);          // two accumulators (64 columns each), alternating
          const uint32_t accf = it > 0 || r > 1 ? 1u : 0u;
          if (mode == 2)
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(d),
                "r"(tmem + 256u), "l"(b_desc), "r"(idesc), "r"(accf), "r"(0u)
                : "memory");
          else
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(d),
                "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accf), "r"(0u)
                : "memory");
        }
      }
      // completion is signalled in both CTAs of the pair
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                       smem_u32(&done)),
                   "h"((unsigned short)3)
                   : "memory");
    }
    __syncwarp();
  }
  mbar_wait_bounded(&done, 0);
  tc_fence_after();
  if (warp == 0) {
    uint32_t v[4];
    tmem_ld4(tmem, v);
    tmem_ld_wait();
    if (__uint_as_float(v[0]) == 123.456f) sink[0] = __uint_as_float(v[1]);
  }
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

cudaError_t launch_tf32_probe(long long iters, int mode, float* sink, int sm_count, double* flop, cudaStream_t stream) {
  if (mode >= 2) {   // pairs of CTAs: an even number of SMs, one M = 128 x N = 128 product per instruction and PAIR
    const int grid = sm_count & ~1;
    {
      ProfScope ps(K_FFMA_PROBE, stream);
      tf32_pair_probe_kernel<<<grid, 128, 0, stream>>>(iters, mode, sink);
    }
    if (flop) *flop = 2.0 * 128.0 * 128.0 * 8.0 * 16.0 * (double)iters * (double)(grid / 2);
    return cudaGetLastError();
  }
  {
    ProfScope ps(K_FFMA_PROBE, stream);
    tf32_probe_kernel<<<sm_count, 128, 0, stream>>>(iters, mode, sink);
  }
  const double n = mode == 0 ? 256.0 : 128.0;
  if (flop) *flop = 2.0 * 128.0 * n * 8.0 * 16.0 * (double)iters * (double)sm_count;
  return cudaGetLastError();
}

}  // namespace dmvae
