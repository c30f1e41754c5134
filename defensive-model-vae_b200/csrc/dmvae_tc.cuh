// dmvae_tc.cuh - tcgen05 / tensor-memory PTX shared by the tensor-core kernels
// (dmvae_decode_tc.cu, dmvae_train_tc.cu).  sm_100a only.
//
// Every dense product on this path is error-compensated 3xTF32,
//     a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo,   x_hi = tf32_rn(x),  x_lo = x - x_hi,
// accumulated in fp32 in tensor memory: one TF32 pass (10-bit mantissa) cannot hold the
// 1e-5 fp32 tolerance of the reference's PyTorch-CPU results.
//
// Operand images.  A shared-memory operand is a grid of 128-byte core matrices: 8 indices
// of the "long" dimension (rows of a batch tile, or output features of a weight) times 4
// consecutive indices of the "short" dimension (16 bytes).  The same image can be read
//   K-major   (contraction along the short dimension: LBO = byte distance between the two
//              4-wide chunks of one 8-deep K step, SBO = distance between 8-index groups)
//   MN-major  (contraction along the long dimension: SBO = distance between 4-wide chunks,
//              LBO = distance between 8-index groups; one MMA consumes one group)
// which is what lets one weight image serve the forward GEMM and the data-gradient GEMM,
// and one activation image serve as either side of the weight-gradient GEMM.
#pragma once

#include "dmvae_common.cuh"

namespace dmvae {

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor], kind::tf32, issued by one thread for the CTA
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem descriptor] * B[smem descriptor]
__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every tcgen05 operation issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
          taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor (bit layout: cute::UMMA::SmemDescriptor).
//   layout_type 0 (no swizzle), K-major: LBO = byte distance between the two 4-wide chunks of an 8-deep
//     K step, SBO = distance between 8-index groups along M / N.
//   layout_type 1 (SWIZZLE_128B_BASE32B), MN-major: the only MN-major layout the tensor core accepts
//     for TF32 (found empirically on B200, scripts/umma_probe*.cu: with any other type the MMA reads
//     zeros).  Element (mn, k) lives at byte
//       (mn / 32) * LBO + (k / 4) * SBO + (k % 4) * 128 + ((((mn / 8) % 4) ^ (k % 4)) * 32) + (mn % 8) * 4
//     i.e. rows of 32 consecutive mn per k, atoms of 4 k, 32-byte units XOR-swizzled by k % 4.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type = 0u) {
  uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
  d |= (uint64_t)layout_type << 61;
  return d;                // base offset 0
}
// float index of element (mn, k) in an MN-major image: lbo / sbo in floats
__host__ __device__ __forceinline__ int mn_image_index(int mn, int k, int lbo_f, int sbo_f) {
  return (mn >> 5) * lbo_f + (k >> 2) * sbo_f + (k & 3) * 32 + ((((mn >> 3) & 3) ^ (k & 3)) << 3) + (mn & 7);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, TF32 x TF32.
constexpr uint32_t UMMA_A_MN = 1u << 15, UMMA_B_MN = 1u << 16;
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N, uint32_t majors = 0u) {
  return (1u << 4) | (2u << 7) | (2u << 10) | majors | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// true in exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// x = hi + lo with hi = round-to-nearest TF32 of x; lo is cut to TF32 by the tensor core itself
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
// The split the tensor core makes by itself: it reads only the TF32 bits of an fp32 word (sign, exponent, ten mantissa
// bits - the word is CUT, not rounded), so a raw fp32 operand already is its own high half  hi = cut(x);  what remains,
// x - cut(x), is exact in fp32 and is rounded to TF32 here so that the cut leaves it alone.  One shared-memory image
// less to write than split_tf32 needs (the raw operand stays in place).
__device__ __forceinline__ float tf32_cut_low(float x) {
  const float lo = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  return __uint_as_float((__float_as_uint(lo) + 0x1000u) & 0xffffe000u);
}
// the same split in two instructions (cvt.rna rounds to nearest, ties away from zero)
__device__ __forceinline__ void split_tf32_cvt(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// TMA bulk store shared -> global (bulk async-group completion)
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

}  // namespace dmvae
