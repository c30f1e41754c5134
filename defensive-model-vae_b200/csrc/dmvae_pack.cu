// dmvae_pack.cu - torch-layout parameters -> kernel-layout ("packed") weight arena.
//
// Replaces nothing in the reference by itself: it is the layout change that lets the
// fused kernels stream every GEMM operand with contiguous TMA bulk copies.  Forward
// operands are stored transposed and zero-padded, Wt[k][Np]; data-gradient operands
// are aligned copies W[n][k] (see Layout in dmvae_common.cuh).
#include "dmvae_common.cuh"

namespace dmvae {

__device__ __forceinline__ float fwd_weight(const Layout& lo, const float* __restrict__ p, int l, int k, int n) {
  if (n >= lo.N[l]) return 0.f;
  if (l == L_HEADS) {
    return n < lo.L ? p[lo.p_w[l] + n * (2 * H) + k] : p[lo.p_wlv + (n - lo.L) * (2 * H) + k];
  }
  return p[lo.p_w[l] + n * lo.K[l] + k];
}
__device__ __forceinline__ float fwd_bias(const Layout& lo, const float* __restrict__ p, int l, int n) {
  if (n >= lo.N[l]) return 0.f;
  if (l == L_HEADS) return n < lo.L ? p[lo.p_b[l] + n] : p[lo.p_blv + (n - lo.L)];
  return p[lo.p_b[l] + n];
}

__global__ void pack_kernel(const __grid_constant__ Layout lo, const float* __restrict__ p, float* __restrict__ q) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nth = gridDim.x * blockDim.x;
  for (int l = 0; l < NUM_LAYERS; ++l) {
    const int Np = lo.Np[l], K = lo.K[l];
    for (int idx = tid; idx < K * Np; idx += nth) q[lo.q_w[l] + idx] = fwd_weight(lo, p, l, idx / Np, idx % Np);
    for (int idx = tid; idx < Np; idx += nth) q[lo.q_b[l] + idx] = fwd_bias(lo, p, l, idx);
    if (lo.r_w[l] < 0) continue;
    if (l == L_HEADS) {
      for (int idx = tid; idx < 2 * lo.L * 2 * H; idx += nth) {
        const int n = idx / (2 * H), k = idx % (2 * H);
        const float w = n < lo.L ? p[lo.p_w[l] + n * 2 * H + k] : p[lo.p_wlv + (n - lo.L) * 2 * H + k];
        if (k < H) q[lo.r_w[l] + n * H + k] = w;
        else q[lo.r_heads_c + n * H + (k - H)] = w;
      }
    } else if (l == L_DEC0) {
      const int Kd = lo.L + H;
      for (int idx = tid; idx < H * H; idx += nth) q[lo.r_w[l] + idx] = p[lo.p_w[l] + (idx / H) * Kd + lo.L + (idx % H)];
      for (int idx = tid; idx < H * lo.Lzp; idx += nth) {
        const int n = idx / lo.Lzp, j = idx % lo.Lzp;
        q[lo.r_dec0z + idx] = j < lo.L ? p[lo.p_w[l] + n * Kd + j] : 0.f;
      }
    } else {
      const int cnt = lo.N[l] * lo.K[l];  // [N][K] plain copy (K == 128 for all of these)
      for (int idx = tid; idx < cnt; idx += nth) q[lo.r_w[l] + idx] = p[lo.p_w[l] + idx];
    }
  }
}

cudaError_t launch_pack(const Layout& lo, const float* params, float* packed, cudaStream_t stream) {
  pack_kernel<<<148, 256, 0, stream>>>(lo, params, packed);
  return cudaGetLastError();
}

}  // namespace dmvae
