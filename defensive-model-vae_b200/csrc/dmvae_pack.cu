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

// Round-to-nearest TF32 (10 explicit mantissa bits) of an fp32 value, as an fp32 value.
__device__ __forceinline__ float tf32_rn(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// Weight (k, n) of a tensor-core layer in torch layout (zero in the padding).
__device__ __forceinline__ float tc_weight(const Layout& lo, const float* __restrict__ p, int t, int k, int n) {
  switch (t) {
    case TC_COND0:  // rows: weight of x0, weight of y0, bias (multiplies the ones column), zeros
      return k < 2 ? p[lo.p_w[L_COND0] + n * 2 + k] : (k == 2 ? p[lo.p_b[L_COND0] + n] : 0.f);
    case TC_COND1: return p[lo.p_w[L_COND1] + n * H + k];
    case TC_ENC0: return k < lo.I ? p[lo.p_w[L_ENC0] + n * lo.I + k] : 0.f;
    case TC_ENC1: return p[lo.p_w[L_ENC1] + n * H + k];
    case TC_ENC2: return p[lo.p_w[L_ENC2] + n * H + k];
    case TC_ENC3: return p[lo.p_w[L_ENC3] + n * H + k];
    case TC_HEADS:  // rows [0,128) multiply h_traj, [128,256) h_c (Training_VAE.py:193); columns mu then logvar
      if (n < lo.L) return p[lo.p_w[L_HEADS] + n * (2 * H) + k];
      return n < 2 * lo.L ? p[lo.p_wlv + (n - lo.L) * (2 * H) + k] : 0.f;
    case TC_DEC0: {  // contraction ordered [h_c (128) ; z (L, zero padded to Lp16)]: the shared-start
                     // generation path skips the h_c steps, and h_c stays in place in tensor memory
      const int Kd = lo.L + H;
      if (k < H) return p[lo.p_w[L_DEC0] + n * Kd + lo.L + k];
      return (k - H) < lo.L ? p[lo.p_w[L_DEC0] + n * Kd + (k - H)] : 0.f;
    }
    case TC_DEC1: return p[lo.p_w[L_DEC1] + n * H + k];
    case TC_DEC2: return p[lo.p_w[L_DEC2] + n * H + k];
    default: return n < lo.I ? p[lo.p_w[L_DEC3] + n * H + k] : 0.f;
  }
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
  // tensor-core planes: [k-step][k-chunk of 4][n-group of 8][8 n][4 k], high then low halves
  for (int t = 0; t < NUM_TC; ++t) {
    const TcLayer c = lo.tc[t];
    const int per_step = c.N * 8;
    for (int idx = tid; idx < c.K * c.N; idx += nth) {
      const int ks = idx / per_step, r3 = idx - ks * per_step;
      const int kc = r3 / (c.N * 4), r4 = r3 - kc * (c.N * 4);
      const int n = (r4 >> 5) * 8 + ((r4 & 31) >> 2);
      const int k = ks * 8 + kc * 4 + (r4 & 3);
      const float w = tc_weight(lo, p, t, k, n);
      const float hi = tf32_rn(w);
      q[c.off_hi + idx] = hi;
      q[c.off_lo + idx] = tf32_rn(w - hi);
    }
    if (c.off_thi < 0) continue;
    // data-gradient planes: [slice of 32 k][8-deep step of n][2 atoms][4 n][32 k, 32-byte units swizzled by n % 4]
    for (int idx = tid; idx < c.Kt * c.N; idx += nth) {
      const int sl = idx / (c.N * 32), r = idx - sl * (c.N * 32);
      const int n = (r >> 7) * 4 + ((r >> 5) & 3);            // atom index * 4 + row inside the atom
      const int unit = (r >> 3) & 3, kk = sl * 32 + ((unit ^ (n & 3)) << 3) + (r & 7);
      const float w = kk < c.K ? tc_weight(lo, p, t, kk, n) : 0.f;
      const float hi = tf32_rn(w);
      q[c.off_thi + idx] = hi;
      q[c.off_tlo + idx] = tf32_rn(w - hi);
    }
  }
}

cudaError_t launch_pack(const Layout& lo, const float* params, float* packed, cudaStream_t stream) {
  pack_kernel<<<148, 256, 0, stream>>>(lo, params, packed);
  return cudaGetLastError();
}

}  // namespace dmvae
