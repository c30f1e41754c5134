// dmvae_dense.cu - one Linear (+ ReLU) layer on its own: y = act(x W^T + b).
//
// The reference's sub-modules are ordinary nn.Sequential / nn.Linear objects, callable on a tensor
// (Training_VAE.py:141-167: model.encoder(x), model.decoder(zc), model.fc_mu(h), model.fc_logvar(h)).  No caller in
// the reference uses them that way - encode / decode / forward do, and those run in the fused kernels - so this is
// not a hot path: a plain FP32 FFMA tile kernel that reads the weights straight from the state_dict arena
// ((out, in) row-major, as nn.Linear stores them), one launch per layer.
#include "dmvae_launch.h"

namespace dmvae {

constexpr int DN_ROWS = 32, DN_COLS = 128, DN_K = 32;

__global__ void __launch_bounds__(DN_COLS) dense_kernel(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ x,
                                                        float* __restrict__ y, long long B, int in, int out, int relu) {
  __shared__ float xs[DN_ROWS][DN_K];
  __shared__ float ws[DN_COLS][DN_K + 1];
  const long long r0 = (long long)blockIdx.x * DN_ROWS;
  const int o0 = blockIdx.y * DN_COLS, t = threadIdx.x;
  float acc[DN_ROWS];
#pragma unroll
  for (int r = 0; r < DN_ROWS; ++r) acc[r] = 0.f;
  for (int k0 = 0; k0 < in; k0 += DN_K) {
    __syncthreads();
    for (int i = t; i < DN_ROWS * DN_K; i += DN_COLS) {
      const int r = i / DN_K, k = i % DN_K;
      xs[r][k] = (r0 + r < B && k0 + k < in) ? x[(r0 + r) * in + k0 + k] : 0.f;
    }
    for (int i = t; i < DN_COLS * DN_K; i += DN_COLS) {
      const int o = i / DN_K, k = i % DN_K;
      ws[o][k] = (o0 + o < out && k0 + k < in) ? W[(size_t)(o0 + o) * in + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < DN_K; ++k) {
      const float w = ws[t][k];
#pragma unroll
      for (int r = 0; r < DN_ROWS; ++r) acc[r] = fmaf(xs[r][k], w, acc[r]);
    }
  }
  if (o0 + t < out) {
    const float bias = b != nullptr ? b[o0 + t] : 0.f;
    for (int r = 0; r < DN_ROWS && r0 + r < B; ++r) {
      const float v = acc[r] + bias;
      y[(r0 + r) * out + o0 + t] = relu ? fmaxf(v, 0.f) : v;
    }
  }
}

cudaError_t launch_dense(const float* W, const float* b, const float* x, float* y, long long B, int in, int out, int relu,
                         cudaStream_t stream) {
  const dim3 grid((unsigned int)((B + DN_ROWS - 1) / DN_ROWS), (unsigned int)((out + DN_COLS - 1) / DN_COLS));
  dense_kernel<<<grid, DN_COLS, 0, stream>>>(W, b, x, y, B, in, out, relu);
  return cudaGetLastError();
}

}  // namespace dmvae
