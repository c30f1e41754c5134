// dmvae_loss.cu - conditional_vae_loss on its own (Training_VAE.py:229-268), for callers
// that use the nn.Module / autograd surface instead of the fused train step.
//
// loss_kernel       five scalars [total, recon, kld, start, time]; one block, fixed
//                   reduction order (the unfused surface is used at the reference's batch
//                   sizes, 16-135 rows; the fused kernel owns the large-batch path).
// loss_grad_kernel  d/d(recon), d/d(mu), d/d(logvar) for upstream gradients of all five
//                   outputs (read on the device, no host sync).
#include "dmvae_common.cuh"
#include "dmvae_launch.h"

namespace dmvae {

constexpr int LOSS_THREADS = 1024;

__global__ void __launch_bounds__(LOSS_THREADS) loss_kernel(int T, int L, long long B, const float* __restrict__ recon,
                                                            const float* __restrict__ x, const float* __restrict__ mu,
                                                            const float* __restrict__ logvar, float w_recon,
                                                            float w_kld, float w_start, float w_time,
                                                            float* __restrict__ losses) {
  __shared__ float red[32][4];
  const int I = 3 * T;
  float s_rec = 0.f, s_kld = 0.f, s_start = 0.f, s_time_sq = 0.f, s_mono = 0.f;
  for (long long m = threadIdx.x; m < B; m += LOSS_THREADS) {
    const float* r = recon + m * I;
    const float* t = x + m * I;
    float prev = 0.f;
    for (int k = 0; k < T; ++k) {
      const float r0 = r[3 * k];
      float d = r0 - t[3 * k];
      s_rec = fmaf(d, d, s_rec);
      if (k == 0) s_time_sq = fmaf(r0, r0, s_time_sq);
      else if (r0 - prev < 0.f) s_mono -= r0 - prev;
      prev = r0;
      for (int c = 1; c < 3; ++c) {
        d = r[3 * k + c] - t[3 * k + c];
        s_rec = fmaf(d, d, s_rec);
        if (k == 0) s_start = fmaf(d, d, s_start);
      }
    }
    for (int j = 0; j < L; ++j) {
      const float a = mu[m * L + j], lv = logvar[m * L + j];
      s_kld += 1.f + lv - a * a - expf(lv);
    }
  }
  const float invB = 1.f / (float)B;
  float v[4] = {s_rec * (invB / (float)I), -0.5f * s_kld * (invB / (float)L), s_start * (invB * 0.5f),
                s_time_sq * invB + (T > 1 ? s_mono * (invB / (float)(T - 1)) : 0.f)};
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
    for (int q = 0; q < 4; ++q) red[warp][q] = v[q];
  __syncthreads();
  if (threadIdx.x == 0) {
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    for (int w = 0; w < LOSS_THREADS / 32; ++w)
      for (int q = 0; q < 4; ++q) t[q] += red[w][q];
    const float start = w_start > 0.f ? t[2] : 0.f;
    const float time = w_time > 0.f ? t[3] : 0.f;
    float total = w_recon * t[0] + w_kld * t[1];
    if (w_start > 0.f) total += w_start * start;
    if (w_time > 0.f) total += w_time * time;
    losses[0] = total; losses[1] = t[0]; losses[2] = t[1]; losses[3] = start; losses[4] = time;
  }
}

// g_out[5]: upstream gradients of (total, recon, kld, start, time).  Effective weights:
//   e_r = g_total*w_r + g_recon, e_k = g_total*w_k + g_kld,
//   e_s = (w_s > 0) ? g_total*w_s + g_start : 0,  e_t likewise.
__global__ void loss_grad_kernel(int T, int L, long long B, const float* __restrict__ recon,
                                 const float* __restrict__ x, const float* __restrict__ mu,
                                 const float* __restrict__ logvar, float w_recon, float w_kld, float w_start,
                                 float w_time, const float* __restrict__ g_out, float* __restrict__ g_recon,
                                 float* __restrict__ g_mu, float* __restrict__ g_logvar) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= B) return;
  const int I = 3 * T;
  const float gt = g_out ? g_out[0] : 1.f;
  const float e_r = gt * w_recon + (g_out ? g_out[1] : 0.f);
  const float e_k = gt * w_kld + (g_out ? g_out[2] : 0.f);
  const float e_s = w_start > 0.f ? gt * w_start + (g_out ? g_out[3] : 0.f) : 0.f;
  const float e_t = w_time > 0.f ? gt * w_time + (g_out ? g_out[4] : 0.f) : 0.f;
  const float invB = 1.f / (float)B;
  const float c_rec = e_r * 2.f * invB / (float)I, c_start = e_s * invB, c_t0 = e_t * 2.f * invB;
  const float c_mono = T > 1 ? e_t * invB / (float)(T - 1) : 0.f;
  if (g_recon != nullptr) {
    const float* r = recon + m * I;
    const float* t = x + m * I;
    float* g = g_recon + m * I;
    float g_prev = 0.f, r_prev = 0.f;
    for (int k = 0; k < T; ++k) {
      const float r0 = r[3 * k];
      float gv = c_rec * (r0 - t[3 * k]);
      if (k == 0) {
        gv = fmaf(c_t0, r0, gv);
      } else {
        if (r0 - r_prev < 0.f) { gv -= c_mono; g_prev += c_mono; }
        g[3 * (k - 1)] = g_prev;
      }
      g_prev = gv; r_prev = r0;
      for (int c = 1; c < 3; ++c) {
        const float d = r[3 * k + c] - t[3 * k + c];
        float gc = c_rec * d;
        if (k == 0) gc = fmaf(c_start, d, gc);
        g[3 * k + c] = gc;
      }
    }
    g[3 * (T - 1)] = g_prev;
  }
  if (g_mu != nullptr && g_logvar != nullptr) {
    const float c_k = e_k * invB / (float)L;
    for (int j = 0; j < L; ++j) {
      const float a = mu[m * L + j], lv = logvar[m * L + j];
      g_mu[m * L + j] = c_k * a;
      g_logvar[m * L + j] = -0.5f * c_k * (1.f - expf(lv));
    }
  }
}

cudaError_t launch_loss(const Layout& lo, long long B, const float* recon, const float* x, const float* mu,
                        const float* logvar, const float w[4], float* losses, cudaStream_t stream) {
  loss_kernel<<<1, LOSS_THREADS, 0, stream>>>(lo.T, lo.L, B, recon, x, mu, logvar, w[0], w[1], w[2], w[3], losses);
  return cudaGetLastError();
}

cudaError_t launch_loss_grad(const Layout& lo, long long B, const float* recon, const float* x, const float* mu,
                             const float* logvar, const float w[4], const float* g_out, float* g_recon, float* g_mu,
                             float* g_logvar, cudaStream_t stream) {
  const int threads = 128;
  const long long blocks = (B + threads - 1) / threads;
  loss_grad_kernel<<<(unsigned)blocks, threads, 0, stream>>>(lo.T, lo.L, B, recon, x, mu, logvar, w[0], w[1], w[2],
                                                             w[3], g_out, g_recon, g_mu, g_logvar);
  return cudaGetLastError();
}

}  // namespace dmvae
