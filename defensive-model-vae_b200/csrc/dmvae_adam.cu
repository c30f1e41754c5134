// dmvae_adam.cu - deterministic slab reduction and the flat fused Adam update (K3).
//
// reduce_kernel   grads[i] = sum over CTAs (fixed order) of that CTA's gradient slab
//                 written by train_kernel, plus the five loss terms.  Replaces the
//                 accumulation autograd performs inside loss.backward()
//                 (Training_VAE.py:362).  With ADAM it also applies the update in the
//                 same thread, so the single-GPU step is train_kernel + this kernel.
// adam_kernel     optimizer.step() of torch.optim.Adam(lr) (Training_VAE.py:332, :363):
//                 restates torch optim/adam.py::_single_tensor_adam (betas, eps defaults;
//                 no weight decay, no amsgrad) over the flat parameter arena:
//                   m  = lerp(m, g, 1-b1);  v = v*b2 + (1-b2)*g*g
//                   p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
//                 with the step-dependent scalars computed on the host in double and cast
//                 to fp32, as torch does for Python-float scalars.
#include "dmvae_common.cuh"
#include "dmvae_launch.h"

namespace dmvae {

struct AdamScalars {
  float w1;         // 1 - beta1
  float b2;         // beta2
  float w2;         // 1 - beta2
  float step_size;  // lr / (1 - beta1^t)
  float bc2_sqrt;   // sqrt(1 - beta2^t)
  float eps;
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamScalars& h) {
  // torch lerp: weight < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w)
  m = h.w1 < 0.5f ? fmaf(h.w1, g - m, m) : g - (g - m) * (1.f - h.w1);
  v = fmaf(h.w2 * g, g, v * h.b2);
  const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;
  p = p - h.step_size * (m / denom);
}

template <bool ADAM>
__global__ void reduce_kernel(const float* __restrict__ slabs, int n_slabs, int slab_stride, int n_params,
                              float w_recon, float w_kld, float w_start, float w_time, float* __restrict__ grads,
                              float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, AdamScalars h) {
  const int n4 = n_params >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* src = reinterpret_cast<const float4*>(slabs) + i;
    const int stride4 = slab_stride >> 2;
#pragma unroll 8
    for (int c = 0; c < n_slabs; ++c) {
      const float4 t = __ldcg(src + (size_t)c * stride4);
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    reinterpret_cast<float4*>(grads)[i] = s;
    if (ADAM) {
      float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      adam_update(pp.x, s.x, mm.x, vv.x, h);
      adam_update(pp.y, s.y, mm.y, vv.y, h);
      adam_update(pp.z, s.z, mm.z, vv.z, h);
      adam_update(pp.w, s.w, mm.w, vv.w, h);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
  } else if (i < n4 + (n_params & 3)) {
    const int e = (n4 << 2) + (i - n4);
    float s = 0.f;
    for (int c = 0; c < n_slabs; ++c) s += __ldcg(slabs + (size_t)c * slab_stride + e);
    grads[e] = s;
    if (ADAM) adam_update(p[e], s, m[e], v[e], h);
  }
  // the five loss terms: last block, first warp; lanes stride the slabs, fixed-order shuffle tree
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x < 32) {
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = threadIdx.x; c < n_slabs; c += 32) {
      const float* tail = slabs + (size_t)c * slab_stride + n_params;
#pragma unroll
      for (int q = 0; q < 4; ++q) t[q] += __ldcg(tail + q);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int q = 0; q < 4; ++q) t[q] += __shfl_xor_sync(0xffffffffu, t[q], o);
    if (threadIdx.x == 0) {
      // conditional_vae_loss returns the python int 0 for a term whose weight is <= 0
      // (Training_VAE.py:246-264) and leaves it out of the total
      const float start = w_start > 0.f ? t[2] : 0.f;
      const float time = w_time > 0.f ? t[3] : 0.f;
      float total = w_recon * t[0] + w_kld * t[1];
      if (w_start > 0.f) total += w_start * start;
      if (w_time > 0.f) total += w_time * time;
      float* out = grads + n_params;
      out[0] = total; out[1] = t[0]; out[2] = t[1]; out[3] = start; out[4] = time;
    }
  }
}

// step_dev != null: the step-dependent scalars are derived from *step_dev + 1 on the device (double
// arithmetic), so that the launch can be captured in a CUDA graph (dmvae_adam_step_dev).
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int n, AdamScalars h0, const long long* __restrict__ step_dev, double lr,
                            double beta1, double beta2) {
  __shared__ AdamScalars hs;
  if (threadIdx.x == 0) {
    hs = h0;
    if (step_dev != nullptr) {
      const double step = (double)(*step_dev + 1);
      hs.step_size = (float)(lr / (1.0 - pow(beta1, step)));
      hs.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, step));
    }
  }
  __syncthreads();
  const AdamScalars h = hs;
  const int n4 = n >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    adam_update(pp.x, gg.x, mm.x, vv.x, h);
    adam_update(pp.y, gg.y, mm.y, vv.y, h);
    adam_update(pp.z, gg.z, mm.z, vv.z, h);
    adam_update(pp.w, gg.w, mm.w, vv.w, h);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  } else if (i < n4 + (n & 3)) {
    const int e = (n4 << 2) + (i - n4);
    adam_update(p[e], g[e], m[e], v[e], h);
  }
}

static AdamScalars make_scalars(const DmvaeAdam& a) {
  // host double arithmetic, exactly as the Python floats in torch/optim/adam.py
  const double b1 = a.beta1, b2 = a.beta2;
  const double bc1 = 1.0 - pow(b1, (double)a.step), bc2 = 1.0 - pow(b2, (double)a.step);  // beta ** step
  AdamScalars h;
  h.w1 = (float)(1.0 - b1);
  h.b2 = (float)b2;
  h.w2 = (float)(1.0 - b2);
  h.step_size = (float)(a.lr / bc1);
  h.bc2_sqrt = (float)sqrt(bc2);
  h.eps = (float)a.eps;
  return h;
}

cudaError_t launch_reduce(const Layout& lo, const float* slabs, int n_slabs, int slab_stride, const float w[4],
                          float* grads, const DmvaeAdam* adam, float* p, float* m, float* v, cudaStream_t stream) {
  const int n = lo.n_params;
  const int threads = 128;
  const int work = (n >> 2) + (n & 3);
  const int blocks = (work + threads - 1) / threads;
  if (adam != nullptr) {
    reduce_kernel<true><<<blocks, threads, 0, stream>>>(slabs, n_slabs, slab_stride, n, w[0], w[1], w[2], w[3], grads,
                                                        p, m, v, make_scalars(*adam));
  } else {
    reduce_kernel<false><<<blocks, threads, 0, stream>>>(slabs, n_slabs, slab_stride, n, w[0], w[1], w[2], w[3], grads,
                                                         nullptr, nullptr, nullptr, AdamScalars{});
  }
  return cudaGetLastError();
}

cudaError_t launch_adam(const Layout& lo, float* p, const float* g, float* m, float* v, const DmvaeAdam& a,
                        cudaStream_t stream, const long long* step_dev) {
  const int n = lo.n_params;
  const int threads = 256;
  const int work = (n >> 2) + (n & 3);
  DmvaeAdam a1 = a;
  if (step_dev != nullptr) a1.step = 1;   // placeholder: the kernel derives the real scalars from the device counter
  adam_kernel<<<(work + threads - 1) / threads, threads, 0, stream>>>(p, g, m, v, n, make_scalars(a1), step_dev, a.lr,
                                                                      a.beta1, a.beta2);
  return cudaGetLastError();
}

}  // namespace dmvae
