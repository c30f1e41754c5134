// dmvae_prof.cu - launch accounting and the FP32 roofline probe.
//
//   * every kernel launch of the library is counted per kernel id (always on, one
//     relaxed atomic add); with profiling enabled each launch is also bracketed by a
//     cudaEvent pair on the launching stream, so bench.py can report the live average
//     duration of the dominant kernel inside its timed region (roofline.achieved);
//   * ffma_probe_kernel measures the FP32 FFMA throughput the SMs actually sustain at
//     the clocks of the moment: the denominator of roofline.frac for the fused kernels,
//     which are compute-bound in FFMA, not HBM-bound.
#include <atomic>
#include <mutex>
#include <vector>

#include "dmvae_common.cuh"
#include "dmvae_prof.h"

namespace dmvae {

namespace {
std::atomic<long long> g_counts[K_COUNT];
std::atomic<int> g_enabled{0};
std::mutex g_mu;
struct Pair { cudaEvent_t a, b; int kernel; };
std::vector<Pair> g_pairs;       // recorded pairs of the current profiling session
std::vector<Pair> g_free;        // recycled events
constexpr size_t MAX_PAIRS = 1 << 16;
}  // namespace

const char* kernel_name(int id) {
  static const char* names[K_COUNT] = {"pack_kernel", "decode_kernel", "train_kernel(fused)", "train_kernel(fwd)",
                                       "train_kernel(bwd)", "reduce_kernel", "reduce_kernel(adam)", "adam_kernel",
                                       "loss_kernel", "loss_grad_kernel", "ffma_probe_kernel", "decode_tc_kernel",
                                       "chain_kernel", "wgrad_kernel", "reduce_tc_kernel", "train_tc_fused_kernel",
                                       "speeds_kernel", "histogram_kernel", "cells_kernel", "mpc_prepare_kernel",
                                       "mpc_track_kernel", "dense_kernel"};
  return id >= 0 && id < K_COUNT ? names[id] : "?";
}

ProfScope::ProfScope(int kernel, cudaStream_t stream) : kernel_(kernel), stream_(stream), slot_(-1) {
  g_counts[kernel].fetch_add(1, std::memory_order_relaxed);
  if (!g_enabled.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_pairs.size() >= MAX_PAIRS) return;
  Pair p;
  if (!g_free.empty()) { p = g_free.back(); g_free.pop_back(); }
  else if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return;
  p.kernel = kernel;
  cudaEventRecord(p.a, stream);
  g_pairs.push_back(p);
  slot_ = (int)g_pairs.size() - 1;
}
ProfScope::~ProfScope() {
  if (slot_ < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (slot_ < (int)g_pairs.size()) cudaEventRecord(g_pairs[slot_].b, stream_);
}

void profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (on) { for (auto& p : g_pairs) g_free.push_back(p); g_pairs.clear(); }
  g_enabled.store(on ? 1 : 0);
}

// Sums the event-pair durations per kernel id; ends the session.
cudaError_t profile_collect(double* ms, long long* launches, int n) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_enabled.store(0);
  for (int i = 0; i < n; ++i) { ms[i] = 0.0; launches[i] = 0; }
  cudaError_t rc = cudaSuccess;
  for (auto& p : g_pairs) {
    cudaError_t e = cudaEventSynchronize(p.b);
    float t = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, p.a, p.b);
    if (e != cudaSuccess) { rc = e; continue; }
    if (p.kernel < n) { ms[p.kernel] += t; launches[p.kernel] += 1; }
  }
  for (auto& p : g_pairs) g_free.push_back(p);
  g_pairs.clear();
  return rc;
}

long long launch_count(int kernel) {
  if (kernel >= 0 && kernel < K_COUNT) return g_counts[kernel].load();
  long long s = 0;
  for (int i = 0; i < K_COUNT; ++i) s += g_counts[i].load();
  return s;
}

// 16 independent FFMA chains per thread, 1024 threads per CTA, 2 CTAs per SM: enough ILP and
// warps to keep all four FMA pipes of every SM sub-partition issuing every cycle.
__global__ void __launch_bounds__(1024, 2) ffma_probe_kernel(long long iters, float* sink) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (float)(threadIdx.x + i) * 1e-3f;
  const float b = 0.999f + 1e-6f * (float)blockIdx.x, c = 1e-4f;
  for (long long it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) sink[0] = s;  // keeps the chains alive without a store in practice
}

cudaError_t launch_ffma_probe(long long iters, float* sink, int sm_count, double* flop, cudaStream_t stream) {
  const int grid = sm_count * 2;
  {
    ProfScope ps(K_FFMA_PROBE, stream);
    ffma_probe_kernel<<<grid, 1024, 0, stream>>>(iters, sink);
  }
  if (flop) *flop = 2.0 * 16.0 * 8.0 * (double)iters * 1024.0 * (double)grid;
  return cudaGetLastError();
}

}  // namespace dmvae
