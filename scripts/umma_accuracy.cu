// Development aid: how accurate is a 128 x 128 x 128 3xTF32 product accumulated in tensor memory, and
// how much is lost to the accumulator itself?  Variants:
//   0  one accumulator, per K step: hi*hi, lo*hi, hi*lo            (what the kernels do)
//   1  one accumulator, per K step: lo*hi, hi*lo, hi*hi
//   2  two accumulators: hi*hi in one, the two correction terms in the other, summed on the CUDA cores
//   3  one accumulator, hi*hi only (plain TF32)
// Reports max |err| / max |ref| against a double-precision host product.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

#include "../defensive-model-vae_b200/csrc/dmvae_tc.cuh"

using namespace dmvae;

constexpr int M = 128, N = 128, K = 128;

__global__ void gemm(int variant, const float* A, const float* B, float* out) {
  extern __shared__ __align__(1024) unsigned char raw[];
  float* sb_hi = reinterpret_cast<float*>(raw);      // K-major image [k step][k chunk][n group][8 n][4 k]
  float* sb_lo = sb_hi + N * K;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    const float w = B[n * K + k];
    uint32_t hi, lw;
    split_tf32(w, hi, lw);
    const int idx = (k / 8) * (N * 8) + ((k % 8) / 4) * (N * 4) + (n / 8) * 32 + (n % 8) * 4 + (k % 4);
    sb_hi[idx] = __uint_as_float(hi);
    sb_lo[idx] = __uint_as_float(lw);
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  // A[m][k]: hi at columns 256.., lo at 384..
  const int m = warp * 32 + lane;
  for (int c = 0; c < K / 4; ++c) {
    uint32_t hi[4], lw[4];
    for (int i = 0; i < 4; ++i) split_tf32(A[m * K + c * 4 + i], hi[i], lw[i]);
    tmem_st4(lane_base + 256 + c * 4, hi[0], hi[1], hi[2], hi[3]);
    tmem_st4(lane_base + 384 + c * 4, lw[0], lw[1], lw[2], lw[3]);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_tf32(M, N);
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t dh = umma_desc(smem_u32(sb_hi) + ks * N * 32, N * 16, 128);
        const uint64_t dl = umma_desc(smem_u32(sb_lo) + ks * N * 32, N * 16, 128);
        const uint32_t ah = tmem + 256 + ks * 8, al = tmem + 384 + ks * 8;
        const uint32_t acc = ks > 0;
        if (variant == 0) {
          umma_tf32_ts(tmem, ah, dh, idesc, acc);
          umma_tf32_ts(tmem, al, dh, idesc, 1u);
          umma_tf32_ts(tmem, ah, dl, idesc, 1u);
        } else if (variant == 1) {
          umma_tf32_ts(tmem, al, dh, idesc, acc);
          umma_tf32_ts(tmem, ah, dl, idesc, 1u);
          umma_tf32_ts(tmem, ah, dh, idesc, 1u);
        } else if (variant == 2) {
          umma_tf32_ts(tmem, ah, dh, idesc, acc);
          umma_tf32_ts(tmem + 128, al, dh, idesc, acc);
          umma_tf32_ts(tmem + 128, ah, dl, idesc, 1u);
        } else {
          umma_tf32_ts(tmem, ah, dh, idesc, acc);
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < N / 16; ++c) {
    uint32_t v[16], w[16];
    tmem_ld16(lane_base + c * 16, v);
    tmem_ld16(lane_base + 128 + c * 16, w);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j)
      out[m * N + c * 16 + j] = __uint_as_float(v[j]) + (variant == 2 ? __uint_as_float(w[j]) : 0.f);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  float *hA = (float*)malloc(M * K * 4), *hB = (float*)malloc(N * K * 4), *hO = (float*)malloc(M * N * 4);
  srand(1);
  for (int pass = 0; pass < 2; ++pass) {
    // pass 0: zero-mean operands (cancellation); pass 1: non-negative activations (post-relu like) with zero-mean weights
    for (int i = 0; i < M * K; ++i) { float u = (float)rand() / RAND_MAX; hA[i] = pass ? 30.f * u : 2.f * u - 1.f; }
    for (int i = 0; i < N * K; ++i) hB[i] = 0.18f * ((float)rand() / RAND_MAX) - 0.09f;
    float *dA, *dB, *dO;
    cudaMalloc(&dA, M * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dO, M * N * 4);
    cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * N * K * 4 + 1024);
    double refmax = 0;
    double* ref = (double*)malloc(M * N * 8);
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double r = 0;
        for (int k = 0; k < K; ++k) r += (double)hA[m * K + k] * (double)hB[n * K + k];
        ref[m * N + n] = r;
        if (fabs(r) > refmax) refmax = fabs(r);
      }
    // fp32 FMA chain reference error
    double e32 = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        float r = 0;
        for (int k = 0; k < K; ++k) r = fmaf(hA[m * K + k], hB[n * K + k], r);
        if (fabs(r - ref[m * N + n]) > e32) e32 = fabs(r - ref[m * N + n]);
      }
    printf("pass %d: fp32 FMA chain err %.3e\n", pass, e32 / refmax);
    for (int v = 0; v < 4; ++v) {
      gemm<<<1, 128, 2 * N * K * 4 + 1024>>>(v, dA, dB, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hO, dO, M * N * 4, cudaMemcpyDeviceToHost);
      double emax = 0, bias = 0;
      for (int i = 0; i < M * N; ++i) {
        const double d = hO[i] - ref[i];
        if (fabs(d) > emax) emax = fabs(d);
        bias += d * (ref[i] >= 0 ? 1 : -1);
      }
      printf("  variant %d: max err %.3e   mean signed err (toward larger |x|) %.3e\n", v, emax / refmax, bias / (M * N) / refmax);
    }
  }
  return 0;
}
