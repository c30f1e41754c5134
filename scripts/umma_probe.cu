// Development aid: verifies the MN-major operand layout of tcgen05.mma kind::tf32 that the
// training kernels rely on (umma_desc / mn_image_index in dmvae_tc.cuh): one 128 x N x 8 MMA per
// variant, shared-memory images built with mn_image_index, result checked against the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe scripts/umma_probe.cu && ./umma_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../defensive-model-vae_b200/csrc/dmvae_tc.cuh"

using namespace dmvae;

constexpr int M = 128, K = 8;

__host__ __device__ inline float aval(int m, int k) { return (float)((m * 3 + k * 5) % 7) - 3.f; }
__host__ __device__ inline float bval(int n, int k) { return (float)((n * 2 + k * 3) % 5) - 2.f; }

// mode 0: SS, A and B MN-major.  mode 1: TS (A from tensor memory), B MN-major.
__global__ void probe(int mode, int N, int a_lbo, int a_sbo, int b_lbo, int b_sbo, float* out) {
  __shared__ __align__(1024) float sa[4096];
  __shared__ __align__(1024) float sb[4096];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 4096; i += blockDim.x) { sa[i] = 0.f; sb[i] = 0.f; }
  __syncthreads();
  for (int i = tid; i < M * K; i += blockDim.x) sa[mn_image_index(i / K, i % K, a_lbo / 4, a_sbo / 4)] = aval(i / K, i % K);
  for (int i = tid; i < N * K; i += blockDim.x) sb[mn_image_index(i / K, i % K, b_lbo / 4, b_sbo / 4)] = bval(i / K, i % K);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&slot, 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  if (mode == 1) {  // A[m][k] into columns 128..135
    for (int c = 0; c < 2; ++c) {
      uint32_t v[4];
      for (int i = 0; i < 4; ++i) v[i] = __float_as_uint(aval(warp * 32 + lane, c * 4 + i));
      tmem_st4(lane_base + 128 + c * 4, v[0], v[1], v[2], v[3]);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    if (elect_one()) {
      const uint64_t da = umma_desc(smem_u32(sa), a_lbo, a_sbo, 1u), db = umma_desc(smem_u32(sb), b_lbo, b_sbo, 1u);
      if (mode == 0) umma_tf32_ss(tmem, da, db, umma_idesc_tf32(M, N, UMMA_A_MN | UMMA_B_MN), 0u);
      else umma_tf32_ts(tmem, tmem + 128, db, umma_idesc_tf32(M, N, UMMA_B_MN), 0u);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < N / 16; ++c) {
    uint32_t v[16];
    tmem_ld16(lane_base + c * 16, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * N + c * 16 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int main() {
  float* d;
  cudaMalloc(&d, M * 128 * 4);
  float* h = (float*)malloc(M * 128 * 4);
  struct Cfg { int mode, N, a_lbo, a_sbo, b_lbo, b_sbo; };
  const Cfg cfg[] = {
      {0, 32, 1024, 512, 1024, 512},    // A: 4 mn-atoms of [2 k-atoms][512 B]; B one atom
      {0, 32, 512, 2048, 512, 512},     // A: [k-atom][mn-atom] order
      {0, 16, 1024, 512, 1024, 512},    // N = 16 inside a 32-wide atom row
      {0, 128, 1024, 512, 512, 2048},   // B with 4 mn-atoms
      {1, 32, 0, 0, 1024, 512},
      {1, 128, 0, 0, 1024, 512},
      {1, 64, 0, 0, 4096, 512},
  };
  int rc = 0;
  for (auto& c : cfg) {
    cudaMemset(d, 0, M * 128 * 4);
    probe<<<1, 128>>>(c.mode, c.N, c.a_lbo, c.a_sbo, c.b_lbo, c.b_sbo, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, M * c.N * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < c.N; ++n) {
        float r = 0.f;
        for (int k = 0; k < K; ++k) r += aval(m, k) * bval(n, k);
        if (h[m * c.N + n] != r) ++bad;
      }
    printf("%s N=%d A(lbo %d, sbo %d) B(lbo %d, sbo %d): %d / %d mismatches\n", c.mode ? "TS" : "SS", c.N, c.a_lbo, c.a_sbo,
           c.b_lbo, c.b_sbo, bad, M * c.N);
    rc |= bad != 0;
  }
  return rc;
}
