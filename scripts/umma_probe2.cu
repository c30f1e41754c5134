// Development aid: learn the MN-major operand layouts of tcgen05.mma kind::tf32 empirically.
// A (MN-major, from shared memory) is all zeros except one float at byte offset X; B (K-major, known
// good) is the 8x8 identity padded to N = 16, so D[m][n] = A[m][n]: the nonzero of D tells which
// (m, k) the hardware reads at offset X, for a given layout type and (LBO, SBO).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../defensive-model-vae_b200/csrc/dmvae_tc.cuh"

using namespace dmvae;

constexpr int M = 128, N = 16, K = 8;
constexpr int NX = 4096;  // offsets 0, 4, ..., 16380

__global__ void probe(uint32_t lbo, uint32_t sbo, uint32_t layout_type, uint32_t majors, int* out) {
  __shared__ __align__(1024) float sa[8192];   // 32 KB
  __shared__ __align__(1024) float sb[N * K];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ int hit[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 8192; i += blockDim.x) sa[i] = 0.f;
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int mn = i / K, k = i % K;
    sb[(k / 4) * (N * 4) + (mn / 8) * 32 + (mn % 8) * 4 + (k % 4)] = (mn == k) ? 1.f : 0.f;
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&slot, 32);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t phase = 0;
  for (int xi = 0; xi < NX; ++xi) {
    if (tid == 0) { sa[xi] = 1.0f; hit[0] = -1; hit[1] = 0; }
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        const uint64_t da = umma_desc(smem_u32(sa), lbo, sbo) | ((uint64_t)layout_type << 61);
        const uint64_t kb = umma_desc(smem_u32(sb), N * 16, 128);
        umma_tf32_ss(tmem, da, kb, umma_idesc_tf32(M, N, majors), 0u);
        umma_commit(&bar);
      }
      __syncwarp();
    }
    mbar_wait(&bar, phase);
    phase ^= 1u;
    tc_fence_after();
    uint32_t v[16];
    tmem_ld16(lane_base, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j)
      if (__uint_as_float(v[j]) != 0.f) { atomicAdd(&hit[1], 1); hit[0] = (warp * 32 + lane) * 16 + j; }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) { out[xi] = hit[1] == 1 ? hit[0] : (hit[1] == 0 ? -1 : -2); sa[xi] = 0.f; }
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tmem, 32);
}

int main() {
  int* d;
  cudaMalloc(&d, NX * 4);
  int* h = (int*)malloc(NX * 4);
  struct Cfg { uint32_t lbo, sbo, type, majors; const char* name; };
  const Cfg cfg[] = {
      {2048, 128, 0, 0, "control: K-major, no swizzle, LBO=2048 SBO=128"},
      {4096, 128, 0, UMMA_A_MN, "MN-major none"},
      {4096, 1024, 2, UMMA_A_MN, "MN-major SW128, LBO=4096 SBO=1024"},
      {1024, 4096, 2, UMMA_A_MN, "MN-major SW128, LBO=1024 SBO=4096"},
      {4096, 1024, 1, UMMA_A_MN, "MN-major SW128_BASE32B, LBO=4096 SBO=1024"},
      {4096, 512, 4, UMMA_A_MN, "MN-major SW64, LBO=4096 SBO=512"},
      {4096, 256, 6, UMMA_A_MN, "MN-major SW32, LBO=4096 SBO=256"},
  };
  for (auto& c : cfg) {
    cudaMemset(d, 0xff, NX * 4);
    probe<<<1, 128>>>(c.lbo, c.sbo, c.type, c.majors, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, NX * 4, cudaMemcpyDeviceToHost);
    int hits = 0, multi = 0;
    for (int i = 0; i < NX; ++i) { hits += h[i] >= 0; multi += h[i] == -2; }
    printf("%s: %d single hits, %d multi\n", c.name, hits, multi);
    int shown = 0;
    for (int i = 0; i < NX && shown < 80; ++i)
      if (h[i] >= 0) { printf("  X=%d->(m=%d,k=%d)", i * 4, h[i] / 16, h[i] % 16); if (++shown % 6 == 0) printf("\n"); }
    printf("\n");
  }
  return 0;
}
