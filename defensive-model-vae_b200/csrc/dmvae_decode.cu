// dmvae_decode.cu - fused batched generation (K1 of SURVEY.md section 2).
//
// Replaces the op chain  condition_encoder -> cat -> decoder -> + start  that the
// reference issues per trajectory (Tools.py:55-63) or per small batch
// (Tools.py:898-912; model code Training_VAE.py:132-137, :158-167, :208-215).
//
// One persistent CTA per SM walks 128-row tiles of the batch.  Activations live in
// shared memory feature-major ([feature][row], row stride 132 floats) and are updated
// in place layer by layer; weights are streamed through the TMA ring by the producer
// warp; the latent tile is either loaded or drawn with Philox in the kernel.  With a
// start point shared by the whole launch the condition encoder is evaluated once per
// CTA and folded into the bias of the first decoder layer, so that layer contracts
// over the latent only.
#include "dmvae_common.cuh"

namespace dmvae {

struct DecodeArgs {
  Layout lo;
  const float* packed;
  const float* z;       // (B, L) or null
  const float* start;   // (B, 2) or (1, 2)
  float* out;           // (B, T, 3)
  float* z_out;         // (B, L) or null
  const float* hc_in;   // (B, 128) for MODE_FROM_HC
  float* hc_out;        // (B, 128) for MODE_COND_ONLY
  unsigned long long seed, sample_offset;
  long long B;
  int mode, add_start, stages;
};

enum DecodeMode { MODE_FULL = 0, MODE_SHARED = 1, MODE_FROM_HC = 2, MODE_COND_ONLY = 3 };

constexpr int DEC_M = 128;
constexpr int DEC_LD = DEC_M + 4;

// act[n][m] = relu(acc + bias[n]) for a 128-wide layer.
template <class C>
__device__ __forceinline__ void store_relu(const float (&acc)[C::TI][C::TJ], float* act, int ld,
                                           const float* __restrict__ bias, int warp, int lane) {
  const int i0 = C::i0(warp, lane), j0 = C::j0(warp, lane);
#pragma unroll
  for (int gj = 0; gj < C::GJ; ++gj)
#pragma unroll
    for (int v = 0; v < C::VJ; ++v) {
      const int j = gj * C::VJ + v;
      const int n = j0 + gj * C::SJ + v;
      const float b = bias[n];
#pragma unroll
      for (int gi = 0; gi < C::GI; ++gi) {
        float4 o;
        o.x = fmaxf(acc[4 * gi + 0][j] + b, 0.f);
        o.y = fmaxf(acc[4 * gi + 1][j] + b, 0.f);
        o.z = fmaxf(acc[4 * gi + 2][j] + b, 0.f);
        o.w = fmaxf(acc[4 * gi + 3][j] + b, 0.f);
        *reinterpret_cast<float4*>(act + n * ld + i0 + gi * C::SI) = o;
      }
    }
}

template <int NP3>
// 9 warps: one SM sub-partition holds 3 of them, so 3*32*regs <= 16384 -> at most 168
// registers per thread (what __launch_bounds__(288, 1) makes ptxas target).
__global__ void __launch_bounds__(BLOCK_THREADS, 1) decode_kernel(const __grid_constant__ DecodeArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Layout& lo = a.lo;
  const int L = lo.L, Lq = lo.Lq, I = lo.I;
  float* ring = reinterpret_cast<float*>(smem_raw);
  float* act = ring + a.stages * STAGE_FLOATS;      // [128][DEC_LD]
  float* zt = act + H * DEC_LD;                     // [Lq][DEC_LD]
  float* st = zt + Lq * DEC_LD;                     // [2][DEC_LD]
  float* hb = st + 2 * DEC_LD;                      // [128] folded dec0 bias (shared start)
  float* tmp = hb + H;                              // [2][128] prologue scratch
  uint64_t* full = reinterpret_cast<uint64_t*>(tmp + 2 * H);
  uint64_t* empty = full + 8;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* __restrict__ pk = a.packed;
  const long long n_tiles = (a.B + DEC_M - 1) / DEC_M;
  const bool shared_start = a.mode == MODE_SHARED;
  const bool with_cond = a.mode == MODE_FULL || a.mode == MODE_COND_ONLY;  // per-row condition encoder in the tile loop

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == CONSUMER_WARPS) {
    // ===================== producer warp: stream the weights of every tile =============
    if (lane == 0) {
      RingStateRt rs(a.stages);
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (with_cond) {
          produce(pk + lo.q_w[L_COND0], 2, H, ring, full, empty, rs);
          produce(pk + lo.q_w[L_COND1], H, H, ring, full, empty, rs);
        }
        if (a.mode == MODE_COND_ONLY) continue;
        produce(pk + lo.q_w[L_DEC0], L, H, ring, full, empty, rs);
        if (!shared_start) produce(pk + lo.q_w[L_DEC0] + L * H, H, H, ring, full, empty, rs);
        produce(pk + lo.q_w[L_DEC1], H, H, ring, full, empty, rs);
        produce(pk + lo.q_w[L_DEC2], H, H, ring, full, empty, rs);
        for (int c = 0; c < lo.NC; ++c)   // trajectories longer than 128 floats: one image per chunk of 128 outputs
          produce(pk + lo.q_w[L_DEC3] + (size_t)c * H * NP3, H, NP3, ring, full, empty, rs);
      }
    }
    return;
  }

  // ========================= consumer warps ================================================
  float sx_shared = 0.f, sy_shared = 0.f;
  if (shared_start) {
    // condition encoder for the one shared start point, folded into dec0's bias:
    //   hb[n] = b_dec0[n] + sum_k Wdec0[n][L+k] * h_c[k]
    sx_shared = a.start[0];
    sy_shared = a.start[1];
    if (tid < H) {
      const float* w0 = pk + lo.q_w[L_COND0];
      float v = pk[lo.q_b[L_COND0] + tid];
      v = fmaf(w0[tid], sx_shared, v);
      v = fmaf(w0[H + tid], sy_shared, v);
      tmp[tid] = fmaxf(v, 0.f);
    }
    consumer_sync();
    if (tid < H) {
      const float* w1 = pk + lo.q_w[L_COND1];
      float v = pk[lo.q_b[L_COND1] + tid];
      for (int k = 0; k < H; ++k) v = fmaf(w1[k * H + tid], tmp[k], v);
      tmp[H + tid] = fmaxf(v, 0.f);
    }
    consumer_sync();
    if (tid < H) {
      const float* wd = pk + lo.q_w[L_DEC0] + L * H;
      float v = pk[lo.q_b[L_DEC0] + tid];
      for (int k = 0; k < H; ++k) v = fmaf(wd[k * H + tid], tmp[H + k], v);
      hb[tid] = v;
    }
    consumer_sync();
  }

  using CM = FwdCfg<DEC_M, 128>;
  using C3 = FwdCfg<DEC_M, NP3>;
  RingStateRt rs(a.stages);

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long m0 = tile * DEC_M;
    const int valid = (int)min((long long)DEC_M, a.B - m0);

    // ---- stage the latent tile (transposed) and the start points ------------------------
    if (a.mode == MODE_COND_ONLY) {
      // no latent needed
    } else if (a.z != nullptr) {
      for (int idx = tid; idx < DEC_M * Lq; idx += CONSUMER_THREADS) {
        const int m = idx / Lq, j = idx - m * Lq;
        float v = 0.f;
        if (m < valid && j < L) v = __ldg(a.z + (m0 + m) * L + j);
        zt[j * DEC_LD + m] = v;
      }
    } else {
      const int nb = Lq >> 2;
      for (int idx = tid; idx < DEC_M * nb; idx += CONSUMER_THREADS) {
        const int jb = idx / DEC_M, m = idx - jb * DEC_M;
        const float4 g = philox_normal4(a.seed, a.sample_offset + (unsigned long long)(m0 + m), (uint32_t)jb, 0u);
        const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = jb * 4 + i;
          const float v = (j < L && m < valid) ? gv[i] : 0.f;
          zt[j * DEC_LD + m] = v;
          if (a.z_out != nullptr && j < L && m < valid) a.z_out[(m0 + m) * L + j] = v;
        }
      }
    }
    if (!shared_start && a.start != nullptr) {
      for (int idx = tid; idx < 2 * DEC_M; idx += CONSUMER_THREADS) {
        const int d = idx / DEC_M, m = idx - d * DEC_M;
        st[d * DEC_LD + m] = m < valid ? __ldg(a.start + (m0 + m) * 2 + d) : 0.f;
      }
    }
    if (a.mode == MODE_FROM_HC) {
      // h_c supplied by the caller (model.decode(z, condition), Training_VAE.py:208-215):
      // transpose (B,128) row-major into the feature-major activation tile.
      for (int idx = tid; idx < DEC_M * (H / 4); idx += CONSUMER_THREADS) {
        const int k4 = idx / DEC_M, m = idx - k4 * DEC_M;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < valid) v = __ldg(reinterpret_cast<const float4*>(a.hc_in + (m0 + m) * H) + k4);
        act[(4 * k4 + 0) * DEC_LD + m] = v.x;
        act[(4 * k4 + 1) * DEC_LD + m] = v.y;
        act[(4 * k4 + 2) * DEC_LD + m] = v.z;
        act[(4 * k4 + 3) * DEC_LD + m] = v.w;
      }
    }
    consumer_sync();

    float acc[CM::TI][CM::TJ];
    if (with_cond) {
      // cond0: (x0, y0) -> 128
      zero_acc<CM>(acc);
      consume<CM>(acc, st, DEC_LD, 2, H, ring, full, empty, rs, warp, lane);
      store_relu<CM>(acc, act, DEC_LD, pk + lo.q_b[L_COND0], warp, lane);
      consumer_sync();
      // cond1: 128 -> 128 (h_c), in place
      zero_acc<CM>(acc);
      consume<CM>(acc, act, DEC_LD, H, H, ring, full, empty, rs, warp, lane);
      consumer_sync();
      store_relu<CM>(acc, act, DEC_LD, pk + lo.q_b[L_COND1], warp, lane);
      consumer_sync();
    }
    if (a.mode == MODE_COND_ONLY) {
      // model.condition_encoder(c) (Training_VAE.py:132-137): write h_c (B,128) row-major
      for (int idx = tid; idx < DEC_M * H; idx += CONSUMER_THREADS) {
        const int k = idx / DEC_M, m = idx - k * DEC_M;
        if (m < valid) a.hc_out[(m0 + m) * H + k] = act[k * DEC_LD + m];
      }
      consumer_sync();
      continue;
    }
    // dec0: [z ; h_c] -> 128
    zero_acc<CM>(acc);
    consume<CM>(acc, zt, DEC_LD, L, H, ring, full, empty, rs, warp, lane);
    if (!shared_start) {
      consume<CM>(acc, act, DEC_LD, H, H, ring, full, empty, rs, warp, lane);
      consumer_sync();
      store_relu<CM>(acc, act, DEC_LD, pk + lo.q_b[L_DEC0], warp, lane);
    } else {
      store_relu<CM>(acc, act, DEC_LD, hb, warp, lane);
    }
    consumer_sync();
    // dec1, dec2
#pragma unroll 1
    for (int l = L_DEC1; l <= L_DEC2; ++l) {
      zero_acc<CM>(acc);
      consume<CM>(acc, act, DEC_LD, H, H, ring, full, empty, rs, warp, lane);
      consumer_sync();
      store_relu<CM>(acc, act, DEC_LD, pk + lo.q_b[l], warp, lane);
      consumer_sync();
    }
    // dec3: 128 -> 3T, no activation, + start on the x / y columns, straight to HBM; 128 output
    // columns per pass (one pass unless 3T > 128)
    for (int c = 0; c < lo.NC; ++c) {
      float o[C3::TI][C3::TJ];
      zero_acc<C3>(o);
      consume<C3>(o, act, DEC_LD, H, NP3, ring, full, empty, rs, warp, lane);
      if (C3::active(warp)) {
        const int i0 = C3::i0(warp, lane), j0 = C3::j0(warp, lane);
        const float* __restrict__ b3 = pk + lo.q_b[L_DEC3] + c * NP3;
#pragma unroll
        for (int gi = 0; gi < C3::GI; ++gi)
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int m = i0 + gi * C3::SI + r;
            if (m >= valid) continue;
            float sx = sx_shared, sy = sy_shared;
            if (!shared_start) { sx = st[m]; sy = st[DEC_LD + m]; }
            float* orow = a.out + (m0 + m) * I;
#pragma unroll
            for (int gj = 0; gj < C3::GJ; ++gj)
#pragma unroll
              for (int v = 0; v < C3::VJ; ++v) {
                const int nl = j0 + gj * C3::SJ + v;
                const int n = c * NP3 + nl;
                if (n >= I) continue;
                float val = o[4 * gi + r][gj * C3::VJ + v] + b3[nl];
                if (a.add_start) {
                  const int d = n % 3;
                  if (d == 1) val = sx + val;
                  else if (d == 2) val = sy + val;
                }
                orow[n] = val;
              }
          }
      }
    }
    {
      consumer_sync();  // act / zt / st are rewritten by the next tile
    }
  }
}

size_t decode_smem_bytes(const Layout& lo, int stages) {
  return (size_t)stages * STAGE_BYTES + (size_t)(H + lo.Lq + 2) * DEC_LD * 4 + 3 * H * 4 + 16 * 8;
}

cudaError_t launch_decode(const Layout& lo, int mode, const float* packed, const float* z, uint64_t seed,
                          uint64_t sample_offset, const float* start, const float* hc_in, float* hc_out, float* out,
                          float* z_out, long long B, int add_start, int sm_count, cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  DecodeArgs a;
  a.lo = lo; a.packed = packed; a.z = z; a.start = start; a.out = out; a.z_out = z_out;
  a.hc_in = hc_in; a.hc_out = hc_out;
  a.seed = seed; a.sample_offset = sample_offset; a.B = B;
  a.mode = mode; a.add_start = (start != nullptr) ? add_start : 0;
  int stages = 4;
  while (stages > 2 && decode_smem_bytes(lo, stages) > 232448) --stages;
  a.stages = stages;
  const size_t smem = decode_smem_bytes(lo, stages);
  const long long n_tiles = (B + DEC_M - 1) / DEC_M;
  const int grid = (int)(n_tiles < sm_count ? n_tiles : sm_count);
  cudaError_t e;
#define DMVAE_LAUNCH_DEC(NP)                                                                          \
  e = cudaFuncSetAttribute(decode_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
  if (e != cudaSuccess) return e;                                                                     \
  decode_kernel<NP><<<grid, BLOCK_THREADS, smem, stream>>>(a);
  if (lo.Ip == 32) { DMVAE_LAUNCH_DEC(32) }
  else if (lo.Ip == 64) { DMVAE_LAUNCH_DEC(64) }
  else { DMVAE_LAUNCH_DEC(128) }
#undef DMVAE_LAUNCH_DEC
  return cudaGetLastError();
}

}  // namespace dmvae
