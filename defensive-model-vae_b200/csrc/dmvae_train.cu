// dmvae_train.cu - fused training pass (K2 of SURVEY.md section 2) and the unfused
// forward / backward pair behind the autograd surface.
//
// Replaces, per step, the reference's  offset transform -> model(...) ->
// conditional_vae_loss(...) -> loss.backward()  (Training_VAE.py:345-362; model
// :180-226, loss :229-268), i.e. ~25 forward ops, ~25 loss ops and the autograd walk.
//
// One persistent CTA per SM walks M-row tiles of the batch (M = 64 or 32).  Per tile:
//   forward   eleven FFMA GEMMs over activations kept feature-major in three shared
//             memory tiles; weights stream through the TMA ring; every hidden
//             activation is also written to a per-CTA stash in global memory (it stays
//             in L2) for the backward pass;
//   loss      one thread per row: the five terms and d(total)/d(recon) in place;
//   backward  per layer a weight-gradient GEMM (contraction over the tile's rows, both
//             operands vectorised along it) accumulated into this CTA's private
//             gradient slab, and a data-gradient GEMM (weights from the ring) whose
//             result is masked by relu' and written over the layer input in place.
// The slabs are summed in a fixed order by reduce_kernel (dmvae_adam.cu), so a step is
// bit-reproducible run to run.
//
// Trajectories longer than 128 floats (3 * seq_len > 128, up to seq_len = 400): the first encoder layer
// contracts over the flattened trajectory and the last decoder layer produces it, 128 features at a time.
// x_rel and recon / d(total)/d(recon) then live feature-major in the CTA's stash (L2) instead of shared
// memory: the encoder accumulates over chunks of x_rel, the decoder writes recon chunk by chunk, the loss
// walks the stash (one thread per row, coalesced across rows), and the backward pass of those two layers
// loops over the same chunks.
#include "dmvae_common.cuh"
#include "dmvae_launch.h"

namespace dmvae {

enum StashSlot { S_HC1 = 0, S_HC, S_E1, S_E2, S_E3, S_E4, S_D1, S_D2, S_D3, N_STASH };

struct TrainArgs {
  Layout lo;
  const float* packed;
  const float* x;        // fused: absolute (B,T,3); forward-only: relative
  const float* start;    // forward-only: (B,2)
  const float* eps;      // (B,L) or null (Philox)
  float* stash;          // activation stash
  float* slabs;          // [gridDim.x][slab_stride]
  float* recon; float* mu; float* logvar; float* hc;                        // forward-only outputs
  const float* g_recon; const float* g_mu; const float* g_logvar; const float* g_hc;  // backward-only inputs
  unsigned long long seed, sample_offset, step;
  long long B;
  long long stash_stride;  // floats per stash unit (per CTA when fused, per tile otherwise)
  float w_recon, w_kld, w_start, w_time, inv_batch;
  int stages, mode, slab_stride;
};

// rows of the small shared-memory tiles (each row is LD floats)
struct SmallRows {
  int xt, gt, ml, zt, ep, st, total;
};
__host__ __device__ inline SmallRows small_rows(const Layout& lo) {
  SmallRows r;
  int o = 0;
  r.xt = o; o += lo.Ip;                 // x_rel, feature-major, rows >= I are zero
  r.gt = o; o += lo.Ip;                 // recon, then d(total)/d(recon)
  r.ml = o; o += round_up(2 * lo.L, 4); // mu rows then logvar rows; later their gradients
  r.zt = o; o += lo.Lq;                 // z, later d/dz
  r.ep = o; o += lo.Lq;                 // eps
  r.st = o; o += 2;                     // start x, y
  r.total = o;
  return r;
}

// floats of the long-trajectory areas of a stash unit: x_rel [Ipt][M], then recon / gradient [Ipt][M]
__host__ __device__ inline size_t long_floats(const Layout& lo, int M) { return lo.NC > 1 ? (size_t)lo.Ipt * M : 0; }

template <int M>
__host__ __device__ inline size_t train_smem_bytes(const Layout& lo, int stages) {
  constexpr int LD = M + 4;
  return (size_t)stages * STAGE_BYTES + (size_t)(3 * H + small_rows(lo).total) * LD * 4 + 64 * 4 + 16 * 8;
}

// ---------------------------------------------------------------------------------------
// epilogues
// ---------------------------------------------------------------------------------------
// tile[n][m] = act(acc + bias[n]) for n < n_rows; optionally also to the dense stash.
template <class C, int M, bool RELU>
__device__ __forceinline__ void store_fwd(const float (&acc)[C::TI][C::TJ], float* tile, const float* __restrict__ bias,
                                          float* stash, int n_rows, int warp, int lane) {
  constexpr int LD = M + 4;
  if (!C::active(warp)) return;
  const int i0 = C::i0(warp, lane), j0 = C::j0(warp, lane);
#pragma unroll
  for (int gj = 0; gj < C::GJ; ++gj)
#pragma unroll
    for (int v = 0; v < C::VJ; ++v) {
      const int j = gj * C::VJ + v;
      const int n = j0 + gj * C::SJ + v;
      if (n >= n_rows) continue;
      const float b = bias[n];
#pragma unroll
      for (int gi = 0; gi < C::GI; ++gi) {
        float4 o = make_float4(acc[4 * gi + 0][j] + b, acc[4 * gi + 1][j] + b, acc[4 * gi + 2][j] + b,
                               acc[4 * gi + 3][j] + b);
        if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4*>(tile + n * LD + i0 + gi * C::SI) = o;
        if (stash != nullptr) *reinterpret_cast<float4*>(stash + n * M + i0 + gi * C::SI) = o;
      }
    }
}

enum EpiMode { EPI_MASK = 0, EPI_STORE = 1, EPI_ACCUM = 2 };
// data-gradient epilogue on tile[k][m], k < k_rows:
//   EPI_MASK  tile = acc where tile > 0 else 0   (relu' of the layer input, in place)
//   EPI_STORE tile = acc
//   EPI_ACCUM tile += acc
template <class C, int M, int MODE>
__device__ __forceinline__ void store_dgrad(const float (&acc)[C::TI][C::TJ], float* tile, int k_rows, int warp,
                                            int lane) {
  constexpr int LD = M + 4;
  if (!C::active(warp)) return;
  const int i0 = C::i0(warp, lane), j0 = C::j0(warp, lane);
#pragma unroll
  for (int gj = 0; gj < C::GJ; ++gj)
#pragma unroll
    for (int v = 0; v < C::VJ; ++v) {
      const int j = gj * C::VJ + v;
      const int k = j0 + gj * C::SJ + v;
      if (k >= k_rows) continue;
#pragma unroll
      for (int gi = 0; gi < C::GI; ++gi) {
        float4* p = reinterpret_cast<float4*>(tile + k * LD + i0 + gi * C::SI);
        float4 o = make_float4(acc[4 * gi + 0][j], acc[4 * gi + 1][j], acc[4 * gi + 2][j], acc[4 * gi + 3][j]);
        if (MODE == EPI_MASK) {
          const float4 a = *p;
          o.x = a.x > 0.f ? o.x : 0.f; o.y = a.y > 0.f ? o.y : 0.f;
          o.z = a.z > 0.f ? o.z : 0.f; o.w = a.w > 0.f ? o.w : 0.f;
        } else if (MODE == EPI_ACCUM) {
          const float4 a = *p;
          o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
        }
        *p = o;
      }
    }
}

// ---------------------------------------------------------------------------------------
// weight gradient: dW[n][k] (+)= sum_m D[n][m] * A[k][m]
// ---------------------------------------------------------------------------------------
// Destination rows in the slab (torch layout): row n starts at b0 + n*ld for n < split,
// at b1 + (n-split)*ld otherwise (fc_mu / fc_logvar are two tensors).
struct RowDst {
  float* b0;
  float* b1;
  int split;
  int ld;
  __device__ __forceinline__ float* row(int n) const { return n < split ? b0 + n * ld : b1 + (n - split) * ld; }
};
__device__ __forceinline__ RowDst plain_dst(float* base, int ld) { return RowDst{base, base, 1 << 30, ld}; }

// Lanes are 8 (k) x 4 (n): lane = tj*8 + ti; thread owns k = kb + ti + 8a (a < TI) and
// n = nb + tj + 4b (b < TJ).  A rows for the 8 ti-lanes are consecutive, so with the
// row stride M+4 the 128-bit loads along m are bank-conflict free; all lanes of a
// quarter-warp share tj, so the D loads are broadcasts; and a slab access with fixed
// (a, b) touches, per n, 8 consecutive floats = one full 32-byte sector.
template <int TI, int TJ, int WI, int WJ, int M>
__device__ __forceinline__ void wgrad(const float* __restrict__ A, int k_valid, int k_alloc,
                                      const float* __restrict__ D, int n_valid, int n_alloc, const RowDst dst,
                                      bool first, int warp, int lane) {
  constexpr int LD = M + 4;
  if (warp >= WI * WJ) return;
  const int ti = lane & 7, tj = lane >> 3;
  const int kb = (warp % WI) * (8 * TI) + ti;
  const int nb = (warp / WI) * (4 * TJ) + tj;
  if (kb >= k_valid || nb >= n_valid) return;  // k and n only grow from here: nothing to own
  const float* ap[TI];
  const float* dp[TJ];
#pragma unroll
  for (int a = 0; a < TI; ++a) ap[a] = A + min(kb + 8 * a, k_alloc - 1) * LD;
#pragma unroll
  for (int b = 0; b < TJ; ++b) dp[b] = D + min(nb + 4 * b, n_alloc - 1) * LD;
  float acc[TI][TJ];
#pragma unroll
  for (int a = 0; a < TI; ++a)
#pragma unroll
    for (int b = 0; b < TJ; ++b) acc[a][b] = 0.f;
#pragma unroll 2
  for (int m = 0; m < M; m += 4) {
    float4 av[TI], dv[TJ];
#pragma unroll
    for (int a = 0; a < TI; ++a) av[a] = *reinterpret_cast<const float4*>(ap[a] + m);
#pragma unroll
    for (int b = 0; b < TJ; ++b) dv[b] = *reinterpret_cast<const float4*>(dp[b] + m);
#pragma unroll
    for (int a = 0; a < TI; ++a)
#pragma unroll
      for (int b = 0; b < TJ; ++b) {
        float s = acc[a][b];
        s = fmaf(av[a].x, dv[b].x, s);
        s = fmaf(av[a].y, dv[b].y, s);
        s = fmaf(av[a].z, dv[b].z, s);
        s = fmaf(av[a].w, dv[b].w, s);
        acc[a][b] = s;
      }
  }
#pragma unroll
  for (int b = 0; b < TJ; ++b) {
    const int n = nb + 4 * b;
    if (n >= n_valid) continue;
    float* r = dst.row(n);
#pragma unroll
    for (int a = 0; a < TI; ++a) {
      const int k = kb + 8 * a;
      if (k >= k_valid) continue;
      r[k] = first ? acc[a][b] : r[k] + acc[a][b];
    }
  }
}

// Dispatch on the (padded) extents: k_pad / n_pad in {32, 64, 128}; at least one is 128.
template <int M>
__device__ __forceinline__ void wgrad_any(const float* A, int k_valid, int k_alloc, int k_pad, const float* D,
                                          int n_valid, int n_alloc, int n_pad, const RowDst dst, bool first, int warp,
                                          int lane) {
  if (k_pad == 128 && n_pad == 128) wgrad<8, 8, 2, 4, M>(A, k_valid, k_alloc, D, n_valid, n_alloc, dst, first, warp, lane);
  else if (k_pad == 128 && n_pad == 64) wgrad<4, 8, 4, 2, M>(A, k_valid, k_alloc, D, n_valid, n_alloc, dst, first, warp, lane);
  else if (k_pad == 128) wgrad<2, 8, 8, 1, M>(A, k_valid, k_alloc, D, n_valid, n_alloc, dst, first, warp, lane);
  else if (k_pad == 64) wgrad<8, 4, 1, 8, M>(A, k_valid, k_alloc, D, n_valid, n_alloc, dst, first, warp, lane);
  else wgrad<4, 4, 1, 8, M>(A, k_valid, k_alloc, D, n_valid, n_alloc, dst, first, warp, lane);
}

// db[n] (+)= sum_m D[n][m]; one warp per row, lanes stride m.
template <int M>
__device__ __forceinline__ void bias_grad(const float* __restrict__ D, int n_valid, const RowDst dst, bool first,
                                          int warp, int lane) {
  constexpr int LD = M + 4;
  for (int n = warp; n < n_valid; n += CONSUMER_WARPS) {
    float s = 0.f;
#pragma unroll
    for (int m = lane; m < M; m += 32) s += D[n * LD + m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      float* p = dst.row(n);
      *p = first ? s : *p + s;
    }
  }
}

// stash (dense [128][M]) -> shared tile [128][M+4] with cp.async (L2 -> smem, no registers)
template <int M>
__device__ __forceinline__ void load_tile_async(float* tile, const float* __restrict__ src, int tid) {
  constexpr int LD = M + 4;
  constexpr int V = M / 4;
  for (int idx = tid; idx < H * V; idx += CONSUMER_THREADS) {
    const int row = idx / V, c = idx - row * V;
    cp_async16(tile + row * LD + 4 * c, src + row * M + 4 * c);
  }
  cp_async_commit();
}

// ---------------------------------------------------------------------------------------
// streamed operands, in the order producer and consumers walk them
// ---------------------------------------------------------------------------------------
enum OpId {
  OP_F_COND0 = 0, OP_F_COND1, OP_F_ENC0, OP_F_ENC1, OP_F_ENC2, OP_F_ENC3, OP_F_HEADS_E, OP_F_HEADS_C, OP_F_DEC0_Z,
  OP_F_DEC0_C, OP_F_DEC1, OP_F_DEC2, OP_F_DEC3, N_FWD_OPS,
  OP_B_DEC3 = N_FWD_OPS, OP_B_DEC2, OP_B_DEC1, OP_B_DEC0_Z, OP_B_DEC0_C, OP_B_HEADS_E, OP_B_HEADS_C, OP_B_COND1,
  OP_B_ENC3, OP_B_ENC2, OP_B_ENC1, N_OPS
};
struct Op { int off, rows, width; };
__device__ __forceinline__ Op op_desc(const Layout& lo, int id) {
  switch (id) {
    case OP_F_COND0: return {lo.q_w[L_COND0], 2, H};
    case OP_F_COND1: return {lo.q_w[L_COND1], H, H};
    case OP_F_ENC0: return {lo.q_w[L_ENC0], lo.I, H};
    case OP_F_ENC1: return {lo.q_w[L_ENC1], H, H};
    case OP_F_ENC2: return {lo.q_w[L_ENC2], H, H};
    case OP_F_ENC3: return {lo.q_w[L_ENC3], H, H};
    case OP_F_HEADS_E: return {lo.q_w[L_HEADS], H, lo.L2p};
    case OP_F_HEADS_C: return {lo.q_w[L_HEADS] + H * lo.L2p, H, lo.L2p};
    case OP_F_DEC0_Z: return {lo.q_w[L_DEC0], lo.L, H};
    case OP_F_DEC0_C: return {lo.q_w[L_DEC0] + lo.L * H, H, H};
    case OP_F_DEC1: return {lo.q_w[L_DEC1], H, H};
    case OP_F_DEC2: return {lo.q_w[L_DEC2], H, H};
    case OP_F_DEC3: return {lo.q_w[L_DEC3], H, lo.Ip};
    case OP_B_DEC3: return {lo.r_w[L_DEC3], lo.I, H};
    case OP_B_DEC2: return {lo.r_w[L_DEC2], H, H};
    case OP_B_DEC1: return {lo.r_w[L_DEC1], H, H};
    case OP_B_DEC0_Z: return {lo.r_dec0z, H, lo.Lzp};
    case OP_B_DEC0_C: return {lo.r_w[L_DEC0], H, H};
    case OP_B_HEADS_E: return {lo.r_w[L_HEADS], 2 * lo.L, H};
    case OP_B_HEADS_C: return {lo.r_heads_c, 2 * lo.L, H};
    case OP_B_COND1: return {lo.r_w[L_COND1], H, H};
    case OP_B_ENC3: return {lo.r_w[L_ENC3], H, H};
    case OP_B_ENC2: return {lo.r_w[L_ENC2], H, H};
    default: return {lo.r_w[L_ENC1], H, H};
  }
}

enum TrainMode { TM_FUSED = 0, TM_FWD = 1, TM_BWD = 2 };

template <int M>
struct Smem {
  static constexpr int LD = M + 4;
  float *ring, *R0, *R1, *R2, *XT, *GT, *ML, *ZT, *EP, *ST, *red;
  uint64_t *full, *empty;
  __device__ Smem(unsigned char* raw, const Layout& lo, int stages) {
    const SmallRows sr = small_rows(lo);
    ring = reinterpret_cast<float*>(raw);
    R0 = ring + stages * STAGE_FLOATS;
    R1 = R0 + H * LD;
    R2 = R1 + H * LD;
    float* small = R2 + H * LD;
    XT = small + sr.xt * LD; GT = small + sr.gt * LD; ML = small + sr.ml * LD;
    ZT = small + sr.zt * LD; EP = small + sr.ep * LD; ST = small + sr.st * LD;
    red = small + sr.total * LD;
    full = reinterpret_cast<uint64_t*>(red + 64);
    empty = full + 8;
  }
};

// forward GEMM over one or two streamed operands with the configuration for width NP
#define DMVAE_CONSUME(CFG, acc, P, opid)                                                                   \
  {                                                                                                        \
    const Op _op = op_desc(lo, opid);                                                                      \
    consume<CFG>(acc, P, LD, _op.rows, _op.width, s.ring, s.full, s.empty, rs, warp, lane);                \
  }

template <int M, int NP>
__device__ __forceinline__ void fwd_small_layer(const Layout& lo, const Smem<M>& s, RingStateRt& rs, const float* P0,
                                                int op0, const float* P1, int op1, float* tile, const float* bias,
                                                int n_rows, int warp, int lane) {
  constexpr int LD = M + 4;
  using C = FwdCfg<M, NP>;
  float acc[C::TI][C::TJ];
  zero_acc<C>(acc);
  DMVAE_CONSUME(C, acc, P0, op0);
  if (P1 != nullptr) DMVAE_CONSUME(C, acc, P1, op1);
  store_fwd<C, M, false>(acc, tile, bias, nullptr, n_rows, warp, lane);
}

template <int M, int NP>
__device__ __forceinline__ void dgrad_small(const Layout& lo, const Smem<M>& s, RingStateRt& rs, const float* P, int op,
                                            float* tile, int k_rows, int warp, int lane) {
  constexpr int LD = M + 4;
  using C = FwdCfg<M, NP>;
  float acc[C::TI][C::TJ];
  zero_acc<C>(acc);
  DMVAE_CONSUME(C, acc, P, op);
  consumer_sync();  // every reader of `tile` (the preceding weight-gradient GEMM) is done
  store_dgrad<C, M, EPI_STORE>(acc, tile, k_rows, warp, lane);
}

template <int M>
__global__ void __launch_bounds__(BLOCK_THREADS, 1) train_kernel(const __grid_constant__ TrainArgs a) {
  constexpr int LD = M + 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Layout& lo = a.lo;
  const Smem<M> s(smem_raw, lo, a.stages);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = lo.L, I = lo.I, T = lo.T;
  const float* __restrict__ pk = a.packed;
  const long long n_tiles = (a.B + M - 1) / M;
  const bool do_fwd = a.mode != TM_BWD, do_bwd = a.mode != TM_FWD;

  if (tid == 0) {
    for (int st = 0; st < a.stages; ++st) {
      mbar_init(&s.full[st], 1);
      mbar_init(&s.empty[st], CONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == CONSUMER_WARPS) {
    // ===================== producer warp ==================================================
    if (lane == 0) {
      RingStateRt rs(a.stages);
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // long trajectories: enc0 / dec3 are walked in chunks of 128 features (consumer side: same order)
        for (int id = do_fwd ? 0 : N_FWD_OPS; id < (do_bwd ? N_OPS : N_FWD_OPS); ++id) {
          const Op op = op_desc(lo, id);
          if (lo.NC > 1 && (id == OP_F_ENC0 || id == OP_F_DEC3 || id == OP_B_DEC3)) {
            for (int c = 0; c < lo.NC; ++c) {
              const int kc = min(128, lo.I - c * 128);
              if (id == OP_F_DEC3) produce(pk + op.off + (size_t)c * H * 128, H, 128, s.ring, s.full, s.empty, rs);
              else produce(pk + op.off + (size_t)c * 128 * H, kc, H, s.ring, s.full, s.empty, rs);
            }
          } else {
            produce(pk + op.off, op.rows, op.width, s.ring, s.full, s.empty, rs);
          }
        }
      }
    }
    return;
  }

  // ========================= consumer warps =================================================
  using CM = FwdCfg<M, 128>;
  RingStateRt rs(a.stages);
  const SmallRows sr = small_rows(lo);
  float* slab = a.slabs != nullptr ? a.slabs + (size_t)blockIdx.x * a.slab_stride : nullptr;
  float loss_acc[4] = {0.f, 0.f, 0.f, 0.f};  // thread 0: recon, kld, start, time (already scaled)
  bool first = true;

  // zero the small tiles once: padding rows must read as zeros
  for (int idx = tid; idx < sr.total * LD; idx += CONSUMER_THREADS) s.XT[idx] = 0.f;
  consumer_sync();

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long m0 = tile * M;
    const int valid = (int)min((long long)M, a.B - m0);
    float* stash = a.stash + (size_t)(a.mode == TM_FUSED ? (long long)blockIdx.x : tile) * a.stash_stride;
    float* stash_small = stash + (size_t)N_STASH * H * M;
    const bool chunked = lo.NC > 1;
    float* XR = stash_small + (size_t)sr.total * LD;      // long trajectories: x_rel [Ipt][M]
    float* RG = XR + long_floats(lo, M);                   //                    recon, then its gradient [Ipt][M]
    // chunk c of a feature-major stash area -> shared tile [128][LD] (rows past the chunk are zero)
    auto load_chunk = [&](float* tile, const float* area, int c) {
      const int kc = min(128, I - c * 128);
      for (int idx = tid; idx < 128 * M; idx += CONSUMER_THREADS) {
        const int n = idx / M, m = idx - n * M;
        tile[n * LD + m] = n < kc ? area[(size_t)(c * 128 + n) * M + m] : 0.f;
      }
    };

    if (do_fwd) {
      // ---- stage x (relative), start, eps ------------------------------------------------
      for (int idx = tid; idx < M * I; idx += CONSUMER_THREADS) {
        const int m = idx / I, n = idx - m * I;
        float v = 0.f;
        if (m < valid) {
          const float* row = a.x + (m0 + m) * I;
          v = __ldg(row + n);
          if (a.mode == TM_FUSED) {  // Training_VAE.py:345-348: x/y columns minus the start point
            const int d = n % 3;
            if (d == 1) v = v - __ldg(row + 1);
            else if (d == 2) v = v - __ldg(row + 2);
          }
        }
        if (chunked) XR[(size_t)n * M + m] = v;
        else s.XT[n * LD + m] = v;
      }
      for (int idx = tid; idx < 2 * M; idx += CONSUMER_THREADS) {
        const int d = idx / M, m = idx - d * M;
        float v = 0.f;
        if (m < valid) v = a.mode == TM_FUSED ? __ldg(a.x + (m0 + m) * I + 1 + d) : __ldg(a.start + (m0 + m) * 2 + d);
        s.ST[d * LD + m] = v;
      }
      if (a.eps != nullptr) {
        for (int idx = tid; idx < M * L; idx += CONSUMER_THREADS) {
          const int m = idx / L, j = idx - m * L;
          s.EP[j * LD + m] = m < valid ? __ldg(a.eps + (m0 + m) * L + j) : 0.f;
        }
      } else {
        const int nb = lo.Lq >> 2;
        for (int idx = tid; idx < M * nb; idx += CONSUMER_THREADS) {
          const int jb = idx / M, m = idx - jb * M;
          const float4 g = philox_normal4(a.seed, a.sample_offset + (unsigned long long)(m0 + m), (uint32_t)jb,
                                          (uint32_t)(a.step + 1));
          const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = jb * 4 + i;
            if (j < lo.Lq) s.EP[j * LD + m] = (j < L && m < valid) ? gv[i] : 0.f;
          }
        }
      }
      consumer_sync();

      float acc[CM::TI][CM::TJ];
      // cond0: start -> R0 (hc1)
      zero_acc<CM>(acc);
      DMVAE_CONSUME(CM, acc, s.ST, OP_F_COND0);
      store_fwd<CM, M, true>(acc, s.R0, pk + lo.q_b[L_COND0], stash + S_HC1 * H * M, H, warp, lane);
      consumer_sync();
      // cond1: R0 -> R2 (hc, stays for heads and dec0)
      zero_acc<CM>(acc);
      DMVAE_CONSUME(CM, acc, s.R0, OP_F_COND1);
      store_fwd<CM, M, true>(acc, s.R2, pk + lo.q_b[L_COND1], stash + S_HC * H * M, H, warp, lane);
      consumer_sync();
      // enc0: XT -> R0 (e1)
      zero_acc<CM>(acc);
      if (!chunked) {
        DMVAE_CONSUME(CM, acc, s.XT, OP_F_ENC0);
      } else {
        for (int c = 0; c < lo.NC; ++c) {
          load_chunk(s.XT, XR, c);
          consumer_sync();
          consume<CM>(acc, s.XT, LD, min(128, I - c * 128), H, s.ring, s.full, s.empty, rs, warp, lane);
          consumer_sync();  // XT is restaged by the next chunk
        }
      }
      store_fwd<CM, M, true>(acc, s.R0, pk + lo.q_b[L_ENC0], stash + S_E1 * H * M, H, warp, lane);
      consumer_sync();
      // enc1: R0 -> R1 (e2)
      zero_acc<CM>(acc);
      DMVAE_CONSUME(CM, acc, s.R0, OP_F_ENC1);
      store_fwd<CM, M, true>(acc, s.R1, pk + lo.q_b[L_ENC1], stash + S_E2 * H * M, H, warp, lane);
      consumer_sync();
      // enc2: R1 -> R0 (e3)
      zero_acc<CM>(acc);
      DMVAE_CONSUME(CM, acc, s.R1, OP_F_ENC2);
      store_fwd<CM, M, true>(acc, s.R0, pk + lo.q_b[L_ENC2], stash + S_E3 * H * M, H, warp, lane);
      consumer_sync();
      // enc3: R0 -> R1 (e4)
      zero_acc<CM>(acc);
      DMVAE_CONSUME(CM, acc, s.R0, OP_F_ENC3);
      store_fwd<CM, M, true>(acc, s.R1, pk + lo.q_b[L_ENC3], stash + S_E4 * H * M, H, warp, lane);
      consumer_sync();
      // heads: [e4 ; hc] -> ML (mu rows, logvar rows)
      if (lo.L2p == 32) fwd_small_layer<M, 32>(lo, s, rs, s.R1, OP_F_HEADS_E, s.R2, OP_F_HEADS_C, s.ML, pk + lo.q_b[L_HEADS], 2 * L, warp, lane);
      else if (lo.L2p == 64) fwd_small_layer<M, 64>(lo, s, rs, s.R1, OP_F_HEADS_E, s.R2, OP_F_HEADS_C, s.ML, pk + lo.q_b[L_HEADS], 2 * L, warp, lane);
      else fwd_small_layer<M, 128>(lo, s, rs, s.R1, OP_F_HEADS_E, s.R2, OP_F_HEADS_C, s.ML, pk + lo.q_b[L_HEADS], 2 * L, warp, lane);
      consumer_sync();
      // reparameterise (Training_VAE.py:199-206): z = mu + eps * exp(0.5 * logvar)
      for (int idx = tid; idx < M * L; idx += CONSUMER_THREADS) {
        const int j = idx / M, m = idx - j * M;
        const float mu = s.ML[j * LD + m], lv = s.ML[(L + j) * LD + m];
        s.ZT[j * LD + m] = mu + s.EP[j * LD + m] * expf(0.5f * lv);
      }
      consumer_sync();
      // dec0: [z ; hc] -> R0 (d1)
      zero_acc<CM>(acc);
      DMVAE_CONSUME(CM, acc, s.ZT, OP_F_DEC0_Z);
      DMVAE_CONSUME(CM, acc, s.R2, OP_F_DEC0_C);
      store_fwd<CM, M, true>(acc, s.R0, pk + lo.q_b[L_DEC0], stash + S_D1 * H * M, H, warp, lane);
      consumer_sync();
      // dec1: R0 -> R1 (d2)
      zero_acc<CM>(acc);
      DMVAE_CONSUME(CM, acc, s.R0, OP_F_DEC1);
      store_fwd<CM, M, true>(acc, s.R1, pk + lo.q_b[L_DEC1], stash + S_D2 * H * M, H, warp, lane);
      consumer_sync();
      // dec2: R1 -> R0 (d3)
      zero_acc<CM>(acc);
      DMVAE_CONSUME(CM, acc, s.R1, OP_F_DEC2);
      store_fwd<CM, M, true>(acc, s.R0, pk + lo.q_b[L_DEC2], stash + S_D3 * H * M, H, warp, lane);
      consumer_sync();
      // dec3: R0 -> GT (recon), no activation; padding rows come out as exact zeros
      if (chunked) {
        for (int c = 0; c < lo.NC; ++c) {
          const int kc = min(128, I - c * 128);
          float o3[CM::TI][CM::TJ];
          zero_acc<CM>(o3);
          consume<CM>(o3, s.R0, LD, H, 128, s.ring, s.full, s.empty, rs, warp, lane);
          store_fwd<CM, M, false>(o3, s.GT, pk + lo.q_b[L_DEC3] + c * 128, nullptr, 128, warp, lane);
          consumer_sync();
          for (int idx = tid; idx < kc * M; idx += CONSUMER_THREADS) {
            const int n = idx / M, m = idx - n * M;
            RG[(size_t)(c * 128 + n) * M + m] = s.GT[n * LD + m];
          }
          consumer_sync();  // GT is rewritten by the next chunk
        }
      } else
      if (lo.Ip == 32) fwd_small_layer<M, 32>(lo, s, rs, s.R0, OP_F_DEC3, nullptr, 0, s.GT, pk + lo.q_b[L_DEC3], lo.Ip, warp, lane);
      else if (lo.Ip == 64) fwd_small_layer<M, 64>(lo, s, rs, s.R0, OP_F_DEC3, nullptr, 0, s.GT, pk + lo.q_b[L_DEC3], lo.Ip, warp, lane);
      else fwd_small_layer<M, 128>(lo, s, rs, s.R0, OP_F_DEC3, nullptr, 0, s.GT, pk + lo.q_b[L_DEC3], lo.Ip, warp, lane);
      consumer_sync();

      if (a.mode == TM_FWD) {
        // outputs of model.forward (Training_VAE.py:226), row-major, + the small tiles for backward
        for (int idx = tid; idx < M * I; idx += CONSUMER_THREADS) {
          const int m = idx / I, n = idx - m * I;
          if (m < valid) a.recon[(m0 + m) * I + n] = chunked ? RG[(size_t)n * M + m] : s.GT[n * LD + m];
        }
        for (int idx = tid; idx < M * L; idx += CONSUMER_THREADS) {
          const int m = idx / L, j = idx - m * L;
          if (m < valid) {
            a.mu[(m0 + m) * L + j] = s.ML[j * LD + m];
            a.logvar[(m0 + m) * L + j] = s.ML[(L + j) * LD + m];
          }
        }
        for (int idx = tid; idx < M * H; idx += CONSUMER_THREADS) {
          const int m = idx / H, k = idx - m * H;
          if (m < valid) a.hc[(m0 + m) * H + k] = s.R2[k * LD + m];
        }
        for (int idx = tid; idx < sr.total * LD; idx += CONSUMER_THREADS) stash_small[idx] = s.XT[idx];
        consumer_sync();
        continue;
      }

      // ---- loss (Training_VAE.py:229-268) and d(total)/d(recon), one thread per row -----------
      {
        float p_rec = 0.f, p_kld = 0.f, p_start = 0.f, p_time = 0.f;
        // recon (then its gradient) and x_rel, feature-major: shared tiles, or the stash areas for long trajectories
        float* Gp = chunked ? RG : s.GT;
        const float* Xp = chunked ? XR : s.XT;
        const int gld = chunked ? M : LD;
        if (tid < M && tid < valid) {
          const int m = tid;
          const float c_rec = a.w_recon * 2.f * a.inv_batch / (float)I;
          const float c_start = a.w_start * a.inv_batch;           // 2 * w / (2B)
          const float c_t0 = a.w_time * 2.f * a.inv_batch;
          const float c_mono = T > 1 ? a.w_time * a.inv_batch / (float)(T - 1) : 0.f;
          float s_rec = 0.f, s_start = 0.f, s_t0 = 0.f, s_mono = 0.f;
          float g_prev = 0.f, r_prev = 0.f;
          for (int t = 0; t < T; ++t) {
            {
              const float r = Gp[(size_t)(3 * t) * gld + m], x = Xp[(size_t)(3 * t) * gld + m];
              const float diff = r - x;
              s_rec = fmaf(diff, diff, s_rec);
              float g = c_rec * diff;
              if (t == 0) {
                s_t0 = r * r;
                g = fmaf(c_t0, r, g);
              } else {
                const float dt = r - r_prev;
                if (dt < 0.f) {  // relu'(0) = 0: strict
                  s_mono -= dt;
                  g -= c_mono;
                  g_prev += c_mono;
                }
                Gp[(size_t)(3 * (t - 1)) * gld + m] = g_prev;
              }
              g_prev = g;
              r_prev = r;
            }
#pragma unroll
            for (int d = 1; d < 3; ++d) {
              const int n = 3 * t + d;
              const float r = Gp[(size_t)n * gld + m], x = Xp[(size_t)n * gld + m];
              const float diff = r - x;
              s_rec = fmaf(diff, diff, s_rec);
              float g = c_rec * diff;
              if (t == 0) {
                s_start = fmaf(diff, diff, s_start);
                g = fmaf(c_start, diff, g);
              }
              Gp[(size_t)n * gld + m] = g;
            }
          }
          Gp[(size_t)(3 * (T - 1)) * gld + m] = g_prev;
          float s_k = 0.f;
          for (int j = 0; j < L; ++j) {
            const float mu = s.ML[j * LD + m], lv = s.ML[(L + j) * LD + m];
            s_k += 1.f + lv - mu * mu - expf(lv);
          }
          p_rec = s_rec * (a.inv_batch / (float)I);
          p_kld = -0.5f * s_k * (a.inv_batch / (float)L);
          p_start = s_start * (a.inv_batch * 0.5f);
          p_time = s_t0 * a.inv_batch + (T > 1 ? s_mono * (a.inv_batch / (float)(T - 1)) : 0.f);
        } else if (tid < M) {
          for (int n = 0; n < I; ++n) Gp[(size_t)n * gld + tid] = 0.f;  // rows past the batch end carry no gradient
        }
        if (warp < (M + 31) / 32) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            p_rec += __shfl_xor_sync(0xffffffffu, p_rec, o);
            p_kld += __shfl_xor_sync(0xffffffffu, p_kld, o);
            p_start += __shfl_xor_sync(0xffffffffu, p_start, o);
            p_time += __shfl_xor_sync(0xffffffffu, p_time, o);
          }
          if (lane == 0) {
            s.red[warp * 4 + 0] = p_rec; s.red[warp * 4 + 1] = p_kld;
            s.red[warp * 4 + 2] = p_start; s.red[warp * 4 + 3] = p_time;
          }
        }
        consumer_sync();
        if (tid == 0)
          for (int w = 0; w < (M + 31) / 32; ++w)
#pragma unroll
            for (int q = 0; q < 4; ++q) loss_acc[q] += s.red[w * 4 + q];
      }
    }  // do_fwd

    // ======================================= backward ==========================================
    if (a.mode == TM_BWD) {
      // small tiles back from the stash; upstream d/d(recon) transposed into GT; d3, d2 into R0, R1
      for (int idx = tid; idx < sr.total * LD; idx += CONSUMER_THREADS) s.XT[idx] = stash_small[idx];
      consumer_sync();
      for (int idx = tid; idx < M * I; idx += CONSUMER_THREADS) {
        const int m = idx / I, n = idx - m * I;
        const float g = (a.g_recon != nullptr && m < valid) ? __ldg(a.g_recon + (m0 + m) * I + n) : 0.f;
        if (chunked) RG[(size_t)n * M + m] = g;
        else s.GT[n * LD + m] = g;
      }
      if (!chunked)
        for (int idx = tid; idx < (lo.Ip - I) * LD; idx += CONSUMER_THREADS) s.GT[I * LD + idx] = 0.f;
      load_tile_async<M>(s.R0, stash + S_D3 * H * M, tid);
      load_tile_async<M>(s.R1, stash + S_D2 * H * M, tid);
      cp_async_wait<0>();
      consumer_sync();
    }

    float acc[CM::TI][CM::TJ];
    // ---- dec3: delta = GT (I rows), input d3 = R0 ------------------------------------------------
    load_tile_async<M>(s.R2, stash + S_D1 * H * M, tid);  // prefetch d1 (R2 = hc is reloaded later)
    zero_acc<CM>(acc);
    if (!chunked) {
      wgrad_any<M>(s.R0, H, H, 128, s.GT, I, lo.Ip, lo.Ip, plain_dst(slab + lo.p_w[L_DEC3], H), first, warp, lane);
      bias_grad<M>(s.GT, I, plain_dst(slab + lo.p_b[L_DEC3], 1), first, warp, lane);
      DMVAE_CONSUME(CM, acc, s.GT, OP_B_DEC3);
    } else {
      consumer_sync();  // the loss pass (or the upstream-gradient copy) has finished writing the stash area
      for (int c = 0; c < lo.NC; ++c) {
        const int kc = min(128, I - c * 128);
        load_chunk(s.GT, RG, c);
        consumer_sync();
        wgrad_any<M>(s.R0, H, H, 128, s.GT, kc, 128, 128, plain_dst(slab + lo.p_w[L_DEC3] + (size_t)c * 128 * H, H), first,
                     warp, lane);
        bias_grad<M>(s.GT, kc, plain_dst(slab + lo.p_b[L_DEC3] + c * 128, 1), first, warp, lane);
        consume<CM>(acc, s.GT, LD, kc, H, s.ring, s.full, s.empty, rs, warp, lane);
        consumer_sync();  // GT is reloaded by the next chunk
      }
    }
    consumer_sync();
    store_dgrad<CM, M, EPI_MASK>(acc, s.R0, H, warp, lane);
    consumer_sync();
    // ---- dec2: delta = R0, input d2 = R1 -----------------------------------------------------------
    wgrad_any<M>(s.R1, H, H, 128, s.R0, H, H, 128, plain_dst(slab + lo.p_w[L_DEC2], H), first, warp, lane);
    bias_grad<M>(s.R0, H, plain_dst(slab + lo.p_b[L_DEC2], 1), first, warp, lane);
    zero_acc<CM>(acc);
    DMVAE_CONSUME(CM, acc, s.R0, OP_B_DEC2);
    consumer_sync();
    store_dgrad<CM, M, EPI_MASK>(acc, s.R1, H, warp, lane);
    cp_async_wait<0>();  // d1 has landed in R2
    consumer_sync();
    // ---- dec1: delta = R1, input d1 = R2 -----------------------------------------------------------
    load_tile_async<M>(s.R0, stash + S_HC * H * M, tid);  // prefetch hc
    wgrad_any<M>(s.R2, H, H, 128, s.R1, H, H, 128, plain_dst(slab + lo.p_w[L_DEC1], H), first, warp, lane);
    bias_grad<M>(s.R1, H, plain_dst(slab + lo.p_b[L_DEC1], 1), first, warp, lane);
    zero_acc<CM>(acc);
    DMVAE_CONSUME(CM, acc, s.R1, OP_B_DEC1);
    consumer_sync();
    store_dgrad<CM, M, EPI_MASK>(acc, s.R2, H, warp, lane);
    cp_async_wait<0>();  // hc has landed in R0
    consumer_sync();
    // ---- dec0: delta = R2, input [z (ZT) ; hc (R0)] ------------------------------------------------
    {
      const int Kd = L + H;
      wgrad_any<M>(s.ZT, L, lo.Lq, pad_width(L), s.R2, H, H, 128, plain_dst(slab + lo.p_w[L_DEC0], Kd), first, warp, lane);
      wgrad_any<M>(s.R0, H, H, 128, s.R2, H, H, 128, plain_dst(slab + lo.p_w[L_DEC0] + L, Kd), first, warp, lane);
      bias_grad<M>(s.R2, H, plain_dst(slab + lo.p_b[L_DEC0], 1), first, warp, lane);
      // d/dz -> ZT (over z), d/dhc (decoder share) -> R1, no relu' yet
      if (lo.Lzp == 32) dgrad_small<M, 32>(lo, s, rs, s.R2, OP_B_DEC0_Z, s.ZT, L, warp, lane);
      else dgrad_small<M, 64>(lo, s, rs, s.R2, OP_B_DEC0_Z, s.ZT, L, warp, lane);
      zero_acc<CM>(acc);
      DMVAE_CONSUME(CM, acc, s.R2, OP_B_DEC0_C);
      store_dgrad<CM, M, EPI_STORE>(acc, s.R1, H, warp, lane);  // R1 (delta of dec1) is dead since the barrier above
      consumer_sync();
    }
    // ---- reparameterisation + KLD backward -> ML rows become d/dmu, d/dlogvar -------------------
    load_tile_async<M>(s.R2, stash + S_E4 * H * M, tid);  // e4 (exposed: all three tiles were busy)
    {
      const float c_k = a.w_kld * a.inv_batch / (float)L;
      for (int idx = tid; idx < M * L; idx += CONSUMER_THREADS) {
        const int j = idx / M, m = idx - j * M;
        float gmu = 0.f, glv = 0.f;
        if (m < valid) {
          const float mu = s.ML[j * LD + m], lv = s.ML[(L + j) * LD + m];
          const float gz = s.ZT[j * LD + m], ep = s.EP[j * LD + m];
          gmu = fmaf(c_k, mu, gz);
          glv = -0.5f * c_k * (1.f - expf(lv)) + 0.5f * gz * ep * expf(0.5f * lv);
          if (a.g_mu != nullptr) gmu += __ldg(a.g_mu + (m0 + m) * L + j);
          if (a.g_logvar != nullptr) glv += __ldg(a.g_logvar + (m0 + m) * L + j);
        }
        s.ML[j * LD + m] = gmu;
        s.ML[(L + j) * LD + m] = glv;
      }
      if (a.g_hc != nullptr) {
        for (int idx = tid; idx < M * H; idx += CONSUMER_THREADS) {
          const int m = idx / H, k = idx - m * H;
          if (m < valid) s.R1[k * LD + m] += __ldg(a.g_hc + (m0 + m) * H + k);
        }
      }
    }
    cp_async_wait<0>();
    consumer_sync();
    // ---- heads: delta = ML (2L rows), input [e4 (R2) ; hc (R0)] ----------------------------------
    {
      const RowDst de{slab + lo.p_w[L_HEADS], slab + lo.p_wlv, L, 2 * H};
      const RowDst dc{slab + lo.p_w[L_HEADS] + H, slab + lo.p_wlv + H, L, 2 * H};
      const RowDst db{slab + lo.p_b[L_HEADS], slab + lo.p_blv, L, 1};
      const int ml_alloc = round_up(2 * L, 4);
      wgrad_any<M>(s.R2, H, H, 128, s.ML, 2 * L, ml_alloc, lo.L2p, de, first, warp, lane);
      wgrad_any<M>(s.R0, H, H, 128, s.ML, 2 * L, ml_alloc, lo.L2p, dc, first, warp, lane);
      bias_grad<M>(s.ML, 2 * L, db, first, warp, lane);
      zero_acc<CM>(acc);
      DMVAE_CONSUME(CM, acc, s.ML, OP_B_HEADS_E);
      consumer_sync();
      store_dgrad<CM, M, EPI_MASK>(acc, s.R2, H, warp, lane);  // delta of enc3 over e4
      zero_acc<CM>(acc);
      DMVAE_CONSUME(CM, acc, s.ML, OP_B_HEADS_C);
      store_dgrad<CM, M, EPI_ACCUM>(acc, s.R1, H, warp, lane);  // encoder share of d/dhc (same thread owns the element)
      consumer_sync();
    }
    // ---- condition branch: delta_hc = R1 * relu'(hc = R0) ---------------------------------------
    for (int idx = tid; idx < H * (M / 4); idx += CONSUMER_THREADS) {
      const int k = idx / (M / 4), c = idx - k * (M / 4);
      float4* g = reinterpret_cast<float4*>(s.R1 + k * LD + 4 * c);
      const float4 h = *reinterpret_cast<const float4*>(s.R0 + k * LD + 4 * c);
      float4 v = *g;
      v.x = h.x > 0.f ? v.x : 0.f; v.y = h.y > 0.f ? v.y : 0.f;
      v.z = h.z > 0.f ? v.z : 0.f; v.w = h.w > 0.f ? v.w : 0.f;
      *g = v;
    }
    consumer_sync();
    load_tile_async<M>(s.R0, stash + S_HC1 * H * M, tid);  // hc1 (exposed)
    cp_async_wait<0>();
    consumer_sync();
    // cond1: delta = R1, input hc1 = R0
    wgrad_any<M>(s.R0, H, H, 128, s.R1, H, H, 128, plain_dst(slab + lo.p_w[L_COND1], H), first, warp, lane);
    bias_grad<M>(s.R1, H, plain_dst(slab + lo.p_b[L_COND1], 1), first, warp, lane);
    zero_acc<CM>(acc);
    DMVAE_CONSUME(CM, acc, s.R1, OP_B_COND1);
    consumer_sync();
    store_dgrad<CM, M, EPI_MASK>(acc, s.R0, H, warp, lane);
    consumer_sync();
    load_tile_async<M>(s.R1, stash + S_E3 * H * M, tid);  // prefetch e3 under cond0
    // cond0: delta = R0 (128 rows), input start (2 rows): one output per thread
    {
      const int n = tid >> 1, d = tid & 1;
      float sacc = 0.f;
      for (int m = 0; m < M; ++m) sacc = fmaf(s.R0[n * LD + m], s.ST[d * LD + m], sacc);
      float* p = slab + lo.p_w[L_COND0] + n * 2 + d;
      *p = first ? sacc : *p + sacc;
      bias_grad<M>(s.R0, H, plain_dst(slab + lo.p_b[L_COND0], 1), first, warp, lane);
    }
    cp_async_wait<0>();
    consumer_sync();
    // ---- enc3: delta = R2, input e3 = R1 ---------------------------------------------------------
    load_tile_async<M>(s.R0, stash + S_E2 * H * M, tid);  // prefetch e2
    wgrad_any<M>(s.R1, H, H, 128, s.R2, H, H, 128, plain_dst(slab + lo.p_w[L_ENC3], H), first, warp, lane);
    bias_grad<M>(s.R2, H, plain_dst(slab + lo.p_b[L_ENC3], 1), first, warp, lane);
    zero_acc<CM>(acc);
    DMVAE_CONSUME(CM, acc, s.R2, OP_B_ENC3);
    consumer_sync();
    store_dgrad<CM, M, EPI_MASK>(acc, s.R1, H, warp, lane);
    cp_async_wait<0>();
    consumer_sync();
    // ---- enc2: delta = R1, input e2 = R0 ---------------------------------------------------------
    load_tile_async<M>(s.R2, stash + S_E1 * H * M, tid);  // prefetch e1
    wgrad_any<M>(s.R0, H, H, 128, s.R1, H, H, 128, plain_dst(slab + lo.p_w[L_ENC2], H), first, warp, lane);
    bias_grad<M>(s.R1, H, plain_dst(slab + lo.p_b[L_ENC2], 1), first, warp, lane);
    zero_acc<CM>(acc);
    DMVAE_CONSUME(CM, acc, s.R1, OP_B_ENC2);
    consumer_sync();
    store_dgrad<CM, M, EPI_MASK>(acc, s.R0, H, warp, lane);
    cp_async_wait<0>();
    consumer_sync();
    // ---- enc1: delta = R0, input e1 = R2 ---------------------------------------------------------
    wgrad_any<M>(s.R2, H, H, 128, s.R0, H, H, 128, plain_dst(slab + lo.p_w[L_ENC1], H), first, warp, lane);
    bias_grad<M>(s.R0, H, plain_dst(slab + lo.p_b[L_ENC1], 1), first, warp, lane);
    zero_acc<CM>(acc);
    DMVAE_CONSUME(CM, acc, s.R0, OP_B_ENC1);
    consumer_sync();
    store_dgrad<CM, M, EPI_MASK>(acc, s.R2, H, warp, lane);
    consumer_sync();
    // ---- enc0: delta = R2, input x_rel = XT (no data gradient) ------------------------------------
    if (!chunked) {
      wgrad_any<M>(s.XT, I, lo.Ip, lo.Ip, s.R2, H, H, 128, plain_dst(slab + lo.p_w[L_ENC0], I), first, warp, lane);
    } else {
      for (int c = 0; c < lo.NC; ++c) {
        load_chunk(s.XT, XR, c);
        consumer_sync();
        wgrad_any<M>(s.XT, min(128, I - c * 128), 128, 128, s.R2, H, H, 128, plain_dst(slab + lo.p_w[L_ENC0] + c * 128, I),
                     first, warp, lane);
        consumer_sync();  // XT is reloaded by the next chunk
      }
    }
    bias_grad<M>(s.R2, H, plain_dst(slab + lo.p_b[L_ENC0], 1), first, warp, lane);
    consumer_sync();  // the next tile restages XT / ST / EP and rewrites R0..R2
    first = false;
  }

  if (do_bwd && tid == 0) {
    float* tail = slab + lo.n_params;
    tail[0] = loss_acc[0]; tail[1] = loss_acc[1]; tail[2] = loss_acc[2]; tail[3] = loss_acc[3];
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
namespace {
constexpr size_t SMEM_LIMIT = 232448;

// Tile height and ring depth for a batch: 64 rows when it fits with >= 3 stages and the
// batch gives every SM a tile, else 32 rows.
void pick_tiling(const Layout& lo, long long B, int sm_count, int* M, int* stages) {
  int st64 = 0;
  for (int st = 4; st >= 2; --st)
    if (train_smem_bytes<64>(lo, st) <= SMEM_LIMIT) { st64 = st; break; }
  const bool fills = (B + 63) / 64 >= sm_count;
  if (st64 >= 3 && fills) { *M = 64; *stages = st64; return; }
  *M = 32;
  *stages = 2;
  for (int st = 4; st >= 2; --st)
    if (train_smem_bytes<32>(lo, st) <= SMEM_LIMIT) { *stages = st; break; }
}
}  // namespace

TrainPlan plan_train(const Layout& lo, long long B, int sm_count, bool per_tile_stash) {
  TrainPlan p;
  pick_tiling(lo, B, sm_count, &p.M, &p.stages);
  p.n_tiles = (B + p.M - 1) / p.M;
  p.grid = (int)(p.n_tiles < sm_count ? p.n_tiles : sm_count);
  if (p.grid < 1) p.grid = 1;
  const SmallRows sr = small_rows(lo);
  p.stash_stride = round_up(N_STASH * H * p.M + sr.total * (p.M + 4) + 2 * (int)long_floats(lo, p.M), 4);
  p.stash_units = per_tile_stash ? p.n_tiles : p.grid;
  p.slab_stride = round_up(lo.n_params + 4, 4);
  p.smem = p.M == 64 ? train_smem_bytes<64>(lo, p.stages) : train_smem_bytes<32>(lo, p.stages);
  return p;
}

cudaError_t launch_train(const Layout& lo, const TrainPlan& plan, int mode, const TrainIO& io, cudaStream_t stream) {
  if (io.B <= 0) return cudaSuccess;
  TrainArgs a;
  a.lo = lo;
  a.packed = io.packed; a.x = io.x; a.start = io.start; a.eps = io.eps;
  a.stash = io.stash; a.slabs = io.slabs;
  a.recon = io.recon; a.mu = io.mu; a.logvar = io.logvar; a.hc = io.hc;
  a.g_recon = io.g_recon; a.g_mu = io.g_mu; a.g_logvar = io.g_logvar; a.g_hc = io.g_hc;
  a.seed = io.seed; a.sample_offset = io.sample_offset; a.step = io.step;
  a.B = io.B;
  a.stash_stride = plan.stash_stride;
  a.w_recon = io.w_recon; a.w_kld = io.w_kld; a.w_start = io.w_start; a.w_time = io.w_time;
  a.inv_batch = io.inv_batch;
  a.stages = plan.stages; a.mode = mode; a.slab_stride = plan.slab_stride;
  cudaError_t e;
  if (plan.M == 64) {
    e = cudaFuncSetAttribute(train_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    if (e != cudaSuccess) return e;
    train_kernel<64><<<plan.grid, BLOCK_THREADS, plan.smem, stream>>>(a);
  } else {
    e = cudaFuncSetAttribute(train_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    if (e != cudaSuccess) return e;
    train_kernel<32><<<plan.grid, BLOCK_THREADS, plan.smem, stream>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace dmvae
