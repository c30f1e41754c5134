// dmvae_common.cuh - shared device/host infrastructure of libdmvae (sm_100a only).
//
//   * Layout      : offsets of the 24 state_dict tensors (torch layout) and of the
//                   kernel-layout ("packed") weight arena.
//   * WeightRing  : weights are streamed L2 -> shared memory by a dedicated producer
//                   warp with TMA bulk copies (cp.async.bulk, completion on an
//                   mbarrier) through a ring of 32 KB stages; the eight consumer warps
//                   wait on the "full" barrier of a stage, run FFMA over it and release
//                   it through the "empty" barrier.
//   * gemm_chunk  : the one FFMA primitive,  C[i][j] += sum_c P[c][i] * Q[c][j],
//                   both operands contraction-major in shared memory, register-tiled
//                   (4*GI) x (VJ*GJ) per thread with 128-bit shared loads.
//   * gemm_nt4    : C[i][j] += sum_c P[i][c] * Q[j][c] (both operands vectorised
//                   along the contraction), used by the weight-gradient GEMMs.
//   * Philox4x32-10 + Box-Muller for in-kernel latent / reparameterisation noise.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dmvae.h"

namespace dmvae {

constexpr int H = 128;             // hidden width (every shipped checkpoint)
constexpr int NUM_LAYERS = 11;     // cond0 cond1 enc0 enc1 enc2 enc3 heads dec0 dec1 dec2 dec3
constexpr int CONSUMER_WARPS = 8;
constexpr int CONSUMER_THREADS = CONSUMER_WARPS * 32;
constexpr int BLOCK_THREADS = CONSUMER_THREADS + 32;  // + one producer warp
constexpr int STAGE_FLOATS = 8192;                    // 32 KB ring stage
constexpr int STAGE_BYTES = STAGE_FLOATS * 4;

// Tensor-core (tcgen05 kind::tf32, 3xTF32 split) weight images, one per dense layer.  Each is
// stored as two planes (TF32 high halves, then low halves); a plane is the exact shared-memory
// image the UMMA descriptors expect, K-step by K-step: [k-step of 8][k-chunk of 4][n-group of 8]
// [8 n][4 k], i.e. 128-byte core matrices, no swizzle.  Any run of consecutive K-steps is a
// contiguous byte range of each plane, so a ring stage is filled by two TMA bulk copies.  Read
// K-major the image is the forward operand W[n][k].  The data-gradient GEMM contracts over n
// instead; it reads W MN-major, which for TF32 the tensor core only accepts in the 32-byte
// swizzled layout of dmvae_tc.cuh, so layers with a data gradient carry a second image ("t"
// planes): [group of gsz 8-deep steps of n][slice of 32 input features k][step][two 512-byte atoms
// [4 n][32 k swizzled]].  A ring stage holds one group for a run of slices (32 output columns each),
// so a data-gradient MMA is as wide as a forward one.
enum TcId {
  TC_COND0 = 0, TC_COND1, TC_ENC0, TC_ENC1, TC_ENC2, TC_ENC3, TC_HEADS, TC_DEC0, TC_DEC1, TC_DEC2, TC_DEC3, NUM_TC
};
struct TcLayer {
  int off_hi;    // float offset of the high plane in the packed arena
  int off_lo;    // ... of the low plane
  int K;         // contraction length (multiple of 8).  cond0 = 8: [x0, y0, 1 (bias row), 0...];
                 // enc0 = Ip (zero rows past I); heads = 256: [h_traj ; h_c];
                 // dec0 = 128 (h_c) + Lp16 (z, zero padded)
  int Kb;        // K plus the bias step (8 more rows) of the forward planes
  int N;         // output width (multiple of 16): 128; heads = NH; dec3 = Ip
  int kps;       // K-steps per 32 KB ring stage (both planes): 512 / N
  int off_thi;   // data-gradient image (high plane), -1 if the layer needs none
  int off_tlo;
  int Kt;        // its output width: K rounded up to 32 (dec0: 128 + Lz32)
  int gsz;       // contraction steps per group: min(4, N / 8)
};

// Training stash: what the forward / backward chain kernel leaves for the weight-gradient
// kernel, per 128-row tile: every layer input X and every pre-activation gradient G as a raw
// fp32 operand image [row group of 8][feature chunk of 4][8 rows][4 features].
enum TcSlot {
  SX_START = 0, SX_HC1, SX_HC, SX_X, SX_E1, SX_E2, SX_E3, SX_E4, SX_Z, SX_D1, SX_D2, SX_D3,
  SG_REC, SG_D3, SG_D2, SG_D1, SG_ML, SG_HC, SG_HC1, SG_E4, SG_E3, SG_E2, SG_E1, NUM_SLOTS
};

enum LayerId { L_COND0 = 0, L_COND1, L_ENC0, L_ENC1, L_ENC2, L_ENC3, L_HEADS, L_DEC0, L_DEC1, L_DEC2, L_DEC3 };

__host__ __device__ constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }

// Offsets in floats.  Torch layout ("p_*") follows the state_dict order
// (Training_VAE.py:132-167); the packed arena ("q_*") holds, per layer, the
// transposed weight Wt[k][Np] (k-major, N padded with zeros to Np) followed by the
// bias padded to Np.  The heads layer packs fc_mu and fc_logvar side by side
// (columns [0,L) = mu, [L,2L) = logvar); rows [0,128) multiply h_traj, [128,256) h_c
// (Training_VAE.py:193).  dec0 rows [0,L) multiply z, [L,L+128) h_c (:214).
struct Layout {
  int T, L, I;      // seq_len, latent_dim, 3*seq_len
  int Ip;           // I padded: 32, 64 or 128; for I > 128 the width of one chunk of the flattened trajectory (128)
  int NC;           // chunks of 128 features the first encoder / last decoder layer is cut into (1 unless I > 128)
  int Ipt;          // NC * Ip: padded total width of the flattened trajectory
  int L2p;          // 2L padded: 32, 64 or 128
  int n_params;
  int n_packed;
  int p_w[NUM_LAYERS];   // heads: fc_mu.weight
  int p_b[NUM_LAYERS];   // heads: fc_mu.bias
  int p_wlv, p_blv;      // fc_logvar.weight / bias
  int K[NUM_LAYERS];     // contraction length of the forward GEMM
  int N[NUM_LAYERS];     // true output width
  int Np[NUM_LAYERS];    // padded output width
  int q_w[NUM_LAYERS];
  int q_b[NUM_LAYERS];
  // data-gradient operands: aligned copies W[n][k] (n-major = contraction-major for
  // dX = dY * W).  r_w[l] < 0 for layers whose input needs no gradient (cond0, enc0).
  // heads: rows are mu then logvar; r_w[heads] = columns [0,128) (the h_traj part,
  // [2L][128]), r_heads_c = columns [128,256) (the h_c part).  dec0: r_w = the h_c part
  // [128][128]; r_dec0z = the z part [128][Lzp], Lzp = L padded to 32/64 with zeros.
  int r_w[NUM_LAYERS];
  int r_heads_c;
  int r_dec0z, Lzp;
  int Lq;            // L rounded up to 4 (rows of the latent tiles in shared memory)
  int Lp8;           // L rounded up to 8 (K granularity of a TF32 MMA)
  int Lp16;          // L rounded up to 16 (N granularity of an M = 128 MMA)
  int NH;            // 2L padded to 16, 32, 64 or 128: width of the heads layer on the tensor cores
  // trajectories longer than 128 floats (NC > 1): generation still runs on the tensor cores; the last decoder layer
  // is cut into NC64 chunks of 64 outputs, each with its own forward image (high plane, then low plane, 8192
  // floats each: [k-step][k-chunk of 4][n-group of 8][8 n][4 k] with N = 64, no bias row)
  int NC64, d3c_off;
  // ... and TRAINING on the tensor cores cuts the last decoder layer into NC chunks of 128 outputs, each a 128 x 128
  // layer of its own: [forward high plane][forward low plane][data-gradient high][data-gradient low], 16384 floats each
  // (no bias step: the loss epilogue adds the bias); zero beyond output I
  int d3t_off;
  TcLayer tc[NUM_TC];
  int slot_off[NUM_SLOTS];  // float offset of a stash slot inside a tile's stash
  int slot_w[NUM_SLOTS];    // features per row of the slot in memory (multiple of 32)
  int slot_n[NUM_SLOTS];    // features that carry data (multiple of 16): the N of an MMA that reads the slot
  int tile_stash;           // floats per tile
};

__host__ __device__ inline int pad_width(int n) { return n <= 32 ? 32 : (n <= 64 ? 64 : 128); }

// Returns 0 on success, DMVAE_ERR_SHAPE when outside the envelope.
inline int make_layout(const DmvaeCfg* c, Layout* lo) {
  if (!c || c->dim != 3 || c->hidden_dim != H) return DMVAE_ERR_SHAPE;
  if (c->latent_dim < 1 || c->latent_dim > 64) return DMVAE_ERR_SHAPE;
  if (c->seq_len < 2 || c->seq_len > 400) return DMVAE_ERR_SHAPE;
  Layout& l = *lo;
  l.T = c->seq_len; l.L = c->latent_dim; l.I = 3 * l.T;
  l.Ip = pad_width(l.I < 128 ? l.I : 128); l.L2p = pad_width(2 * l.L);
  l.NC = (l.I + 127) / 128;
  l.Ipt = l.NC * l.Ip;
  const int Ks[NUM_LAYERS] = {2, H, l.I, H, H, H, 2 * H, l.L + H, H, H, H};
  const int Ns[NUM_LAYERS] = {H, H, H, H, H, H, 2 * l.L, H, H, H, l.I};
  int p = 0, q = 0;
  for (int i = 0; i < NUM_LAYERS; ++i) {
    l.K[i] = Ks[i]; l.N[i] = Ns[i];
    l.Np[i] = (i == L_HEADS) ? l.L2p : (i == L_DEC3 ? l.Ip : H);
    if (i == L_HEADS) {
      l.p_w[i] = p; p += l.L * 2 * H;
      l.p_b[i] = p; p += l.L;
      l.p_wlv = p; p += l.L * 2 * H;
      l.p_blv = p; p += l.L;
    } else {
      l.p_w[i] = p; p += Ns[i] * Ks[i];
      l.p_b[i] = p; p += Ns[i];
    }
    // dec3 with I > 128: one [k][128] image per chunk of 128 outputs, then the bias of all chunks
    const int np_total = (i == L_DEC3) ? l.Ipt : l.Np[i];
    l.q_w[i] = q; q += round_up(Ks[i] * np_total, 4);
    l.q_b[i] = q; q += np_total;
  }
  l.n_params = p;
  q = round_up(q, 4);
  l.Lq = round_up(l.L, 4);
  for (int i = 0; i < NUM_LAYERS; ++i) {
    if (i == L_COND0 || i == L_ENC0) { l.r_w[i] = -1; continue; }
    l.r_w[i] = q;
    if (i == L_HEADS) {
      q += 2 * l.L * H;
      l.r_heads_c = q;
      q += 2 * l.L * H;
    } else if (i == L_DEC3) {
      q += l.I * H;
    } else {
      q += H * H;
    }
    q = round_up(q, 4);
  }
  l.Lzp = l.L <= 32 ? 32 : 64;
  l.r_dec0z = q; q += H * l.Lzp;
  q = round_up(q, 32);  // 128-byte aligned stages for the TMA bulk copies
  l.Lp8 = round_up(l.L, 8);
  l.Lp16 = round_up(l.L, 16);
  l.NH = 2 * l.L <= 16 ? 16 : l.L2p;
  for (int t = 0; t < NUM_TC; ++t) {
    TcLayer& c = l.tc[t];
    c.K = H; c.N = H;
    if (l.NC > 1 && t == TC_DEC3) {
      // long trajectories: dec3 is chunked (d3c images for generation, d3t images for training, below)
      c.K = 0; c.Kb = 0; c.N = 16; c.kps = 32; c.off_hi = c.off_lo = q; c.off_thi = c.off_tlo = -1; c.Kt = 0; c.gsz = 0;
      continue;
    }
    if (t == TC_COND0) c.K = 8;
    else if (t == TC_ENC0) c.K = l.Ipt;   // long trajectories: NC x 128 contraction rows, walked chunk by chunk
    else if (t == TC_HEADS) { c.K = 2 * H; c.N = l.NH; }
    else if (t == TC_DEC0) c.K = H + l.Lp16;
    else if (t == TC_DEC3) c.N = l.Ip;
    c.kps = STAGE_FLOATS / (16 * c.N);
    // forward planes carry one more K step: the bias row (k = K) and seven zero rows, multiplied by a
    // constant ones column of the A operand (cond0 has its bias row inside its K = 8)
    c.Kb = t == TC_COND0 ? c.K : c.K + 8;
    c.off_hi = q; q += c.Kb * c.N;
    c.off_lo = q; q += c.Kb * c.N;
    c.off_thi = c.off_tlo = -1; c.Kt = 0; c.gsz = 0;
    if (t != TC_COND0 && t != TC_ENC0) {
      c.Kt = (t == TC_DEC0) ? H + round_up(l.L, 32) : c.K;
      c.gsz = c.N / 8 < 4 ? c.N / 8 : 4;
      c.off_thi = q; q += c.Kt * c.N;
      c.off_tlo = q; q += c.Kt * c.N;
    }
  }
  l.NC64 = (l.I + 63) / 64;
  l.d3c_off = q;
  if (l.NC > 1) {
    TcLayer& c = l.tc[TC_DEC3];
    c.K = H; c.Kb = 0; c.N = 64; c.kps = 8; c.off_hi = q; c.off_lo = q + 8192;
    q += l.NC64 * 16384;
  }
  l.d3t_off = q;
  if (l.NC > 1) q += l.NC * 65536;
  {
    int o = 0;
    for (int sl = 0; sl < NUM_SLOTS; ++sl) {
      int w = H;
      if (sl == SX_START) w = 16;
      else if (sl == SX_X || sl == SG_REC) w = l.Ipt;   // long trajectories: NC images of 128 features, back to back
      else if (sl == SX_Z) w = l.Lp16;
      else if (sl == SG_ML) w = l.NH;
      l.slot_off[sl] = o; l.slot_n[sl] = w; l.slot_w[sl] = round_up(w, 32);
      o += 128 * l.slot_w[sl];
    }
    l.tile_stash = o;
  }
  l.n_packed = round_up(q, 4);
  return DMVAE_OK;
}

// The last decoder layer as the training chain sees it: chunk c of 128 outputs as a 128 x 128 layer (one chunk - the
// layer's own images - up to 128 features).
__host__ __device__ inline TcLayer dec3_chunk_layer(const Layout& lo, int c) {
  if (lo.NC == 1) return lo.tc[TC_DEC3];
  TcLayer t;
  t.off_hi = lo.d3t_off + c * 65536; t.off_lo = t.off_hi + 16384; t.off_thi = t.off_hi + 32768; t.off_tlo = t.off_hi + 49152;
  t.K = H; t.Kb = H; t.N = H; t.kps = STAGE_FLOATS / (16 * H); t.Kt = H; t.gsz = 4;
  return t;
}

// --------------------------------------------------------------------------------------
// PTX helpers: mbarrier, TMA bulk copy, named barrier, cp.async
// --------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni DONE_%=;\n\t"
      "bra.uni WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(addr), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on `bar`.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// Barrier among the consumer warps only (the producer warp never joins it).
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_THREADS) : "memory"); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// --------------------------------------------------------------------------------------
// Weight ring
// --------------------------------------------------------------------------------------
// Ring position (stage index + phase parity); producer and every consumer warp keep
// their own copy and advance it once per chunk.
struct RingStateRt {
  int stage = 0, stages;
  uint32_t phase = 0;
  __device__ explicit RingStateRt(int s) : stages(s) {}
  __device__ __forceinline__ void advance() {
    if (++stage == stages) { stage = 0; phase ^= 1u; }
  }
};

__device__ __forceinline__ void produce(const float* src, int rows, int width, float* ring, uint64_t* full,
                                        uint64_t* empty, RingStateRt& rs) {
  const int rpc = STAGE_FLOATS / width;
  for (int r0 = 0; r0 < rows; r0 += rpc) {
    const int n = min(rpc, rows - r0);
    const uint32_t bytes = (uint32_t)(n * width * 4);
    mbar_wait(&empty[rs.stage], rs.phase ^ 1u);
    mbar_arrive_expect_tx(&full[rs.stage], bytes);
    tma_load_1d(ring + rs.stage * STAGE_FLOATS, src + (size_t)r0 * width, bytes, &full[rs.stage]);
    rs.advance();
  }
}

// --------------------------------------------------------------------------------------
// FFMA GEMM primitives
// --------------------------------------------------------------------------------------
// Thread tile (4*GI) x (VJ*GJ); lanes are 8 (i) x 4 (j) with lane = tj*8 + ti so that
// every quarter-warp of a 128-bit P load reads 128 contiguous bytes and every
// quarter-warp of a Q load reads one address (broadcast).  i-groups are 32 apart,
// j-groups 4*VJ apart.  Warps are WI x WJ.
template <int GI_, int GJ_, int VJ_, int WI_, int WJ_>
struct TileCfg {
  static constexpr int GI = GI_, GJ = GJ_, VJ = VJ_, WI = WI_, WJ = WJ_;
  static constexpr int TI = 4 * GI, TJ = VJ * GJ;
  static constexpr int SI = 32, SJ = 4 * VJ;
  static constexpr int I = WI * GI * 32, J = WJ * GJ * SJ;
  static constexpr int ACTIVE_WARPS = WI * WJ;
  static_assert(ACTIVE_WARPS <= CONSUMER_WARPS, "too many warps");
  __device__ static __forceinline__ bool active(int warp) { return warp < ACTIVE_WARPS; }
  __device__ static __forceinline__ int i0(int warp, int lane) { return (warp % WI) * (GI * 32) + (lane & 7) * 4; }
  __device__ static __forceinline__ int j0(int warp, int lane) { return (warp / WI) * (GJ * SJ) + (lane >> 3) * VJ; }
};

// Forward / data-gradient tiles: I = rows of the batch tile, J = output features.
template <int M, int N> struct FwdCfg;
template <> struct FwdCfg<128, 128> : TileCfg<2, 2, 4, 2, 4> {};
template <> struct FwdCfg<128, 64> : TileCfg<2, 1, 4, 2, 4> {};
template <> struct FwdCfg<128, 32> : TileCfg<1, 2, 2, 4, 2> {};
template <> struct FwdCfg<64, 128> : TileCfg<1, 2, 4, 2, 4> {};
template <> struct FwdCfg<64, 64> : TileCfg<1, 1, 4, 2, 4> {};
template <> struct FwdCfg<64, 32> : TileCfg<1, 1, 2, 2, 4> {};
template <> struct FwdCfg<32, 128> : TileCfg<1, 1, 4, 1, 8> {};
template <> struct FwdCfg<32, 64> : TileCfg<1, 1, 2, 1, 8> {};
template <> struct FwdCfg<32, 32> : TileCfg<1, 1, 1, 1, 8> {};

template <int V> struct VecLoad;
template <> struct VecLoad<4> {
  __device__ static __forceinline__ void ld(float* d, const float* s) {
    const float4 v = *reinterpret_cast<const float4*>(s);
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
};
template <> struct VecLoad<2> {
  __device__ static __forceinline__ void ld(float* d, const float* s) {
    const float2 v = *reinterpret_cast<const float2*>(s);
    d[0] = v.x; d[1] = v.y;
  }
};
template <> struct VecLoad<1> {
  __device__ static __forceinline__ void ld(float* d, const float* s) { d[0] = *s; }
};

// acc[i][j] += sum_{c<rows} P[c*ldp + i] * Q[c*ldq + j]   (P, Q already offset to the
// thread's i0 / j0).
template <class C>
__device__ __forceinline__ void gemm_chunk(float (&acc)[C::TI][C::TJ], const float* __restrict__ P, int ldp,
                                           const float* __restrict__ Q, int ldq, int rows) {
#pragma unroll 4
  for (int c = 0; c < rows; ++c) {
    float a[C::TI], b[C::TJ];
#pragma unroll
    for (int g = 0; g < C::GI; ++g) VecLoad<4>::ld(&a[4 * g], P + c * ldp + g * C::SI);
#pragma unroll
    for (int g = 0; g < C::GJ; ++g) VecLoad<C::VJ>::ld(&b[C::VJ * g], Q + c * ldq + g * C::SJ);
#pragma unroll
    for (int i = 0; i < C::TI; ++i)
#pragma unroll
      for (int j = 0; j < C::TJ; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

template <class C>
__device__ __forceinline__ void zero_acc(float (&acc)[C::TI][C::TJ]) {
#pragma unroll
  for (int i = 0; i < C::TI; ++i)
#pragma unroll
    for (int j = 0; j < C::TJ; ++j) acc[i][j] = 0.f;
}

// Consume one streamed operand (`rows` contraction rows of `width` floats, cut into
// ring chunks): acc += P^T * Q.  Every consumer warp waits for and releases every chunk,
// also warps that own no output of this tile shape, so the ring protocol does not
// depend on the tile configuration.
template <class C>
__device__ __forceinline__ void consume(float (&acc)[C::TI][C::TJ], const float* __restrict__ P, int ldp, int rows,
                                        int width, const float* ring, uint64_t* full, uint64_t* empty,
                                        RingStateRt& rs, int warp, int lane) {
  const int rpc = STAGE_FLOATS / width;
  const bool act = C::active(warp);
  const int i0 = C::i0(warp, lane), j0 = C::j0(warp, lane);
  for (int r0 = 0; r0 < rows; r0 += rpc) {
    const int n = min(rpc, rows - r0);
    mbar_wait(&full[rs.stage], rs.phase);
    if (act) gemm_chunk<C>(acc, P + r0 * ldp + i0, ldp, ring + rs.stage * STAGE_FLOATS + j0, width, n);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[rs.stage]);
    rs.advance();
  }
}


// --------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) + Box-Muller
// --------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) {  // (0, 1)
  return (float)(x >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f);
}
// Four standard normals for (sample, block); stream distinguishes z / eps / step.
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t sample, uint32_t block, uint32_t stream) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)sample, (uint32_t)(sample >> 32), block, stream),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float r0 = sqrtf(-2.0f * logf(u01(r.x))), r1 = sqrtf(-2.0f * logf(u01(r.z)));
  float s0, c0, s1, c1;
  sincospif(2.0f * u01(r.y), &s0, &c0);
  sincospif(2.0f * u01(r.w), &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}


// --------------------------------------------------------------------------------------
// Shuffled walk over a resident data set (the reference's DataLoader(shuffle=True), Training_VAE.py:327)
// --------------------------------------------------------------------------------------
// pi_{seed, epoch}: a bijection of [0, n) evaluated per element, so that a kernel can pick the rows of its batch
// without any permutation array: a keyed mixing function that is a bijection of k-bit words (k = bits of n - 1:
// xor-shift, multiplication by an odd constant and key addition are each one), cycle-walked until the value falls
// inside [0, n) (fewer than two evaluations on average).  Position p of epoch e reads row resident_row(seed, e, p, n);
// every row is read exactly once per epoch and the order changes with the epoch.  Host and device evaluate the same code.
__host__ __device__ inline uint32_t resident_mix(uint32_t x, uint32_t mask, int k, uint32_t k0, uint32_t k1) {
  const int s1 = (k + 1) / 2, s2 = (k + 2) / 3 > 0 ? (k + 2) / 3 : 1;
  x = (x + k0) & mask;
  x ^= x >> s1;
  x = (x * 0x9E3779B1u) & mask;
  x ^= x >> s2;
  x = (x + k1) & mask;
  x = (x * 0x85EBCA6Bu) & mask;
  x ^= x >> s1;
  x = (x * 0xC2B2AE35u) & mask;
  x ^= x >> s2;
  return x;
}
__host__ __device__ inline uint32_t resident_row(uint64_t seed, uint64_t epoch, uint32_t pos, uint32_t n) {
  if (n <= 1u) return 0u;
  int k = 1;
  while (k < 32 && (1u << k) < n) ++k;
  const uint32_t mask = k >= 32 ? 0xffffffffu : ((1u << k) - 1u);
  // two key words from (seed, epoch): one splitmix64 step each
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (epoch + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const uint32_t k0 = (uint32_t)z, k1 = (uint32_t)(z >> 32);
  uint32_t x = pos;
  do { x = resident_mix(x, mask, k, k0, k1); } while (x >= n);
  return x;
}

}  // namespace dmvae
