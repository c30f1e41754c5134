// dmvae_decode_tc.cu - batched generation on the 5th-generation tensor cores (K1, tcgen05).
//
// Same contract as decode_kernel (dmvae_decode.cu): replaces  condition_encoder -> cat ->
// decoder -> + start  of Tools.py:55-63 / Tools.py:898-912 (model code Training_VAE.py:132-137,
// :158-167, :208-215) for a whole batch.  The 128-wide hidden layers are real dense
// contractions (M = 128-row tile, N = K = 128) and run as tcgen05.mma kind::tf32; because one
// TF32 pass (10-bit mantissa) cannot hold the 1e-5 decode tolerance, every product is
// error-compensated (3xTF32):   a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo,   x_hi = tf32(x),
// x_lo = tf32(x - x_hi), accumulated in fp32 in tensor memory.
//
// One persistent CTA per SM walks 128-row tiles.  The activations of a tile never leave
// tensor memory:
//     TMEM columns [  0,128)  D      fp32 accumulator of the layer in flight (lane = row)
//                  [128,320)  A_hi   TF32 high halves of the layer input: 128 hidden + <=64 latent
//                  [320,512)  A_lo   TF32 low halves
//   MMA warp        walks the layer program convergently, one elected lane issues
//                   D = A(TMEM) x B(smem)  per 8-deep K step, three terms
//   producer warp   streams the B operands (weights, pre-split and pre-laid-out by pack_kernel in
//                   the UMMA K-major core-matrix image) L2 -> smem ring with TMA bulk copies
//   8 hidden warps  tcgen05.ld D -> bias + ReLU -> hi/lo split -> tcgen05.st A for the next layer
//                   (warp w owns TMEM lanes 32*(w%4).., column half w/4)
//   4 output warps  off the critical path: draw / load the latents and start points of the NEXT
//                   tile into the spare A columns while the hidden layers run; drain the last
//                   layer D -> registers -> smem -> one TMA bulk store per tile.
// With a per-row start point the first condition-encoder layer (2 -> 128) is one more tiny MMA
// over the columns [x0, y0, 1, 0...] (the bias rides on the ones column); with a shared start
// point the whole condition encoder is folded once per CTA into the bias of dec0.
// Trajectories longer than 128 floats (3 * seq_len > 128, up to seq_len = 400): the last layer runs as a
// sequence of N = 64 chunks that alternate between the two halves of the accumulator, so the output warps
// drain chunk c (bias, start-point add, staging tile, coalesced stores) while the MMAs of chunk c + 1 run.
#include "dmvae_common.cuh"
#include "dmvae_tc.cuh"

namespace dmvae {

constexpr int TC_EPI_WARPS = 8;                  // hidden-layer epilogue warps
constexpr int TC_EPI_THREADS = TC_EPI_WARPS * 32;
constexpr int TC_OUT_WARPS = 4;                  // staging + output warps
constexpr int TC_OUT_THREADS = TC_OUT_WARPS * 32;
constexpr int TC_PRODUCER_WARP = TC_EPI_WARPS + TC_OUT_WARPS;
constexpr int TC_MMA_WARP = TC_PRODUCER_WARP + 1;
constexpr int TC_THREADS = (TC_MMA_WARP + 1) * 32;
constexpr int TC_MAX_OPS = 6;
constexpr int TC_M = 128;
constexpr uint32_t TM_D = 0, TM_AHI = 128, TM_ALO = 320, TM_Z = 128 /* within an A region */, TM_COLS = 512;

struct TcArgs {
  Layout lo;
  const float* packed;
  const float* z;       // (B, L) or null (Philox)
  const float* start;   // (B, 2) or (1, 2)
  float* out;           // (B, T, 3)
  float* z_out;         // (B, L) or null
  unsigned long long seed, sample_offset;
  long long B;
  int shared_start, add_start, stages, out_bufs, bulk_ok;
  long long* trace;   // development aid: per-op clock64 stamps of CTA 0 (null in production)
};

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 2, %0;" ::"n"(TC_EPI_THREADS) : "memory"); }
__device__ __forceinline__ void out_sync() { asm volatile("bar.sync 3, %0;" ::"n"(TC_OUT_THREADS) : "memory"); }

// ---------------------------------------------------------------------------------------
// the per-tile GEMM program, identical for the producer and the MMA warp
// ---------------------------------------------------------------------------------------
struct TcOp {
  int off_hi;   // float offset of the layer's high plane in the packed arena
  int off_lo;   // ... low plane
  int k0;       // first K step of the plane this op reads
  int ksteps;   // 8-deep K steps
  int kps;      // K steps per stage
  int N;
  int a_col;    // first A column inside the A regions
  int d_col;    // first accumulator column
};
__device__ __forceinline__ TcOp tc_op(const TcLayer& c, int k0, int ksteps, int a_col) {
  return TcOp{c.off_hi, c.off_lo, k0, ksteps, c.kps, c.N, a_col, 0};
}
// shared start:  dec0 (latent rows only) -> dec1 -> dec2 -> dec3
// per-row start: cond0 -> cond1 -> dec0 -> dec1 -> dec2 -> dec3
// long trajectories: dec3 = NC64 ops of 64 outputs each, on alternating halves of the accumulator
__device__ __forceinline__ int tc_hidden_ops(bool shared_start) { return shared_start ? 3 : 5; }
template <bool kChunked>
__device__ __forceinline__ TcOp tc_get_op(const Layout& lo, bool shared_start, int o) {
  const int nh = tc_hidden_ops(shared_start);
  if (o >= nh) {
    if (!kChunked) return tc_op(lo.tc[TC_DEC3], 0, H / 8, 0);
    const int c = o - nh;
    return TcOp{lo.d3c_off + c * 16384, lo.d3c_off + c * 16384 + 8192, 0, H / 8, 8, 64, 0, (c & 1) * 64};
  }
  if (shared_start) {  // skip the K steps of dec0's h_c rows
    if (o == 0) return tc_op(lo.tc[TC_DEC0], H / 8, lo.Lp8 / 8, (int)TM_Z);
    return tc_op(lo.tc[o == 1 ? TC_DEC1 : TC_DEC2], 0, H / 8, 0);
  }
  if (o == 0) return tc_op(lo.tc[TC_COND0], 0, 1, (int)TM_Z + lo.Lp8);
  if (o == 1) return tc_op(lo.tc[TC_COND1], 0, H / 8, 0);
  if (o == 2) return tc_op(lo.tc[TC_DEC0], 0, (H + lo.Lp8) / 8, 0);
  return tc_op(lo.tc[o == 3 ? TC_DEC1 : TC_DEC2], 0, H / 8, 0);
}

struct TcSmem {
  float *ring, *outs, *bias, *tmp, *b3;
  uint64_t *full, *empty, *d_hid, *d_out, *z_free, *a_ready, *tile_ready, *d_free;
  uint32_t* tmem_slot;
};
constexpr int TC_CHUNK_LD = 68;   // row stride (floats) of a 128 x 64 output staging tile: conflict-free 128-bit row writes
// bias: one 128-float row per op in program order (the last row = dec3's bias, Ip entries); b3: dec3's bias of a
// long trajectory (NC64 * 64 entries); outs: whole output tiles (out_bufs of them), or two chunk staging tiles
__host__ __device__ inline size_t tc_out_floats(const Layout& lo, int out_bufs) {
  return lo.NC > 1 ? (size_t)2 * TC_M * TC_CHUNK_LD : (size_t)out_bufs * TC_M * lo.I;
}
__host__ __device__ inline size_t tc_smem_floats(const Layout& lo, int stages, int out_bufs) {
  return (size_t)stages * STAGE_FLOATS + tc_out_floats(lo, out_bufs) + TC_MAX_OPS * H + 2 * H + (lo.NC > 1 ? lo.NC64 * 64 : 0);
}
__host__ __device__ inline size_t tc_smem_bytes(const Layout& lo, int stages, int out_bufs) {
  return tc_smem_floats(lo, stages, out_bufs) * 4 + 32 * 8 + 16 + 1024;
}

// kChunked: long trajectories (lo.NC > 1); a separate instantiation keeps the short-trajectory kernel free of the
// chunk path's registers
template <bool kChunked>
__global__ void __launch_bounds__(TC_THREADS, 1) decode_tc_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  const Layout& lo = a.lo;
  const int L = lo.L, I = lo.I, Lp8 = lo.Lp8;
  TcSmem s;
  {
    // 1024-byte aligned base: UMMA descriptors address shared memory in 16-byte units
    const uint32_t base = smem_u32(smem_dyn);
    unsigned char* p = smem_dyn + ((1024u - (base & 1023u)) & 1023u);
    s.ring = reinterpret_cast<float*>(p);
    s.outs = s.ring + (size_t)a.stages * STAGE_FLOATS;
    s.bias = s.outs + tc_out_floats(lo, a.out_bufs);
    s.tmp = s.bias + TC_MAX_OPS * H;
    s.b3 = s.tmp + 2 * H;
    s.full = reinterpret_cast<uint64_t*>(s.b3 + (lo.NC > 1 ? lo.NC64 * 64 : 0));
    s.empty = s.full + 8;
    s.d_hid = s.empty + 8;
    s.d_out = s.d_hid + 1;        // [2]: long trajectories signal the two halves of the accumulator separately
    s.z_free = s.d_out + 2;
    s.a_ready = s.z_free + 1;
    s.tile_ready = s.a_ready + 1;
    s.d_free = s.tile_ready + 1;  // [2]: output warps -> MMA, half of the accumulator drained (long trajectories)
    s.tmem_slot = reinterpret_cast<uint32_t*>(s.d_free + 2);
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* __restrict__ pk = a.packed;
  const long long n_tiles = (a.B + TC_M - 1) / TC_M;
  const bool shared_start = a.shared_start != 0;

  if (tid == 0) {
    for (int st = 0; st < a.stages; ++st) {
      mbar_init(&s.full[st], 1);
      mbar_init(&s.empty[st], 1);
    }
    // Every barrier completes exactly one phase per hand-shake with its consumers, so no waiter
    // can fall two phases behind:  d_hid  MMA -> hidden warps (one phase per hidden layer),
    // d_out  MMA -> output warps (last layer), z_free  MMA -> output warps (latent columns read),
    // a_ready  hidden warps -> MMA,  tile_ready  output warps -> MMA.
    mbar_init(s.d_hid, 1);
    mbar_init(&s.d_out[0], 1);
    mbar_init(&s.d_out[1], 1);
    mbar_init(&s.d_free[0], TC_OUT_WARPS);
    mbar_init(&s.d_free[1], TC_OUT_WARPS);
    mbar_init(s.z_free, 1);
    mbar_init(s.a_ready, TC_EPI_WARPS);
    mbar_init(s.tile_ready, TC_OUT_WARPS);
    mbar_fence_init();
  }
  if (warp == TC_PRODUCER_WARP) tmem_alloc(s.tmem_slot, TM_COLS);  // whole warp; this warp also frees it
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s.tmem_slot;

  const int n_hidden = tc_hidden_ops(shared_start);
  constexpr bool chunked = kChunked;
  // the hidden layers' ops (and the single last-layer op of a short trajectory) are built once; the chunk ops of a
  // long trajectory are derived from the chunk index
  TcOp ops_fixed[TC_MAX_OPS];
  for (int o = 0; o < TC_MAX_OPS; ++o) ops_fixed[o] = tc_get_op<kChunked>(lo, shared_start, o < n_hidden + 1 ? o : n_hidden);
  auto get_op = [&](int o) -> TcOp { return (!kChunked || o < n_hidden) ? ops_fixed[o] : tc_get_op<kChunked>(lo, shared_start, o); };
  const int n_last = chunked ? lo.NC64 : 1;        // ops of the last layer
  const int n_ops = n_hidden + n_last;
  const int zop = shared_start ? 0 : 2;  // the op that reads the staged latent columns last (dec0)

  if (warp == TC_PRODUCER_WARP) {
    // ===================== producer warp: weights L2 -> smem ring ============================
    if (lane == 0) {
      RingStateRt rs(a.stages);
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int o = 0; o < n_ops; ++o) {
          const TcOp op = get_op(o);
          for (int k0 = 0; k0 < op.ksteps; k0 += op.kps) {
            const int nks = min(op.kps, op.ksteps - k0);
            const int fl = nks * op.N * 8;                       // floats per plane
            const size_t src = (size_t)(op.k0 + k0) * op.N * 8;  // K steps are contiguous in a plane
            float* dst = s.ring + rs.stage * STAGE_FLOATS;
            mbar_wait(&s.empty[rs.stage], rs.phase ^ 1u);
            mbar_arrive_expect_tx(&s.full[rs.stage], (uint32_t)(2 * fl * 4));
            tma_load_1d(dst, pk + op.off_hi + src, (uint32_t)(fl * 4), &s.full[rs.stage]);
            tma_load_1d(dst + fl, pk + op.off_lo + src, (uint32_t)(fl * 4), &s.full[rs.stage]);
            rs.advance();
          }
        }
    }
  } else if (warp == TC_MMA_WARP) {
    // ===================== MMA warp ===========================================================
    // The whole warp walks the program convergently so that every operand of tcgen05.mma is
    // warp-uniform (uniform registers, no per-thread fix-up code); one elected lane issues.
    RingStateRt rs(a.stages);
    uint32_t a_phase = 0, t_phase = 0, f_phase[2] = {0u, 0u};
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
      for (int o = 0; o < n_ops; ++o) {
        const TcOp op = get_op(o);
        const uint32_t idesc = umma_idesc_tf32(TC_M, op.N);
        const uint32_t kstep_bytes = (uint32_t)op.N * 32u;  // one 8-deep K step of B: N x 8 TF32
        const uint64_t desc_hi_bits = umma_desc(0u, (uint32_t)op.N * 16u, 128u);
        const int chunk = o - n_hidden;   // >= 0: an op of the last layer
        if (o == 0) {  // latents / start points of this tile staged, D of the previous tile drained
          mbar_wait(s.tile_ready, t_phase);
          t_phase ^= 1u;
        } else if (chunk <= 0) {       // the previous layer's output is in A
          mbar_wait(s.a_ready, a_phase);
          a_phase ^= 1u;
        }
        if (chunked && chunk >= 2) {   // this half of the accumulator was last written by chunk - 2: drained?
          mbar_wait(&s.d_free[chunk & 1], f_phase[chunk & 1]);
          f_phase[chunk & 1] ^= 1u;
        }
        tc_fence_after();
        const long long tix = (tile - blockIdx.x) / gridDim.x;
        const bool tr = a.trace != nullptr && blockIdx.x == 0 && tix < 4 && lane == 0 && o < 8;
        if (tr) a.trace[(tix * 8 + o) * 4 + 0] = clock64();
        uint32_t acc = 0;
        uint32_t a_hi = tmem + TM_AHI + (uint32_t)op.a_col, a_lo = tmem + TM_ALO + (uint32_t)op.a_col;
        for (int k0 = 0; k0 < op.ksteps; k0 += op.kps) {
          const int nks = min(op.kps, op.ksteps - k0);
          mbar_wait(&s.full[rs.stage], rs.phase);
          tc_fence_after();
          const uint32_t b_hi = smem_u32(s.ring + rs.stage * STAGE_FLOATS);
          uint64_t dh = desc_hi_bits | (uint64_t)(b_hi >> 4);
          uint64_t dl = desc_hi_bits | (uint64_t)((b_hi + (uint32_t)nks * kstep_bytes) >> 4);
          const uint64_t dinc = (uint64_t)(kstep_bytes >> 4);
          if (elect_one()) {
#pragma unroll 4
            for (int ks = 0; ks < nks; ++ks) {
              umma_tf32_ts(tmem + TM_D + (uint32_t)op.d_col, a_hi, dh, idesc, acc);   // a_hi * b_hi
              umma_tf32_ts(tmem + TM_D + (uint32_t)op.d_col, a_lo, dh, idesc, 1u);    // a_lo * b_hi
              umma_tf32_ts(tmem + TM_D + (uint32_t)op.d_col, a_hi, dl, idesc, 1u);    // a_hi * b_lo
              acc = 1u;
              a_hi += 8u; a_lo += 8u; dh += dinc; dl += dinc;
            }
            umma_commit(&s.empty[rs.stage]);  // the stage is free once these MMAs have read it
          }
          __syncwarp();
          a_hi = tmem + TM_AHI + (uint32_t)op.a_col + 8u * (uint32_t)(k0 + nks);
          a_lo = tmem + TM_ALO + (uint32_t)op.a_col + 8u * (uint32_t)(k0 + nks);
          acc = 1u;
          rs.advance();
        }
        if (elect_one()) {
          if (o == zop) umma_commit(s.z_free);                       // the staged columns have been read
          // accumulator (or, for a chunk, its half) complete -> epilogue
          umma_commit(chunk < 0 ? s.d_hid : &s.d_out[chunked ? (chunk & 1) : 0]);
        }
        __syncwarp();
        if (tr) a.trace[(tix * 8 + o) * 4 + 1] = clock64();
      }
  } else if (warp < TC_EPI_WARPS) {
    // ===================== hidden-layer epilogue warps ======================================
    const int q = warp & 3, h = warp >> 2;     // TMEM lane quarter, column half
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);

    // bias rows in program order (broadcast reads in the epilogues)
    for (int i = tid; i < H; i += TC_EPI_THREADS) {
      int r = 0;
      if (!shared_start) {
        s.bias[r++ * H + i] = 0.f;                                  // cond0: the bias rides on the ones column
        s.bias[r++ * H + i] = pk[lo.q_b[L_COND1] + i];
      }
      s.bias[r++ * H + i] = pk[lo.q_b[L_DEC0] + i];
      s.bias[r++ * H + i] = pk[lo.q_b[L_DEC1] + i];
      s.bias[r++ * H + i] = pk[lo.q_b[L_DEC2] + i];
      s.bias[r++ * H + i] = i < lo.Ip ? pk[lo.q_b[L_DEC3] + i] : 0.f;
    }
    if (chunked)
      for (int i = tid; i < lo.NC64 * 64; i += TC_EPI_THREADS) s.b3[i] = i < I ? pk[lo.q_b[L_DEC3] + i] : 0.f;
    if (shared_start) {
      // one start point for the whole launch: condition encoder once per CTA, folded into
      // the bias of dec0:  hb[n] = b_dec0[n] + sum_k Wdec0[n][L+k] * h_c[k]
      const float sx = a.start[0], sy = a.start[1];
      if (tid < H) {
        const float* w0 = pk + lo.q_w[L_COND0];
        float v = pk[lo.q_b[L_COND0] + tid];
        v = fmaf(w0[tid], sx, v);
        v = fmaf(w0[H + tid], sy, v);
        s.tmp[tid] = fmaxf(v, 0.f);
      }
      epi_sync();
      if (tid < H) {
        const float* w1 = pk + lo.q_w[L_COND1];
        float v = pk[lo.q_b[L_COND1] + tid];
        for (int k = 0; k < H; ++k) v = fmaf(w1[k * H + tid], s.tmp[k], v);
        s.tmp[H + tid] = fmaxf(v, 0.f);
      }
      epi_sync();
      if (tid < H) {
        const float* wd = pk + lo.q_w[L_DEC0] + L * H;
        float v = pk[lo.q_b[L_DEC0] + tid];
        for (int k = 0; k < H; ++k) v = fmaf(wd[k * H + tid], s.tmp[H + k], v);
        s.bias[tid] = v;
      }
    }
    epi_sync();
    // the output warps read the dec3 bias row: everybody meets once before the tile loop
    asm volatile("bar.sync 4, %0;" ::"n"(TC_EPI_THREADS + TC_OUT_THREADS) : "memory");

    uint32_t d_phase = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int o = 0; o < n_hidden; ++o) {
        const float* __restrict__ bias = s.bias + o * H + h * 64;
        mbar_wait(s.d_hid, d_phase);
        d_phase ^= 1u;
        tc_fence_after();
        const long long tix = (tile - blockIdx.x) / gridDim.x;
        const bool tr = a.trace != nullptr && blockIdx.x == 0 && tix < 4 && tid == 0;
        if (tr) a.trace[(tix * 8 + o) * 4 + 2] = clock64();
        uint32_t v[4][16];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld16(lane_base + TM_D + h * 64 + c * 16, v[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t hi[16], lw[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) split_tf32(fmaxf(__uint_as_float(v[c][j]) + bias[c * 16 + j], 0.f), hi[j], lw[j]);
          tmem_st16(lane_base + TM_AHI + h * 64 + c * 16, hi);
          tmem_st16(lane_base + TM_ALO + h * 64 + c * 16, lw);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s.a_ready);
        if (tr) a.trace[(tix * 8 + o) * 4 + 3] = clock64();
      }
    }
  } else {
    // ===================== staging + output warps ===========================================
    const int q = warp - TC_EPI_WARPS;         // TMEM lane quarter (warp 8..11 -> warp % 4 = 0..3)
    const int otid = tid - TC_EPI_THREADS;
    const int m = q * 32 + lane;               // row of the tile owned by this thread
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    float sx_cur = 0.f, sy_cur = 0.f, sx_nxt = 0.f, sy_nxt = 0.f;
    if (shared_start) { sx_cur = sx_nxt = a.start[0]; sy_cur = sy_nxt = a.start[1]; }

    // latents (and, per-row, the start point as [x0, y0, 1, 0, 0, 0, 0, 0]) of a tile -> spare A columns
    auto stage_tile = [&](long long tile) {
      const long long row = tile * TC_M + m;
      const bool row_ok = row < a.B;
      for (int jb = 0; jb < Lp8 / 4; ++jb) {
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        if (row_ok) {
          if (a.z != nullptr) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (jb * 4 + i < L) g[i] = __ldg(a.z + row * L + jb * 4 + i);
          } else {
            const float4 r = philox_normal4(a.seed, a.sample_offset + (unsigned long long)row, (uint32_t)jb, 0u);
            g[0] = r.x; g[1] = r.y; g[2] = r.z; g[3] = r.w;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (jb * 4 + i >= L) g[i] = 0.f;
              else if (a.z_out != nullptr) a.z_out[row * L + jb * 4 + i] = g[i];
            }
          }
        }
        uint32_t hi[4], lw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(g[i], hi[i], lw[i]);
        tmem_st4(lane_base + TM_AHI + TM_Z + 4 * jb, hi[0], hi[1], hi[2], hi[3]);
        tmem_st4(lane_base + TM_ALO + TM_Z + 4 * jb, lw[0], lw[1], lw[2], lw[3]);
      }
      if (!shared_start) {
        sx_nxt = row_ok ? __ldg(a.start + row * 2) : 0.f;
        sy_nxt = row_ok ? __ldg(a.start + row * 2 + 1) : 0.f;
        uint32_t xh, xl, yh, yl;
        split_tf32(sx_nxt, xh, xl);
        split_tf32(sy_nxt, yh, yl);
        const uint32_t one = __float_as_uint(1.0f);
        tmem_st4(lane_base + TM_AHI + TM_Z + Lp8, xh, yh, one, 0u);
        tmem_st4(lane_base + TM_AHI + TM_Z + Lp8 + 4, 0u, 0u, 0u, 0u);
        tmem_st4(lane_base + TM_ALO + TM_Z + Lp8, xl, yl, 0u, 0u);
        tmem_st4(lane_base + TM_ALO + TM_Z + Lp8 + 4, 0u, 0u, 0u, 0u);
      }
      tmem_st_wait();
    };
    auto release_tile = [&]() {  // staged columns written, D drained by this warp -> MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s.tile_ready);
    };

    asm volatile("bar.sync 4, %0;" ::"n"(TC_EPI_THREADS + TC_OUT_THREADS) : "memory");  // bias rows are in smem
    const float* __restrict__ b3 = chunked ? s.b3 : s.bias + n_hidden * H;
    // start-point add per column class n % 3 (0: time, 1: x, 2: y), Tools.py:61-63
    uint32_t d_phase = 0, dc_phase[2] = {0u, 0u};
    int out_buf = 0;
    unsigned int gchunk = 0;   // chunks drained so far (long trajectories): parity = staging tile in use
    stage_tile(blockIdx.x);
    sx_cur = sx_nxt; sy_cur = sy_nxt;
    release_tile();
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long m0 = tile * TC_M;
      const int valid = (int)min((long long)TC_M, a.B - m0);
      const long long next = tile + gridDim.x;
      mbar_wait(s.z_free, d_phase);  // the staged columns of this tile have been consumed:
      if (next < n_tiles) {          // stage the next tile while the hidden layers run
        tc_fence_after();
        stage_tile(next);
      }
      if (chunked) {
        // ---- long trajectory: the last layer arrives as NC64 chunks of 64 columns on alternating halves of D
        d_phase ^= 1u;
        const float add1 = a.add_start ? sx_cur : 0.f, add2 = a.add_start ? sy_cur : 0.f;
        float* gdst = a.out + m0 * I;
        const bool vec_ok = (I & 3) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15u) == 0;
        const bool vec2_ok = (I & 1) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 7u) == 0;
        for (int c = 0; c < lo.NC64; ++c, ++gchunk) {
          const int half = c & 1;
          mbar_wait(&s.d_out[half], dc_phase[half]);
          dc_phase[half] ^= 1u;
          tc_fence_after();
          float* tile_st = s.outs + (size_t)(gchunk & 1u) * TC_M * TC_CHUNK_LD;
          float* st = tile_st + m * TC_CHUNK_LD;
          const int nb = c * 64;
#pragma unroll 1
          for (int hf = 0; hf < 2; ++hf) {   // 32 columns at a time (register budget)
            uint32_t o[2][16];
            tmem_ld16(lane_base + TM_D + half * 64 + hf * 32, o[0]);
            tmem_ld16(lane_base + TM_D + half * 64 + hf * 32 + 16, o[1]);
            tmem_ld_wait();
            if (hf == 1) {
              // this half of D is in registers: chunk c + 2 (or the next tile) may overwrite it
              if (c + 2 < lo.NC64) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.d_free[half]);
              }
              if (c == lo.NC64 - 1) release_tile();
            }
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              const int n0 = nb + hf * 32 + cc * 16;
              // start-point add per column class n % 3 (0: time, 1: x, 2: y): the three values in the order the
              // columns of this group of 16 meet them
              const int r0 = n0 % 3;
              const float p0 = r0 == 0 ? 0.f : (r0 == 1 ? add1 : add2);
              const float p1 = r0 == 0 ? add1 : (r0 == 1 ? add2 : 0.f);
              const float p2 = r0 == 0 ? add2 : (r0 == 1 ? 0.f : add1);
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const float4 b4 = *reinterpret_cast<const float4*>(b3 + n0 + j4 * 4);
                const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
                float v4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int j = j4 * 4 + i;
                  const float rel = __uint_as_float(o[cc][j]) + bb[i];
                  v4[i] = rel + (j % 3 == 0 ? p0 : (j % 3 == 1 ? p1 : p2));
                }
                *reinterpret_cast<float4*>(st + hf * 32 + cc * 16 + j4 * 4) = make_float4(v4[0], v4[1], v4[2], v4[3]);
              }
            }
          }
          out_sync();   // the staging tile is complete; the other one may still be read by slower threads
          const int ncol = min(64, I - nb);
          if (vec_ok && ncol == 64) {   // rows are 16-byte aligned: 128-bit copies, 2 rows x 256 bytes per warp instruction
            for (int idx = otid; idx < TC_M * 16; idx += TC_OUT_THREADS) {
              const int row = idx >> 4, c4 = idx & 15;
              if (row < valid)
                *reinterpret_cast<float4*>(gdst + (size_t)row * I + nb + c4 * 4) =
                    *reinterpret_cast<const float4*>(tile_st + row * TC_CHUNK_LD + c4 * 4);
            }
          } else if (vec2_ok && (ncol & 1) == 0) {   // rows are 8-byte aligned: 64-bit copies
            for (int idx = otid; idx < TC_M * 32; idx += TC_OUT_THREADS) {
              const int row = idx >> 5, c2 = idx & 31;
              if (row < valid && c2 * 2 < ncol)
                *reinterpret_cast<float2*>(gdst + (size_t)row * I + nb + c2 * 2) =
                    *reinterpret_cast<const float2*>(tile_st + row * TC_CHUNK_LD + c2 * 2);
            }
          } else {
            for (int idx = otid; idx < TC_M * 64; idx += TC_OUT_THREADS) {   // 32 lanes = 128 contiguous bytes of a row
              const int row = idx >> 6, col = idx & 63;
              if (row < valid && col < ncol) gdst[(size_t)row * I + nb + col] = tile_st[row * TC_CHUNK_LD + col];
            }
          }
        }
        sx_cur = sx_nxt; sy_cur = sy_nxt;
        continue;
      }
      mbar_wait(s.d_out, d_phase);
      d_phase ^= 1u;
      tc_fence_after();
      const long long tix = (tile - blockIdx.x) / gridDim.x;
      const bool tr = a.trace != nullptr && blockIdx.x == 0 && tix < 4 && otid == 0;
      if (tr) a.trace[(tix * 8 + n_ops - 1) * 4 + 2] = clock64();

      float* stage_out = s.outs + (size_t)out_buf * TC_M * I;
      if (otid == 0) {  // the bulk store that last read this buffer has drained it
        if (a.out_bufs > 1) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
      }
      out_sync();
      const float add0 = 0.f, add1 = a.add_start ? sx_cur : 0.f, add2 = a.add_start ? sy_cur : 0.f;
      const int n_chunks = lo.Ip >> 4;  // 2, 4 or 8
      for (int c0 = 0; c0 < n_chunks; c0 += 4) {
        uint32_t o[4][16];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c0 + c < n_chunks) tmem_ld16(lane_base + TM_D + (c0 + c) * 16, o[c]);
        tmem_ld_wait();
        if (c0 + 4 >= n_chunks) {  // D is in registers: the next tile may start
          release_tile();
          if (tr) a.trace[(tix * 8 + n_ops - 1) * 4 + 3] = clock64();
        }
        if (m < valid) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c0 + c < n_chunks) {
              const int nb = (c0 + c) * 16;
              const int r0 = nb % 3;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int n = nb + j;
                if (n < I) {
                  const int d = (r0 + j) % 3;
                  const float rel = __uint_as_float(o[c][j]) + b3[n];
                  stage_out[m * I + n] = rel + (d == 0 ? add0 : (d == 1 ? add1 : add2));
                }
              }
            }
        }
      }
      sx_cur = sx_nxt; sy_cur = sy_nxt;
      const uint32_t bytes = (uint32_t)valid * (uint32_t)I * 4u;
      float* gdst = a.out + m0 * I;
      if (a.bulk_ok && (bytes & 15u) == 0) {
        fence_proxy_async_smem();
        out_sync();
        if (otid == 0) tma_store_1d(gdst, stage_out, bytes);
      } else {
        out_sync();
        for (int idx = otid; idx < valid * I; idx += TC_OUT_THREADS) gdst[idx] = stage_out[idx];
        out_sync();  // stage_out may be rewritten two tiles from now (or next tile with one buffer)
      }
      if (a.out_bufs > 1) out_buf ^= 1;
    }
    if (otid == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_PRODUCER_WARP) tmem_dealloc(tmem, TM_COLS);
}

static long long* g_tc_trace = nullptr;
void set_decode_tc_trace(long long* p) { g_tc_trace = p; }

// false when the configuration needs more spare A columns than tensor memory has
bool decode_tc_supported(const Layout& lo, bool shared_start) {
  return shared_start ? lo.Lp8 <= 64 : lo.Lp8 + 8 <= 64;
}

cudaError_t launch_decode_tc(const Layout& lo, bool shared_start, const float* packed, const float* z, uint64_t seed,
                             uint64_t sample_offset, const float* start, float* out, float* z_out, long long B,
                             int add_start, int sm_count, cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  TcArgs a;
  a.lo = lo; a.packed = packed; a.z = z; a.start = start; a.out = out; a.z_out = z_out;
  a.seed = seed; a.sample_offset = sample_offset; a.B = B;
  a.shared_start = shared_start ? 1 : 0;
  a.add_start = add_start;
  a.trace = g_tc_trace;
  int stages = 4, out_bufs = 2;
  constexpr size_t LIMIT = 232448;
  if (tc_smem_bytes(lo, stages, out_bufs) > LIMIT) out_bufs = 1;
  while (stages > 2 && tc_smem_bytes(lo, stages, out_bufs) > LIMIT) --stages;
  a.stages = stages; a.out_bufs = out_bufs;
  // TMA bulk stores need a 16-byte aligned destination; tile offsets are multiples of 128 rows x 12 T bytes
  a.bulk_ok = ((reinterpret_cast<uintptr_t>(out) & 15u) == 0 && ((TC_M * lo.I * 4) & 15) == 0) ? 1 : 0;
  const size_t smem = tc_smem_bytes(lo, stages, out_bufs);
  const long long n_tiles = (B + TC_M - 1) / TC_M;
  const int grid = (int)(n_tiles < sm_count ? n_tiles : sm_count);
  if (lo.NC > 1) {
    const cudaError_t e = cudaFuncSetAttribute(decode_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    decode_tc_kernel<true><<<grid, TC_THREADS, smem, stream>>>(a);
  } else {
    const cudaError_t e = cudaFuncSetAttribute(decode_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    decode_tc_kernel<false><<<grid, TC_THREADS, smem, stream>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace dmvae
