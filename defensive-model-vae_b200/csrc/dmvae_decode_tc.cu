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
//   MMA warp      one elected thread issues  D = A(TMEM) x B(smem)  per 8-deep K step, three terms
//   producer warp streams the B operands (weights, pre-split and pre-laid-out by pack_kernel in
//                 the UMMA K-major core-matrix image) L2 -> smem ring with TMA bulk copies
//   8 epilogue warps  tcgen05.ld D -> bias + ReLU -> hi/lo split -> tcgen05.st A for the next
//                 layer (warp w owns TMEM lanes 32*(w%4).., column half w/4); draw / load the
//                 latents; the last layer goes registers -> smem -> one TMA bulk store per tile.
// The narrow layers stay FFMA: cond0 (2 -> 128) in the epilogue threads, and with a shared start
// point the whole condition encoder is folded once per CTA into the bias of dec0.
#include "dmvae_common.cuh"

namespace dmvae {

constexpr int TC_EPI_WARPS = 8;
constexpr int TC_EPI_THREADS = TC_EPI_WARPS * 32;
constexpr int TC_THREADS = TC_EPI_THREADS + 64;  // + producer warp + MMA warp
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
};

// ---------------------------------------------------------------------------------------
// tcgen05 / TMEM PTX
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor], kind::tf32, issued by one thread for the CTA
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every tcgen05 operation issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
          taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 bytes stored as
// 128 contiguous bytes; LBO = byte distance between the two 4-wide K chunks of one MMA, SBO = byte
// distance between consecutive 8-row groups along N (bit layout: cute::UMMA::SmemDescriptor).
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
  return d;                // base offset 0, layout type 0 = no swizzle
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, TF32 x TF32, both K-major.
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 2, %0;" ::"n"(TC_EPI_THREADS) : "memory"); }

// x = hi + lo with hi = round-to-nearest TF32 of x; lo is cut to TF32 by the tensor core itself
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// TMA bulk store shared -> global (bulk async-group completion)
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------
// the per-tile GEMM program, identical for the producer and the MMA warp
// ---------------------------------------------------------------------------------------
struct TcOp {
  int off;      // float offset of the first stage in the packed arena
  int ksteps;   // 8-deep K steps
  int kps;      // K steps per stage
  int N;
  int a_col;    // first A column inside the A regions
};
__device__ __forceinline__ int tc_program(const Layout& lo, bool shared_start, TcOp (&ops)[NUM_TC]) {
  int n = 0;
  if (!shared_start) ops[n++] = TcOp{lo.tc[TC_COND1].off, H / 8, lo.tc[TC_COND1].kps, H, 0};
  if (shared_start) {  // only the latent rows of dec0: skip the stages of the h_c rows
    const TcLayer& c = lo.tc[TC_DEC0];
    ops[n++] = TcOp{c.off + (H / 8 / c.kps) * STAGE_FLOATS, lo.Lp8 / 8, c.kps, H, (int)TM_Z};
  } else {
    ops[n++] = TcOp{lo.tc[TC_DEC0].off, (H + lo.Lp8) / 8, lo.tc[TC_DEC0].kps, H, 0};
  }
  ops[n++] = TcOp{lo.tc[TC_DEC1].off, H / 8, lo.tc[TC_DEC1].kps, H, 0};
  ops[n++] = TcOp{lo.tc[TC_DEC2].off, H / 8, lo.tc[TC_DEC2].kps, H, 0};
  ops[n++] = TcOp{lo.tc[TC_DEC3].off, H / 8, lo.tc[TC_DEC3].kps, lo.Ip, 0};
  return n;
}

struct TcSmem {
  float *ring, *outs, *bias, *w0, *hb, *tmp;
  uint64_t *full, *empty, *d_full, *a_ready;
  uint32_t* tmem_slot;
};
// bias rows: 0 cond0, 1 cond1, 2 dec0, 3 dec1, 4 dec2, 5 dec3 (each 128 floats)
__host__ __device__ inline size_t tc_smem_floats(const Layout& lo, int stages, int out_bufs) {
  return (size_t)stages * STAGE_FLOATS + (size_t)out_bufs * TC_M * lo.I + 6 * H + 2 * H + H + 2 * H;
}
__host__ __device__ inline size_t tc_smem_bytes(const Layout& lo, int stages, int out_bufs) {
  return tc_smem_floats(lo, stages, out_bufs) * 4 + 32 * 8 + 16 + 1024;
}

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
    s.bias = s.outs + (size_t)a.out_bufs * TC_M * I;
    s.w0 = s.bias + 6 * H;
    s.hb = s.w0 + 2 * H;
    s.tmp = s.hb + H;
    s.full = reinterpret_cast<uint64_t*>(s.tmp + 2 * H);
    s.empty = s.full + 8;
    s.d_full = s.empty + 8;
    s.a_ready = s.d_full + 1;
    s.tmem_slot = reinterpret_cast<uint32_t*>(s.a_ready + 1);
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
    mbar_init(s.d_full, 1);
    mbar_init(s.a_ready, TC_EPI_WARPS);
    mbar_fence_init();
  }
  if (warp == TC_EPI_WARPS) tmem_alloc(s.tmem_slot, TM_COLS);  // whole warp; this warp also frees it
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s.tmem_slot;

  TcOp ops[NUM_TC];
  const int n_ops = tc_program(lo, shared_start, ops);

  if (warp == TC_EPI_WARPS) {
    // ===================== producer warp: weights L2 -> smem ring ============================
    if (lane == 0) {
      RingStateRt rs(a.stages);
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int o = 0; o < n_ops; ++o) {
          const TcOp op = ops[o];
          for (int k0 = 0, st = 0; k0 < op.ksteps; k0 += op.kps, ++st) {
            const int nks = min(op.kps, op.ksteps - k0);
            const uint32_t bytes = (uint32_t)(2 * nks * op.N * 8 * 4);
            mbar_wait(&s.empty[rs.stage], rs.phase ^ 1u);
            mbar_arrive_expect_tx(&s.full[rs.stage], bytes);
            tma_load_1d(s.ring + rs.stage * STAGE_FLOATS, pk + op.off + (size_t)st * STAGE_FLOATS, bytes, &s.full[rs.stage]);
            rs.advance();
          }
        }
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // ===================== MMA warp: one thread issues every tcgen05.mma =====================
    if (lane == 0) {
      RingStateRt rs(a.stages);
      uint32_t a_phase = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int o = 0; o < n_ops; ++o) {
          const TcOp op = ops[o];
          const uint32_t idesc = umma_idesc_tf32(TC_M, op.N);
          const uint32_t kstep_bytes = (uint32_t)op.N * 32u;  // one 8-deep K step of B: N x 8 TF32
          mbar_wait(s.a_ready, a_phase);  // the layer input is in A, and D has been drained
          a_phase ^= 1u;
          tc_fence_after();
          uint32_t acc = 0;
          for (int k0 = 0; k0 < op.ksteps; k0 += op.kps) {
            const int nks = min(op.kps, op.ksteps - k0);
            mbar_wait(&s.full[rs.stage], rs.phase);
            tc_fence_after();
            const uint32_t b_hi = smem_u32(s.ring + rs.stage * STAGE_FLOATS);
            const uint32_t b_lo = b_hi + (uint32_t)nks * kstep_bytes;
            for (int ks = 0; ks < nks; ++ks) {
              const uint32_t col = (uint32_t)op.a_col + 8u * (uint32_t)(k0 + ks);
              const uint64_t dh = umma_desc_kmajor(b_hi + ks * kstep_bytes, (uint32_t)op.N * 16u, 128u);
              const uint64_t dl = umma_desc_kmajor(b_lo + ks * kstep_bytes, (uint32_t)op.N * 16u, 128u);
              umma_tf32_ts(tmem + TM_D, tmem + TM_AHI + col, dh, idesc, acc);   // a_hi * b_hi
              umma_tf32_ts(tmem + TM_D, tmem + TM_ALO + col, dh, idesc, 1u);    // a_lo * b_hi
              umma_tf32_ts(tmem + TM_D, tmem + TM_AHI + col, dl, idesc, 1u);    // a_hi * b_lo
              acc = 1u;
            }
            umma_commit(&s.empty[rs.stage]);  // the stage is free once these MMAs have read it
            rs.advance();
          }
          umma_commit(s.d_full);  // accumulator complete -> epilogue
        }
    }
  } else {
    // ===================== epilogue warps ====================================================
    const int q = warp & 3, h = warp >> 2;     // TMEM lane quarter, column half
    const int m = q * 32 + lane;               // row of the tile owned by this thread
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const float* __restrict__ bias_of[6] = {s.bias, s.bias + H, s.bias + 2 * H, s.bias + 3 * H, s.bias + 4 * H, s.bias + 5 * H};

    // biases and cond0 weights to shared memory (broadcast reads in the epilogues)
    for (int i = tid; i < H; i += TC_EPI_THREADS) {
      s.bias[0 * H + i] = pk[lo.q_b[L_COND0] + i];
      s.bias[1 * H + i] = pk[lo.q_b[L_COND1] + i];
      s.bias[2 * H + i] = pk[lo.q_b[L_DEC0] + i];
      s.bias[3 * H + i] = pk[lo.q_b[L_DEC1] + i];
      s.bias[4 * H + i] = pk[lo.q_b[L_DEC2] + i];
      s.bias[5 * H + i] = i < lo.Ip ? pk[lo.q_b[L_DEC3] + i] : 0.f;
      s.w0[i] = pk[lo.q_w[L_COND0] + i];
      s.w0[H + i] = pk[lo.q_w[L_COND0] + H + i];
    }
    float sx_sh = 0.f, sy_sh = 0.f;
    if (shared_start) {
      // one start point for the whole launch: condition encoder once per CTA, folded into
      // the bias of dec0:  hb[n] = b_dec0[n] + sum_k Wdec0[n][L+k] * h_c[k]
      sx_sh = a.start[0];
      sy_sh = a.start[1];
      if (tid < H) {
        const float* w0 = pk + lo.q_w[L_COND0];
        float v = pk[lo.q_b[L_COND0] + tid];
        v = fmaf(w0[tid], sx_sh, v);
        v = fmaf(w0[H + tid], sy_sh, v);
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
        s.hb[tid] = v;
      }
    }
    epi_sync();
    if (shared_start) bias_of[2] = s.hb;

    // writes the layer input of a tile: latents (both modes) and cond0 (per-row start) -> A
    float sx = sx_sh, sy = sy_sh;   // start point of the row this thread owns, for the tile being STAGED
    auto stage_tile = [&](long long tile) {
      const long long m0 = tile * TC_M;
      const bool row_ok = m0 + m < a.B;
      const long long row = m0 + m;
      for (int jb = h; jb < Lp8 / 4; jb += 2) {
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
        sx = row_ok ? __ldg(a.start + row * 2) : 0.f;
        sy = row_ok ? __ldg(a.start + row * 2 + 1) : 0.f;
        // cond0 (Training_VAE.py:132-133): relu(W0 [x0, y0] + b0), 64 of the 128 features per thread
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t hi[16], lw[16];
          const int n0 = h * 64 + c * 16;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float v = s.bias[n0 + j];
            v = fmaf(s.w0[n0 + j], sx, v);
            v = fmaf(s.w0[H + n0 + j], sy, v);
            split_tf32(fmaxf(v, 0.f), hi[j], lw[j]);
          }
          tmem_st16(lane_base + TM_AHI + n0, hi);
          tmem_st16(lane_base + TM_ALO + n0, lw);
        }
      }
    };
    auto publish = [&]() {  // A written / D read by this warp -> MMA warp
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s.a_ready);
    };

    uint32_t d_phase = 0;
    int out_buf = 0;
    if ((long long)blockIdx.x < n_tiles) {
      stage_tile(blockIdx.x);
      publish();
    }
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long m0 = tile * TC_M;
      const int valid = (int)min((long long)TC_M, a.B - m0);
      const float out_sx = sx, out_sy = sy;  // start point of THIS tile's row (stage_tile overwrites sx, sy)
      // ---- hidden layers: D -> relu(D + bias) -> A ------------------------------------------
      for (int o = 0; o < n_ops - 1; ++o) {
        const float* __restrict__ bias = bias_of[shared_start ? o + 2 : o + 1];
        mbar_wait(s.d_full, d_phase);
        d_phase ^= 1u;
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int n0 = h * 64 + c * 16;
          uint32_t v[16], hi[16], lw[16];
          tmem_ld16(lane_base + TM_D + n0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) split_tf32(fmaxf(__uint_as_float(v[j]) + bias[n0 + j], 0.f), hi[j], lw[j]);
          tmem_st16(lane_base + TM_AHI + n0, hi);
          tmem_st16(lane_base + TM_ALO + n0, lw);
        }
        publish();
      }
      // ---- last layer: D (Ip columns) -> registers; release D; stage the next tile; store ----
      mbar_wait(s.d_full, d_phase);
      d_phase ^= 1u;
      tc_fence_after();
      const int half = lo.Ip >> 1;          // columns per column-half: 16, 32 or 64
      uint32_t o[4][16];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c * 16 < half) tmem_ld16(lane_base + TM_D + h * half + c * 16, o[c]);
      tmem_ld_wait();
      const long long next = tile + gridDim.x;
      if (next < n_tiles) stage_tile(next);
      publish();  // also for the last tile: keeps the arrival count per phase uniform (nobody waits on it)

      float* stage_out = s.outs + (size_t)out_buf * TC_M * I;
      if (a.out_bufs > 1) {
        if (tid == 0) tma_store_wait_read<1>();  // the store that last read this buffer has drained it
      } else {
        if (tid == 0) tma_store_wait_read<0>();
      }
      epi_sync();
      if (m < valid) {
        const float* __restrict__ b3 = bias_of[5];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c * 16 < half) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int n = h * half + c * 16 + j;
              if (n < I) {
                float val = __uint_as_float(o[c][j]) + b3[n];
                if (a.add_start) {
                  const int d = n % 3;
                  if (d == 1) val = out_sx + val;
                  else if (d == 2) val = out_sy + val;
                }
                stage_out[m * I + n] = val;
              }
            }
          }
      }
      const uint32_t bytes = (uint32_t)valid * (uint32_t)I * 4u;
      float* gdst = a.out + m0 * I;
      if (a.bulk_ok && (bytes & 15u) == 0) {
        fence_proxy_async_smem();
        epi_sync();
        if (tid == 0) tma_store_1d(gdst, stage_out, bytes);
      } else {
        epi_sync();
        for (int idx = tid; idx < valid * I; idx += TC_EPI_THREADS) gdst[idx] = stage_out[idx];
        epi_sync();  // stage_out may be rewritten two tiles from now (or next tile with one buffer)
      }
      if (a.out_bufs > 1) out_buf ^= 1;
    }
    if (tid == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_EPI_WARPS) tmem_dealloc(tmem, TM_COLS);
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
  cudaError_t e = cudaFuncSetAttribute(decode_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  decode_tc_kernel<<<grid, TC_THREADS, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace dmvae
