// dmvae_train_tc.cu - the training pass on the 5th-generation tensor cores (tcgen05, 3xTF32).
//
// Replaces, per step, the reference's  offset transform -> model(...) -> conditional_vae_loss(...)
// -> loss.backward()  (Training_VAE.py:345-362; model :180-226, loss :229-268) with two kernels:
//
//   chain_kernel   one persistent CTA per SM walks 128-row tiles.  The activations of a tile
//                  never leave tensor memory: forward (cond0, cond1, enc0-3, heads, reparameterise,
//                  dec0-3), the five loss terms, and the whole data-gradient chain back to the
//                  first layers run as  D(TMEM) = A(TMEM) x B(smem)  with the layer input / the
//                  pre-activation gradient resident as the A operand (TF32 high and low halves)
//                  and the weights streamed L2 -> smem by TMA.  One weight image serves both
//                  directions: K-major it is W (forward), MN-major the same bytes are W^T (data
//                  gradient), one 8-column output slice per K step.  relu' masks stay in registers
//                  between the forward and the backward half of a tile.  Every layer input X and
//                  every pre-activation gradient G is also written once, as a raw fp32 operand
//                  image, to the tile's stash in global memory.
//   wgrad_kernel   the weight gradients  dW = G^T X  summed over all rows: a streaming GEMM whose
//                  contraction runs over the batch rows.  CTAs are specialised by role (a group of
//                  layers whose accumulators fit the 512 tensor-memory columns together) and by
//                  tile subset; X and G stream through a TMA ring, are split into TF32 halves in
//                  shared memory, and both operands are read MN-major.  Bias gradients ride along
//                  as products with a constant ones operand.  Accumulators stay in tensor memory
//                  across all tiles of a CTA and are written once, to the CTA's partial slab.
//   reduce_tc_kernel  fixed-order sum of the partial slabs (+ the loss partials), optionally with
//                  the Adam update (torch optim/adam.py::_single_tensor_adam) in the same thread.
//
// Envelope of this path: latent_dim <= 32 at any seq_len (the reference uses 8 and 10/12).  Up to 3*seq_len = 64 the
// trajectory is one operand; beyond that the <true> instantiations of the chain and weight-gradient bodies walk the
// first encoder / last decoder layer in chunks of 128 features.  latent_dim 33..64 runs the FFMA kernel of dmvae_train.cu.
#include "dmvae_common.cuh"
#include "dmvae_launch.h"
#include "dmvae_pack.cuh"
#include "dmvae_tc.cuh"

#include <type_traits>

namespace dmvae {

// 256-bit global accesses (one full 32-byte sector per lane)
__device__ __forceinline__ void st_global_v8(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_v8(const float* p, float (&v)[8]) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p)
               : "memory");
}

__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// =========================================================================================
// chain kernel
// =========================================================================================
constexpr int CH_EPI_WARPS = 8;
constexpr int CH_EPI_THREADS = CH_EPI_WARPS * 32;
constexpr int CH_CP = CH_EPI_WARPS / 4;            // warps per tensor-memory lane quarter: they share the columns of a phase
constexpr int CH_CPT = 32 / CH_CP;                 // columns per thread and 32-column phase
constexpr int CH_MW = 4 * CH_CPT / 32;             // relu' mask words per thread and layer
constexpr int CH_PRODUCER_WARP = CH_EPI_WARPS;
constexpr int CH_MMA_WARP = CH_EPI_WARPS + 1;
constexpr int CH_SIGNAL_WARP = CH_EPI_WARPS + 2;   // publishes the per-tile progress counter of the fused launch
constexpr int CH_THREADS = (CH_SIGNAL_WARP + 1) * 32;
constexpr int CH_M = 128;
constexpr int CH_MAX_OPS = 56;            // 24 + 3 per extra 128-feature chunk of a long trajectory (at most 10 chunks)
// Tensor-memory columns: TWO accumulators and the A operand (TF32 high and low halves, 128 columns each).  Consecutive
// 128 x 128 layers alternate between the accumulators, so the products of layer l + 1 (into the other accumulator) run
// while the epilogue warps still read layer l's.  The small operands and accumulators of the odd-shaped steps (heads,
// z, the start point, d/dz, d/d(mu, logvar)) live in whichever accumulator is idle at that point of the program:
constexpr uint32_t CT_D0 = 0, CT_D1 = 128, CT_AHI = 256, CT_ALO = 384, CT_COLS = 512;
constexpr uint32_t CT_HEADS = CT_D0;                                  // forward: (mu, logvar) accumulator, NH <= 64 columns
constexpr uint32_t CT_START_HI = CT_D0 + 64, CT_START_LO = CT_D0 + 72; // forward: [x0, y0, 1, 0 x 5], the A operand of cond0
constexpr uint32_t CT_Z_HI = CT_D0 + 64, CT_Z_LO = CT_D0 + 96;         // forward: z (Lp16 <= 32), the A operand of dec0's z rows
constexpr uint32_t CT_DZ = CT_D0;                                     // backward: d/dz accumulator (32 columns)
constexpr uint32_t CT_GML_HI = CT_D0, CT_GML_LO = CT_D0 + 64;          // backward: d/d(mu, logvar) (NH <= 64), A operand of the heads' data gradients

struct COp {
  int off_hi, off_lo;  // planes of the layer image in the packed arena (floats): forward or data-gradient image
  int N;               // output width of the layer (forward N; the contraction length of its data gradient)
  int k0, nk;          // forward: K-step range of the image; data gradient: slice range (32 outputs each)
  int kps;             // forward: K steps per ring stage; data gradient: contraction steps per group
  int S;               // data gradient: slices in the whole image
  int dgrad;           // 0 forward (B read K-major); 1 data gradient (B read MN-major)
  int gps;             // data gradient: groups per ring stage (a small image travels as ONE stage)
  int a_hi, a_lo;      // tensor-memory columns of the first contraction step of the A operand (TF32 high / low halves)
  int d_col;           // tensor-memory column of the accumulator (dgrad: of the first slice)
  int acc;             // accumulate onto what the accumulator already holds
  int wait_a;          // wait for the epilogue warps before issuing (A operand written, D drained)
  int commit_d;        // signal the epilogue warps when the accumulator is complete
  int pipe;            // 128 x 128 layer behind a four-phase epilogue, writing the accumulator that epilogue does NOT read:
                       // its ring stage k (a 32-deep quarter of the contraction) is issued as soon as the epilogue has
                       // written quarter k of the A operand (a_ready[k]) instead of after the whole epilogue
};

struct ChainArgs {
  const float* packed;
  const float* x;     // (B, T, 3) absolute; with x_batches > 0: a resident set of x_batches such batches, back to back
  long long x_batches;   // resident set: the pass reads batch (*step_dev mod x_batches) - no per-step copy or host work
  int x_shuffle;         // resident set: position p of epoch e = *step_dev / x_batches reads row resident_row(seed, e, p, rows)
  unsigned long long x_shuffle_seed;
  const float* eps;   // (B, L) or null (Philox)
  float* stash;       // [n_tiles][tile_stash]
  float* loss_part;   // [grid][epilogue warps][4 terms]
  unsigned long long seed, sample_offset, step;
  const long long* step_dev;   // when set: the step index is *step_dev + 1 (graph-capturable launches)
  long long B;
  float w_recon, w_kld, w_start, w_time, inv_batch;
  int stages;
  COp ops[CH_MAX_OPS];   // the per-tile GEMM program (chain_program / chain_program_long, built on the host)
  int n_ops;
  int n_chunks;          // 0: trajectories of up to 64 floats (3 T <= 64); else their 128-feature chunks (chain_program_long)
  int n_epis;            // epilogues per tile: CH_EPIS + 3 (n_chunks - 1)
  long long* trace;   // development aid: clock64 stamps of CTA 0, one of its tiles (null in production)
  int trace_tile;     // which of CTA 0's tiles (0 = its first)
  int* ready;         // when set: per-tile counter, + (epilogue warps) per epilogue, once the stash images of that
                      // epilogue are written (the weight-gradient CTAs of train_tc_fused_kernel wait on it)
};
struct ChainKArgs {
  Layout lo;
  ChainArgs c;
};

__host__ __device__ inline COp c_op(const TcLayer& c, int k0, int nk, int dgrad, int a_hi, int a_lo, int d_col, int acc,
                                    int wait_a, int commit_d, int pipe) {
  if (dgrad) {
    // groups of a small image (few slices) share one ring stage: one wait instead of one per group
    const int ngroups = (c.N >> 3) / c.gsz;
    const int group_bytes = 2 * nk * c.gsz * 1024;
    const int gps = (!pipe && ngroups * group_bytes <= STAGE_BYTES) ? ngroups : 1;
    return COp{c.off_thi, c.off_tlo, c.N, k0, nk, c.gsz, c.Kt / 32, 1, gps, a_hi, a_lo, d_col, acc, wait_a, commit_d, pipe};
  }
  return COp{c.off_hi, c.off_lo, c.N, k0, nk, c.kps, 0, 0, 1, a_hi, a_lo, d_col, acc, wait_a, commit_d, pipe};
}

// The per-tile GEMM program, walked identically by the producer and the MMA warp.  Every op with
// commit_d is followed by exactly one epilogue of the tile body below (same order, c_epi).
//
// Forward order: the encoder first (its input x_rel is staged at the START of a tile, before any accumulator
// exists to wait for), its share of the heads kept aside, then the condition encoder, whose output h_c stays in the
// A operand for the three products that read it (its share of the heads, and the h_c rows of the first decoder
// layer, issued back to back); the reparameterisation runs on the CUDA cores UNDER the decoder product and only adds
// the z rows afterwards.  Backward: d/dz first, so that the reparameterisation backward runs under the decoder share of
// d h_c.  No bias travels through the tensor cores: every epilogue adds its own.  "pipe" marks the layers whose
// products overlap the epilogue before them (they write the other accumulator); cond1 and its data gradient cannot
// (the idle accumulator holds the heads' small operands at that point) and wait for the whole epilogue.
__host__ __device__ inline int chain_program(const Layout& lo, COp* ops) {
  const int zs = lo.Lp16 / 8;
  const int A = (int)CT_AHI, Al = (int)CT_ALO, D0 = (int)CT_D0, D1 = (int)CT_D1;
  int n = 0;
  auto fwd = [&](int t, int k0, int nk, int a_hi, int a_lo, int d_col, int acc, int wait_a, int commit_d, int pipe) {
    ops[n++] = c_op(lo.tc[t], k0, nk, 0, a_hi, a_lo, d_col, acc, wait_a, commit_d, pipe);
  };
  auto bwd = [&](int t, int k0, int nk, int a_hi, int a_lo, int d_col, int acc, int wait_a, int commit_d, int pipe) {
    ops[n++] = c_op(lo.tc[t], k0, nk, 1, a_hi, a_lo, d_col, acc, wait_a, commit_d, pipe);
  };
  fwd(TC_ENC0, 0, lo.Ip / 8, A, Al, D0, 0, 1, 1, 0);                       // x_rel -> e1
  fwd(TC_ENC1, 0, 16, A, Al, D1, 0, 1, 1, 1);
  fwd(TC_ENC2, 0, 16, A, Al, D0, 0, 1, 1, 1);
  fwd(TC_ENC3, 0, 16, A, Al, D1, 0, 1, 1, 1);                              // -> e4
  fwd(TC_HEADS, 0, 16, A, Al, (int)CT_HEADS, 0, 1, 0, 0);                  // h_traj share of the heads (kept aside, no epilogue)
  fwd(TC_COND0, 0, 1, (int)CT_START_HI, (int)CT_START_LO, D1, 0, 0, 1, 0); // start -> hc1 (bias row inside its K = 8)
  fwd(TC_COND1, 0, 16, A, Al, D1, 0, 1, 1, 0);                             // hc1 -> hc
  fwd(TC_HEADS, 16, 16, A, Al, (int)CT_HEADS, 1, 1, 1, 0);                 // + hc share -> mu, logvar: the reparameterisation epilogue
  fwd(TC_DEC0, 0, 16, A, Al, D1, 0, 0, 0, 0);                              // hc rows of dec0, running under that epilogue
  fwd(TC_DEC0, 16, zs, (int)CT_Z_HI, (int)CT_Z_LO, D1, 1, 1, 1, 0);        // + z rows -> d1
  fwd(TC_DEC1, 0, 16, A, Al, D0, 0, 1, 1, 1);
  fwd(TC_DEC2, 0, 16, A, Al, D1, 0, 1, 1, 1);
  fwd(TC_DEC3, 0, 16, A, Al, D0, 0, 1, 1, 0);                              // -> recon
  bwd(TC_DEC3, 0, 4, A, Al, D1, 0, 1, 1, 0);                               // d recon -> d d3 (4 slices of 32 columns)
  bwd(TC_DEC2, 0, 4, A, Al, D0, 0, 1, 1, 1);
  bwd(TC_DEC1, 0, 4, A, Al, D1, 0, 1, 1, 1);
  bwd(TC_DEC0, 4, 1, A, Al, (int)CT_DZ, 0, 1, 1, 0);                       // d d1 -> d z (one slice): the reparameterisation backward
  bwd(TC_DEC0, 0, 4, A, Al, D1, 0, 0, 0, 0);                               //      -> d hc (decoder share), under that epilogue
  bwd(TC_HEADS, 4, 4, (int)CT_GML_HI, (int)CT_GML_LO, D1, 1, 1, 1, 0);     // d (mu, logvar) -> + encoder share of d hc
  bwd(TC_COND1, 0, 4, A, Al, D1, 0, 1, 1, 0);                              // d hc -> d hc1
  bwd(TC_HEADS, 0, 4, (int)CT_GML_HI, (int)CT_GML_LO, D1, 0, 1, 1, 0);     // d (mu, logvar) -> d e4
  bwd(TC_ENC3, 0, 4, A, Al, D0, 0, 1, 1, 1);
  bwd(TC_ENC2, 0, 4, A, Al, D1, 0, 1, 1, 1);
  bwd(TC_ENC1, 0, 4, A, Al, D0, 0, 1, 1, 1);                               // -> d e1
  return n;
}

// Trajectories of more than 64 floats (3 T > 64): the first encoder layer contracts over NC chunks of 128 features and
// the last decoder layer produces NC chunks of 128 outputs - each chunk a 128 x 128 product of its own (dec3_chunk_layer),
// with the A operand re-staged between chunks by an epilogue of its own:
//   enc0 chunk c     D0 (+)= x_rel[:, chunk c] W0[chunk c]      then: stage x_rel chunk c + 1 (last chunk: the usual epilogue)
//   dec3 chunk c     D0   = d3 W3[chunk c]                      then: the loss over these 128 features (carrying the time
//                                                                     walk from chunk to chunk); the gradient goes to the stash
//   b_dec3 chunk c   D1 (+)= g_rec[:, chunk c] W3[chunk c]^T    then: stage g_rec chunk c + 1 from the stash (last: usual)
// Everything between is the program of chain_program.
__host__ __device__ inline int chain_program_long(const Layout& lo, COp* ops) {
  const int zs = lo.Lp16 / 8, NC = lo.NC;
  const int A = (int)CT_AHI, Al = (int)CT_ALO, D0 = (int)CT_D0, D1 = (int)CT_D1;
  int n = 0;
  auto fwd = [&](const TcLayer& c, int k0, int nk, int a_hi, int a_lo, int d_col, int acc, int wait_a, int commit_d, int pipe) {
    ops[n++] = c_op(c, k0, nk, 0, a_hi, a_lo, d_col, acc, wait_a, commit_d, pipe);
  };
  auto bwd = [&](const TcLayer& c, int k0, int nk, int a_hi, int a_lo, int d_col, int acc, int wait_a, int commit_d, int pipe) {
    ops[n++] = c_op(c, k0, nk, 1, a_hi, a_lo, d_col, acc, wait_a, commit_d, pipe);
  };
  for (int c = 0; c < NC; ++c) fwd(lo.tc[TC_ENC0], 16 * c, 16, A, Al, D0, c > 0, 1, 1, 0);
  fwd(lo.tc[TC_ENC1], 0, 16, A, Al, D1, 0, 1, 1, 1);
  fwd(lo.tc[TC_ENC2], 0, 16, A, Al, D0, 0, 1, 1, 1);
  fwd(lo.tc[TC_ENC3], 0, 16, A, Al, D1, 0, 1, 1, 1);
  fwd(lo.tc[TC_HEADS], 0, 16, A, Al, (int)CT_HEADS, 0, 1, 0, 0);
  fwd(lo.tc[TC_COND0], 0, 1, (int)CT_START_HI, (int)CT_START_LO, D1, 0, 0, 1, 0);
  fwd(lo.tc[TC_COND1], 0, 16, A, Al, D1, 0, 1, 1, 0);
  fwd(lo.tc[TC_HEADS], 16, 16, A, Al, (int)CT_HEADS, 1, 1, 1, 0);
  fwd(lo.tc[TC_DEC0], 0, 16, A, Al, D1, 0, 0, 0, 0);
  fwd(lo.tc[TC_DEC0], 16, zs, (int)CT_Z_HI, (int)CT_Z_LO, D1, 1, 1, 1, 0);
  fwd(lo.tc[TC_DEC1], 0, 16, A, Al, D0, 0, 1, 1, 1);
  fwd(lo.tc[TC_DEC2], 0, 16, A, Al, D1, 0, 1, 1, 1);
  for (int c = 0; c < NC; ++c) fwd(dec3_chunk_layer(lo, c), 0, 16, A, Al, D0, 0, 1, 1, 0);
  for (int c = 0; c < NC; ++c) bwd(dec3_chunk_layer(lo, c), 0, 4, A, Al, D1, c > 0, 1, 1, 0);
  bwd(lo.tc[TC_DEC2], 0, 4, A, Al, D0, 0, 1, 1, 1);
  bwd(lo.tc[TC_DEC1], 0, 4, A, Al, D1, 0, 1, 1, 1);
  bwd(lo.tc[TC_DEC0], 4, 1, A, Al, (int)CT_DZ, 0, 1, 1, 0);
  bwd(lo.tc[TC_DEC0], 0, 4, A, Al, D1, 0, 0, 0, 0);
  bwd(lo.tc[TC_HEADS], 4, 4, (int)CT_GML_HI, (int)CT_GML_LO, D1, 1, 1, 1, 0);
  bwd(lo.tc[TC_COND1], 0, 4, A, Al, D1, 0, 1, 1, 0);
  bwd(lo.tc[TC_HEADS], 0, 4, (int)CT_GML_HI, (int)CT_GML_LO, D1, 0, 1, 1, 0);
  bwd(lo.tc[TC_ENC3], 0, 4, A, Al, D0, 0, 1, 1, 1);
  bwd(lo.tc[TC_ENC2], 0, 4, A, Al, D1, 0, 1, 1, 1);
  bwd(lo.tc[TC_ENC1], 0, 4, A, Al, D0, 0, 1, 1, 1);
  return n;
}
// Epilogue e of a tile with nc chunks (nc = 1: the list c_epi itself) -> its row of c_epi, or one of the chunk
// epilogues: EP_STAGE_X / EP_LOSS / EP_STAGE_G with the chunk they work on
enum EpiType { EP_HIDDEN = 0, EP_HEADS, EP_LOSS, EP_DGRAD, EP_BDEC0, EP_STAGE_X, EP_STAGE_G };
struct EpiRef { int type, row, chunk; };
__host__ __device__ inline EpiRef epi_at(int e, int nc) {
  if (e < nc - 1) return EpiRef{EP_STAGE_X, -1, e + 1};
  if (e < nc + 9) return EpiRef{-1, e - (nc - 1), 0};                    // rows 0..9: E1 .. D3
  if (e < 2 * nc + 9) return EpiRef{EP_LOSS, 10, e - (nc + 9)};
  if (e < 3 * nc + 8) return EpiRef{EP_STAGE_G, -1, e - (2 * nc + 9) + 1};
  return EpiRef{-1, 11 + e - (3 * nc + 8), 0};                            // rows 11..20: the data gradients
}

// relu' mask slots (CH_MW words per epilogue thread and slot, in shared memory)
enum MaskSlot { MK_HC1 = 0, MK_HC, MK_E1, MK_E2, MK_E3, MK_E4, MK_D1, MK_D2, MK_D3, MK_COUNT };
// The epilogues of a tile, in the order of the ops that signal them (chain_program).  The tile body is a
// loop over this table rather than 21 inlined epilogues: the code stays small.
constexpr int CH_EPIS = 21;
constexpr int CH_EPI_FIRST_DGRAD = 11;
__constant__ int c_epi[CH_EPIS][7] = {
    // type, mask slot, stash slot (the image the epilogue completes), write the A operand, layer whose bias the
    // epilogue adds (-1: none - cond0 carries its bias row inside its K = 8), accumulator it reads, 1: it also writes
    // the start-point columns (the A operand of cond0, which follows the encoder)
    {EP_HIDDEN, MK_E1, SX_E1, 1, L_ENC0, CT_D0, 0},   {EP_HIDDEN, MK_E2, SX_E2, 1, L_ENC1, CT_D1, 0},
    {EP_HIDDEN, MK_E3, SX_E3, 1, L_ENC2, CT_D0, 0},   {EP_HIDDEN, MK_E4, SX_E4, 1, L_ENC3, CT_D1, 1},
    {EP_HIDDEN, MK_HC1, SX_HC1, 1, -1, CT_D1, 0},     {EP_HIDDEN, MK_HC, SX_HC, 1, L_COND1, CT_D1, 0},
    {EP_HEADS, 0, SX_Z, 0, -1, CT_HEADS, 0},          {EP_HIDDEN, MK_D1, SX_D1, 1, L_DEC0, CT_D1, 0},
    {EP_HIDDEN, MK_D2, SX_D2, 1, L_DEC1, CT_D0, 0},   {EP_HIDDEN, MK_D3, SX_D3, 1, L_DEC2, CT_D1, 0},
    {EP_LOSS, 0, SG_REC, 0, -1, CT_D0, 0},
    {EP_DGRAD, MK_D3, SG_D3, 1, -1, CT_D1, 0},        {EP_DGRAD, MK_D2, SG_D2, 1, -1, CT_D0, 0},
    {EP_DGRAD, MK_D1, SG_D1, 1, -1, CT_D1, 0},        {EP_BDEC0, 0, SG_ML, 0, -1, CT_DZ, 0},
    {EP_DGRAD, MK_HC, SG_HC, 1, -1, CT_D1, 0},        {EP_DGRAD, MK_HC1, SG_HC1, 0, -1, CT_D1, 0},
    {EP_DGRAD, MK_E4, SG_E4, 1, -1, CT_D1, 0},        {EP_DGRAD, MK_E3, SG_E3, 1, -1, CT_D0, 0},
    {EP_DGRAD, MK_E2, SG_E2, 1, -1, CT_D1, 0},        {EP_DGRAD, MK_E1, SG_E1, 0, -1, CT_D0, 0},
};

// trajectories of more than 64 floats: chunks of 128 features (chain_program_long); nothing of x or recon is kept in
// shared memory then
__host__ __device__ inline bool chain_long(const Layout& lo) { return lo.Ip > 64; }
__host__ __device__ inline size_t chain_scratch_floats(const Layout& lo) { return (lo.Ip == 32 || chain_long(lo)) ? 0 : (size_t)lo.Ip * 128; }
__host__ __device__ inline size_t chain_x_floats(const Layout& lo) { return chain_long(lo) ? 0 : (size_t)round_up(128 * lo.I, 4); }
__host__ __device__ inline size_t chain_smem_floats(const Layout& lo, int stages) {
  return (size_t)stages * STAGE_FLOATS + chain_scratch_floats(lo) /* recon scratch of the wide-row loss */ + chain_x_floats(lo) /* x tile */ +
         (size_t)lo.NH * 128 /* mu, logvar */ + (size_t)lo.Lp16 * 128 /* eps */ + MK_COUNT * CH_MW * CH_EPI_THREADS /* masks */ +
         NUM_LAYERS * 128 /* biases */;
}
__host__ __device__ inline size_t chain_smem_bytes(const Layout& lo, int stages) {
  return chain_smem_floats(lo, stages) * 4 + 32 * 8 + 16 + 1024;
}

template <int N> struct TmemVec;
template <> struct TmemVec<16> {
  __device__ static __forceinline__ void ld(uint32_t t, uint32_t (&v)[16]) { tmem_ld16(t, v); }
  __device__ static __forceinline__ void st(uint32_t t, const uint32_t (&v)[16]) { tmem_st16(t, v); }
};
template <> struct TmemVec<8> {
  __device__ static __forceinline__ void ld(uint32_t t, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(t)
                 : "memory");
  }
  __device__ static __forceinline__ void st(uint32_t t, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(t), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
  }
};

// cta / ncta: index of this CTA among the chain CTAs and their number (the whole grid for chain_kernel)
// LNG: trajectories of more than 64 floats (chain_program_long); a compile-time switch, so that the short-trajectory
// instantiation carries none of the chunk walks (as a run-time flag they cost the batch-4096 step 5.6 us)
template <bool LNG>
__device__ __forceinline__ void chain_body(const Layout& lo, const ChainArgs& a, const int cta, const int ncta,
                                           unsigned char* smem_dyn) {
  const int L = lo.L, I = lo.I, T = lo.T, Ip = lo.Ip, NH = lo.NH, Lp16 = lo.Lp16;
  float *ring, *scratch, *xbuf, *mlb, *epb, *bias_s;
  uint32_t* masks;
  const COp* ops = a.ops;   // kernel parameter space: uniform reads by the producer and the MMA warp
  // d_ready: the accumulator of the op that commits is complete - and so is every product issued before it, so its A
  // operand may be overwritten (MMA -> epilogue); a_ready[k]: quarter k of the A operand is written and quarter k
  // of the accumulator has been read (epilogue -> MMA); stash_done: every epilogue warp has issued the stash stores
  // of the current epilogue (epilogue -> signal warp)
  // x_full: the tile's trajectories (and, for the first tile, the biases) have landed in shared memory (TMA -> epilogue)
  uint64_t *full, *empty, *d_ready, *a_ready, *stash_done, *x_full;
  uint32_t* tmem_slot;
  {
    const uint32_t base = smem_u32(smem_dyn);
    unsigned char* p = smem_dyn + ((1024u - (base & 1023u)) & 1023u);
    ring = reinterpret_cast<float*>(p);
    scratch = ring + (size_t)a.stages * STAGE_FLOATS;
    xbuf = scratch + chain_scratch_floats(lo);
    mlb = xbuf + chain_x_floats(lo);
    epb = mlb + (size_t)NH * 128;
    masks = reinterpret_cast<uint32_t*>(epb + (size_t)Lp16 * 128);
    bias_s = reinterpret_cast<float*>(masks + MK_COUNT * CH_MW * CH_EPI_THREADS);
    full = reinterpret_cast<uint64_t*>(bias_s + NUM_LAYERS * 128);
    empty = full + 8;
    d_ready = empty + 8;
    a_ready = d_ready + 1;
    stash_done = a_ready + 4;
    x_full = stash_done + 1;
    tmem_slot = reinterpret_cast<uint32_t*>(x_full + 1);
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* __restrict__ pk = a.packed;
  const long long n_tiles = (a.B + CH_M - 1) / CH_M;
  constexpr bool lng = LNG;                 // trajectories of more than 64 floats, walked in chunks of 128 features
  const int NCK = LNG ? a.n_chunks : 1;

  if (tid == 0) {
    if (a.trace != nullptr && cta == 0) a.trace[176] = global_ns();
    for (int st = 0; st < a.stages; ++st) {
      mbar_init(&full[st], 1);
      mbar_init(&empty[st], 1);
    }
    mbar_init(d_ready, 1);
    for (int k = 0; k < 4; ++k) mbar_init(&a_ready[k], CH_EPI_WARPS);
    mbar_init(stash_done, CH_EPI_WARPS);
    mbar_init(x_full, 1);
    mbar_fence_init();
  }
  if (warp == CH_PRODUCER_WARP) tmem_alloc(tmem_slot, CT_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_ops = a.n_ops;

  if (warp == CH_PRODUCER_WARP) {
    // ===================== producer warp: weight planes L2 -> smem ring =======================
    if (lane == 0) {
      RingStateRt rs(a.stages);
      for (long long tile = cta; tile < n_tiles; tile += ncta)
        for (int o = 0; o < n_ops; ++o) {
          const COp op = ops[o];
          if (!op.dgrad) {
            for (int k = 0; k < op.nk; k += op.kps) {  // nks K steps of N*8 floats, contiguous in a plane
              const int nks = min(op.kps, op.nk - k);
              const int fl = nks * op.N * 8;
              const size_t src = (size_t)(op.k0 + k) * op.N * 8;
              float* dst = ring + rs.stage * STAGE_FLOATS;
              mbar_wait(&empty[rs.stage], rs.phase ^ 1u);
              mbar_arrive_expect_tx(&full[rs.stage], (uint32_t)(2 * fl * 4));
              tma_load_1d(dst, pk + op.off_hi + src, (uint32_t)(fl * 4), &full[rs.stage]);
              tma_load_1d(dst + fl, pk + op.off_lo + src, (uint32_t)(fl * 4), &full[rs.stage]);
              rs.advance();
            }
          } else {
            const int ngroups = (op.N >> 3) / op.kps;
            const int fl = op.nk * op.kps * 256;       // nk slices x gsz steps x 1 KB, contiguous inside a group
            for (int jg = 0; jg < ngroups; jg += op.gps) {
              float* dst = ring + rs.stage * STAGE_FLOATS;
              mbar_wait(&empty[rs.stage], rs.phase ^ 1u);
              mbar_arrive_expect_tx(&full[rs.stage], (uint32_t)(op.gps * 2 * fl * 4));
              for (int g = 0; g < op.gps; ++g) {       // group g of the stage: [high plane][low plane]
                const size_t src = (size_t)(((jg + g) * op.S + op.k0) * op.kps) * 256;
                tma_load_1d(dst + (size_t)g * 2 * fl, pk + op.off_hi + src, (uint32_t)(fl * 4), &full[rs.stage]);
                tma_load_1d(dst + (size_t)g * 2 * fl + fl, pk + op.off_lo + src, (uint32_t)(fl * 4), &full[rs.stage]);
              }
              rs.advance();
            }
          }
        }
    }
  } else if (warp == CH_MMA_WARP) {
    // ===================== MMA warp (warp-uniform walk, one elected lane issues) ===============
    RingStateRt rs(a.stages);
    uint32_t a_phase = 0;
    for (long long tile = cta; tile < n_tiles; tile += ncta)
      for (int o = 0; o < n_ops; ++o) {
        const COp op = ops[o];
        const bool tr = a.trace != nullptr && (!lng || NCK <= 2) && cta == 0 && tile == cta + (long long)a.trace_tile * ncta && lane == 0;
        if (op.wait_a && !op.pipe) {
          for (int k = 0; k < 4; ++k) mbar_wait(&a_ready[k], a_phase);
        }
        tc_fence_after();
        const uint32_t d = tmem + (uint32_t)op.d_col;
        if (!op.dgrad) {
          const uint32_t unit_bytes = (uint32_t)op.N * 32u;  // one K step of a plane
          // D[128 x N] (+)= A[:, 8 (k + ks) ..] x W^T, B K-major: LBO = chunk stride N*16, SBO = 128
          const uint32_t idesc = umma_idesc_tf32(CH_M, op.N);
          const uint64_t dbits = umma_desc(0u, (uint32_t)op.N * 16u, 128u);
          const uint64_t dinc = (uint64_t)(unit_bytes >> 4);
          int sidx = 0;
          for (int k = 0; k < op.nk; k += op.kps, ++sidx) {
            const int nks = min(op.kps, op.nk - k);
            if (op.pipe) {   // a pipelined layer has four stages: stage k reads quarter k of the A operand
              mbar_wait(&a_ready[sidx], a_phase);
              tc_fence_after();
            }
            mbar_wait(&full[rs.stage], rs.phase);
            tc_fence_after();
            if (tr && k == 0) { a.trace[o * 4 + 0] = clock64(); if (o == 0) a.trace[178] = global_ns(); }
            const uint32_t b_hi = smem_u32(ring + rs.stage * STAGE_FLOATS);
            const uint32_t b_lo = b_hi + (uint32_t)nks * unit_bytes;
            uint64_t dh = dbits | (uint64_t)(b_hi >> 4), dl = dbits | (uint64_t)(b_lo >> 4);
            uint32_t a_hi = tmem + (uint32_t)(op.a_hi + 8 * k), a_lo = tmem + (uint32_t)(op.a_lo + 8 * k);
            uint32_t accf = (op.acc || k > 0) ? 1u : 0u;
            if (elect_one()) {
#pragma unroll 4
              for (int ks = 0; ks < nks; ++ks) {
                umma_tf32_ts(d, a_hi, dh, idesc, accf);
                umma_tf32_ts(d, a_lo, dh, idesc, 1u);
                umma_tf32_ts(d, a_hi, dl, idesc, 1u);
                accf = 1u;
                a_hi += 8u; a_lo += 8u; dh += dinc; dl += dinc;
              }
              umma_commit(&empty[rs.stage]);
            }
            __syncwarp();
            rs.advance();
          }
        } else {
          // D[128 x 32 nk] (+)= G[128 x N] x W[:, those inputs]; B MN-major (32-byte swizzle): LBO = slice
          // stride (next 32 outputs) = gsz KB, SBO = 512 (next 4 of the contraction); step st of the group
          // starts st KB into a slice
          const int ngroups = (op.N >> 3) / op.kps;
          const uint32_t plane = (uint32_t)(op.nk * op.kps) * 1024u;
          const uint32_t idesc = umma_idesc_tf32(CH_M, 32 * op.nk, UMMA_B_MN);
          const uint64_t dbits = umma_desc(0u, (uint32_t)op.kps * 1024u, 512u, 1u);
          uint32_t accf = op.acc ? 1u : 0u;
          for (int jg = 0; jg < ngroups; jg += op.gps) {
            if (op.pipe) {   // four groups: group k reads quarter k of the A operand (gps == 1)
              mbar_wait(&a_ready[jg], a_phase);
              tc_fence_after();
            }
            mbar_wait(&full[rs.stage], rs.phase);
            tc_fence_after();
            if (tr && jg == 0) a.trace[o * 4 + 0] = clock64();
            const uint32_t s0 = smem_u32(ring + rs.stage * STAGE_FLOATS);
            if (elect_one()) {
              for (int g = 0; g < op.gps; ++g) {
                const uint32_t b_hi = s0 + (uint32_t)g * 2u * plane;
                uint64_t dh = dbits | (uint64_t)(b_hi >> 4), dl = dbits | (uint64_t)((b_hi + plane) >> 4);
                uint32_t a_hi = tmem + (uint32_t)(op.a_hi + 8 * (jg + g) * op.kps);
                uint32_t a_lo = tmem + (uint32_t)(op.a_lo + 8 * (jg + g) * op.kps);
                for (int st = 0; st < op.kps; ++st) {
                  umma_tf32_ts(d, a_hi, dh, idesc, accf);
                  umma_tf32_ts(d, a_lo, dh, idesc, 1u);
                  umma_tf32_ts(d, a_hi, dl, idesc, 1u);
                  accf = 1u;
                  a_hi += 8u; a_lo += 8u; dh += 64u; dl += 64u;  // + 1024 bytes
                }
              }
              umma_commit(&empty[rs.stage]);
            }
            __syncwarp();
            rs.advance();
          }
        }
        if (op.wait_a) a_phase ^= 1u;   // this op consumed the a_ready phase of the epilogue before it
        if (op.commit_d) {              // fires when everything issued so far is complete
          if (elect_one()) umma_commit(d_ready);
          __syncwarp();
        }
        if (tr) a.trace[o * 4 + 1] = clock64();
      }
  } else if (warp == CH_SIGNAL_WARP) {
    // ===================== signal warp (fused launch only) ======================================
    // Publishes the chain's progress to the weight-gradient CTAs: once all epilogue warps have issued the
    // stash stores of an epilogue (their arrival on stash_done releases those stores to this thread), one thread
    // makes them visible device-wide - also to the bulk copies (async proxy) of the other CTAs - and bumps the
    // tile's counter.  The device-scope fence is cumulative over the stores it has observed through the barrier;
    // keeping it in this warp takes it off the epilogue warps' critical path (it costs several hundred cycles).
    if (lane == 0 && a.ready != nullptr) {
      uint32_t phase = 0;
      for (long long tile = cta; tile < n_tiles; tile += ncta)
        for (int e = 0; e < a.n_epis; ++e) {
          mbar_wait(stash_done, phase);
          phase ^= 1u;
          fence_proxy_async_global();
          __threadfence();
          atomicAdd(a.ready + tile, CH_EPI_WARPS);
        }
    }
  } else {
    // ===================== epilogue warps =======================================================
    const int q = warp & 3, cp = warp >> 2;  // tensor-memory lane quarter, column part of a phase
    const int m = q * 32 + lane;             // row of the tile owned by this thread
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t d_phase = 0;
    uint32_t* my_mask = masks + tid;   // [(slot * CH_MW + word) * CH_EPI_THREADS + tid]
    float loss_acc[4] = {0.f, 0.f, 0.f, 0.f};
    float* my_ml = mlb + m;          // [n * 128 + m]
    float* my_ep = epb + m;

    int epi_no = 0;
    bool tr_tile = false;
    // ---- hand-shakes with the MMA warp.  Every op that commits completes d_ready once, and every epilogue completes
    // a_ready[0..3] once: one parity bit per direction.
    auto wait_d = [&]() {                  // the accumulator is complete; every earlier product too (A may be overwritten)
      mbar_wait(d_ready, d_phase);
      tc_fence_after();
      if (tr_tile && tid == 0) a.trace[128 + epi_no * 2] = clock64();
    };
    // the warp's stash stores of this epilogue are issued: tell the signal warp.  Always BEFORE the warp's last
    // arrival on a_ready of the same epilogue, so that no warp can be an epilogue ahead of a warp that has not
    // reported yet (the next epilogue starts only after all a_ready arrivals).
    auto report_stash = [&]() {
      if (a.ready != nullptr) {
        __syncwarp();
        if (lane == 0) mbar_arrive(stash_done);
      }
    };
    auto arrive_q = [&](int k) {           // this warp has read accumulator quarter k and written quarter k of A
      if (k == 3) report_stash();
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_ready[k]);
    };
    auto finish_epilogue = [&]() {
      d_phase ^= 1u;
      if (tr_tile && tid == 0) a.trace[128 + epi_no * 2 + 1] = clock64();
      ++epi_no;
    };
    auto arrive_all = [&](bool epilogue = true) {
      if (epilogue) report_stash();
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) mbar_arrive(&a_ready[k]);
      }
    };
    auto release_a = [&]() {
      arrive_all();
      finish_epilogue();
    };
    // a 4-feature chunk of this thread's row in a stash image (MN-major, 32-byte swizzle: dmvae_tc.cuh)
    auto stash_ptr = [&](float* tile_stash, int slot, int chunk) -> float4* {
      return reinterpret_cast<float4*>(tile_stash + lo.slot_off[slot] + mn_image_index(4 * chunk, m, 128, lo.slot_w[slot] * 4));
    };
    // the batch of this pass: the caller's, or the one the device-side step counter selects in a resident set
    const unsigned long long x_step = a.x_batches > 0 ? (unsigned long long)(*a.step_dev) : 0ull;
    const unsigned long long x_b = a.x_batches > 0 ? x_step % (unsigned long long)a.x_batches : 0ull;
    const unsigned long long x_epoch = a.x_batches > 0 ? x_step / (unsigned long long)a.x_batches : 0ull;
    const bool shuffled = a.x_batches > 0 && a.x_shuffle != 0;
    const float* __restrict__ x_batch = a.x + (shuffled ? 0 : (size_t)x_b * (size_t)a.B * I);
    // long trajectories: a thread reads its row of the tile straight from global memory (the tile does not fit shared
    // memory); null past the batch end
    const float* xrow = nullptr;
    auto set_row = [&](long long tile) {
      const long long row = tile * CH_M + m;
      xrow = nullptr;
      if (row < a.B) {
        if (shuffled) {
          const uint32_t src_row = resident_row(a.x_shuffle_seed, x_epoch, (uint32_t)(x_b * (unsigned long long)a.B + (unsigned long long)row),
                                                (uint32_t)(a.x_batches * a.B));
          xrow = a.x + (size_t)src_row * I;
        } else {
          xrow = x_batch + (size_t)row * I;
        }
      }
    };
    auto x_at = [&](int n) -> float { return lng ? (xrow != nullptr ? __ldg(xrow + n) : 0.f) : xbuf[m * I + n]; };
    // long trajectories: features [n, n + 4) of this thread's row (n a multiple of 4), zero beyond I or the batch end; one
    // 128-bit or two 64-bit loads where the row's address allows (a warp-wide scalar load of 32 rows occupies the load
    // unit as long as a 128-bit one: a quarter of the instructions)
    auto x4_at = [&](int n) -> float4 {
      if (xrow == nullptr || n >= I) return make_float4(0.f, 0.f, 0.f, 0.f);
      const float* p = xrow + n;
      if (n + 4 <= I) {
        const uintptr_t ad = reinterpret_cast<uintptr_t>(p);
        if ((ad & 15u) == 0) return __ldg(reinterpret_cast<const float4*>(p));
        if ((ad & 7u) == 0) {
          const float2 u = __ldg(reinterpret_cast<const float2*>(p)), w = __ldg(reinterpret_cast<const float2*>(p) + 1);
          return make_float4(u.x, u.y, w.x, w.y);
        }
        return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
      }
      return make_float4(__ldg(p), n + 1 < I ? __ldg(p + 1) : 0.f, n + 2 < I ? __ldg(p + 2) : 0.f, 0.f);
    };
    // start point of the tile's row: the 16-wide stash image ([x0, y0, 1, 0...]; the same columns become the A operand
    // of cond0 in the epilogue that precedes it, stage_start_cols)
    auto stage_start = [&](long long tile) {
      if (cp == 0) {
        const float sx = x_at(1), sy = x_at(2);   // zeros past the batch end
        float* ts = a.stash + (size_t)tile * lo.tile_stash;
        *stash_ptr(ts, SX_START, 0) = make_float4(sx, sy, 1.0f, 0.f);
#pragma unroll
        for (int c = 1; c < 8; ++c) *stash_ptr(ts, SX_START, c) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto stage_start_cols = [&]() {
      if (cp == 0) {
        const float sx = x_at(1), sy = x_at(2);
        uint32_t xh, xl, yh, yl;
        split_tf32(sx, xh, xl);
        split_tf32(sy, yh, yl);
        tmem_st4(lane_base + CT_START_HI, xh, yh, __float_as_uint(1.0f), 0u);
        tmem_st4(lane_base + CT_START_HI + 4, 0u, 0u, 0u, 0u);
        tmem_st4(lane_base + CT_START_LO, xl, yl, 0u, 0u);
        tmem_st4(lane_base + CT_START_LO + 4, 0u, 0u, 0u, 0u);
      }
    };
    // encoder input of the tile's row: x_rel = x - start on the x, y columns (Training_VAE.py:345-348), zero beyond I
    // -> the A operand of enc0 and the stash image.  Staged before the tile's first MMA; the warps of a lane quarter
    // share the 4-column chunks of a row.
    // chunk: long trajectories, the 128 features [128 chunk, 128 chunk + 128) and their own stash image
    auto stage_xrel = [&](long long tile, int chunk) {
      float* ts = a.stash + (size_t)tile * lo.tile_stash;
      const float sx = x_at(1), sy = x_at(2);
      const int width = lng ? 128 : Ip, n0 = chunk * 128;
      float* img = ts + lo.slot_off[SX_X] + (size_t)chunk * (128 * 128);
      if (lng) {
        // A thread reads its own row from global memory (__ldg is a volatile asm: loads stay in program order, and the
        // first USE of one stalls the thread for a full memory latency).  One tile of the trace: 36 k cycles per chunk
        // with the loads of a group issued right before their use - so the 16 loads of four groups are issued
        // together, then used.
#pragma unroll 1
        for (int c16 = 0; c16 < 128 / (16 * CH_CP) * CH_CP; c16 += CH_CP) {   // batches of four groups of this thread
          float xv[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c4 = (c16 * 4) + cp + q * CH_CP;
            const float4 g = x4_at(n0 + c4 * 4);
            xv[q * 4] = g.x; xv[q * 4 + 1] = g.y; xv[q * 4 + 2] = g.z; xv[q * 4 + 3] = g.w;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c4 = (c16 * 4) + cp + q * CH_CP;
            uint32_t hi[4], lw[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int n = n0 + c4 * 4 + i;
              float val = xv[q * 4 + i];
              if (n < I) {
                const int d = n % 3;
                if (d == 1) val = val - sx;
                else if (d == 2) val = val - sy;
              }
              xv[q * 4 + i] = val;
              split_tf32(val, hi[i], lw[i]);
            }
            tmem_st4(lane_base + CT_AHI + c4 * 4, hi[0], hi[1], hi[2], hi[3]);
            tmem_st4(lane_base + CT_ALO + c4 * 4, lw[0], lw[1], lw[2], lw[3]);
            *reinterpret_cast<float4*>(img + mn_image_index(4 * c4, m, 128, width * 4)) =
                make_float4(xv[q * 4], xv[q * 4 + 1], xv[q * 4 + 2], xv[q * 4 + 3]);
          }
        }
        return;
      }
#pragma unroll 2
      for (int c4 = cp; c4 < width / 4; c4 += CH_CP) {
        float xv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int n = n0 + c4 * 4 + i;
          float val = 0.f;
          if (n < I) {
            val = x_at(n);
            const int d = n % 3;
            if (d == 1) val = val - sx;
            else if (d == 2) val = val - sy;
          }
          xv[i] = val;
        }
        uint32_t hi[4], lw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(xv[i], hi[i], lw[i]);
        tmem_st4(lane_base + CT_AHI + c4 * 4, hi[0], hi[1], hi[2], hi[3]);
        tmem_st4(lane_base + CT_ALO + c4 * 4, lw[0], lw[1], lw[2], lw[3]);
        *reinterpret_cast<float4*>(img + mn_image_index(4 * c4, m, 128, width * 4)) = make_float4(xv[0], xv[1], xv[2], xv[3]);
      }
    };
    // reparameterisation noise of the tile's row (injected, or Philox keyed by the global row index): nothing depends
    // on it before the heads epilogue, so it is drawn while the first encoder product runs
    auto stage_eps = [&](long long tile) {
      const long long row = tile * CH_M + m;
      const bool row_ok = row < a.B;
#pragma unroll 1
      for (int jb = cp; jb < Lp16 / 4; jb += CH_CP) {
        float e4[4] = {0.f, 0.f, 0.f, 0.f};
        if (row_ok && jb * 4 < L) {
          if (a.eps != nullptr) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (jb * 4 + i < L) e4[i] = __ldg(a.eps + row * L + jb * 4 + i);
          } else {
            const unsigned long long step = a.step_dev != nullptr ? (unsigned long long)(*a.step_dev + 1) : a.step;
            const float4 r = philox_normal4(a.seed, a.sample_offset + (unsigned long long)row, (uint32_t)jb,
                                            (uint32_t)(step + 1));
            e4[0] = r.x; e4[1] = r.y; e4[2] = r.z; e4[3] = r.w;
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) my_ep[(jb * 4 + i) * 128] = jb * 4 + i < L ? e4[i] : 0.f;
      }
    };

    // The tile's trajectories (128 x I floats, contiguous in global memory) -> shared memory, zero past the batch end.
    // A full tile travels as ONE bulk copy issued by one thread (completion on x_full); a ragged or misaligned one is
    // loaded by all threads, which then arrive on the same barrier, so that the wait below is the same either way.
    uint32_t x_phase = 0;
    auto load_x = [&](long long tile, bool with_biases) {
      if (lng) {   // long trajectories are read from global memory row by row (set_row): only the biases come here
        if (with_biases)
          for (int i = tid; i < NUM_LAYERS * 128; i += CH_EPI_THREADS) {
            const int l = i >> 7, n = i & 127;
            if (n < lo.Np[l]) bias_s[i] = __ldg(pk + lo.q_b[l] + n);
          }
        asm volatile("bar.sync 1, %0;" ::"n"(CH_EPI_THREADS) : "memory");
        if (tid == 0) mbar_arrive(x_full);
        return;
      }
      const long long base = tile * CH_M * I;
      const long long left = a.B * I - base;
      const int nval = (int)(left < (long long)CH_M * I ? left : (long long)CH_M * I);
      const float* src = x_batch + base;
      const bool bulk = !shuffled && nval == CH_M * I && (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
      if (shuffled) {
        // the rows of this tile are scattered over the resident set: position p = batch * B + row of the epoch's
        // permutation (the reference's DataLoader(shuffle=True) draws a new row order per epoch, Training_VAE.py:327).
        // Two threads per row; all loads of a thread are issued before its first store.
        const int r = tid >> 1, half = tid & 1;
        const long long row = tile * CH_M + r;
        const int n0 = half * ((I + 1) / 2), n1 = half ? I : (I + 1) / 2;
        float v[32];
        if (row < a.B) {
          const uint32_t src_row = resident_row(a.x_shuffle_seed, x_epoch, (uint32_t)(x_b * (unsigned long long)a.B + (unsigned long long)row),
                                                (uint32_t)(a.x_batches * a.B));
          const float* sp = a.x + (size_t)src_row * I;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = n0 + j < n1 ? __ldg(sp + n0 + j) : 0.f;
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < n1) xbuf[r * I + n0 + j] = v[j];
        if (with_biases)
          for (int i = tid; i < NUM_LAYERS * 128; i += CH_EPI_THREADS) {
            const int l = i >> 7, n = i & 127;
            if (n < lo.Np[l]) bias_s[i] = __ldg(pk + lo.q_b[l] + n);
          }
        asm volatile("bar.sync 1, %0;" ::"n"(CH_EPI_THREADS) : "memory");
        if (tid == 0) mbar_arrive(x_full);
      } else if (bulk) {
        if (tid == 0) {
          uint32_t bytes = (uint32_t)(CH_M * I * 4);
          if (with_biases)
            for (int l = 0; l < NUM_LAYERS; ++l) bytes += (uint32_t)(lo.Np[l] * 4);
          fence_proxy_async_smem();   // earlier generic-proxy reads of the x tile are ordered before the copy overwrites it
          mbar_arrive_expect_tx(x_full, bytes);
          tma_load_1d(xbuf, src, (uint32_t)(CH_M * I * 4), x_full);
          if (with_biases)
            for (int l = 0; l < NUM_LAYERS; ++l) tma_load_1d(bias_s + l * 128, pk + lo.q_b[l], (uint32_t)(lo.Np[l] * 4), x_full);
        }
      } else {
        for (int i = tid; i < CH_M * I; i += CH_EPI_THREADS) xbuf[i] = i < nval ? __ldg(x_batch + base + i) : 0.f;
        if (with_biases)
          for (int i = tid; i < NUM_LAYERS * 128; i += CH_EPI_THREADS) {
            const int l = i >> 7, n = i & 127;
            if (n < lo.Np[l]) bias_s[i] = __ldg(pk + lo.q_b[l] + n);
          }
        asm volatile("bar.sync 1, %0;" ::"n"(CH_EPI_THREADS) : "memory");
        if (tid == 0) mbar_arrive(x_full);
      }
    };
    auto wait_x = [&]() {
      mbar_wait(x_full, x_phase);
      x_phase ^= 1u;
    };
    if (a.trace != nullptr && cta == 0 && tid == 0) a.trace[249] = global_ns();
    // the (padded) biases of all layers are read by every epilogue and stay in shared memory for the whole launch:
    // the padding is zeroed here, the values arrive with the first x tile
    for (int i = tid; i < NUM_LAYERS * 128; i += CH_EPI_THREADS)
      if ((i & 127) >= lo.Np[i >> 7]) bias_s[i] = 0.f;
    load_x(cta, true);
    wait_x();
    if (a.trace != nullptr && cta == 0 && tid == 0) a.trace[250] = global_ns();
    set_row(cta);
    stage_start(cta);
    stage_xrel(cta, 0);
    arrive_all(false);   // not an epilogue: the staged images are covered by the report of the tile's first epilogue
    if (a.trace != nullptr && cta == 0 && tid == 0) a.trace[251] = global_ns();

    for (long long tile = cta; tile < n_tiles; tile += ncta) {
      const long long row = tile * CH_M + m;
      const bool row_ok = row < a.B;
      float* ts = a.stash + (size_t)tile * lo.tile_stash;
      tr_tile = a.trace != nullptr && (!lng || NCK <= 2) && cta == 0 && tile == cta + (long long)a.trace_tile * ncta;
      epi_no = 0;
      float lw_prev[16], lw_sum[4], lw_rt;   // state of the chunked loss walk (epi_loss_chunk)
      int lw_pos;
      stage_eps(tile);   // under the first encoder product

      // this thread's features of a 128-wide stash image: base of its row, then per 8-feature unit
      // (mn_image_index with the row part hoisted; f = c*32 + cp*CH_CPT + up*8)
      const int row_part = (m >> 2) * 512 + (m & 3) * 32;
      const int swz = m & 3;
      // A 128-wide epilogue walks the accumulator in four phases of 32 columns (quarter c); the warps of a lane
      // quarter take CH_CPT columns each: this thread's columns of phase c are c*32 + cp*CH_CPT .. + CH_CPT - 1.
      auto unit_off = [&](int c, int up) -> int {   // 8 features c*32 + cp*CH_CPT + up*8 ..: one swizzled 32-byte unit
        return c * 128 + ((((cp * (CH_CPT / 8)) + up) ^ swz) << 3);
      };
      // hidden layer: + bias, relu -> mask, stash image, A operand.  Each phase hands its quarter of A (and of D) to
      // the MMA warp.  A mask word holds 32 consecutive columns of the thread, first column in the top bit.
      auto epi_hidden = [&](int ms, int xslot, int bias_l, uint32_t dcol, int with_start) {
        float* xs = ts + lo.slot_off[xslot] + row_part;
        const float* bias = bias_l >= 0 ? bias_s + bias_l * 128 + cp * CH_CPT : nullptr;
        const uint32_t dbase = lane_base + dcol + cp * CH_CPT;
        // one phase: CH_CPT accumulator values -> (+ bias) relu -> mask bits, stash image, quarter c of the A operand
        auto phase = [&](const uint32_t (&v)[CH_CPT], int c, uint32_t& mword) {
          float bv[CH_CPT];
          if (bias != nullptr) {
#pragma unroll
            for (int j4 = 0; j4 < CH_CPT / 4; ++j4) {
              const float4 b4 = *(reinterpret_cast<const float4*>(bias + c * 32) + j4);
              bv[4 * j4] = b4.x; bv[4 * j4 + 1] = b4.y; bv[4 * j4 + 2] = b4.z; bv[4 * j4 + 3] = b4.w;
            }
          }
          uint32_t hi[CH_CPT], lw[CH_CPT];
#pragma unroll
          for (int up = 0; up < CH_CPT / 8; ++up) {
            float xv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int j = up * 8 + i;
              uint32_t bits = v[j];
              if (bias != nullptr) bits = __float_as_uint(__uint_as_float(bits) + bv[j]);
              // positive <=> the negated bit pattern is negative as an integer: shift its sign into the mask
              mword = __funnelshift_l(0u - bits, mword, 1);
              const float x = fmaxf(__uint_as_float(bits), 0.f);
              split_tf32(x, hi[j], lw[j]);
              xv[i] = x;
            }
            st_global_v8(xs + unit_off(c, up), xv);
          }
          TmemVec<CH_CPT>::st(lane_base + CT_AHI + c * 32 + cp * CH_CPT, hi);
          TmemVec<CH_CPT>::st(lane_base + CT_ALO + c * 32 + cp * CH_CPT, lw);
          arrive_q(c);
        };
        wait_d();
        if (with_start) stage_start_cols();
        uint32_t mword = 0u;
        // rolled over the two halves of the accumulator (half the code of the four-phase unrolled form)
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint32_t v0[CH_CPT], v1[CH_CPT];
          TmemVec<CH_CPT>::ld(dbase + half * 64, v0);
          tmem_ld_wait();
          TmemVec<CH_CPT>::ld(dbase + half * 64 + 32, v1);
          phase(v0, 2 * half, mword);
          tmem_ld_wait();
          phase(v1, 2 * half + 1, mword);
          if (CH_MW == 2 || half == 1) my_mask[(ms * CH_MW + (CH_MW == 2 ? half : 0)) * CH_EPI_THREADS] = mword;
        }
        finish_epilogue();
      };
      // data gradient: D -> relu' mask of the layer input -> stash image (-> A operand); same four phases.
      // defer: the caller arrives for all four quarters itself (the last epilogue stages the next tile first)
      auto epi_dgrad = [&](int ms, int gslot, bool write_a, bool defer, uint32_t dcol) {
        float* gs = ts + lo.slot_off[gslot] + row_part;
        const uint32_t dbase = lane_base + dcol + cp * CH_CPT;
        auto phase = [&](const uint32_t (&v)[CH_CPT], int c, uint32_t& mword) {
          uint32_t hi[CH_CPT], lw[CH_CPT];
          float gv[CH_CPT];
#pragma unroll
          for (int j = 0; j < CH_CPT; ++j) {
            // all-ones / all-zeros from the top mask bit, then AND: no branch, no select on a predicate
            const uint32_t keep = (uint32_t)((int32_t)mword >> 31);
            mword <<= 1;
            gv[j] = __uint_as_float(v[j] & keep);
            split_tf32(gv[j], hi[j], lw[j]);
          }
#pragma unroll
          for (int up = 0; up < CH_CPT / 8; ++up) {
            const float g8[8] = {gv[up * 8], gv[up * 8 + 1], gv[up * 8 + 2], gv[up * 8 + 3],
                                 gv[up * 8 + 4], gv[up * 8 + 5], gv[up * 8 + 6], gv[up * 8 + 7]};
            st_global_v8(gs + unit_off(c, up), g8);
          }
          if (write_a) {
            TmemVec<CH_CPT>::st(lane_base + CT_AHI + c * 32 + cp * CH_CPT, hi);
            TmemVec<CH_CPT>::st(lane_base + CT_ALO + c * 32 + cp * CH_CPT, lw);
          }
          if (!defer) arrive_q(c);
        };
        wait_d();
        uint32_t mword = 0u;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint32_t v0[CH_CPT], v1[CH_CPT];
          if (CH_MW == 2 || half == 0) mword = my_mask[(ms * CH_MW + (CH_MW == 2 ? half : 0)) * CH_EPI_THREADS];
          TmemVec<CH_CPT>::ld(dbase + half * 64, v0);
          tmem_ld_wait();
          TmemVec<CH_CPT>::ld(dbase + half * 64 + 32, v1);
          phase(v0, 2 * half, mword);
          tmem_ld_wait();
          phase(v1, 2 * half + 1, mword);
        }
      };
      auto epi_heads = [&]() {
      // heads: mu, logvar (Training_VAE.py:193-196); z = mu + eps * exp(0.5 logvar) (:199-206); z becomes the A
      // operand of the z rows of dec0.  The h_c rows of dec0 are already running on the tensor cores (h_c is still
      // the main A operand: this epilogue must not touch it).  The KLD term of the loss (:243) is summed here,
      // where mu and logvar are at hand.  The warps of a lane quarter share the 4-latent blocks of a row.
      wait_d();
      if (cp == 0) {
        const float* bias = bias_s + L_HEADS * 128;
        for (int c = 0; c < NH / 16; ++c) {
          uint32_t v[16];
          tmem_ld16(lane_base + CT_HEADS + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) my_ml[(c * 16 + j) * 128] = __uint_as_float(v[j]) + bias[c * 16 + j];
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(CH_EPI_THREADS) : "memory");   // mu, logvar of the row are in shared memory
      if (tr_tile && tid == 0) a.trace[246] = clock64();
      float s_k = 0.f;
#pragma unroll 1
      for (int jb = cp; jb < lo.slot_w[SX_Z] / 4; jb += CH_CP) {   // past Lp16 / 4: the zero padding of the stash image
        float zv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = jb * 4 + i;
          float z = 0.f;
          if (j < L) {
            const float mu = my_ml[j * 128], lv = my_ml[(L + j) * 128];
            const float e_half = expf(0.5f * lv);          // exp(lv) = exp(lv / 2)^2: one exponential per element
            z = fmaf(my_ep[j * 128], e_half, mu);
            s_k += 1.f + lv - mu * mu - e_half * e_half;
          }
          zv[i] = z;
        }
        if (jb < Lp16 / 4) {
          uint32_t hi[4], lw[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) split_tf32(zv[i], hi[i], lw[i]);
          tmem_st4(lane_base + CT_Z_HI + jb * 4, hi[0], hi[1], hi[2], hi[3]);
          tmem_st4(lane_base + CT_Z_LO + jb * 4, lw[0], lw[1], lw[2], lw[3]);
        }
        *stash_ptr(ts, SX_Z, jb) = make_float4(zv[0], zv[1], zv[2], zv[3]);
      }
      if (row_ok) loss_acc[1] += -0.5f * s_k * (a.inv_batch / (float)L);
      release_a();
      };
      auto epi_loss = [&](uint32_t dcol) {
      // ---------------------------------- loss (Training_VAE.py:229-268) -----------------------
      // recon -> the recon / start / time terms and d(total)/d(recon), which becomes the A operand of dec3's
      // data gradient; one thread per row walks the time steps in order.  Rows of up to 32 features (seq_len <= 10)
      // stay in registers (compile-time feature indices); wider ones go through a scratch tile in shared memory.
      wait_d();
      auto row_loss = [&](auto ip_tag) {
        constexpr int IP = decltype(ip_tag)::value;
        float r[IP];   // recon, overwritten in place by d(total)/d(recon) three features behind the walk
        {
          const float* bias = bias_s + L_DEC3 * 128;
#pragma unroll
          for (int c = 0; c < IP / 16; ++c) {
            uint32_t v[16];
            tmem_ld16(lane_base + dcol + c * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 b4 = *(reinterpret_cast<const float4*>(bias + c * 16) + j4);
              r[c * 16 + 4 * j4] = __uint_as_float(v[4 * j4]) + b4.x;
              r[c * 16 + 4 * j4 + 1] = __uint_as_float(v[4 * j4 + 1]) + b4.y;
              r[c * 16 + 4 * j4 + 2] = __uint_as_float(v[4 * j4 + 2]) + b4.z;
              r[c * 16 + 4 * j4 + 3] = __uint_as_float(v[4 * j4 + 3]) + b4.w;
            }
          }
        }
        if (tr_tile && tid == 0) a.trace[248] = clock64();
        const float c_rec = a.w_recon * 2.f * a.inv_batch / (float)I;
        const float c_start = a.w_start * a.inv_batch;
        const float c_t0 = a.w_time * 2.f * a.inv_batch;
        const float c_mono = T > 1 ? a.w_time * a.inv_batch / (float)(T - 1) : 0.f;
        const float* xr = xbuf + m * I;
        const float sx = xr[1], sy = xr[2];
        float s_rec = 0.f, s_start = 0.f, s_t0 = 0.f, s_mono = 0.f;
        // the gradient of feature n is final once feature n + 3 (the next time step of the same column) has been
        // visited - the monotone-time term of step t + 1 adds to the time gradient of step t - and recon[n] is read
        // for the last time there: three gradients wait in `pend` and replace recon three features behind
        float pend[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int n = 0; n < IP; ++n) {
          const int d = n % 3;                        // compile-time: 0 time, 1 x, 2 y
          const bool on = n < I && row_ok;            // rows past the batch end carry no loss and no gradient
          const float xv = n < I ? xr[n] : 0.f;
          const float target = d == 0 ? xv : (d == 1 ? xv - sx : xv - sy);
          const float diff = r[n] - target;
          float gn = c_rec * diff;
          if (on) s_rec = fmaf(diff, diff, s_rec);
          if (n == 0) {
            s_t0 = r[0] * r[0];
            gn = fmaf(c_t0, r[0], gn);
          } else if (d == 0) {
            const float dt = r[n] - r[n - 3];
            if (on && dt < 0.f) {   // relu'(0) = 0: strict
              s_mono -= dt;
              gn -= c_mono;
              pend[0] += c_mono;
            }
          } else if (n < 3) {
            if (on) s_start = fmaf(diff, diff, s_start);
            gn = fmaf(c_start, diff, gn);
          }
          if (n >= 3) r[n - 3] = pend[d];
          pend[d] = on ? gn : 0.f;
        }
#pragma unroll
        for (int n = IP - 3; n < IP; ++n) r[n] = pend[n % 3];
        if (row_ok) {
          loss_acc[0] += s_rec * (a.inv_batch / (float)I);
          loss_acc[2] += s_start * (a.inv_batch * 0.5f);
          loss_acc[3] += s_t0 * a.inv_batch + (T > 1 ? s_mono * (a.inv_batch / (float)(T - 1)) : 0.f);
        }
#pragma unroll
        for (int c4 = 0; c4 < IP / 4; ++c4) {
          uint32_t hi[4], lw[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) split_tf32(r[c4 * 4 + i], hi[i], lw[i]);
          tmem_st4(lane_base + CT_AHI + c4 * 4, hi[0], hi[1], hi[2], hi[3]);
          tmem_st4(lane_base + CT_ALO + c4 * 4, lw[0], lw[1], lw[2], lw[3]);
          *stash_ptr(ts, SG_REC, c4) = make_float4(r[c4 * 4], r[c4 * 4 + 1], r[c4 * 4 + 2], r[c4 * 4 + 3]);
        }
      };
      if (cp == 0 && Ip == 32) row_loss(std::integral_constant<int, 32>{});
      if (cp == 0 && Ip != 32) {   // wider rows: through a scratch tile in shared memory
        float* rb = scratch + m;   // [n * 128 + m]
        const float* bias = bias_s + L_DEC3 * 128;
        for (int c = 0; c < Ip / 16; ++c) {
          uint32_t v[16];
          tmem_ld16(lane_base + dcol + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) rb[(c * 16 + j) * 128] = __uint_as_float(v[j]) + bias[c * 16 + j];
        }
        if (tr_tile && tid == 0) a.trace[248] = clock64();
        if (row_ok) {
          const float c_rec = a.w_recon * 2.f * a.inv_batch / (float)I;
          const float c_start = a.w_start * a.inv_batch;
          const float c_t0 = a.w_time * 2.f * a.inv_batch;
          const float c_mono = T > 1 ? a.w_time * a.inv_batch / (float)(T - 1) : 0.f;
          const float* xr = xbuf + m * I;
          const float sx = xr[1], sy = xr[2];
          float s_rec = 0.f, s_start = 0.f, s_t0 = 0.f, s_mono = 0.f;
          float g_prev = 0.f, r_prev = 0.f;
          for (int t = 0; t < T; ++t) {
            {
              const float r = rb[(3 * t) * 128], xv = xr[3 * t];
              const float diff = r - xv;
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
                rb[(3 * (t - 1)) * 128] = g_prev;
              }
              g_prev = g;
              r_prev = r;
            }
#pragma unroll
            for (int d = 1; d < 3; ++d) {
              const int n = 3 * t + d;
              const float r = rb[n * 128], xv = xr[n] - (d == 1 ? sx : sy);
              const float diff = r - xv;
              s_rec = fmaf(diff, diff, s_rec);
              float g = c_rec * diff;
              if (t == 0) {
                s_start = fmaf(diff, diff, s_start);
                g = fmaf(c_start, diff, g);
              }
              rb[n * 128] = g;
            }
          }
          rb[(3 * (T - 1)) * 128] = g_prev;
          loss_acc[0] += s_rec * (a.inv_batch / (float)I);
          loss_acc[2] += s_start * (a.inv_batch * 0.5f);
          loss_acc[3] += s_t0 * a.inv_batch + (T > 1 ? s_mono * (a.inv_batch / (float)(T - 1)) : 0.f);
        }
        for (int c4 = 0; c4 < Ip / 4; ++c4) {
          float gv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int n = c4 * 4 + i;
            gv[i] = (row_ok && n < I) ? rb[n * 128] : 0.f;   // rows past the batch end carry no gradient
          }
          uint32_t hi[4], lw[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) split_tf32(gv[i], hi[i], lw[i]);
          tmem_st4(lane_base + CT_AHI + c4 * 4, hi[0], hi[1], hi[2], hi[3]);
          tmem_st4(lane_base + CT_ALO + c4 * 4, lw[0], lw[1], lw[2], lw[3]);
          *stash_ptr(ts, SG_REC, c4) = make_float4(gv[0], gv[1], gv[2], gv[3]);
        }
      }
      release_a();
      };
      auto epi_bdec0 = [&]() {
      // dec0: d/dz is in its small accumulator; reparameterisation + KLD backward (Training_VAE.py:243):
      //   d/dmu = w_k mu / (B L) + g_z ;  d/dlogvar = -0.5 w_k (1 - e^lv) / (B L) + 0.5 g_z eps e^(lv/2)
      // -> the (mu, logvar) gradient becomes the A operand of the heads' data gradients (it replaces d/dz in the idle
      // accumulator).  The decoder share of d/dhc is being computed into the other accumulator meanwhile.
      if (tr_tile && tid == 0) a.trace[240] = clock64();
      wait_d();
      if (tr_tile && tid == 0) a.trace[241] = clock64();
      if (cp == 0) {
        const float c_k = a.w_kld * a.inv_batch / (float)L;
        // every thread first reads ALL of its row's d/dz (the gradient columns written below overlap them)
#pragma unroll 1
        for (int jb = 0; jb < (L + 3) / 4; ++jb) {
          uint32_t v[4];
          tmem_ld4(lane_base + CT_DZ + jb * 4, v);
          tmem_ld_wait();
          float gm[4], gl[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = jb * 4 + i;
            const bool on = j < L && row_ok;
            const float gz = __uint_as_float(v[i]);
            const float mu = j < L ? my_ml[j * 128] : 0.f, lv = j < L ? my_ml[(L + j) * 128] : 0.f, ep = j < L ? my_ep[j * 128] : 0.f;
            const float e_half = expf(0.5f * lv);          // exp(lv) = exp(lv / 2)^2: one exponential per element
            gm[i] = on ? fmaf(c_k, mu, gz) : 0.f;
            gl[i] = on ? -0.5f * c_k * (1.f - e_half * e_half) + 0.5f * gz * ep * e_half : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = jb * 4 + i;
            if (j < L) {
              my_ml[j * 128] = gm[i];
              my_ml[(L + j) * 128] = gl[i];
            }
          }
        }
        if (tr_tile && tid == 0) a.trace[242] = clock64();
#pragma unroll 1
        for (int c4 = 0; c4 < lo.slot_w[SG_ML] / 4; ++c4) {   // past NH / 4: the zero padding of the stash image
          float gv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int n = c4 * 4 + i;
            gv[i] = n < 2 * L ? my_ml[n * 128] : 0.f;
          }
          if (c4 < NH / 4) {
            uint32_t hi[4], lw[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) split_tf32(gv[i], hi[i], lw[i]);
            tmem_st4(lane_base + CT_GML_HI + c4 * 4, hi[0], hi[1], hi[2], hi[3]);
            tmem_st4(lane_base + CT_GML_LO + c4 * 4, lw[0], lw[1], lw[2], lw[3]);
          }
          *stash_ptr(ts, SG_ML, c4) = make_float4(gv[0], gv[1], gv[2], gv[3]);
        }
        if (tr_tile && tid == 0) a.trace[243] = clock64();
      }
      release_a();
      if (tr_tile && tid == 0) a.trace[245] = clock64();
      };

      // ---- long trajectories: the chunk epilogues --------------------------------------------------------------
      // x_rel chunk c -> the A operand of the first encoder layer's next product (the previous one has completed: wait_d)
      auto epi_stage_x = [&](int c) {
        wait_d();
        stage_xrel(tile, c);
        release_a();
      };
      // the recon gradient of chunk c, back from the stash -> the A operand of the last decoder layer's data gradient
      auto stage_g = [&](int c) {
        const float* img = ts + lo.slot_off[SG_REC] + (size_t)c * (128 * 128);
#pragma unroll 2
        for (int c4 = cp; c4 < 32; c4 += CH_CP) {
          const float4 g = *reinterpret_cast<const float4*>(img + mn_image_index(4 * c4, m, 128, 512));
          uint32_t hi[4], lw[4];
          split_tf32(g.x, hi[0], lw[0]); split_tf32(g.y, hi[1], lw[1]); split_tf32(g.z, hi[2], lw[2]); split_tf32(g.w, hi[3], lw[3]);
          tmem_st4(lane_base + CT_AHI + c4 * 4, hi[0], hi[1], hi[2], hi[3]);
          tmem_st4(lane_base + CT_ALO + c4 * 4, lw[0], lw[1], lw[2], lw[3]);
        }
      };
      auto epi_stage_g = [&](int c) {
        wait_d();
        stage_g(c);
        release_a();
      };
      // The loss (Training_VAE.py:229-268) over the 128 recon features of chunk c: one thread per row walks them in
      // groups of 16 straight from tensor memory.  The walk's state lives in the thread from chunk to chunk: the
      // sums, the previous time value, and the gradients of the last finished group - a time step whose successor runs
      // backwards gets c_mono added three features later, possibly from the next group or chunk, so a group is written
      // to the stash only when the next one has been walked.
      auto epi_loss_chunk = [&](int c, uint32_t dcol) {
        static_assert(CH_CP == 2, "the loss walk of a chunk is shared by the two threads of a row");
        wait_d();
        {
          // The two threads of a row take one half of the chunk each (groups [4 cp, 4 cp + 4) of 16 features): the walk is
          // bound by the instruction stream of ONE warp per scheduler (30 k cycles per chunk with half of the epilogue
          // warps idle).  The monotonicity term couples a time feature with the one three features earlier; across
          // the two seams (feature 64 of the chunk, and the chunk's end) every thread evaluates what it needs from the
          // accumulator columns next to its range - r = accumulator + bias, the same two operands and the same
          // addition on either side of a seam, so both threads see the same number.
          const float c_rec = a.w_recon * 2.f * a.inv_batch / (float)I;
          const float c_start = a.w_start * a.inv_batch;
          const float c_t0 = a.w_time * 2.f * a.inv_batch;
          const float c_mono = T > 1 ? a.w_time * a.inv_batch / (float)(T - 1) : 0.f;
          const float sx = x_at(1), sy = x_at(2);
          const float* bias = pk + lo.q_b[L_DEC3];
          float* img0 = ts + lo.slot_off[SG_REC];
          auto write_group = [&](int pos, const float (&g)[16]) {   // pos = 8 chunk + group
            float* img = img0 + (size_t)(pos >> 3) * (128 * 128);
            const int c4 = (pos & 7) * 4;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              *reinterpret_cast<float4*>(img + mn_image_index(4 * (c4 + q4), m, 128, 512)) =
                  make_float4(g[4 * q4], g[4 * q4 + 1], g[4 * q4 + 2], g[4 * q4 + 3]);
          };
          // r of feature n0 + k (k < 4) of this chunk: columns [col0, col0 + 4) of the accumulator + bias
          auto r4_at = [&](int col0, float (&r)[4]) {
            uint32_t v4[4];
            tmem_ld4(lane_base + dcol + col0, v4);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int n = c * 128 + col0 + k;
              r[k] = __uint_as_float(v4[k]) + (n < I ? __ldg(bias + n) : 0.f);
            }
          };
          // (static indices only: a run-time index would move the arrays to local memory)
          auto pick3 = [](const float (&r)[4], int i) { return i == 0 ? r[0] : (i == 1 ? r[1] : (i == 2 ? r[2] : r[3])); };
          auto bump = [&](int o, float x) {   // lw_prev[13 + o] += x
            if (o == 0) lw_prev[13] += x;
            else if (o == 1) lw_prev[14] += x;
            else lw_prev[15] += x;
          };
          const int g0 = 4 * cp;                       // this thread's groups [g0, g0 + 4)
          if (c == 0) {
            lw_sum[0] = lw_sum[1] = lw_sum[2] = lw_sum[3] = 0.f;
            if (cp == 0) lw_rt = 0.f;
            lw_pos = -1;
          }
          if (cp == 1) {
            if (c > 0) {
              // the last group of the previous chunk is still pending: a time step of THIS chunk that runs backwards
              // adds c_mono to its predecessor there (feature n_f - 3, position 13 + (n_f - 128 c) of that group)
              const int o = (3 - (c * 128) % 3) % 3, n_f = c * 128 + o;
              float r[4];
              r4_at(0, r);
              const float dt = pick3(r, o) - lw_rt;    // lw_rt: the last time feature of the previous chunk (this thread's)
              if (n_f < I && row_ok && dt < 0.f) bump(o, c_mono);
              write_group(lw_pos, lw_prev);
              lw_pos = -1;
            }
            // the time feature before feature 64 of the chunk (the other thread's last one)
            const int o = (c * 128 + 64) % 3, k = o == 0 ? 3 : o;      // n_p = 128 c + 64 - k, column 64 - k
            float r[4];
            r4_at(60, r);
            lw_rt = pick3(r, 4 - k);
          }
#pragma unroll 1
          for (int grp = g0; grp < g0 + 4; ++grp) {
            const int nb = c * 128 + grp * 16;
            // all loads of the group first (see stage_xrel: 114 k cycles per chunk with a load right before its use),
            // the accumulator columns under them
            float xg[16], bg[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 g = x4_at(nb + 4 * q);
              xg[4 * q] = g.x; xg[4 * q + 1] = g.y; xg[4 * q + 2] = g.z; xg[4 * q + 3] = g.w;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) bg[j] = nb + j < I ? __ldg(bias + nb + j) : 0.f;
            uint32_t v[16];
            tmem_ld16(lane_base + dcol + grp * 16, v);
            tmem_ld_wait();
            float gcur[16];
            const int ph = nb % 3;
            const bool seam = grp == g0;   // the predecessor of this group's first time feature belongs to the other thread
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int n = nb + j;
              int d = ph + (j % 3);
              d = d >= 3 ? d - 3 : d;                       // n % 3: 0 time, 1 x, 2 y
              const bool inside = n < I;
              const bool on = inside && row_ok;
              const float r = __uint_as_float(v[j]) + bg[j];
              const float xv = xg[j];
              const float target = d == 0 ? xv : (d == 1 ? xv - sx : xv - sy);
              const float diff = r - target;
              float gn = c_rec * diff;
              if (on) lw_sum[0] = fmaf(diff, diff, lw_sum[0]);
              if (d == 0) {
                if (n == 0) {
                  if (on) lw_sum[2] = r * r;
                  gn = fmaf(c_t0, r, gn);
                } else {
                  const float dt = r - lw_rt;
                  if (on && dt < 0.f) {   // relu'(0) = 0: strict
                    lw_sum[3] -= dt;
                    gn -= c_mono;
                    if (j >= 3) gcur[j >= 3 ? j - 3 : 0] += c_mono;
                    else if (!seam) lw_prev[13 + j] += c_mono;   // (across a seam the owner of that feature adds it)
                  }
                }
                lw_rt = r;
              } else if (n < 3) {
                if (on) lw_sum[1] = fmaf(diff, diff, lw_sum[1]);
                gn = fmaf(c_start, diff, gn);
              }
              gcur[j] = on ? gn : 0.f;
            }
            if (lw_pos >= 0) write_group(lw_pos, lw_prev);
#pragma unroll
            for (int j = 0; j < 16; ++j) lw_prev[j] = gcur[j];
            lw_pos = c * 8 + grp;
          }
          if (cp == 0) {
            // the first time feature of the other half (feature 128 c + 64 + o): running backwards, it adds c_mono to this
            // half's last time feature, position 13 + o of the group that is still pending here
            const int o = (3 - (c * 128 + 64) % 3) % 3, n_a = c * 128 + 64 + o;
            float r[4];
            r4_at(64, r);
            const float dt = pick3(r, o) - lw_rt;
            if (n_a < I && row_ok && dt < 0.f) bump(o, c_mono);
            write_group(lw_pos, lw_prev);
            lw_pos = -1;
            // ... and the last time feature of the chunk, which this thread's first one of the NEXT chunk is compared with
            // (its accumulator will be gone by then)
            if (c < NCK - 1) {
              const int k = (c * 128 + 128) % 3 == 0 ? 3 : (c * 128 + 128) % 3;   // feature 128 c + 128 - k, column 128 - k
              r4_at(124, r);
              lw_rt = pick3(r, 4 - k);
            }
          } else if (c == NCK - 1) {
            write_group(lw_pos, lw_prev);   // nothing follows the last chunk
          }
          if (c == NCK - 1 && row_ok) {
            loss_acc[0] += lw_sum[0] * (a.inv_batch / (float)I);
            loss_acc[2] += lw_sum[1] * (a.inv_batch * 0.5f);
            loss_acc[3] += lw_sum[2] * a.inv_batch + (T > 1 ? lw_sum[3] * (a.inv_batch / (float)(T - 1)) : 0.f);
          }
        }
        if (c == NCK - 1) {
          // every product of the last decoder layer is complete: its input d3 may go.  The first chunk of the gradient
          // comes back from the stash (both threads of a row have written their halves of it)
          asm volatile("bar.sync 1, %0;" ::"n"(CH_EPI_THREADS) : "memory");
          stage_g(0);
        }
        release_a();
      };

      const long long next = tile + ncta;
#pragma unroll 1
      for (int e = 0; e < a.n_epis; ++e) {
        const EpiRef er = epi_at(e, NCK);
        if (er.type == EP_STAGE_X) { epi_stage_x(er.chunk); continue; }
        if (er.type == EP_STAGE_G) { epi_stage_g(er.chunk); continue; }
        const int er_row = er.row;
        const int ty = c_epi[er_row][0];
        const uint32_t dcol = (uint32_t)c_epi[er_row][5];
        if (ty == EP_LOSS && lng) { epi_loss_chunk(er.chunk, dcol); continue; }
        if (ty == EP_HIDDEN) {
          epi_hidden(c_epi[er_row][1], c_epi[er_row][2], c_epi[er_row][4], dcol, c_epi[er_row][6]);
        } else if (ty == EP_DGRAD) {
          const bool last = e == a.n_epis - 1;
          epi_dgrad(c_epi[er_row][1], c_epi[er_row][2], c_epi[er_row][3] != 0, last, dcol);
          // after the first data gradient every warp is past the loss (that MMA could not finish before all of
          // them had released its A operand): the x tile can be replaced by the next tile's; after the last one
          // the next tile's start point and encoder input are staged, and only then is the A operand handed over
          if (next < n_tiles) {
            if (er_row == CH_EPI_FIRST_DGRAD && !lng) {
              asm volatile("bar.sync 1, %0;" ::"n"(CH_EPI_THREADS) : "memory");   // every warp is done with this tile's x
              load_x(next, false);
            } else if (last) {
              if (!lng) wait_x();
              set_row(next);
              stage_start(next);
              stage_xrel(next, 0);
            }
          }
          if (last) arrive_all();
          finish_epilogue();
        } else if (ty == EP_HEADS) {
          epi_heads();
        } else if (ty == EP_LOSS) {
          epi_loss(dcol);
        } else {
          epi_bdec0();
        }
      }
    }

    if (a.trace != nullptr && cta == 0 && tid == 0) a.trace[179] = global_ns();
    // loss partials: fixed-order tree inside the warp, one slot per (CTA, epilogue warp)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int t4 = 0; t4 < 4; ++t4) loss_acc[t4] += __shfl_xor_sync(0xffffffffu, loss_acc[t4], o);
    if (lane == 0) {
      float* dst = a.loss_part + ((size_t)cta * CH_EPI_WARPS + warp) * 4;
      dst[0] = loss_acc[0]; dst[1] = loss_acc[1]; dst[2] = loss_acc[2]; dst[3] = loss_acc[3];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == CH_PRODUCER_WARP) tmem_dealloc(tmem, CT_COLS);
  if (a.trace != nullptr && cta == 0 && tid == 0) a.trace[177] = global_ns();
}

template <bool LNG>
__global__ void __launch_bounds__(CH_THREADS, 1) chain_kernel(const __grid_constant__ ChainKArgs k) {
  extern __shared__ unsigned char smem_dyn[];
  chain_body<LNG>(k.lo, k.c, (int)blockIdx.x, (int)gridDim.x, smem_dyn);
}

// =========================================================================================
// weight-gradient kernel
// =========================================================================================
constexpr int WG_WORK_WARPS = 8;
constexpr int WG_WORK_THREADS = WG_WORK_WARPS * 32;
constexpr int WG_PRODUCER_WARP = WG_WORK_WARPS;
constexpr int WG_MMA_WARP = WG_WORK_WARPS + 1;
constexpr int WG_THREADS = (WG_MMA_WARP + 2) * 32;   // + one idle warp: the fused launch runs both bodies with the chain's block size
constexpr int WG_ROWS = 16;                 // batch rows per ring stage (two 8-deep contraction steps)
constexpr int WG_STAGE_FLOATS = 8192;       // [A_hi 2048][B_hi <= 2048][A_lo 2048][B_lo <= 2048]
constexpr int WG_MAX_OPS = 16;   // long trajectories: one op per 128-feature chunk of enc0 / dec3
constexpr int WG_ROLES = 3;
constexpr int WG_STAGES = 6;                // even: the work warps split alternate stages in two groups
constexpr int WG_SPLIT_THREADS = WG_WORK_THREADS / 2;

enum WKind { WK_COND0 = 0, WK_COND1, WK_ENC0, WK_ENC1, WK_ENC2, WK_ENC3, WK_HEADS_E, WK_HEADS_C, WK_DEC0_C, WK_DEC0_Z,
             WK_DEC1, WK_DEC2, WK_DEC3 };

struct WOp {
  int kind;
  int slotA, slotB;  // A: the 128-wide image that supplies the M dimension (tensor-memory lanes); B supplies N
  int FB;            // width of B
  int d_col;         // accumulator column
  int bias;          // 0 none; 1: A x ones -> column 0 of a 16-wide accumulator (lane = feature);
                     // 2: ones x B -> lane 0 of an FB-wide accumulator
  int d_col_b;
  int chunk;         // long trajectories: the 128-feature chunk of x_rel / the recon gradient this op reads (B side), else -1
};

// Layer groups whose accumulators fit tensor memory together; 128-wide layers are spread so that
// the MMA time of the roles is comparable.
// tensor-memory columns of the decoder-side role without enc2: dec3 (+ bias row), dec0 (h_c part + bias column, z part),
// heads (h_traj part + bias row, h_c part); enc2 needs 128 + 16 more
__host__ __device__ inline bool wgrad_enc2_in_role2(const Layout& lo) {
  return 2 * lo.Ip + (H + 16) + lo.Lp16 + 3 * lo.NH + (H + 16) <= 512;
}
__host__ __device__ inline int wgrad_program(const Layout& lo, int role, WOp* ops) {
  int n = 0, col = 0;
  auto add = [&](int kind, int sa, int sb, int FB, int bias) {
    WOp w;
    w.kind = kind; w.slotA = sa; w.slotB = sb; w.FB = FB; w.d_col = col; col += FB;
    w.bias = bias; w.d_col_b = col; w.chunk = -1;
    if (bias == 1) col += 16;
    else if (bias == 2) col += FB;
    ops[n++] = w;
  };
  // within a role the ops follow the order in which the chain kernel completes their gradient images
  // (epilogue numbers in c_epi), so that a CTA running next to the chain can start each one early.  The late
  // encoder layers are spread over the roles: enc2 joins the decoder-side role when its accumulator fits there
  // (that role is otherwise done long before the chain ends), else it stays with enc1.
  const bool enc2_in_role2 = wgrad_enc2_in_role2(lo);
  if (role == 0) {
    add(WK_COND1, SG_HC, SX_HC1, H, 1);       // 16
    add(WK_COND0, SG_HC1, SX_START, 16, 0);   // 17; columns 0, 1 = dW, column 2 = db (the ones column of the start image)
    if (!enc2_in_role2) add(WK_ENC2, SG_E3, SX_E2, H, 1);   // 19
    add(WK_ENC1, SG_E2, SX_E1, H, 1);         // 20
  } else if (role == 1) {
    add(WK_DEC2, SG_D3, SX_D2, H, 1);         // 12
    add(WK_DEC1, SG_D2, SX_D1, H, 1);         // 13
    add(WK_ENC3, SG_E4, SX_E3, H, 1);         // 18
    add(WK_ENC0, SG_E1, SX_X, lo.Ip, 1);      // 21
  } else {
    add(WK_DEC3, SX_D3, SG_REC, lo.Ip, 2);    // 11; transposed: lanes = input features, columns = output index
    add(WK_DEC0_C, SG_D1, SX_HC, H, 1);       // 14
    add(WK_DEC0_Z, SG_D1, SX_Z, lo.Lp16, 0);
    add(WK_HEADS_E, SX_E4, SG_ML, lo.NH, 2);  // 15; transposed: lanes = input features, columns = (mu, logvar) index
    add(WK_HEADS_C, SX_HC, SG_ML, lo.NH, 0);
    if (enc2_in_role2) add(WK_ENC2, SG_E3, SX_E2, H, 1);    // 19
  }
  return n;
}
// Long trajectories (chain_program_long): enc0 and dec3 become one op per 128-feature chunk, more accumulators than
// tensor memory holds.  Every op is written out (added to the CTA's slab) as soon as it is complete, and the ops
// alternate between two column sets, so that the write-out of one runs under the products of the next.
__host__ __device__ inline int wgrad_program_long(const Layout& lo, int role, WOp* ops) {
  int n = 0;
  auto add = [&](int kind, int sa, int sb, int FB, int bias, int chunk) {
    WOp w;
    w.kind = kind; w.slotA = sa; w.slotB = sb; w.FB = FB; w.d_col = (n & 1) * 256; w.bias = bias; w.d_col_b = w.d_col + 128;
    w.chunk = chunk;
    ops[n++] = w;
  };
  const bool enc2_in_role2 = wgrad_enc2_in_role2(lo);
  if (role == 0) {
    add(WK_COND1, SG_HC, SX_HC1, H, 1, -1);
    add(WK_COND0, SG_HC1, SX_START, 16, 0, -1);
    if (!enc2_in_role2) add(WK_ENC2, SG_E3, SX_E2, H, 1, -1);
    add(WK_ENC1, SG_E2, SX_E1, H, 1, -1);
  } else if (role == 1) {
    add(WK_DEC2, SG_D3, SX_D2, H, 1, -1);
    add(WK_DEC1, SG_D2, SX_D1, H, 1, -1);
    add(WK_ENC3, SG_E4, SX_E3, H, 1, -1);
    for (int c = 0; c < lo.NC; ++c) add(WK_ENC0, SG_E1, SX_X, H, c == 0 ? 1 : 0, c);
  } else {
    for (int c = 0; c < lo.NC; ++c) add(WK_DEC3, SX_D3, SG_REC, H, 2, c);
    add(WK_DEC0_C, SG_D1, SX_HC, H, 1, -1);
    add(WK_DEC0_Z, SG_D1, SX_Z, lo.Lp16, 0, -1);
    add(WK_HEADS_E, SX_E4, SG_ML, lo.NH, 2, -1);
    add(WK_HEADS_C, SX_HC, SG_ML, lo.NH, 0, -1);
    if (enc2_in_role2) add(WK_ENC2, SG_E3, SX_E2, H, 1, -1);
  }
  return n;
}
// relative cost of a role's ops per tile (for the division of the SMs between the roles)
__host__ __device__ inline int wgrad_role_cols(const Layout& lo, int role) {
  WOp ops[WG_MAX_OPS];
  if (chain_long(lo)) {
    const int n = wgrad_program_long(lo, role, ops);
    int total = 0;
    for (int o = 0; o < n; ++o) total += ops[o].FB + (ops[o].bias == 1 ? 16 : (ops[o].bias == 2 ? ops[o].FB / 4 : 0));
    return total;
  }
  const int n = wgrad_program(lo, role, ops);
  const WOp& w = ops[n - 1];   // columns are handed out in op order
  return w.bias == 0 ? w.d_col + w.FB : (w.bias == 1 ? w.d_col_b + 16 : w.d_col_b + w.FB);
}

struct WgradArgs {
  const float* stash;
  float* slabs;        // [units of all roles][slab_stride]
  long long n_tiles;
  int slab_stride;
  int role_end[WG_ROLES];   // CTA index ranges: role r owns [role_end[r-1], role_end[r])
  int unit_tiles[WG_ROLES]; // tiles accumulated in tensor memory before the accumulators are written out
  int unit_begin[WG_ROLES]; // first slab of the role; unit u of role r writes slab unit_begin[r] + u
  WOp ops[WG_ROLES][WG_MAX_OPS];   // the per-tile program of every role (wgrad_program, built on the host)
  int n_ops[WG_ROLES];
  const int* ready;         // when set: per-tile epilogue counters of the chain CTAs running beside this kernel's
  int nc;                   // long trajectories: their 128-feature chunks (wgrad_program_long; a CTA adds every op of
                            // every tile it walks to ONE slab of its own); 0: trajectories of up to 64 floats
  long long* trace;         // development aid: %globaltimer stamps of the CTAs of tile 0 (null in production)
};
struct WgradKArgs {
  Layout lo;
  WgradArgs w;
};

// pulls a byte range (multiple of 16) into L2 without a destination
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// number of the chain epilogue that completes a stash image; nc: chunks of a long trajectory (epi_at), else 1
__device__ inline int slot_epilogue(int slot, int nc) {
  if (slot == SX_START) return 0;                   // staged before the tile's first epilogue
  if (slot == SX_X) return nc - 1;                  // ... the last chunk of x_rel by the last staging epilogue
  if (slot == SG_REC) return 2 * nc + 8;            // the last chunk of the loss writes the last gradients
  for (int r = 0; r < CH_EPIS; ++r)
    if (c_epi[r][2] == slot) return r < 10 ? r + nc - 1 : r + 3 * (nc - 1);
  return CH_EPIS + 3 * (nc - 1) - 1;
}
// spin (one lane) until the chain has completed epilogue `epi` of a tile, then order the bulk copies behind it
// The chain CTAs come first in the grid and never wait, so the counter always moves; should it not (a fault in the chain
// body), the launch must end with an error instead of hanging the GPU: after two seconds the waiting thread traps.
__device__ __forceinline__ void wait_tile_ready(const int* flag, int epi) {
  const int need = (epi + 1) * CH_EPI_WARPS;
  long long t0 = 0;
  for (int spin = 0;; ++spin) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v >= need) break;
    __nanosleep(100);
    if ((spin & 1023) == 1023) {
      const long long now = global_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ll) __trap();
    }
  }
  fence_proxy_async_global();
}

constexpr int WG_WO_FLOATS = WG_WORK_WARPS * 32 * 20;   // write-out staging: per work warp 32 rows x 16 columns, row stride 20
constexpr size_t WG_SMEM_BYTES = (size_t)WG_STAGES * WG_STAGE_FLOATS * 4 + 4096 /* ones (A side) */ + 1024 /* ones (B side) */ +
                                 WG_WO_FLOATS * 4 + 32 * 8 + 16 + 1024;
static_assert(WG_SMEM_BYTES <= 232448, "weight-gradient kernel: shared memory");

// cta: index of this CTA among the weight-gradient CTAs
template <bool LNG>
__device__ __forceinline__ void wgrad_body(const Layout& lo, const WgradArgs& a, const int cta, unsigned char* smem_dyn) {
  float *ring, *ones_a, *ones_b, *wo_stage;
  uint64_t *raw_full, *split_full, *empty, *d_done, *d_free;
  uint32_t* tmem_slot;
  {
    const uint32_t base = smem_u32(smem_dyn);
    unsigned char* p = smem_dyn + ((1024u - (base & 1023u)) & 1023u);
    ring = reinterpret_cast<float*>(p);
    ones_a = ring + (size_t)WG_STAGES * WG_STAGE_FLOATS;
    ones_b = ones_a + 1024;
    wo_stage = ones_b + 256;
    raw_full = reinterpret_cast<uint64_t*>(wo_stage + WG_WO_FLOATS);
    split_full = raw_full + 8;
    empty = split_full + 8;
    d_done = empty + 8;     // two, used alternately: the work warps may be one write-out behind the MMA warp
    d_free = d_done + 2;
    tmem_slot = reinterpret_cast<uint32_t*>(d_free + 1);
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int role = 0;
  while (role < WG_ROLES - 1 && cta >= a.role_end[role]) ++role;
  const int role_begin = role == 0 ? 0 : a.role_end[role - 1];
  const int role_ctas = a.role_end[role] - role_begin;
  const int my_index = cta - role_begin;
  // A unit = a run of consecutive tiles whose gradients are accumulated in tensor memory and then written to
  // the unit's own partial slab: bounds the number of (round-toward-zero) accumulations per accumulator.
  const long long ut = a.unit_tiles[role];
  const long long n_units = (a.n_tiles + ut - 1) / ut;

  if (tid == 0) {
    if (a.trace != nullptr && my_index == 0) a.trace[180 + role * 16] = global_ns();
    for (int st = 0; st < WG_STAGES; ++st) {
      mbar_init(&raw_full[st], 1);
      mbar_init(&split_full[st], WG_WORK_WARPS / 2);
      mbar_init(&empty[st], 1);
    }
    mbar_init(&d_done[0], 1);
    mbar_init(&d_done[1], 1);
    mbar_init(d_free, WG_WORK_WARPS);
    mbar_fence_init();
  }
  // constant "ones" operands (MN-major images of one 8-row group): feature 0 is 1 in every row
  for (int i = tid; i < 1024 + 256; i += WG_THREADS) ones_a[i] = 0.f;
  __syncthreads();
  if (tid < 8) {
    ones_a[mn_image_index(0, tid, 128, 512)] = 1.0f;   // 128 features: 4 feature atoms per 4-row atom
    ones_b[mn_image_index(0, tid, 128, 128)] = 1.0f;   // one feature atom
  }
  if (warp == WG_PRODUCER_WARP) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const WOp* ops = a.ops[role];
  const int n_ops = a.n_ops[role];
  // Next to a running chain (one tile per unit) every op is written out as soon as its own products are complete,
  // while the CTA waits for the chain to finish the images of its next op; otherwise once per unit.
  const int nck = LNG ? a.nc : 1;
  // Long trajectories: every op of every tile is added to the CTA's one slab as soon as it is complete (the
  // accumulators of all chunks do not fit tensor memory).  Else next to a running chain with one tile per unit: ...
  const bool per_op = ut == 1 && (a.ready != nullptr || LNG);
  // Long trajectories after a chain kernel, several tiles per unit: op by op over the unit's tiles, an op's accumulator
  // kept in tensor memory over those tiles and added to the slab once per unit instead of once per tile (the
  // read-add-write of every op of every tile was 16 % of the sampled stalls of the work warps at T = 100)
  const bool lng_units = LNG && ut > 1;
  // Several tiles per unit next to a running chain: op by op over the unit's tiles (the early images of ALL its tiles
  // before the late images of the first), so that only the last ops are left when the chain ends.  After a chain
  // kernel: tile by tile (the next tile's images are being prefetched meanwhile).
  const bool op_major = (a.ready != nullptr && !per_op) || lng_units;
  constexpr int chunks = CH_M / WG_ROWS;  // 8 ring stages per (tile, op)

  if (warp == WG_PRODUCER_WARP) {
    if (lane == 0) {
      RingStateRt rs(WG_STAGES);
      for (long long unit = my_index; unit < n_units; unit += role_ctas) {
        const long long t0 = unit * ut;
        const int nt = (int)(min((unit + 1) * ut, a.n_tiles) - t0);
        for (int it = 0; it < nt * n_ops; ++it) {
          const int o = op_major ? it / nt : it % n_ops;
          const long long tile = t0 + (op_major ? it % nt : it / n_ops);
          const float* ts = a.stash + (size_t)tile * lo.tile_stash;
          const WOp op = ops[o];
          const float* srcA = ts + lo.slot_off[op.slotA];
          const float* srcB = ts + lo.slot_off[op.slotB] + (LNG && op.chunk > 0 ? (size_t)op.chunk * (128 * 128) : 0);
          const int FBm = LNG && op.chunk >= 0 ? 128 : lo.slot_w[op.slotB];  // width of the B image in memory
          const uint32_t bytesA = WG_ROWS * H * 4, bytesB = (uint32_t)(WG_ROWS * FBm * 4);
          if (a.ready != nullptr) wait_tile_ready(a.ready + tile, max(slot_epilogue(op.slotA, nck), slot_epilogue(op.slotB, nck)));
          if (a.trace != nullptr && tile == 0) a.trace[180 + role * 16 + 1 + o] = global_ns();
          if (a.ready == nullptr) {
            // After a chain kernel the stash of a large batch streams from HBM, and six 16-row stages in flight do
            // not cover that latency: the images of the NEXT op (of this tile, or the first op of the next tile
            // of this CTA) are pulled into L2 while this op's stages go through the ring.
            int on = o + 1;
            long long tn = tile;
            if (op_major) {   // next in the walk: the same op of the unit's next tile, else the next op of its first tile
              on = o;
              tn = tile + 1;
              if (tn == t0 + nt) {
                on = o + 1;
                tn = t0;
                if (on == n_ops) {
                  on = 0;
                  tn = (unit + role_ctas) * ut;
                }
              }
            } else if (on == n_ops) {
              on = 0;
              tn = tile + 1 < t0 + nt ? tile + 1 : (unit + role_ctas) * ut;
            }
            if (tn < a.n_tiles) {
              const float* tsn = a.stash + (size_t)tn * lo.tile_stash;
              l2_prefetch(tsn + lo.slot_off[ops[on].slotA], CH_M * H * 4);
              if (LNG && ops[on].chunk >= 0) l2_prefetch(tsn + lo.slot_off[ops[on].slotB] + (size_t)ops[on].chunk * (128 * 128), CH_M * 128 * 4);
              else l2_prefetch(tsn + lo.slot_off[ops[on].slotB], (uint32_t)(CH_M * lo.slot_w[ops[on].slotB] * 4));
            }
          }
          for (int c = 0; c < chunks; ++c) {
            float* dst = ring + rs.stage * WG_STAGE_FLOATS;
            mbar_wait(&empty[rs.stage], rs.phase ^ 1u);
            mbar_arrive_expect_tx(&raw_full[rs.stage], bytesA + bytesB);
            tma_load_1d(dst, srcA + (size_t)c * WG_ROWS * H, bytesA, &raw_full[rs.stage]);
            tma_load_1d(dst + 2048, srcB + (size_t)c * WG_ROWS * FBm, bytesB, &raw_full[rs.stage]);
            rs.advance();
          }
        }
      }
    }
  } else if (warp == WG_MMA_WARP) {
    RingStateRt rs(WG_STAGES);
    uint32_t free_phase = 0;
    uint32_t n_commit = 0;   // completions signalled so far: number k goes to d_done[k & 1]
    for (long long unit = my_index; unit < n_units; unit += role_ctas) {
      if (unit != my_index) {  // the previous unit's accumulators have been read out
        mbar_wait(d_free, free_phase);
        free_phase ^= 1u;
        tc_fence_after();
      }
      const long long t0 = unit * ut;
      const int nt = (int)(min((unit + 1) * ut, a.n_tiles) - t0);
      for (int it = 0; it < nt * n_ops; ++it) {
        const int o = op_major ? it / nt : it % n_ops;
        const bool first_tile = (op_major ? it % nt : it / n_ops) == 0;
        const WOp op = ops[o];
        const uint32_t idesc = umma_idesc_tf32(128, op.FB, UMMA_A_MN | UMMA_B_MN);
        const uint32_t idesc_b1 = umma_idesc_tf32(128, 16, UMMA_A_MN | UMMA_B_MN);
        // MN-major images: LBO = feature-atom stride (512 bytes), SBO = 4-row-atom stride (width * 16 bytes)
        const uint32_t FBm = LNG && op.chunk >= 0 ? 128u : (uint32_t)lo.slot_w[op.slotB];
        const uint64_t a_bits = umma_desc(0u, 512u, H * 16u, 1u), b_bits = umma_desc(0u, 512u, FBm * 16u, 1u);
        const uint64_t ones_a_desc = umma_desc(smem_u32(ones_a), 512u, 2048u, 1u);
        const uint64_t ones_b_desc = umma_desc(smem_u32(ones_b), 512u, 512u, 1u);
        for (int c = 0; c < chunks; ++c) {
          mbar_wait(&split_full[rs.stage], rs.phase);
          tc_fence_after();
          const uint32_t s0 = smem_u32(ring + rs.stage * WG_STAGE_FLOATS);
          if (elect_one()) {
#pragma unroll
            for (int g = 0; g < WG_ROWS / 8; ++g) {
              const uint64_t a_hi = a_bits | (uint64_t)((s0 + (uint32_t)g * (H * 32u)) >> 4);
              const uint64_t a_lo = a_bits | (uint64_t)((s0 + 16384u + (uint32_t)g * (H * 32u)) >> 4);
              const uint64_t b_hi = b_bits | (uint64_t)((s0 + 8192u + (uint32_t)g * (FBm * 32u)) >> 4);
              const uint64_t b_lo = b_bits | (uint64_t)((s0 + 24576u + (uint32_t)g * (FBm * 32u)) >> 4);
              const uint32_t accf = (first_tile && c == 0 && g == 0) ? 0u : 1u;
              umma_tf32_ss(tmem + (uint32_t)op.d_col, a_hi, b_hi, idesc, accf);
              umma_tf32_ss(tmem + (uint32_t)op.d_col, a_lo, b_hi, idesc, 1u);
              umma_tf32_ss(tmem + (uint32_t)op.d_col, a_hi, b_lo, idesc, 1u);
              if (op.bias == 1) {
                umma_tf32_ss(tmem + (uint32_t)op.d_col_b, a_hi, ones_b_desc, idesc_b1, accf);
                umma_tf32_ss(tmem + (uint32_t)op.d_col_b, a_lo, ones_b_desc, idesc_b1, 1u);
              } else if (op.bias == 2) {
                umma_tf32_ss(tmem + (uint32_t)op.d_col_b, ones_a_desc, b_hi, idesc, accf);
                umma_tf32_ss(tmem + (uint32_t)op.d_col_b, ones_a_desc, b_lo, idesc, 1u);
              }
            }
            umma_commit(&empty[rs.stage]);
          }
          __syncwarp();
          rs.advance();
        }
        // this op's accumulator is final (one tile per unit, or the unit's last tile of the op): the work warps write
        // it out while the next op's products run
        if (per_op || (lng_units && it % nt == nt - 1)) {
          if (elect_one()) umma_commit(&d_done[n_commit & 1u]);
          __syncwarp();
          ++n_commit;
        }
      }
      if (!per_op && !lng_units) {
        if (elect_one()) umma_commit(&d_done[n_commit & 1u]);
        __syncwarp();
        ++n_commit;
      }
    }
  } else if (warp < WG_WORK_WARPS) {
    // ===================== work warps: TF32 split of the raw stages, then the final write-out ======
    RingStateRt rs(WG_STAGES);
    uint32_t n_done = 0;     // completions consumed so far (the MMA warp's n_commit)
    auto wait_done = [&]() {
      mbar_wait(&d_done[n_done & 1u], (n_done >> 1) & 1u);
      ++n_done;
      tc_fence_after();
    };
    // one op's accumulator -> the unit's partial slab
    auto write_op = [&](int o, long long unit) {
    const int q = warp & 3, hh = warp >> 2;
    const int ln = q * 32 + lane;  // tensor-memory lane = feature index of the A side
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    // long trajectories: one slab per CTA, the first unit writes it, the others add to it
    const bool add_to = LNG && unit != my_index;
    float* slab = a.slabs + (size_t)(a.unit_begin[role] + (LNG ? my_index : unit)) * a.slab_stride;
    float* stg = wo_stage + warp * (32 * 20);   // this warp's staging tile of the write-out
    const int f0 = LNG && ops[o].chunk > 0 ? ops[o].chunk * 128 : 0;   // first feature of the op's chunk of the trajectory
    // Adding to the slab: a reduction at the L2 (red.global.add.f32, nothing comes back), not load - add - store: the
    // slabs of a large batch are evicted by the stash that streams past them, and a thread that waits for the old value
    // of every element pays an HBM round trip per element (28 % of this kernel's stall samples at T = 100).  An element
    // is only ever touched by one thread, unit after unit: same-address accesses of one thread keep their order, so the
    // sum is the same bits as before.
    auto put = [&](float* dst, float val) {
      if (add_to) atomicAdd(dst, val);
      else *dst = val;
    };
    const int L = lo.L, I = lo.I;
    {
      const WOp op = ops[o];
      // columns of this op are shared between the two warps of a lane quarter in 16-column chunks
      for (int c = hh; c < op.FB / 16; c += 2) {
        uint32_t v[16];
        tmem_ld16(lane_base + (uint32_t)(op.d_col + c * 16), v);
        tmem_ld_wait();
        // layers stored [lane = output feature][column = input feature]: row stride and first float of the
        // chunk in the row of tensor-memory lane 0
        float* row0 = nullptr;
        int rstride = H, ncols = 16;
        switch (op.kind) {
          case WK_COND1: row0 = slab + lo.p_w[L_COND1] + c * 16; break;
          case WK_ENC0: row0 = slab + lo.p_w[L_ENC0] + f0 + c * 16; rstride = I; ncols = min(16, I - f0 - c * 16); break;
          case WK_ENC1: row0 = slab + lo.p_w[L_ENC1] + c * 16; break;
          case WK_ENC2: row0 = slab + lo.p_w[L_ENC2] + c * 16; break;
          case WK_ENC3: row0 = slab + lo.p_w[L_ENC3] + c * 16; break;
          case WK_DEC0_C: row0 = slab + lo.p_w[L_DEC0] + L + c * 16; rstride = L + H; break;
          case WK_DEC0_Z: row0 = slab + lo.p_w[L_DEC0] + c * 16; rstride = L + H; ncols = min(16, L - c * 16); break;
          case WK_DEC1: row0 = slab + lo.p_w[L_DEC1] + c * 16; break;
          case WK_DEC2: row0 = slab + lo.p_w[L_DEC2] + c * 16; break;
          default: break;
        }
        if (row0 != nullptr) {
          // A lane holds 16 floats of its own row: stored directly, every instruction would touch 32 different
          // lines.  Through the warp's staging tile (row stride 20 floats: conflict-free 128-bit accesses both
          // ways) an instruction instead writes 8 rows x 64 contiguous bytes.
          float4* mine = reinterpret_cast<float4*>(stg + lane * 20);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4)
            mine[j4] = make_float4(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1]), __uint_as_float(v[4 * j4 + 2]),
                                   __uint_as_float(v[4 * j4 + 3]));
          __syncwarp();
          const int c4 = lane & 3;
          const bool vec = ncols == 16 && (rstride & 3) == 0 && (reinterpret_cast<uintptr_t>(row0) & 15u) == 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = (lane >> 2) + 8 * i;
            const float4 val = *reinterpret_cast<const float4*>(stg + r * 20 + c4 * 4);
            float* dst = row0 + (size_t)(q * 32 + r) * rstride + c4 * 4;
            if (vec && !add_to) {
              *reinterpret_cast<float4*>(dst) = val;
            } else {
              if (c4 * 4 + 0 < ncols) put(dst, val.x);
              if (c4 * 4 + 1 < ncols) put(dst + 1, val.y);
              if (c4 * 4 + 2 < ncols) put(dst + 2, val.z);
              if (c4 * 4 + 3 < ncols) put(dst + 3, val.w);
            }
          }
          __syncwarp();   // the tile is rewritten by the next chunk
          continue;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = c * 16 + j;
          const float val = __uint_as_float(v[j]);
          switch (op.kind) {
            case WK_COND0:
              if (col < 2) put(slab + lo.p_w[L_COND0] + ln * 2 + col, val);
              else if (col == 2) put(slab + lo.p_b[L_COND0] + ln, val);
              break;
            case WK_HEADS_E:
            case WK_HEADS_C: {  // transposed: for one column the warp writes 32 consecutive floats
              const int koff = op.kind == WK_HEADS_C ? H : 0;
              if (col < L) put(slab + lo.p_w[L_HEADS] + col * (2 * H) + koff + ln, val);
              else if (col < 2 * L) put(slab + lo.p_wlv + (col - L) * (2 * H) + koff + ln, val);
              break;
            }
            default:  // WK_DEC3, transposed
              if (f0 + col < I) put(slab + lo.p_w[L_DEC3] + (size_t)(f0 + col) * H + ln, val);
              break;
          }
        }
      }
      if (op.bias == 1 && hh == 0) {
        uint32_t v[16];
        tmem_ld16(lane_base + (uint32_t)op.d_col_b, v);
        tmem_ld_wait();
        const float val = __uint_as_float(v[0]);
        int l = L_COND1;
        switch (op.kind) {
          case WK_COND1: l = L_COND1; break;
          case WK_ENC0: l = L_ENC0; break;
          case WK_ENC1: l = L_ENC1; break;
          case WK_ENC2: l = L_ENC2; break;
          case WK_ENC3: l = L_ENC3; break;
          case WK_DEC0_C: l = L_DEC0; break;
          case WK_DEC1: l = L_DEC1; break;
          default: l = L_DEC2; break;
        }
        put(slab + lo.p_b[l] + ln, val);
      } else if (op.bias == 2 && q == 0) {
        for (int c = hh; c < op.FB / 16; c += 2) {
          uint32_t v[16];
          tmem_ld16(lane_base + (uint32_t)(op.d_col_b + c * 16), v);
          tmem_ld_wait();
          if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = c * 16 + j;
              const float val = __uint_as_float(v[j]);
              if (op.kind == WK_HEADS_E) {
                if (col < L) put(slab + lo.p_b[L_HEADS] + col, val);
                else if (col < 2 * L) put(slab + lo.p_blv + (col - L), val);
              } else if (f0 + col < I) {
                put(slab + lo.p_b[L_DEC3] + f0 + col, val);
              }
            }
          }
        }
      }
        }
    };
    for (long long unit = my_index; unit < n_units; unit += role_ctas) {
    const int nt = (int)(min((unit + 1) * ut, a.n_tiles) - unit * ut);
    int pending = -1;   // per_op: the op whose accumulator is complete (or about to be) and not yet written out
    for (int it = 0; it < nt * n_ops; ++it) {
      {
        const int o = op_major ? it / nt : it % n_ops;
        const int nB4 = WG_ROWS * (LNG && ops[o].chunk >= 0 ? 128 : lo.slot_w[ops[o].slotB]) / 4;  // float4 of the B part
        for (int c = 0; c < chunks; ++c, rs.advance()) {
          // the two halves of the work warps take alternate stages (the ring depth is even, so a stage always has the
          // same half): two stages are being split at any time, which hides the shared-memory round trip, the proxy
          // fence and the hand-over of one behind the other
          if ((rs.stage & 1) != (warp >> 2)) continue;
          float4* st4 = reinterpret_cast<float4*>(ring + rs.stage * WG_STAGE_FLOATS);
          mbar_wait(&raw_full[rs.stage], rs.phase);
          const int n4 = 512 + nB4;   // A part [0, 512), B part [512, 512 + nB4)
          for (int i0 = tid & (WG_SPLIT_THREADS - 1); i0 < n4; i0 += 4 * WG_SPLIT_THREADS) {
            float4 x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)   // all loads of a batch are issued before the first store
              if (i0 + u * WG_SPLIT_THREADS < n4) x[u] = st4[i0 + u * WG_SPLIT_THREADS];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int idx = i0 + u * WG_SPLIT_THREADS;
              if (idx < n4)   // the raw words stay where they are: they ARE the high halves (split_tf32_cut)
                st4[idx + 1024] = make_float4(tf32_cut_low(x[u].x), tf32_cut_low(x[u].y), tf32_cut_low(x[u].z), tf32_cut_low(x[u].w));
            }
          }
          fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) mbar_arrive(&split_full[rs.stage]);
        }
        if (per_op || (lng_units && it % nt == 0)) {   // (several tiles per unit: after the op's FIRST tile)
          // the stages of op o are split and on their way through the tensor cores: meanwhile the accumulator of the
          // op before it (complete by now, or soon) goes to the slab - a write-out costs almost as much as the
          // products of an op, and the accumulators of a role do not share columns
          if (pending >= 0) {
            wait_done();
            write_op(pending, unit);
            // long trajectories: op o + 1 writes the columns this op has just been read from, and its first stage is
            // handed over by HALF of the work warps - all of them must be done reading first
            if (LNG) asm volatile("bar.sync 1, %0;" ::"n"(WG_WORK_THREADS) : "memory");
          }
          pending = o;
        }
      }
    }
    if ((per_op || lng_units) && pending >= 0) {
      wait_done();
      if (a.trace != nullptr && unit == 0 && tid == 0) a.trace[180 + role * 16 + 8] = global_ns();
      write_op(pending, unit);
    }

    // ---- write-out: accumulators -> the unit's partial slab (torch layout of each tensor) ----------
    if (!per_op && !lng_units) {
      wait_done();
      if (a.trace != nullptr && unit == 0 && tid == 0) a.trace[180 + role * 16 + 8] = global_ns();
      for (int o = 0; o < n_ops; ++o) write_op(o, unit);
    }
    // accumulators read: the MMA warp may start the next unit
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(d_free);
    if (a.trace != nullptr && unit == 0 && tid == 0) a.trace[180 + role * 16 + 9] = global_ns();
    }  // units
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WG_PRODUCER_WARP) tmem_dealloc(tmem, 512);
}

template <bool LNG>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const __grid_constant__ WgradKArgs k) {
  extern __shared__ unsigned char smem_dyn[];
  wgrad_body<LNG>(k.lo, k.w, (int)blockIdx.x, smem_dyn);
}

// Small batches leave most SMs without a chain tile: one launch runs the chain CTAs (the first chain_grid blocks)
// and the weight-gradient CTAs side by side, and a weight-gradient CTA picks up each stash image of its tile as
// soon as the tile's counter says the chain has completed it.  The chain never waits for the other side, and
// its blocks come first in the grid, so the wait cannot deadlock.
struct FusedTcArgs {
  Layout lo;
  ChainArgs c;
  WgradArgs w;
  int chain_grid;
};
static_assert(CH_THREADS == WG_THREADS, "the fused launch runs both bodies with one block size");
template <bool LNG>
__global__ void __launch_bounds__(CH_THREADS, 1) train_tc_fused_kernel(const __grid_constant__ FusedTcArgs k) {
  extern __shared__ unsigned char smem_dyn[];
  if ((int)blockIdx.x < k.chain_grid) chain_body<LNG>(k.lo, k.c, (int)blockIdx.x, k.chain_grid, smem_dyn);
  else wgrad_body<LNG>(k.lo, k.w, (int)blockIdx.x - k.chain_grid, smem_dyn);
}

// =========================================================================================
// reduction of the partial slabs (+ Adam)
// =========================================================================================
constexpr int REDUCE_TC_THREADS = 256;
struct AdamScalarsTc {
  float w1, b2, w2, step_size, bc2_sqrt, eps;
};
struct DpCtl {
  int owned_from;              // smallest world size that uses the owner scheme
  long long timeout_ns;        // how long a thread polls before it gives up
  unsigned int* status;        // the status word at the end of this rank's inbox
};
struct ReduceTcArgs {
  int tensor_end[24];   // end offset (floats) of each state_dict tensor
  int tensor_role[24];
  int role_begin[WG_ROLES], role_count[WG_ROLES];
  int n_params, slab_stride, chain_grid;
  float w_recon, w_kld, w_start, w_time;
  AdamScalarsTc h;
  int adam;
  // graph-capturable launches: the scalars that depend on the step index are derived in the kernel from
  // *step_dev + 1 (double arithmetic, as the host path and torch's Python floats)
  const long long* step_dev;
  double lr, beta1, beta2;
  // with the update: the kernel-layout weight arena is refreshed in the same thread (scatter_param), and the
  // last block to finish advances the device-side step counter (`done` counts finished blocks, self-resetting)
  float* packed;
  long long* step_inc;
  unsigned int* done;
  int* tile_flags;            // per-tile progress counters of the fused launch before this kernel: reset here for the next pass
  int n_tile_flags;
  // data parallel (dp.world > 1): the gradient exchange over peer memory runs inside this kernel, between the
  // slab sum and the update (dp_all_sum)
  DmvaeDpPeers dp;
  int dp_stride;
  DpCtl dp_ctl;              // owner-scheme threshold, poll time-out and status word of dp_all_sum
  unsigned int epoch_host;   // the step index when there is no device-side counter
  long long* trace;          // development aid: %globaltimer stamps of block 0 (slots 230..)
};

// Data-parallel exchange, "low-latency" style: a gradient travels as ONE 8-byte word {step index : value bits}
// written straight into the peer's inbox with a single 64-bit store (one register, one NVLink write - unlike a
// two-element vector access, which the PTX memory model treats as two scalar accesses, the word cannot be
// observed half-written), so the receiver polls the word itself and no fence or separate flag is needed.
__device__ __forceinline__ void dp_push(unsigned long long* dst, float v, unsigned int epoch) {
  const unsigned long long w = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(dst), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long dp_load(const unsigned long long* src) {
  unsigned long long w;
  asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(w) : "l"(src) : "memory");
  return w;
}
// Polls until the word carries `epoch`.  A dead or lagging peer (or ranks that disagree about the scheme) must not
// hang the GPU: after timeout_ns the thread raises the rank's status word, returns NaN for the element - the
// parameters of the rank turn NaN, loudly - and the kernel runs to its end (dmvae_dp_status reports the word).
__device__ __noinline__ unsigned long long dp_wait_slow(const unsigned long long* src, unsigned int epoch, const DpCtl& c) {
  const long long t0 = global_ns();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
      const unsigned long long w = dp_load(src);
      if ((unsigned int)(w >> 32) == epoch) return w;
    }
    if (global_ns() - t0 > c.timeout_ns) {
      atomicExch(c.status, 0x80000000u | epoch);
      return ((unsigned long long)epoch << 32) | 0x7fc00000ull;
    }
  }
}
__device__ __forceinline__ unsigned long long dp_wait(unsigned long long w, const unsigned long long* src, unsigned int epoch,
                                                      const DpCtl& c) {
  if ((unsigned int)(w >> 32) == epoch) return w;
  return dp_wait_slow(src, epoch, c);
}
// Sum of `mine` over the ranks; idx = element index inside the exchange.  Every rank receives the same bits.
//   up to 2 ranks (world < owned_from): every rank pushes its word to every peer and adds the ranks in rank order (one NVLink hop);
//   more ranks:    an element has an owner ((idx / 256) mod world, i.e. block-wise round robin).  The others push
//                  their word to the owner only; the owner adds the ranks in rank order and pushes the sum to
//                  everyone's sum area.  Two hops, but (world + 6) / 8 MB instead of (world - 1) MB out of every
//                  rank per step - at 8 GPUs the all-to-all version spent ~20 us per step on it.
// Inbox of a rank: [source rank 0..world-1][step parity][dp_stride] words, then [sum][step parity][dp_stride], then
// the status word.
__device__ __forceinline__ float dp_all_sum(const DmvaeDpPeers& dp, int stride, int idx, float mine, unsigned int epoch,
                                            const DpCtl& c) {
  typedef unsigned long long u64;
  const size_t par = (size_t)(epoch & 1u) * stride + idx;
  const u64* in = reinterpret_cast<const u64*>(dp.inbox[dp.rank]) + par;
  const int owner = dp.world < c.owned_from ? -1 : (idx >> 8) % dp.world;
  if (owner >= 0 && owner != dp.rank) {
    dp_push(reinterpret_cast<u64*>(dp.inbox[owner]) + (size_t)dp.rank * 2 * stride + par, mine, epoch);
    const u64* sum_in = in + (size_t)dp.world * 2 * stride;
    return __uint_as_float((unsigned int)dp_wait(dp_load(sum_in), sum_in, epoch, c));
  }
  if (owner < 0)
    for (int p = 0; p < dp.world; ++p)
      if (p != dp.rank) dp_push(reinterpret_cast<u64*>(dp.inbox[p]) + (size_t)dp.rank * 2 * stride + par, mine, epoch);
  // all peers' words are requested at once (independent loads), then those that had not arrived yet are polled
  u64 w[DMVAE_MAX_PEERS];
#pragma unroll
  for (int p = 0; p < DMVAE_MAX_PEERS; ++p)
    if (p < dp.world && p != dp.rank) w[p] = dp_load(in + (size_t)p * 2 * stride);
  float g = 0.f;
#pragma unroll
  for (int p = 0; p < DMVAE_MAX_PEERS; ++p) {
    if (p >= dp.world) break;
    if (p == dp.rank) { g += mine; continue; }
    g += __uint_as_float((unsigned int)dp_wait(w[p], in + (size_t)p * 2 * stride, epoch, c));
  }
  if (owner >= 0)
    for (int p = 0; p < dp.world; ++p)
      if (p != dp.rank) dp_push(reinterpret_cast<u64*>(dp.inbox[p]) + (size_t)dp.world * 2 * stride + par, g, epoch);
  return g;
}

__global__ void reduce_tc_kernel(const __grid_constant__ Layout lo, const float* __restrict__ slabs,
                                 const float* __restrict__ loss_part, const __grid_constant__ ReduceTcArgs r,
                                 float* __restrict__ grads, float* __restrict__ p, float* __restrict__ m,
                                 float* __restrict__ v) {
  // One thread per parameter.  (Four lanes per parameter, each summing a quarter of the slabs, was measured at twice
  // the time: the update and the scatter then run with a quarter of the lanes on four times the warps.)
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < r.n_tile_flags) r.tile_flags[e] = 0;   // every CTA of the fused launch has exited: its counters start over
  __shared__ AdamScalarsTc hs;
  __shared__ unsigned int epoch_s;
  const bool tr = r.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  if (tr) r.trace[230] = global_ns();
  if (threadIdx.x == 0) {
    hs = r.h;
    epoch_s = r.epoch_host;
    if (r.step_dev != nullptr) {
      const double step = (double)(*r.step_dev + 1);
      epoch_s = (unsigned int)(*r.step_dev + 1);
      if (r.adam) {
        hs.step_size = (float)(r.lr / (1.0 - pow(r.beta1, step)));
        hs.bc2_sqrt = (float)sqrt(1.0 - pow(r.beta2, step));
      }
    }
  }
  __syncthreads();
  const bool dp_on = r.dp.world > 1;
  float s = 0.f;
  if (e < r.n_params) {
    int t = 0;
    while (t < 23 && e >= r.tensor_end[t]) ++t;
    const int role = r.tensor_role[t];
    const float* src = slabs + (size_t)r.role_begin[role] * r.slab_stride + e;
    const int n = r.role_count[role];
    // fixed order; 16 loads in flight per thread (the sum stays one sequential chain of adds)
#pragma unroll 16
    for (int c = 0; c < n; ++c) s += __ldcg(src + (size_t)c * r.slab_stride);
    if (tr) r.trace[231] = global_ns();
    if (dp_on) s = dp_all_sum(r.dp, r.dp_stride, e, s, epoch_s, r.dp_ctl);
    if (tr) r.trace[233] = global_ns() + (long long)(s == 123.456f);
  }
  if (e < r.n_params) {
    grads[e] = s;
    if (r.adam) {
      // torch optim/adam.py::_single_tensor_adam (see dmvae_adam.cu)
      float pp = p[e], mm = m[e], vv = v[e];
      mm = hs.w1 < 0.5f ? fmaf(hs.w1, s - mm, mm) : s - (s - mm) * (1.f - hs.w1);
      vv = fmaf(hs.w2 * s, s, vv * hs.b2);
      const float denom = sqrtf(vv) / hs.bc2_sqrt + hs.eps;
      pp = pp - hs.step_size * (mm / denom);
      p[e] = pp; m[e] = mm; v[e] = vv;
      if (r.packed != nullptr) scatter_param(lo, e, pp, r.packed);
    }
  }
  if (tr) r.trace[234] = global_ns();
  if (r.step_inc != nullptr) {   // every block has read *step_dev (above, before its first barrier) by the time it arrives here
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(r.done, 1u) == gridDim.x - 1) {
        *r.done = 0u;
        *r.step_inc += 1;
      }
    }
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x < 32) {
    // the five loss terms: lanes stride the (CTA, epilogue warp) partials, fixed-order shuffle tree
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = threadIdx.x; c < r.chain_grid * CH_EPI_WARPS; c += 32) {
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) t[qd] += __ldcg(loss_part + (size_t)c * 4 + qd);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) t[qd] += __shfl_xor_sync(0xffffffffu, t[qd], o);
    float mine[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    {
      const float start = r.w_start > 0.f ? t[2] : 0.f;
      const float time = r.w_time > 0.f ? t[3] : 0.f;
      float total = r.w_recon * t[0] + r.w_kld * t[1];
      if (r.w_start > 0.f) total += r.w_start * start;
      if (r.w_time > 0.f) total += r.w_time * time;
      mine[0] = total; mine[1] = t[0]; mine[2] = t[1]; mine[3] = start; mine[4] = time;   // identical in every lane
    }
    const int lane = (int)threadIdx.x;
    float val = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) val = lane == i ? mine[i] : val;
    if (dp_on && lane < 5) val = dp_all_sum(r.dp, r.dp_stride, r.n_params + lane, val, epoch_s, r.dp_ctl);   // the global batch's
    if (lane < 5) grads[r.n_params + lane] = val;
  }
}

// =========================================================================================
// host side
// =========================================================================================
// the extra A columns (64) hold the widest small operand (NH = 2 * latent_dim padded) and, in their last 8, the ones
// column, which is rewritten at the start of every tile
// latent_dim <= 32 (the widest small operand, NH = 2 * latent_dim padded, is 64 columns); any trajectory length: up
// to 64 floats in one piece (chain_program), longer ones in chunks of 128 features (chain_program_long)
bool train_tc_supported(const Layout& lo) { return lo.NH <= 64 && lo.Lp16 <= 32; }

static long long* g_chain_trace = nullptr;
static int g_chain_trace_tile = 0;
void set_chain_trace(long long* p, int tile) { g_chain_trace = p; g_chain_trace_tile = tile; }
static thread_local bool g_tc_overlap = true;   // per host thread, like dmvae_set_train_impl
void set_train_tc_overlap(bool on) { g_tc_overlap = on; }

TrainTcPlan plan_train_tc(const Layout& lo, long long B, int sm_count, int overlap) {
  TrainTcPlan p;
  p.n_tiles = (B + CH_M - 1) / CH_M;
  p.chain_grid = (int)(p.n_tiles < sm_count ? p.n_tiles : sm_count);
  p.chain_stages = 5;   // one more than a 128 x 128 layer holds: the first stage of the next op is always in flight
  while (p.chain_stages > 2 && chain_smem_bytes(lo, p.chain_stages) > 232448) --p.chain_stages;
  p.chain_smem = chain_smem_bytes(lo, p.chain_stages);
  // Chain and weight-gradient CTAs side by side in ONE launch (train_tc_fused_kernel), every CTA on its own SM, a
  // chain CTA per tile:
  //   up to a quarter of the SMs in tiles: one weight-gradient CTA per tile and role; the weight gradients overlap the
  //     second half of the chain;
  //   up to 68 tiles on 148 SMs ("streamed"): the other SMs are divided between the roles, and a weight-gradient CTA
  //     follows a unit of several tiles op by op (+14..18 % at batch 4864 .. 8192 over the two kernels, whose chain
  //     leaves more than half of the GPU idle; break-even at 74 tiles).  More tiles than that and a chain CTA would
  //     have to walk several tiles while the weight gradients (which cost an SM as much as the chain does) pile up
  //     behind it: measured slower than the two kernels with chain CTAs of 2..8 tiles at batch 12 288 .. 65 536.
  const bool want = overlap < 0 ? g_tc_overlap : overlap != 0;
  p.overlap = want && (long long)(1 + WG_ROLES) * p.n_tiles <= sm_count;
  p.streamed = want && !p.overlap && 2 * p.n_tiles <= sm_count - 12 && p.n_tiles <= WS_DONE_SLOT;
  if (p.streamed) p.overlap = true;
  // CTAs per role in proportion to the accumulator columns (= MMA time per tile), at most one per tile.
  // Tiles are grouped into units of at most 4 (64 K steps: bounds the tensor-memory accumulation depth);
  // every unit writes its own partial slab.
  int cols[WG_ROLES], total = 0;
  for (int r = 0; r < WG_ROLES; ++r) { cols[r] = wgrad_role_cols(lo, r); total += cols[r]; }
  const int wg_sms = p.streamed ? sm_count - p.chain_grid : sm_count;
  int used = 0, slabs = 0;
  for (int r = 0; r < WG_ROLES; ++r) {
    long long n = (long long)wg_sms * cols[r] / total;
    if (n < 1) n = 1;
    if (n > p.n_tiles || (p.overlap && !p.streamed)) n = p.n_tiles;
    p.role_count[r] = (int)n;
    p.role_begin[r] = used;
    used += (int)n;
    // unit size 1..4 that minimises the busiest CTA: units go round-robin, a write-out costs ~0.4 tiles
    int best_u = 1;
    double best_cost = 1e30;
    for (int u = 1; u <= 4; ++u) {
      const long long units = (p.n_tiles + u - 1) / u;
      const double cost = (double)((units + n - 1) / n) * (u + 0.4);
      if (cost < best_cost - 1e-9 || (cost < best_cost + 1e-9 && u > best_u)) { best_cost = cost; best_u = u; }
    }
    // long trajectories: an op's accumulator goes straight to the slab of its CTA - next to a running chain after every
    // tile, after a chain kernel once per unit of tiles (the same op of up to 4 tiles accumulated in tensor memory)
    if (chain_long(lo) && p.overlap) best_u = 1;
    // ... and every CTA of the role must walk at least one unit: its slab is summed whether it wrote it or not
    while (chain_long(lo) && best_u > 1 && (p.n_tiles + best_u - 1) / best_u < n) --best_u;
    p.unit_tiles[r] = best_u;
    p.unit_count[r] = chain_long(lo) ? (int)n : (int)((p.n_tiles + p.unit_tiles[r] - 1) / p.unit_tiles[r]);
    p.unit_begin[r] = slabs;
    slabs += p.unit_count[r];
  }
  p.wgrad_grid = used;
  p.n_slabs = slabs;
  p.slab_stride = round_up(lo.n_params, 4);
  p.stash_floats = (size_t)p.n_tiles * lo.tile_stash;
  p.slab_floats = (size_t)p.n_slabs * p.slab_stride;
  p.loss_floats = (size_t)p.chain_grid * CH_EPI_WARPS * 4;
  return p;
}

static ChainArgs chain_args(const Layout& lo, const TrainTcPlan& plan, const TrainIO& io, float* stash, float* loss_part) {
  ChainArgs a;
  a.packed = io.packed; a.x = io.x; a.eps = io.eps; a.stash = stash; a.loss_part = loss_part;
  a.x_batches = io.step_dev != nullptr ? io.x_batches : 0;
  a.x_shuffle = io.x_shuffle;
  a.x_shuffle_seed = io.x_shuffle_seed;
  a.seed = io.seed; a.sample_offset = io.sample_offset; a.step = io.step; a.B = io.B;
  a.w_recon = io.w_recon; a.w_kld = io.w_kld; a.w_start = io.w_start; a.w_time = io.w_time; a.inv_batch = io.inv_batch;
  a.stages = plan.chain_stages;
  a.step_dev = io.step_dev;
  a.trace = g_chain_trace;
  a.trace_tile = g_chain_trace_tile;
  a.ready = nullptr;
  for (int o = 0; o < CH_MAX_OPS; ++o) a.ops[o] = COp{};
  a.n_chunks = chain_long(lo) ? lo.NC : 0;
  a.n_ops = chain_long(lo) ? chain_program_long(lo, a.ops) : chain_program(lo, a.ops);
  a.n_epis = CH_EPIS + 3 * (a.n_chunks > 0 ? a.n_chunks - 1 : 0);
  return a;
}
static WgradArgs wgrad_args(const Layout& lo, const TrainTcPlan& plan, const float* stash, float* slabs) {
  WgradArgs a;
  a.stash = stash; a.slabs = slabs; a.n_tiles = plan.n_tiles; a.slab_stride = plan.slab_stride;
  for (int r = 0; r < WG_ROLES; ++r) {
    a.role_end[r] = plan.role_begin[r] + plan.role_count[r];
    a.unit_tiles[r] = plan.unit_tiles[r];
    a.unit_begin[r] = plan.unit_begin[r];
    for (int o = 0; o < WG_MAX_OPS; ++o) a.ops[r][o] = WOp{};
    a.n_ops[r] = chain_long(lo) ? wgrad_program_long(lo, r, a.ops[r]) : wgrad_program(lo, r, a.ops[r]);
  }
  a.nc = chain_long(lo) ? lo.NC : 0;
  a.ready = nullptr;
  a.trace = g_chain_trace;
  return a;
}
// The opt-in shared-memory limit is an attribute of (kernel, device): it is set on every launch, as the other
// launchers do (a cached "already set" would be wrong on a second device of the process and racy between threads).
template <typename Kernel>
static cudaError_t set_smem_limit(Kernel kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t launch_chain(const Layout& lo, const TrainTcPlan& plan, const TrainIO& io, float* stash, float* loss_part,
                         cudaStream_t stream) {
  ChainKArgs k;
  k.lo = lo;
  k.c = chain_args(lo, plan, io, stash, loss_part);
  auto kernel = k.c.n_chunks > 0 ? chain_kernel<true> : chain_kernel<false>;
  const cudaError_t e = set_smem_limit(kernel, plan.chain_smem);
  if (e != cudaSuccess) return e;
  kernel<<<plan.chain_grid, CH_THREADS, plan.chain_smem, stream>>>(k);
  return cudaGetLastError();
}

cudaError_t launch_wgrad(const Layout& lo, const TrainTcPlan& plan, const float* stash, float* slabs, cudaStream_t stream) {
  WgradKArgs k;
  k.lo = lo;
  k.w = wgrad_args(lo, plan, stash, slabs);
  auto kernel = k.w.nc > 0 ? wgrad_kernel<true> : wgrad_kernel<false>;
  const cudaError_t e = set_smem_limit(kernel, WG_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  kernel<<<plan.wgrad_grid, WG_THREADS, WG_SMEM_BYTES, stream>>>(k);
  return cudaGetLastError();
}

// plan.overlap: chain and weight-gradient CTAs in one launch; `flags` = the zeroed per-tile counters in the workspace
cudaError_t launch_chain_wgrad_fused(const Layout& lo, const TrainTcPlan& plan, const TrainIO& io, float* stash, float* slabs,
                                     float* loss_part, int* flags, cudaStream_t stream) {
  FusedTcArgs k;
  k.lo = lo;
  k.c = chain_args(lo, plan, io, stash, loss_part);
  k.w = wgrad_args(lo, plan, stash, slabs);
  k.c.ready = flags;
  k.w.ready = flags;
  k.chain_grid = plan.chain_grid;
  const size_t smem = plan.chain_smem > WG_SMEM_BYTES ? plan.chain_smem : WG_SMEM_BYTES;
  auto kernel = k.c.n_chunks > 0 ? train_tc_fused_kernel<true> : train_tc_fused_kernel<false>;
  cudaError_t e = set_smem_limit(kernel, smem);
  if (e != cudaSuccess) return e;
  kernel<<<plan.chain_grid + plan.wgrad_grid, CH_THREADS, smem, stream>>>(k);
  return cudaGetLastError();
}

cudaError_t launch_reduce_tc(const Layout& lo, const TrainTcPlan& plan, const float* slabs, const float* loss_part,
                             const float w[4], float* grads, const DmvaeAdam* adam, float* p, float* m, float* v,
                             const long long* step_dev, float* packed, long long* step_inc, unsigned int* done,
                             const DmvaeDpPeers* dp, int* tile_flags, cudaStream_t stream) {
  ReduceTcArgs r;
  r.tile_flags = tile_flags;
  r.n_tile_flags = (tile_flags != nullptr && plan.overlap) ? (int)plan.n_tiles : 0;   // n_tiles < n_params always (one thread each)
  if (dp != nullptr) r.dp = *dp;
  else { r.dp = DmvaeDpPeers{}; r.dp.world = 1; }
  r.dp_stride = dp_exchange_stride(lo);
  r.dp_ctl.owned_from = r.dp.owned_from > 0 ? r.dp.owned_from : 3;
  r.dp_ctl.timeout_ns = (long long)(r.dp.timeout_ms > 0 ? r.dp.timeout_ms : 2000) * 1000000ll;
  r.dp_ctl.status = r.dp.world > 1 ? reinterpret_cast<unsigned int*>(static_cast<char*>(r.dp.inbox[r.dp.rank]) +
                                                                     dp_status_offset(lo, r.dp.world))
                                   : nullptr;
  r.epoch_host = adam != nullptr ? (unsigned int)adam->step : 0u;
  r.trace = g_chain_trace;
  r.packed = adam != nullptr ? packed : nullptr;
  r.step_inc = (adam != nullptr && done != nullptr) ? step_inc : nullptr;
  r.done = done;
  // state_dict order: cond0 (w,b) cond1 enc0 enc1 enc2 enc3 fc_mu fc_logvar dec0 dec1 dec2 dec3
  const int role_of_pair[12] = {0, 0, 1, 0, wgrad_enc2_in_role2(lo) ? 2 : 0, 1, 2, 2, 2, 1, 1, 2};
  const int w_off[12] = {lo.p_w[L_COND0], lo.p_w[L_COND1], lo.p_w[L_ENC0], lo.p_w[L_ENC1], lo.p_w[L_ENC2], lo.p_w[L_ENC3],
                         lo.p_w[L_HEADS], lo.p_wlv,        lo.p_w[L_DEC0], lo.p_w[L_DEC1], lo.p_w[L_DEC2], lo.p_w[L_DEC3]};
  const int b_off[12] = {lo.p_b[L_COND0], lo.p_b[L_COND1], lo.p_b[L_ENC0], lo.p_b[L_ENC1], lo.p_b[L_ENC2], lo.p_b[L_ENC3],
                         lo.p_b[L_HEADS], lo.p_blv,        lo.p_b[L_DEC0], lo.p_b[L_DEC1], lo.p_b[L_DEC2], lo.p_b[L_DEC3]};
  for (int i = 0; i < 12; ++i) {
    r.tensor_end[2 * i] = b_off[i];                                     // the weight ends where its bias starts
    r.tensor_end[2 * i + 1] = i < 11 ? w_off[i + 1] : lo.n_params;      // the bias ends where the next weight starts
    r.tensor_role[2 * i] = r.tensor_role[2 * i + 1] = role_of_pair[i];
  }
  for (int k = 0; k < WG_ROLES; ++k) { r.role_begin[k] = plan.unit_begin[k]; r.role_count[k] = plan.unit_count[k]; }
  r.n_params = lo.n_params; r.slab_stride = plan.slab_stride; r.chain_grid = plan.chain_grid;
  r.w_recon = w[0]; r.w_kld = w[1]; r.w_start = w[2]; r.w_time = w[3];
  r.adam = adam != nullptr ? 1 : 0;
  r.h = AdamScalarsTc{};
  r.step_dev = step_dev;
  r.lr = adam != nullptr ? adam->lr : 0.0; r.beta1 = adam != nullptr ? adam->beta1 : 0.0; r.beta2 = adam != nullptr ? adam->beta2 : 0.0;
  if (adam != nullptr) {
    const double b1 = adam->beta1, b2 = adam->beta2;
    const double bc1 = 1.0 - pow(b1, (double)adam->step), bc2 = 1.0 - pow(b2, (double)adam->step);
    r.h.w1 = (float)(1.0 - b1); r.h.b2 = (float)b2; r.h.w2 = (float)(1.0 - b2);
    r.h.step_size = (float)(adam->lr / bc1); r.h.bc2_sqrt = (float)sqrt(bc2); r.h.eps = (float)adam->eps;
  }
  reduce_tc_kernel<<<(lo.n_params + REDUCE_TC_THREADS - 1) / REDUCE_TC_THREADS, REDUCE_TC_THREADS, 0, stream>>>(lo, slabs, loss_part, r, grads, p, m, v);
  return cudaGetLastError();
}

}  // namespace dmvae
