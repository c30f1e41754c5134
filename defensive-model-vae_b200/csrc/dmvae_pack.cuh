// dmvae_pack.cuh - the kernel-layout ("packed") weight arena, element by element.
//
// Two views of the same mapping (Layout in dmvae_common.cuh describes the arena):
//   pack_element    gather: packed element -> the parameter it holds (pack_kernel: a full repack, one thread
//                   per packed element; also writes the zero padding)
//   scatter_param   scatter: one parameter -> every packed element that holds it (the optimizer kernels update
//                   the arena in the same thread that updates the parameter, so a training step needs no repack)
// Both are __host__ __device__: tests/test_pack_scatter.py compiles them for the host and checks that the scatter
// of every parameter reproduces the gather of the whole arena.
#pragma once
#include <cassert>

#include "dmvae_common.cuh"

namespace dmvae {

__host__ __device__ inline float fwd_weight(const Layout& lo, const float* __restrict__ p, int l, int k, int n) {
  if (n >= lo.N[l]) return 0.f;
  if (l == L_HEADS) {
    return n < lo.L ? p[lo.p_w[l] + n * (2 * H) + k] : p[lo.p_wlv + (n - lo.L) * (2 * H) + k];
  }
  return p[lo.p_w[l] + n * lo.K[l] + k];
}
__host__ __device__ inline float fwd_bias(const Layout& lo, const float* __restrict__ p, int l, int n) {
  if (n >= lo.N[l]) return 0.f;
  if (l == L_HEADS) return n < lo.L ? p[lo.p_b[l] + n] : p[lo.p_blv + (n - lo.L)];
  return p[lo.p_b[l] + n];
}

// Round-to-nearest TF32 (10 explicit mantissa bits) of an fp32 value, as an fp32 value.
__host__ __device__ inline float tf32_rn(float x) {
  union { float f; uint32_t u; } v;
  v.f = x;
  v.u = (v.u + 0x1000u) & 0xffffe000u;
  return v.f;
}

// Bias n of a tensor-core layer (zero in the padding).
__host__ __device__ inline float tc_bias(const Layout& lo, const float* __restrict__ p, int t, int n) {
  switch (t) {
    case TC_COND1: return p[lo.p_b[L_COND1] + n];
    case TC_ENC0: return p[lo.p_b[L_ENC0] + n];
    case TC_ENC1: return p[lo.p_b[L_ENC1] + n];
    case TC_ENC2: return p[lo.p_b[L_ENC2] + n];
    case TC_ENC3: return p[lo.p_b[L_ENC3] + n];
    case TC_HEADS: return n < lo.L ? p[lo.p_b[L_HEADS] + n] : (n < 2 * lo.L ? p[lo.p_blv + (n - lo.L)] : 0.f);
    case TC_DEC0: return p[lo.p_b[L_DEC0] + n];
    case TC_DEC1: return p[lo.p_b[L_DEC1] + n];
    case TC_DEC2: return p[lo.p_b[L_DEC2] + n];
    case TC_DEC3: return n < lo.I ? p[lo.p_b[L_DEC3] + n] : 0.f;
    default: return 0.f;
  }
}

// Weight (k, n) of a tensor-core layer in torch layout (zero in the padding).
__host__ __device__ inline float tc_weight(const Layout& lo, const float* __restrict__ p, int t, int k, int n) {
  switch (t) {
    case TC_COND0:  // rows: weight of x0, weight of y0, bias (multiplies the ones column), zeros
      return k < 2 ? p[lo.p_w[L_COND0] + n * 2 + k] : (k == 2 ? p[lo.p_b[L_COND0] + n] : 0.f);
    case TC_COND1: return p[lo.p_w[L_COND1] + n * H + k];
    case TC_ENC0: return k < lo.I ? p[lo.p_w[L_ENC0] + n * lo.I + k] : 0.f;
    case TC_ENC1: return p[lo.p_w[L_ENC1] + n * H + k];
    case TC_ENC2: return p[lo.p_w[L_ENC2] + n * H + k];
    case TC_ENC3: return p[lo.p_w[L_ENC3] + n * H + k];
    case TC_HEADS:  // rows [0,128) multiply h_traj, [128,256) h_c (Training_VAE.py:193); columns mu then logvar
      if (n < lo.L) return p[lo.p_w[L_HEADS] + n * (2 * H) + k];
      return n < 2 * lo.L ? p[lo.p_wlv + (n - lo.L) * (2 * H) + k] : 0.f;
    case TC_DEC0: {  // contraction ordered [h_c (128) ; z (L, zero padded to Lp16)]: the shared-start
                     // generation path skips the h_c steps, and h_c stays in place in tensor memory
      const int Kd = lo.L + H;
      if (k < H) return p[lo.p_w[L_DEC0] + n * Kd + lo.L + k];
      return (k - H) < lo.L ? p[lo.p_w[L_DEC0] + n * Kd + (k - H)] : 0.f;
    }
    case TC_DEC1: return p[lo.p_w[L_DEC1] + n * H + k];
    case TC_DEC2: return p[lo.p_w[L_DEC2] + n * H + k];
    default: return n < lo.I ? p[lo.p_w[L_DEC3] + n * H + k] : 0.f;
  }
}

// The arena is cut into segments (one per image / bias row); element idx of a segment:
enum PackSeg { PS_FW = 0, PS_FB, PS_RW, PS_RW_HEADS, PS_RW_DEC0C, PS_RW_DEC0Z, PS_TC, PS_TT, PS_D3C, PS_D3T };

// index of (k, n) inside a tensor-core forward plane: [k-step][k-chunk of 4][n-group of 8][8 n][4 k]
__host__ __device__ inline int tc_plane_index(const TcLayer& c, int k, int n) {
  return (k >> 3) * (c.N * 8) + ((k >> 2) & 1) * (c.N * 4) + (n >> 3) * 32 + (n & 7) * 4 + (k & 3);
}
// index of (k, n) inside a data-gradient plane: [group of gsz steps of n][slice of 32 k][step][2 atoms][4 n]
// [32 k, 32-byte units swizzled by n % 4]
__host__ __device__ inline int tc_tplane_index(const TcLayer& c, int k, int n) {
  const int S = c.Kt / 32, sl = k >> 5, step = n >> 3;
  const int unit = ((k >> 3) & 3) ^ (n & 3);
  const int r = ((n >> 2) & 1) * 128 + (n & 3) * 32 + unit * 8 + (k & 7);
  return (((step / c.gsz) * S + sl) * c.gsz + step % c.gsz) * 256 + r;
}

__host__ __device__ inline void pack_element(const Layout& lo, int type, int l, int idx, const float* __restrict__ p,
                                             float* __restrict__ q) {
  switch (type) {
    case PS_FW: {
      if (l == L_DEC3 && lo.NC > 1) {  // [chunk of 128 outputs][k][128]
        const int c = idx / (H * 128), r = idx - c * (H * 128);
        q[lo.q_w[l] + idx] = fwd_weight(lo, p, l, r / 128, c * 128 + (r % 128));
        break;
      }
      const int Np = lo.Np[l];
      q[lo.q_w[l] + idx] = fwd_weight(lo, p, l, idx / Np, idx % Np);
      break;
    }
    case PS_FB: q[lo.q_b[l] + idx] = fwd_bias(lo, p, l, idx); break;
    case PS_RW: q[lo.r_w[l] + idx] = p[lo.p_w[l] + idx]; break;  // [N][K] plain copy (K == 128 for all of these)
    case PS_RW_HEADS: {
      const int n = idx / (2 * H), k = idx % (2 * H);
      const float w = n < lo.L ? p[lo.p_w[l] + n * 2 * H + k] : p[lo.p_wlv + (n - lo.L) * 2 * H + k];
      if (k < H) q[lo.r_w[l] + n * H + k] = w;
      else q[lo.r_heads_c + n * H + (k - H)] = w;
      break;
    }
    case PS_RW_DEC0C: q[lo.r_w[l] + idx] = p[lo.p_w[l] + (idx / H) * (lo.L + H) + lo.L + (idx % H)]; break;
    case PS_RW_DEC0Z: {
      const int n = idx / lo.Lzp, j = idx % lo.Lzp;
      q[lo.r_dec0z + idx] = j < lo.L ? p[lo.p_w[l] + n * (lo.L + H) + j] : 0.f;
      break;
    }
    case PS_TC: {
      // tensor-core planes: [k-step][k-chunk of 4][n-group of 8][8 n][4 k], high then low halves
      const TcLayer c = lo.tc[l];
      const int per_step = c.N * 8;
      const int ks = idx / per_step, r3 = idx - ks * per_step;
      const int kc = r3 / (c.N * 4), r4 = r3 - kc * (c.N * 4);
      const int n = (r4 >> 5) * 8 + ((r4 & 31) >> 2);
      const int k = ks * 8 + kc * 4 + (r4 & 3);
      const float w = k < c.K ? tc_weight(lo, p, l, k, n) : (k == c.K ? tc_bias(lo, p, l, n) : 0.f);
      const float hi = tf32_rn(w);
      q[c.off_hi + idx] = hi;
      q[c.off_lo + idx] = tf32_rn(w - hi);
      break;
    }
    case PS_D3C: {
      // last decoder layer of a long trajectory: one N = 64 forward image per chunk of 64 outputs, high plane
      // (8192 floats) then low plane
      const int c = idx >> 13, r = idx & 8191;
      const int ks = r >> 9, r3 = r & 511, kc = r3 >> 8, r4 = r3 & 255;
      const int n = c * 64 + (r4 >> 5) * 8 + ((r4 & 31) >> 2);
      const int k = ks * 8 + kc * 4 + (r4 & 3);
      const float w = n < lo.I ? p[lo.p_w[L_DEC3] + n * H + k] : 0.f;
      const float hi = tf32_rn(w);
      q[lo.d3c_off + c * 16384 + r] = hi;
      q[lo.d3c_off + c * 16384 + 8192 + r] = tf32_rn(w - hi);
      break;
    }
    case PS_D3T: {
      // last decoder layer of a long trajectory, training: per chunk of 128 outputs a forward image and a
      // data-gradient image of a 128 x 128 layer (dec3_chunk_layer); idx walks (chunk, k, n)
      const int c = idx >> 14, r = idx & 16383;
      const int k = r >> 7, n = r & 127;
      const int ng = c * 128 + n;
      const float w = ng < lo.I ? p[lo.p_w[L_DEC3] + ng * H + k] : 0.f;
      const float hi = tf32_rn(w), lw = tf32_rn(w - hi);
      const TcLayer t = dec3_chunk_layer(lo, c);
      const int fi = tc_plane_index(t, k, n), ti = tc_tplane_index(t, k, n);
      q[t.off_hi + fi] = hi; q[t.off_lo + fi] = lw;
      q[t.off_thi + ti] = hi; q[t.off_tlo + ti] = lw;
      break;
    }
    default: {
      // data-gradient planes: [group of gsz steps of n][slice of 32 k][step][2 atoms][4 n][32 k, 32-byte units
      // swizzled by n % 4]
      const TcLayer c = lo.tc[l];
      const int S = c.Kt / 32;
      const int blk = idx >> 8, r = idx & 255;
      const int st = blk % c.gsz, t2 = blk / c.gsz;
      const int sl = t2 % S, jg = t2 / S;
      const int n = (jg * c.gsz + st) * 8 + (r >> 7) * 4 + ((r >> 5) & 3);
      const int unit = (r >> 3) & 3, kk = sl * 32 + ((unit ^ (n & 3)) << 3) + (r & 7);
      const float w = kk < c.K ? tc_weight(lo, p, l, kk, n) : 0.f;
      const float hi = tf32_rn(w);
      q[c.off_thi + idx] = hi;
      q[c.off_tlo + idx] = tf32_rn(w - hi);
      break;
    }
  }
}

// The segments of a full repack, in arena order.
struct PackPlan {
  int n;
  int type[64], id[64], count[64], block0[65];
};
constexpr int PACK_THREADS = 256;
inline PackPlan make_pack_plan(const Layout& lo) {
  PackPlan plan;
  plan.n = 0;
  int blocks = 0;
  auto add = [&](int type, int id, int count) {
    if (count <= 0) return;
    assert(plan.n < 64 && "PackPlan holds at most 64 segments");
    plan.type[plan.n] = type; plan.id[plan.n] = id; plan.count[plan.n] = count; plan.block0[plan.n] = blocks;
    blocks += (count + PACK_THREADS - 1) / PACK_THREADS;
    ++plan.n;
  };
  for (int l = 0; l < NUM_LAYERS; ++l) {
    add(PS_FW, l, lo.K[l] * (l == L_DEC3 ? lo.Ipt : lo.Np[l]));
    add(PS_FB, l, l == L_DEC3 ? lo.Ipt : lo.Np[l]);
    if (lo.r_w[l] < 0) continue;
    if (l == L_HEADS) add(PS_RW_HEADS, l, 2 * lo.L * 2 * H);
    else if (l == L_DEC0) { add(PS_RW_DEC0C, l, H * H); add(PS_RW_DEC0Z, l, H * lo.Lzp); }
    else add(PS_RW, l, lo.N[l] * lo.K[l]);
  }
  for (int t = 0; t < NUM_TC; ++t) {
    add(PS_TC, t, lo.tc[t].Kb * lo.tc[t].N);
    if (lo.tc[t].off_thi >= 0) add(PS_TT, t, lo.tc[t].Kt * lo.tc[t].N);
  }
  if (lo.NC > 1) add(PS_D3C, 0, lo.NC64 * 8192);
  if (lo.NC > 1) add(PS_D3T, 0, lo.NC * 16384);
  plan.block0[plan.n] = blocks;
  return plan;
}

// Writes the new value `val` of parameter `e` (offset in the state_dict-ordered arena) to every packed element
// that holds it.  Padding elements never change, so an arena that was packed once stays exact.
__host__ __device__ inline void scatter_param(const Layout& lo, int e, float val, float* __restrict__ q) {
  // which tensor: layer l, weight (n, k) or bias n; the heads layer holds fc_mu then fc_logvar
  int l = NUM_LAYERS - 1;
  while (l > 0 && e < lo.p_w[l]) --l;
  bool is_bias;
  int n, k = 0;
  if (l == L_HEADS) {
    const int L = lo.L;
    if (e < lo.p_b[l]) { is_bias = false; n = (e - lo.p_w[l]) / (2 * H); k = (e - lo.p_w[l]) % (2 * H); }
    else if (e < lo.p_wlv) { is_bias = true; n = e - lo.p_b[l]; }
    else if (e < lo.p_blv) { is_bias = false; n = L + (e - lo.p_wlv) / (2 * H); k = (e - lo.p_wlv) % (2 * H); }
    else { is_bias = true; n = L + (e - lo.p_blv); }
  } else if (e < lo.p_b[l]) {
    is_bias = false; n = (e - lo.p_w[l]) / lo.K[l]; k = (e - lo.p_w[l]) % lo.K[l];
  } else {
    is_bias = true; n = e - lo.p_b[l];
  }
  const TcLayer c = lo.tc[l];   // TcId and LayerId enumerate the layers in the same order
  const bool tc = c.Kb > 0;     // the layer has a tensor-core forward image
  const float hi = tf32_rn(val), lw = tf32_rn(val - hi);
  if (is_bias) {
    // FFMA image: bias row behind the weights (dec3, long trajectories: [chunk][128] = plain index n)
    q[lo.q_b[l] + n] = val;
    if (tc) {  // tensor cores: the bias row is K step K/8 of the forward planes (cond0: row 2 of its one K step)
      const int idx = tc_plane_index(c, l == L_COND0 ? 2 : c.K, n);
      q[c.off_hi + idx] = hi;
      q[c.off_lo + idx] = lw;
    }
    return;
  }
  // FFMA forward image Wt[k][Np]
  if (l == L_DEC3 && lo.NC > 1) q[lo.q_w[l] + (n >> 7) * (H * 128) + k * 128 + (n & 127)] = val;
  else q[lo.q_w[l] + k * lo.Np[l] + n] = val;
  // FFMA data-gradient image W[n][k]
  if (l == L_HEADS) {
    if (k < H) q[lo.r_w[l] + n * H + k] = val;
    else q[lo.r_heads_c + n * H + (k - H)] = val;
  } else if (l == L_DEC0) {
    if (k < lo.L) q[lo.r_dec0z + n * lo.Lzp + k] = val;
    else q[lo.r_w[l] + n * H + (k - lo.L)] = val;
  } else if (lo.r_w[l] >= 0) {
    q[lo.r_w[l] + n * lo.K[l] + k] = val;
  }
  if (l == L_DEC3 && lo.NC > 1) {   // chunked N = 64 images of a long trajectory's last layer
    const int idx = (k >> 3) * 512 + ((k >> 2) & 1) * 256 + ((n & 63) >> 3) * 32 + (n & 7) * 4 + (k & 3);
    q[lo.d3c_off + (n >> 6) * 16384 + idx] = hi;
    q[lo.d3c_off + (n >> 6) * 16384 + 8192 + idx] = lw;
    const TcLayer t = dec3_chunk_layer(lo, n >> 7);   // ... and the training images of the same chunk
    const int fi = tc_plane_index(t, k, n & 127), ti = tc_tplane_index(t, k, n & 127);
    q[t.off_hi + fi] = hi; q[t.off_lo + fi] = lw;
    q[t.off_thi + ti] = hi; q[t.off_tlo + ti] = lw;
    return;
  }
  if (!tc) return;
  // tensor-core planes; dec0 contracts [h_c ; z]
  const int kt = l == L_DEC0 ? (k < lo.L ? H + k : k - lo.L) : k;
  {
    const int idx = tc_plane_index(c, kt, n);
    q[c.off_hi + idx] = hi;
    q[c.off_lo + idx] = lw;
  }
  if (c.off_thi >= 0) {
    const int idx = tc_tplane_index(c, kt, n);
    q[c.off_thi + idx] = hi;
    q[c.off_tlo + idx] = lw;
  }
}

}  // namespace dmvae
