// dmvae_api.cu - the C ABI declared in include/dmvae.h.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>

#include "dmvae_common.cuh"
#include "dmvae_launch.h"
#include "dmvae_prof.h"
#include "../../include/dmvae_debug.h"

// count (and, when profiling, time) the kernel launched by `call`
#define PROF(kernel, stream, call) [&] { dmvae::ProfScope _ps((kernel), (stream)); return (call); }()

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(DMVAE_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

struct DeviceInfo {
  int ok = 0;  // 0 unknown, 1 good, -1 bad
  int sm_count = 0;
  int dev = -1;
  char why[256];
};
thread_local DeviceInfo g_dev;

// Every compute entry point goes through this: sm_100 or nothing.
int require_device(int* sm_count) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(DMVAE_ERR_DEVICE, "no CUDA device: %s (libdmvae has no CPU path)", cudaGetErrorString(e));
  if (g_dev.ok == 0 || g_dev.dev != dev) {
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return fail(DMVAE_ERR_DEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    g_dev.dev = dev;
    g_dev.sm_count = prop.multiProcessorCount;
    if (prop.major != 10) {
      g_dev.ok = -1;
      snprintf(g_dev.why, sizeof(g_dev.why), "device %d (%s) is sm_%d%d; libdmvae is built for sm_100a only", dev,
               prop.name, prop.major, prop.minor);
    } else {
      g_dev.ok = 1;
    }
  }
  if (g_dev.ok < 0) return fail(DMVAE_ERR_DEVICE, "%s", g_dev.why);
  if (sm_count) *sm_count = g_dev.sm_count;
  return DMVAE_OK;
}

// Per host thread (include/dmvae.h): a thread that switches kernels for its own calls does not disturb another
// thread's stream.
// 0 = tensor cores (tcgen05, 3xTF32; default), 1 = FP32 FFMA kernel
thread_local int g_decode_impl = 0;
// 0 = tensor cores above 128 rows (default), 1 = FFMA kernels, 2 / 3 = tensor cores at any batch size
thread_local int g_train_impl = 0;
// batches up to this size are latency-bound on either kernel family: the default keeps the FFMA kernels' tighter
// gradients for them (the reference trains on 16..135 rows per step, Training_VAE.py:278)
constexpr long long TRAIN_TC_MIN_ROWS = 129;

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int layout_or_fail(const DmvaeCfg* cfg, dmvae::Layout* lo) {
  if (dmvae::make_layout(cfg, lo) != DMVAE_OK) {
    if (!cfg) return fail(DMVAE_ERR_ARG, "cfg is null");
    return fail(DMVAE_ERR_SHAPE,
                "unsupported configuration seq_len=%d dim=%d latent_dim=%d hidden_dim=%d "
                "(need dim=3, hidden_dim=128, 1<=latent_dim<=64, 2<=seq_len<=400)",
                cfg->seq_len, cfg->dim, cfg->latent_dim, cfg->hidden_dim);
  }
  return DMVAE_OK;
}

}  // namespace

extern "C" {

int dmvae_abi_version(void) { return DMVAE_ABI_VERSION; }
const char* dmvae_last_error(void) { return g_err; }

int dmvae_device_sm_count(void) {
  int n = 0;
  const int rc = require_device(&n);
  return rc == DMVAE_OK ? n : rc;
}

int64_t dmvae_param_count(const DmvaeCfg* cfg) {
  dmvae::Layout lo;
  const int rc = layout_or_fail(cfg, &lo);
  return rc == DMVAE_OK ? lo.n_params : rc;
}

int64_t dmvae_param_offset(const DmvaeCfg* cfg, int index) {
  dmvae::Layout lo;
  const int rc = layout_or_fail(cfg, &lo);
  if (rc != DMVAE_OK) return rc;
  if (index < 0 || index > 24) return fail(DMVAE_ERR_ARG, "tensor index %d out of range [0,24]", index);
  if (index == 24) return lo.n_params;
  // state_dict order: layers 0..5 (w,b), fc_mu (w,b), fc_logvar (w,b), dec0..dec3 (w,b)
  const int pair = index / 2, is_bias = index % 2;
  if (pair < 6) return is_bias ? lo.p_b[pair] : lo.p_w[pair];
  if (pair == 6) return is_bias ? lo.p_b[dmvae::L_HEADS] : lo.p_w[dmvae::L_HEADS];
  if (pair == 7) return is_bias ? lo.p_blv : lo.p_wlv;
  const int l = pair - 1;  // pairs 8..11 -> layers 7..10
  return is_bias ? lo.p_b[l] : lo.p_w[l];
}

int64_t dmvae_packed_count(const DmvaeCfg* cfg) {
  dmvae::Layout lo;
  const int rc = layout_or_fail(cfg, &lo);
  return rc == DMVAE_OK ? lo.n_packed : rc;
}

int dmvae_pack_weights(const DmvaeCfg* cfg, const float* params, float* packed, void* stream) {
  dmvae::Layout lo;
  int rc = layout_or_fail(cfg, &lo);
  if (rc != DMVAE_OK) return rc;
  if (!params || !packed || !aligned16(packed)) return fail(DMVAE_ERR_ARG, "pack_weights: null or misaligned pointer");
  if ((rc = require_device(nullptr)) != DMVAE_OK) return rc;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cudaError_t e = PROF(dmvae::K_PACK, st, dmvae::launch_pack(lo, params, packed, st));
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "pack_weights");
}

int dmvae_decode(const DmvaeCfg* cfg, const float* packed, const float* z, uint64_t seed, uint64_t sample_offset,
                 const float* start, int start_is_shared, float* out, float* z_out, int64_t B, int add_start,
                 void* stream) {
  dmvae::Layout lo;
  int rc = layout_or_fail(cfg, &lo);
  if (rc != DMVAE_OK) return rc;
  if (B < 0) return fail(DMVAE_ERR_ARG, "decode: negative batch");
  if (B == 0) return DMVAE_OK;
  if (!packed || !start || !out || !aligned16(packed)) return fail(DMVAE_ERR_ARG, "decode: null or misaligned pointer");
  int sms = 0;
  if ((rc = require_device(&sms)) != DMVAE_OK) return rc;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (g_decode_impl == 0 && dmvae::decode_tc_supported(lo, start_is_shared != 0))
    e = PROF(dmvae::K_DECODE_TC, st,
             dmvae::launch_decode_tc(lo, start_is_shared != 0, packed, z, seed, sample_offset, start, out, z_out, B,
                                     add_start ? 1 : 0, sms, st));
  else
    e = PROF(dmvae::K_DECODE, st,
             dmvae::launch_decode(lo, start_is_shared ? 1 : 0, packed, z, seed, sample_offset, start, nullptr, nullptr,
                                  out, z_out, B, add_start ? 1 : 0, sms, st));
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "decode");
}

int dmvae_cond_encode(const DmvaeCfg* cfg, const float* packed, const float* start, float* h_c, int64_t B,
                      void* stream) {
  dmvae::Layout lo;
  int rc = layout_or_fail(cfg, &lo);
  if (rc != DMVAE_OK) return rc;
  if (B < 0) return fail(DMVAE_ERR_ARG, "cond_encode: negative batch");
  if (B == 0) return DMVAE_OK;
  if (!packed || !start || !h_c || !aligned16(packed)) return fail(DMVAE_ERR_ARG, "cond_encode: null or misaligned pointer");
  int sms = 0;
  if ((rc = require_device(&sms)) != DMVAE_OK) return rc;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cudaError_t e = PROF(dmvae::K_DECODE, st, dmvae::launch_decode(lo, 3, packed, nullptr, 0, 0, start, nullptr, h_c,
                                                                        nullptr, nullptr, B, 0, sms, st));
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "cond_encode");
}

int dmvae_decode_from_condition(const DmvaeCfg* cfg, const float* packed, const float* z, const float* h_c, float* out,
                                int64_t B, void* stream) {
  dmvae::Layout lo;
  int rc = layout_or_fail(cfg, &lo);
  if (rc != DMVAE_OK) return rc;
  if (B < 0) return fail(DMVAE_ERR_ARG, "decode_from_condition: negative batch");
  if (B == 0) return DMVAE_OK;
  if (!packed || !z || !h_c || !out || !aligned16(packed) || !aligned16(h_c))
    return fail(DMVAE_ERR_ARG, "decode_from_condition: null or misaligned pointer");
  int sms = 0;
  if ((rc = require_device(&sms)) != DMVAE_OK) return rc;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cudaError_t e = PROF(dmvae::K_DECODE, st, dmvae::launch_decode(lo, 2, packed, z, 0, 0, nullptr, h_c, nullptr, out,
                                                                        nullptr, B, 0, sms, st));
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "decode_from_condition");
}


// ---------------------------------------------------------------------------- training
namespace {
struct Workspace {
  float* slabs;
  float* stash;
};
// workspace = [header][grid slabs][stash units], all 16-byte aligned (the header belongs to the tensor-core kernels)
Workspace carve(void* ws, const dmvae::TrainPlan& p) {
  Workspace w;
  w.slabs = static_cast<float*>(ws) + dmvae::WS_HEADER_FLOATS;
  w.stash = w.slabs + (size_t)p.grid * p.slab_stride;
  return w;
}
size_t workspace_floats(const dmvae::TrainPlan& p) {
  return (size_t)p.grid * p.slab_stride + (size_t)p.stash_units * p.stash_stride;
}
}  // namespace

int64_t dmvae_grad_count(const DmvaeCfg* cfg) {
  dmvae::Layout lo;
  const int rc = layout_or_fail(cfg, &lo);
  return rc == DMVAE_OK ? lo.n_params + 5 : rc;
}

int64_t dmvae_train_workspace_bytes(const DmvaeCfg* cfg, int64_t B) {
  dmvae::Layout lo;
  int rc = layout_or_fail(cfg, &lo);
  if (rc != DMVAE_OK) return rc;
  if (B < 0) return fail(DMVAE_ERR_ARG, "train_workspace_bytes: negative batch");
  int sms = 0;
  if ((rc = require_device(&sms)) != DMVAE_OK) return rc;
  const dmvae::TrainPlan p = dmvae::plan_train(lo, B > 0 ? B : 1, sms, false);
  size_t floats = workspace_floats(p);
  if (dmvae::train_tc_supported(lo)) {  // covers both implementations (dmvae_set_train_impl)
    for (int overlap = 0; overlap < 2; ++overlap) {   // either launch scheme (dmvae_set_train_impl 0 / 2)
      const dmvae::TrainTcPlan t = dmvae::plan_train_tc(lo, B > 0 ? B : 1, sms, overlap);
      const size_t tc = t.stash_floats + t.slab_floats + t.loss_floats;
      if (tc > floats) floats = tc;
    }
  }
  return (int64_t)((floats + dmvae::WS_HEADER_FLOATS) * sizeof(float));
}

int64_t dmvae_stash_bytes(const DmvaeCfg* cfg, int64_t B) {
  dmvae::Layout lo;
  int rc = layout_or_fail(cfg, &lo);
  if (rc != DMVAE_OK) return rc;
  if (B < 0) return fail(DMVAE_ERR_ARG, "stash_bytes: negative batch");
  int sms = 0;
  if ((rc = require_device(&sms)) != DMVAE_OK) return rc;
  const dmvae::TrainPlan p = dmvae::plan_train(lo, B > 0 ? B : 1, sms, true);
  return (int64_t)((size_t)p.stash_units * p.stash_stride * sizeof(float));
}

static int train_common(const DmvaeCfg* cfg, const float* packed, const float* x, const float* eps, uint64_t seed,
                        uint64_t sample_offset, uint64_t step, const DmvaeLossWeights* w, float inv_batch, int64_t B,
                        void* workspace, float* grads, const DmvaeAdam* adam, float* params, float* m, float* v,
                        float* packed_rw, void* stream, const char* what, long long* step_dev = nullptr,
                        const DmvaeDpPeers* dp = nullptr, int64_t x_batches = 0, int x_shuffle = 0,
                        uint64_t x_shuffle_seed = 0) {
  dmvae::Layout lo;
  int rc = layout_or_fail(cfg, &lo);
  if (rc != DMVAE_OK) return rc;
  if (B <= 0) return fail(DMVAE_ERR_ARG, "%s: batch must be positive", what);
  if (!packed || !x || !w || !workspace || !grads || !aligned16(packed) || !aligned16(workspace) || !aligned16(grads))
    return fail(DMVAE_ERR_ARG, "%s: null or misaligned pointer", what);
  if (adam && (!params || !m || !v || !aligned16(params) || !aligned16(m) || !aligned16(v)))
    return fail(DMVAE_ERR_ARG, "%s: null or misaligned optimizer state", what);
  int sms = 0;
  if ((rc = require_device(&sms)) != DMVAE_OK) return rc;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  dmvae::TrainIO io;
  io.packed = packed; io.x = x; io.eps = eps; io.x_batches = x_batches;
  io.x_shuffle = x_shuffle; io.x_shuffle_seed = x_shuffle_seed;
  io.seed = seed; io.sample_offset = sample_offset; io.step = step; io.step_dev = step_dev; io.B = B;
  io.w_recon = w->recon; io.w_kld = w->kld; io.w_start = w->start; io.w_time = w->time; io.inv_batch = inv_batch;
  const float wv[4] = {w->recon, w->kld, w->start, w->time};
  cudaError_t e;
  if (step_dev != nullptr && !dmvae::train_tc_supported(lo))
    return fail(DMVAE_ERR_SHAPE, "%s: the device-side step counter needs the tensor-core path (latent_dim <= 32)", what);
  if (dp != nullptr && !dmvae::train_tc_supported(lo))
    return fail(DMVAE_ERR_SHAPE, "%s: the peer-memory exchange needs the tensor-core path (latent_dim <= 32)", what);
  const bool want_tc = g_train_impl >= 2 || (g_train_impl == 0 && B >= TRAIN_TC_MIN_ROWS);
  if ((want_tc || step_dev != nullptr || dp != nullptr) && dmvae::train_tc_supported(lo)) {
    // tensor cores: forward/loss/backward chain -> weight gradients -> partial-slab reduction (+ Adam)
    const dmvae::TrainTcPlan tp = dmvae::plan_train_tc(lo, B, sms);
    // the header (zero at allocation) holds the per-tile counters and the finished-block counter: both reset themselves
    int* flags = static_cast<int*>(workspace);
    float* stash = static_cast<float*>(workspace) + dmvae::WS_HEADER_FLOATS;
    float* slabs = stash + tp.stash_floats;
    float* loss_part = slabs + tp.slab_floats;
    if (tp.overlap) {
      e = PROF(dmvae::K_TRAIN_TC_FUSED, st, dmvae::launch_chain_wgrad_fused(lo, tp, io, stash, slabs, loss_part, flags, st));
      if (e != cudaSuccess) return cuda_fail(e, what);
    } else {
      e = PROF(dmvae::K_CHAIN, st, dmvae::launch_chain(lo, tp, io, stash, loss_part, st));
      if (e != cudaSuccess) return cuda_fail(e, what);
      e = PROF(dmvae::K_WGRAD, st, dmvae::launch_wgrad(lo, tp, stash, slabs, st));
      if (e != cudaSuccess) return cuda_fail(e, what);
    }
    // with the update, the reduction kernel also refreshes `packed` and advances the device-side step counter
    unsigned int* done = reinterpret_cast<unsigned int*>(flags + dmvae::WS_DONE_SLOT);
    e = PROF(dmvae::K_REDUCE_TC, st,
             dmvae::launch_reduce_tc(lo, tp, slabs, loss_part, wv, grads, adam, params, m, v, step_dev, packed_rw, step_dev, done,
                                     dp, flags, st));
    if (e != cudaSuccess) return cuda_fail(e, what);
    return DMVAE_OK;
  } else {
    const dmvae::TrainPlan plan = dmvae::plan_train(lo, B, sms, false);
    const Workspace ws = carve(workspace, plan);
    io.stash = ws.stash; io.slabs = ws.slabs;
    e = PROF(dmvae::K_TRAIN_FUSED, st, dmvae::launch_train(lo, plan, 0, io, st));
    if (e != cudaSuccess) return cuda_fail(e, what);
    e = PROF(adam ? dmvae::K_REDUCE_ADAM : dmvae::K_REDUCE, st,
             dmvae::launch_reduce(lo, ws.slabs, plan.grid, plan.slab_stride, wv, grads, adam, params, m, v, st));
    if (e != cudaSuccess) return cuda_fail(e, what);
  }
  if (adam && packed_rw) {
    e = PROF(dmvae::K_PACK, st, dmvae::launch_pack(lo, params, packed_rw, st, step_dev));
    if (e != cudaSuccess) return cuda_fail(e, what);
  }
  return DMVAE_OK;
}

int dmvae_train_step_dev(const DmvaeCfg* cfg, float* params, float* packed, float* m, float* v, const float* x,
                         const float* eps, uint64_t seed, uint64_t sample_offset, const DmvaeLossWeights* w,
                         float inv_batch, int64_t B, const DmvaeAdam* adam, int64_t* step_dev, void* workspace,
                         float* grads, void* stream) {
  if (!adam || !step_dev) return fail(DMVAE_ERR_ARG, "train_step_dev: adam or step_dev is null");
  if (!packed) return fail(DMVAE_ERR_ARG, "train_step_dev: packed is null");
  return train_common(cfg, packed, x, eps, seed, sample_offset, 0, w, inv_batch, B, workspace, grads, adam, params, m, v,
                      packed, stream, "train_step_dev", reinterpret_cast<long long*>(step_dev));
}

int64_t dmvae_dp_inbox_bytes(const DmvaeCfg* cfg, int world) {
  dmvae::Layout lo;
  const int rc = layout_or_fail(cfg, &lo);
  if (rc != DMVAE_OK) return rc;
  if (world < 1 || world > DMVAE_MAX_PEERS) return fail(DMVAE_ERR_ARG, "dp_inbox_bytes: 1..%d ranks", DMVAE_MAX_PEERS);
  return (int64_t)dmvae::dp_inbox_bytes(lo, world);   // [source | sum][parity][stride] x {step : value}, status word
}
static int check_peers(const DmvaeDpPeers* peers, const char* what) {
  if (!peers || peers->world < 1 || peers->world > DMVAE_MAX_PEERS || peers->rank < 0 || peers->rank >= peers->world)
    return fail(DMVAE_ERR_ARG, "%s: peers must name 1..%d ranks and this rank among them", what, DMVAE_MAX_PEERS);
  if (peers->owned_from != 0 && (peers->owned_from < 2 || peers->owned_from > DMVAE_MAX_PEERS + 1))
    return fail(DMVAE_ERR_ARG, "%s: peers->owned_from must be 0 (default) or 2..%d", what, DMVAE_MAX_PEERS + 1);
  if (peers->timeout_ms < 0) return fail(DMVAE_ERR_ARG, "%s: peers->timeout_ms must not be negative", what);
  for (int p = 0; p < peers->world; ++p)
    if (!peers->inbox[p] || !aligned16(peers->inbox[p]))
      return fail(DMVAE_ERR_ARG, "%s: null or misaligned inbox of rank %d", what, p);
  return DMVAE_OK;
}
int dmvae_dp_status(const DmvaeCfg* cfg, const DmvaeDpPeers* peers, void* stream) {
  dmvae::Layout lo;
  int rc = layout_or_fail(cfg, &lo);
  if (rc != DMVAE_OK) return rc;
  if ((rc = check_peers(peers, "dp_status")) != DMVAE_OK) return rc;
  if ((rc = require_device(nullptr)) != DMVAE_OK) return rc;
  unsigned int word = 0;
  const char* src = static_cast<const char*>(peers->inbox[peers->rank]) + dmvae::dp_status_offset(lo, peers->world);
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemcpyAsync(&word, src, sizeof(word), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(e, "dp_status");
  if (word != 0u)
    return fail(DMVAE_ERR_TIMEOUT, "dp_status: rank %d gave up waiting for a peer's gradients in step %u (peer dead, a step behind, "
                "or the ranks disagree about world / owned_from); its parameters hold NaN", peers->rank, word & 0x7fffffffu);
  return DMVAE_OK;
}
int dmvae_train_step_dp(const DmvaeCfg* cfg, float* params, float* packed, float* m, float* v, const float* x,
                        const float* eps, uint64_t seed, uint64_t sample_offset, const DmvaeLossWeights* w,
                        float inv_batch, int64_t B, const DmvaeAdam* adam, int64_t* step_dev, void* workspace,
                        float* grads, const DmvaeDpPeers* peers, void* stream) {
  if (!adam || !packed) return fail(DMVAE_ERR_ARG, "train_step_dp: adam or packed is null");
  const int prc = check_peers(peers, "train_step_dp");
  if (prc != DMVAE_OK) return prc;
  if (!step_dev && adam->step < 1) return fail(DMVAE_ERR_ARG, "train_step_dp: step must be >= 1");
  return train_common(cfg, packed, x, eps, seed, sample_offset, step_dev ? 0 : (uint64_t)adam->step, w, inv_batch, B,
                      workspace, grads, adam, params, m, v, packed, stream, "train_step_dp",
                      reinterpret_cast<long long*>(step_dev), peers);
}

int64_t dmvae_resident_row(uint64_t shuffle_seed, int64_t epoch, int64_t pos, int64_t n_rows) {
  if (n_rows < 1 || n_rows > 0xffffffffll || pos < 0 || pos >= n_rows || epoch < 0)
    return fail(DMVAE_ERR_ARG, "resident_row: need 0 <= pos < n_rows <= 2^32 - 1 and epoch >= 0");
  return (int64_t)dmvae::resident_row(shuffle_seed, (uint64_t)epoch, (uint32_t)pos, (uint32_t)n_rows);
}

int dmvae_train_step_resident(const DmvaeCfg* cfg, float* params, float* packed, float* m, float* v, const float* x_set,
                              int64_t n_batches, int shuffle, uint64_t shuffle_seed, uint64_t seed, uint64_t sample_offset,
                              const DmvaeLossWeights* w,
                              float inv_batch, int64_t B, const DmvaeAdam* adam, int64_t* step_dev, void* workspace,
                              float* grads, const DmvaeDpPeers* peers, void* stream) {
  if (!adam || !step_dev || !packed) return fail(DMVAE_ERR_ARG, "train_step_resident: adam, step_dev or packed is null");
  if (n_batches < 1) return fail(DMVAE_ERR_ARG, "train_step_resident: the resident set holds at least one batch");
  if (shuffle && (n_batches > 0xffffffffll / (B > 0 ? B : 1)))
    return fail(DMVAE_ERR_ARG, "train_step_resident: a shuffled set holds at most 2^32 - 1 rows");
  if (peers) {
    const int prc = check_peers(peers, "train_step_resident");
    if (prc != DMVAE_OK) return prc;
  }
  return train_common(cfg, packed, x_set, nullptr, seed, sample_offset, 0, w, inv_batch, B, workspace, grads, adam, params, m, v,
                      packed, stream, "train_step_resident", reinterpret_cast<long long*>(step_dev),
                      (peers && peers->world > 1) ? peers : nullptr, n_batches, shuffle ? 1 : 0, shuffle_seed);
}

int dmvae_train_fwd_bwd(const DmvaeCfg* cfg, const float* packed, const float* x, const float* eps, uint64_t seed,
                        uint64_t sample_offset, uint64_t step, const DmvaeLossWeights* w, float inv_batch, int64_t B,
                        void* workspace, float* grads, void* stream) {
  return train_common(cfg, packed, x, eps, seed, sample_offset, step, w, inv_batch, B, workspace, grads, nullptr,
                      nullptr, nullptr, nullptr, nullptr, stream, "train_fwd_bwd");
}

int dmvae_train_step(const DmvaeCfg* cfg, float* params, float* packed, float* m, float* v, const float* x,
                     const float* eps, uint64_t seed, uint64_t sample_offset, const DmvaeLossWeights* w,
                     float inv_batch, int64_t B, const DmvaeAdam* adam, void* workspace, float* grads, void* stream) {
  if (!adam) return fail(DMVAE_ERR_ARG, "train_step: adam is null");
  return train_common(cfg, packed, x, eps, seed, sample_offset, (uint64_t)adam->step, w, inv_batch, B, workspace, grads,
                      adam, params, m, v, packed, stream, "train_step");
}

static int adam_common(const DmvaeCfg* cfg, float* params, const float* grads, float* m, float* v, const DmvaeAdam* adam,
                       long long* step_dev, float* packed, void* stream, const char* what) {
  dmvae::Layout lo;
  int rc = layout_or_fail(cfg, &lo);
  if (rc != DMVAE_OK) return rc;
  if (!params || !grads || !m || !v || !adam || !aligned16(params) || !aligned16(grads) || !aligned16(m) || !aligned16(v))
    return fail(DMVAE_ERR_ARG, "%s: null or misaligned pointer", what);
  if (!step_dev && adam->step < 1) return fail(DMVAE_ERR_ARG, "%s: step must be >= 1", what);
  if ((rc = require_device(nullptr)) != DMVAE_OK) return rc;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = PROF(dmvae::K_ADAM, st, dmvae::launch_adam(lo, params, grads, m, v, *adam, st, step_dev));
  if (e != cudaSuccess) return cuda_fail(e, what);
  if (packed) {
    e = PROF(dmvae::K_PACK, st, dmvae::launch_pack(lo, params, packed, st, step_dev));
    if (e != cudaSuccess) return cuda_fail(e, what);
  }
  return DMVAE_OK;
}

int dmvae_adam_step(const DmvaeCfg* cfg, float* params, const float* grads, float* m, float* v, const DmvaeAdam* adam,
                    float* packed, void* stream) {
  return adam_common(cfg, params, grads, m, v, adam, nullptr, packed, stream, "adam_step");
}

int dmvae_adam_step_dev(const DmvaeCfg* cfg, float* params, const float* grads, float* m, float* v, const DmvaeAdam* adam,
                        int64_t* step_dev, float* packed, void* stream) {
  if (!step_dev || !packed) return fail(DMVAE_ERR_ARG, "adam_step_dev: step_dev and packed are required");
  return adam_common(cfg, params, grads, m, v, adam, reinterpret_cast<long long*>(step_dev), packed, stream, "adam_step_dev");
}

int dmvae_train_fwd_bwd_dev(const DmvaeCfg* cfg, const float* packed, const float* x, const float* eps, uint64_t seed,
                            uint64_t sample_offset, const int64_t* step_dev, const DmvaeLossWeights* w, float inv_batch,
                            int64_t B, void* workspace, float* grads, void* stream) {
  if (!step_dev) return fail(DMVAE_ERR_ARG, "train_fwd_bwd_dev: step_dev is null");
  return train_common(cfg, packed, x, eps, seed, sample_offset, 0, w, inv_batch, B, workspace, grads, nullptr, nullptr,
                      nullptr, nullptr, nullptr, stream, "train_fwd_bwd_dev",
                      const_cast<long long*>(reinterpret_cast<const long long*>(step_dev)));
}

int dmvae_forward(const DmvaeCfg* cfg, const float* packed, const float* x_rel, const float* start, const float* eps,
                  float* recon, float* mu, float* logvar, float* h_c, void* stash, int64_t B, void* stream) {
  dmvae::Layout lo;
  int rc = layout_or_fail(cfg, &lo);
  if (rc != DMVAE_OK) return rc;
  if (B < 0) return fail(DMVAE_ERR_ARG, "forward: negative batch");
  if (B == 0) return DMVAE_OK;
  if (!packed || !x_rel || !start || !eps || !recon || !mu || !logvar || !h_c || !stash || !aligned16(packed) ||
      !aligned16(stash))
    return fail(DMVAE_ERR_ARG, "forward: null or misaligned pointer");
  int sms = 0;
  if ((rc = require_device(&sms)) != DMVAE_OK) return rc;
  const dmvae::TrainPlan plan = dmvae::plan_train(lo, B, sms, true);
  dmvae::TrainIO io;
  io.packed = packed; io.x = x_rel; io.start = start; io.eps = eps; io.stash = static_cast<float*>(stash);
  io.recon = recon; io.mu = mu; io.logvar = logvar; io.hc = h_c; io.B = B;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cudaError_t e = PROF(dmvae::K_TRAIN_FWD, st, dmvae::launch_train(lo, plan, 1, io, st));
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "forward");
}

int dmvae_backward(const DmvaeCfg* cfg, const float* packed, const float* g_recon, const float* g_mu,
                   const float* g_logvar, const float* g_hc, const void* stash, void* workspace, float* grads,
                   int64_t B, void* stream) {
  dmvae::Layout lo;
  int rc = layout_or_fail(cfg, &lo);
  if (rc != DMVAE_OK) return rc;
  if (B <= 0) return fail(DMVAE_ERR_ARG, "backward: batch must be positive");
  if (!packed || !stash || !workspace || !grads || !aligned16(packed) || !aligned16(stash) || !aligned16(workspace) ||
      !aligned16(grads))
    return fail(DMVAE_ERR_ARG, "backward: null or misaligned pointer");
  int sms = 0;
  if ((rc = require_device(&sms)) != DMVAE_OK) return rc;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dmvae::TrainPlan plan = dmvae::plan_train(lo, B, sms, true);
  dmvae::TrainIO io;
  io.packed = packed; io.stash = const_cast<float*>(static_cast<const float*>(stash));
  io.slabs = static_cast<float*>(workspace) + dmvae::WS_HEADER_FLOATS;
  io.g_recon = g_recon; io.g_mu = g_mu; io.g_logvar = g_logvar; io.g_hc = g_hc; io.B = B;
  io.inv_batch = 0.f;  // the KLD / loss seeds arrive through the upstream gradients
  cudaError_t e = PROF(dmvae::K_TRAIN_BWD, st, dmvae::launch_train(lo, plan, 2, io, st));
  if (e != cudaSuccess) return cuda_fail(e, "backward");
  const float wv[4] = {0.f, 0.f, 0.f, 0.f};
  e = PROF(dmvae::K_REDUCE, st,
           dmvae::launch_reduce(lo, io.slabs, plan.grid, plan.slab_stride, wv, grads, nullptr, nullptr, nullptr, nullptr, st));
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "backward(reduce)");
}

int dmvae_loss(const DmvaeCfg* cfg, const float* recon, const float* x, const float* mu, const float* logvar,
               const DmvaeLossWeights* w, int64_t B, float* losses, void* stream) {
  dmvae::Layout lo;
  int rc = layout_or_fail(cfg, &lo);
  if (rc != DMVAE_OK) return rc;
  if (B <= 0) return fail(DMVAE_ERR_ARG, "loss: batch must be positive");
  if (!recon || !x || !mu || !logvar || !w || !losses) return fail(DMVAE_ERR_ARG, "loss: null pointer");
  if ((rc = require_device(nullptr)) != DMVAE_OK) return rc;
  const float wv[4] = {w->recon, w->kld, w->start, w->time};
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cudaError_t e = PROF(dmvae::K_LOSS, st, dmvae::launch_loss(lo, B, recon, x, mu, logvar, wv, losses, st));
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "loss");
}

int dmvae_loss_backward(const DmvaeCfg* cfg, const float* recon, const float* x, const float* mu, const float* logvar,
                        const DmvaeLossWeights* w, int64_t B, const float* g_out, float* g_recon, float* g_mu,
                        float* g_logvar, void* stream) {
  dmvae::Layout lo;
  int rc = layout_or_fail(cfg, &lo);
  if (rc != DMVAE_OK) return rc;
  if (B <= 0) return fail(DMVAE_ERR_ARG, "loss_backward: batch must be positive");
  if (!recon || !x || !mu || !logvar || !w) return fail(DMVAE_ERR_ARG, "loss_backward: null pointer");
  if ((rc = require_device(nullptr)) != DMVAE_OK) return rc;
  const float wv[4] = {w->recon, w->kld, w->start, w->time};
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cudaError_t e = PROF(dmvae::K_LOSS_GRAD, st,
                             dmvae::launch_loss_grad(lo, B, recon, x, mu, logvar, wv, g_out, g_recon, g_mu, g_logvar, st));
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "loss_backward");
}

int dmvae_set_decode_impl(int impl) {
  if (impl != 0 && impl != 1) return fail(DMVAE_ERR_ARG, "set_decode_impl: 0 (tensor cores) or 1 (FFMA)");
  g_decode_impl = impl;
  return DMVAE_OK;
}

int dmvae_set_train_impl(int impl) {
  if (impl < 0 || impl > 3)
    return fail(DMVAE_ERR_ARG, "set_train_impl: 0 (default), 1 (FFMA), 2 (tensor cores, two launches) or 3 (tensor cores)");
  g_train_impl = impl;
  if (impl != 1) dmvae::set_train_tc_overlap(impl != 2);
  return DMVAE_OK;
}

int dmvae_debug_decode_trace(void* device_int64x128) {
  dmvae::set_decode_tc_trace(static_cast<long long*>(device_int64x128));
  return DMVAE_OK;
}

int dmvae_debug_train_trace(void* device_int64x256) {
  dmvae::set_chain_trace(static_cast<long long*>(device_int64x256), 0);
  return DMVAE_OK;
}
int dmvae_debug_train_trace_tile(void* device_int64x256, int tile) {
  dmvae::set_chain_trace(static_cast<long long*>(device_int64x256), tile);
  return DMVAE_OK;
}

// ---------------------------------------------------------------------------- instrumentation
const char* dmvae_kernel_name(int kernel) { return dmvae::kernel_name(kernel); }
int64_t dmvae_launch_count(int kernel) { return dmvae::launch_count(kernel); }

int dmvae_profile_begin(void) {
  const int rc = require_device(nullptr);
  if (rc != DMVAE_OK) return rc;
  dmvae::profile_enable(1);
  return DMVAE_OK;
}

static_assert(DMVAE_KERNEL_COUNT == dmvae::K_COUNT, "include/dmvae.h and dmvae_prof.h disagree on the kernel count");
int dmvae_profile_end(double* ms_by_kernel, int64_t* launches_by_kernel, int n) {
  if (!ms_by_kernel || !launches_by_kernel || n < 1 || n > DMVAE_KERNEL_COUNT)
    return fail(DMVAE_ERR_ARG, "profile_end: need two arrays of 1..%d entries", DMVAE_KERNEL_COUNT);
  long long cnt[DMVAE_KERNEL_COUNT];
  const cudaError_t e = dmvae::profile_collect(ms_by_kernel, cnt, n);
  for (int i = 0; i < n; ++i) launches_by_kernel[i] = cnt[i];
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "profile_end");
}

int dmvae_ffma_probe(int64_t iters, float* sink, double* flop_out, void* stream) {
  if (!sink || iters < 1) return fail(DMVAE_ERR_ARG, "ffma_probe: null sink or iters < 1");
  int sms = 0;
  const int rc = require_device(&sms);
  if (rc != DMVAE_OK) return rc;
  const cudaError_t e = dmvae::launch_ffma_probe(iters, sink, sms, flop_out, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "ffma_probe");
}

// ---------------------------------------------------------------------------- validation metrics
static int check_traj(const float* traj, int64_t n, int32_t T, int32_t layout, const char* what) {
  if (!traj || n < 1 || T < 2 || T > 400 || (layout != 0 && layout != 1))
    return fail(DMVAE_ERR_ARG, "%s: need trajectories (n >= 1, 2 <= seq_len <= 400, 3) and layout 0 ([t, x, y]) or 1 ([x, y, t])", what);
  return DMVAE_OK;
}

int dmvae_waypoint_speeds(const float* traj, int64_t n, int32_t seq_len, int32_t layout, float* speeds, float* minmax, void* stream) {
  int rc = check_traj(traj, n, seq_len, layout, "waypoint_speeds");
  if (rc != DMVAE_OK) return rc;
  if (!speeds || !minmax) return fail(DMVAE_ERR_ARG, "waypoint_speeds: null output");
  if ((rc = require_device(nullptr)) != DMVAE_OK) return rc;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cudaError_t e = PROF(dmvae::K_SPEEDS, st, dmvae::launch_speeds(traj, n, seq_len, layout, speeds, minmax, st));
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "waypoint_speeds");
}

int dmvae_histogram(const float* values, int64_t m, const double* edges, int32_t n_bins, uint64_t* counts, void* stream) {
  if (!values || !edges || !counts || m < 0 || n_bins < 1 || n_bins > 256)
    return fail(DMVAE_ERR_ARG, "histogram: null pointer, negative size or n_bins outside 1..256");
  for (int i = 0; i < n_bins; ++i)
    if (!(edges[i] <= edges[i + 1])) return fail(DMVAE_ERR_ARG, "histogram: edges must increase monotonically");
  int sms = 0;
  const int rc = require_device(&sms);
  if (rc != DMVAE_OK) return rc;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cudaError_t e = PROF(dmvae::K_HISTOGRAM, st,
                             dmvae::launch_histogram(values, m, edges, n_bins, reinterpret_cast<unsigned long long*>(counts), sms, st));
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "histogram");
}

int dmvae_trajectories_per_cell(const float* traj, int64_t n, int32_t seq_len, int32_t layout, double x0, double x_step, int32_t nx_edges,
                                double y0, double y_step, int32_t ny_edges, uint64_t* counts, void* stream) {
  int rc = check_traj(traj, n, seq_len, layout, "trajectories_per_cell");
  if (rc != DMVAE_OK) return rc;
  if (!counts || nx_edges < 2 || ny_edges < 2 || !(x_step > 0.0) || !(y_step > 0.0) || (int64_t)(nx_edges - 1) * (ny_edges - 1) > (1 << 24))
    return fail(DMVAE_ERR_ARG, "trajectories_per_cell: null counts, fewer than two edges, non-positive step or more than 2^24 cells");
  int sms = 0;
  if ((rc = require_device(&sms)) != DMVAE_OK) return rc;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cudaError_t e = PROF(dmvae::K_CELLS, st,
                             dmvae::launch_cells(traj, n, seq_len, layout, x0, x_step, nx_edges, y0, y_step, ny_edges,
                                                 reinterpret_cast<unsigned long long*>(counts), sms, st));
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "trajectories_per_cell");
}

// ---------------------------------------------------------------------------- sub-modules called on their own
int dmvae_dense(const float* weight, const float* bias, const float* x, float* y, int64_t B, int32_t in_features, int32_t out_features,
                int32_t relu, void* stream) {
  if (!weight || !x || !y || B < 1 || in_features < 1 || out_features < 1 || in_features > 4096 || out_features > 4096)
    return fail(DMVAE_ERR_ARG, "dense: null pointer, B < 1 or a feature count outside 1..4096");
  const int rc = require_device(nullptr);
  if (rc != DMVAE_OK) return rc;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cudaError_t e = PROF(dmvae::K_DENSE, st, dmvae::launch_dense(weight, bias, x, y, B, in_features, out_features, relu != 0, st));
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "dense");
}

// ---------------------------------------------------------------------------- batched MPC path tracker
static int check_mpc(const DmvaeMpcCfg* c, const char* what) {
  if (!c) return fail(DMVAE_ERR_ARG, "%s: null cfg", what);
  if (c->n_way < 2 || c->n_way > DMVAE_MPC_MAX_WAY)
    return fail(DMVAE_ERR_SHAPE, "%s: n_way %d outside 2..%d (the reference needs at least two waypoints, MPC_Tracking.py:114-115)",
                what, c->n_way, DMVAE_MPC_MAX_WAY);
  if (c->horizon < 1 || c->horizon > DMVAE_MPC_MAX_HORIZON || c->blocks < 1 || c->blocks > c->horizon)
    return fail(DMVAE_ERR_ARG, "%s: need 1 <= blocks <= horizon <= %d (the reference raises when the control horizon exceeds the prediction horizon)",
                what, DMVAE_MPC_MAX_HORIZON);
  if (!(c->wheelbase > 0.0) || !(c->max_steer > 0.0) || !(c->max_steer < 1.5) || !(c->max_accel > 0.0) || !(c->q_theta >= 0.0) ||
      !(c->q_v >= 0.0) || !(c->r_accel > 0.0) || !(c->r_steer > 0.0) || c->max_iter < 1 || !(c->tol >= 0.0))
    return fail(DMVAE_ERR_ARG, "%s: wheelbase, limits and increment weights must be positive (max_steer < 1.5 rad), max_iter >= 1", what);
  return DMVAE_OK;
}

int64_t dmvae_mpc_workspace_bytes(const DmvaeMpcCfg* cfg, int64_t n) {
  if (check_mpc(cfg, "mpc_workspace_bytes") != DMVAE_OK || n < 1) return -1;
  return (int64_t)dmvae::mpc_workspace_bytes(*cfg, n);
}

int dmvae_mpc_prepare(const DmvaeMpcCfg* cfg, const void* waypoints, const double* initial_state, int64_t n, void* workspace,
                      double* state, int32_t* status, double* profile, void* stream) {
  int rc = check_mpc(cfg, "mpc_prepare");
  if (rc != DMVAE_OK) return rc;
  if (!waypoints || !initial_state || !workspace || !state || !status || n < 1)
    return fail(DMVAE_ERR_ARG, "mpc_prepare: null pointer or n < 1");
  if ((rc = require_device(nullptr)) != DMVAE_OK) return rc;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cudaError_t e = PROF(dmvae::K_MPC_PREPARE, st,
                             dmvae::launch_mpc_prepare(*cfg, waypoints, initial_state, static_cast<double*>(workspace), n, state, status, profile, st));
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "mpc_prepare");
}

int dmvae_mpc_track(const DmvaeMpcCfg* cfg, void* workspace, int64_t n, double dt, const int32_t* n_steps, const int32_t* status,
                    int32_t step_begin, int32_t step_count, double* state, double* states_out, double* controls_out,
                    int64_t out_rows, int32_t* iters_out, void* stream) {
  int rc = check_mpc(cfg, "mpc_track");
  if (rc != DMVAE_OK) return rc;
  if (!workspace || !n_steps || !state || n < 1 || !(dt > 0.0) || step_begin < 0 || step_count < 0)
    return fail(DMVAE_ERR_ARG, "mpc_track: null pointer, n < 1, dt <= 0 or a negative step range");
  if ((states_out || controls_out) && out_rows < (int64_t)step_begin + step_count + 1)
    return fail(DMVAE_ERR_ARG, "mpc_track: out_rows %lld cannot hold the states up to step %d", (long long)out_rows, step_begin + step_count);
  if ((rc = require_device(nullptr)) != DMVAE_OK) return rc;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cudaError_t e = PROF(dmvae::K_MPC_TRACK, st,
                             dmvae::launch_mpc_track(*cfg, static_cast<double*>(workspace), n, dt, n_steps, status, step_begin, step_count, state,
                                                     states_out, controls_out, out_rows, iters_out, st));
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "mpc_track");
}

int dmvae_mpc_windows(const DmvaeMpcCfg* cfg, const void* workspace, int64_t n, double dt, const double* times, int32_t n_times,
                      const int32_t* status, double* out, void* stream) {
  int rc = check_mpc(cfg, "mpc_windows");
  if (rc != DMVAE_OK) return rc;
  if (!workspace || !times || !out || n < 1 || n_times < 1 || n_times > 64 || !(dt > 0.0))
    return fail(DMVAE_ERR_ARG, "mpc_windows: null pointer, n < 1, dt <= 0 or n_times outside 1..64");
  if ((rc = require_device(nullptr)) != DMVAE_OK) return rc;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* times_dev = nullptr;
  cudaError_t e = cudaMallocAsync(&times_dev, sizeof(double) * n_times, st);
  if (e != cudaSuccess) return cuda_fail(e, "mpc_windows");
  e = cudaMemcpyAsync(times_dev, times, sizeof(double) * n_times, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = dmvae::launch_mpc_windows(*cfg, static_cast<const double*>(workspace), n, dt, times_dev, n_times, status, out, st);
  cudaFreeAsync(times_dev, st);
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "mpc_windows");
}

int dmvae_tf32_probe(int64_t iters, int mode, float* sink, double* flop_out, void* stream) {
  if (!sink || iters < 1 || mode < 0 || mode > 3) return fail(DMVAE_ERR_ARG, "tf32_probe: null sink, iters < 1 or mode not 0..3");
  int sms = 0;
  const int rc = require_device(&sms);
  if (rc != DMVAE_OK) return rc;
  const cudaError_t e = dmvae::launch_tf32_probe(iters, mode, sink, sms, flop_out, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "tf32_probe");
}

}  // extern "C"
