// dmvae_api.cu - the C ABI declared in include/dmvae.h.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "dmvae_common.cuh"
#include "dmvae_launch.h"

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

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int layout_or_fail(const DmvaeCfg* cfg, dmvae::Layout* lo) {
  if (dmvae::make_layout(cfg, lo) != DMVAE_OK) {
    if (!cfg) return fail(DMVAE_ERR_ARG, "cfg is null");
    return fail(DMVAE_ERR_SHAPE,
                "unsupported configuration seq_len=%d dim=%d latent_dim=%d hidden_dim=%d "
                "(need dim=3, hidden_dim=128, 1<=latent_dim<=64, 2<=seq_len, 3*seq_len<=128)",
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
  const cudaError_t e = dmvae::launch_pack(lo, params, packed, static_cast<cudaStream_t>(stream));
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
  const cudaError_t e =
      dmvae::launch_decode(lo, start_is_shared ? 1 : 0, packed, z, seed, sample_offset, start, nullptr, nullptr, out,
                           z_out, B, add_start ? 1 : 0, sms, static_cast<cudaStream_t>(stream));
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
  const cudaError_t e = dmvae::launch_decode(lo, 3, packed, nullptr, 0, 0, start, nullptr, h_c, nullptr, nullptr, B, 0,
                                             sms, static_cast<cudaStream_t>(stream));
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
  const cudaError_t e = dmvae::launch_decode(lo, 2, packed, z, 0, 0, nullptr, h_c, nullptr, out, nullptr, B, 0, sms,
                                             static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? DMVAE_OK : cuda_fail(e, "decode_from_condition");
}

}  // extern "C"
