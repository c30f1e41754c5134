// Host-side check (no GPU): scatter_param over every parameter reproduces the full gather repack
// (pack_element over every packed element), bit for bit, for a list of (seq_len, latent_dim).
//   nvcc -std=c++17 -I <csrc> pack_scatter_check.cu -o pack_scatter_check && ./pack_scatter_check T L [T L ...]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dmvae_pack.cuh"

using namespace dmvae;

static int check(int T, int L) {
  DmvaeCfg cfg;
  cfg.seq_len = T; cfg.dim = 3; cfg.latent_dim = L; cfg.hidden_dim = 128;
  Layout lo;
  if (make_layout(&cfg, &lo) != DMVAE_OK) { std::printf("T=%d L=%d: layout rejected\n", T, L); return 1; }
  std::vector<float> p(lo.n_params);
  uint32_t s = 12345u + 977u * T + 31u * L;
  for (auto& v : p) {  // values with low-order mantissa bits, both signs
    s = s * 1664525u + 1013904223u;
    v = ((int)(s >> 8) - (1 << 23)) * (1.0f / (1 << 23)) * 0.37f;
  }
  // gather: the whole arena from scratch (garbage first: every element a segment owns must be written)
  std::vector<float> qg(lo.n_packed, -777.f), qs;
  const PackPlan plan = make_pack_plan(lo);
  for (int sg = 0; sg < plan.n; ++sg)
    for (int idx = 0; idx < plan.count[sg]; ++idx) pack_element(lo, plan.type[sg], plan.id[sg], idx, p.data(), qg.data());
  // scatter onto an arena packed from DIFFERENT parameters: every non-padding element must be overwritten
  std::vector<float> p0(lo.n_params);
  for (int i = 0; i < lo.n_params; ++i) p0[i] = -p[i] * 1.7f + 0.01f;
  qs.assign(lo.n_packed, -777.f);
  for (int sg = 0; sg < plan.n; ++sg)
    for (int idx = 0; idx < plan.count[sg]; ++idx) pack_element(lo, plan.type[sg], plan.id[sg], idx, p0.data(), qs.data());
  for (int e = 0; e < lo.n_params; ++e) scatter_param(lo, e, p[e], qs.data());
  long long bad = 0;
  for (int i = 0; i < lo.n_packed; ++i)
    if (std::memcmp(&qg[i], &qs[i], 4) != 0) {
      if (bad < 5) std::printf("T=%d L=%d: packed[%d] gather %g scatter %g\n", T, L, i, qg[i], qs[i]);
      ++bad;
    }
  std::printf("T=%d L=%d: n_params %d n_packed %d mismatches %lld\n", T, L, lo.n_params, lo.n_packed, bad);
  return bad != 0;
}

int main(int argc, char** argv) {
  int rc = 0;
  for (int i = 1; i + 1 < argc; i += 2) rc |= check(std::atoi(argv[i]), std::atoi(argv[i + 1]));
  return rc;
}
