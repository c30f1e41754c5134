// dmvae_launch.h - host-side launchers implemented next to each kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dmvae_common.cuh"

namespace dmvae {

// step_inc: optional device counter incremented by the kernel (graph-capturable training step)
cudaError_t launch_pack(const Layout& lo, const float* params, float* packed, cudaStream_t stream,
                        long long* step_inc = nullptr);

// mode: 0 per-row start, 1 shared start, 2 decode from a supplied h_c, 3 condition encoder only
cudaError_t launch_decode(const Layout& lo, int mode, const float* packed, const float* z, uint64_t seed,
                          uint64_t sample_offset, const float* start, const float* hc_in, float* hc_out, float* out,
                          float* z_out, long long B, int add_start, int sm_count, cudaStream_t stream);

// Tensor-core generation kernel (tcgen05 kind::tf32, 3xTF32): per-row or shared start point.
cudaError_t launch_decode_tc(const Layout& lo, bool shared_start, const float* packed, const float* z, uint64_t seed,
                             uint64_t sample_offset, const float* start, float* out, float* z_out, long long B,
                             int add_start, int sm_count, cudaStream_t stream);

bool decode_tc_supported(const Layout& lo, bool shared_start);
void set_decode_tc_trace(long long* device_buffer);  // development aid (128 int64), null = off

// Tiling of one training pass over B rows.
struct TrainPlan {
  int M;                   // rows per tile: 64 or 32
  int stages;              // ring depth
  int grid;                // CTAs = min(n_tiles, SMs) = number of gradient slabs written
  long long n_tiles;
  long long stash_stride;  // floats per stash unit
  long long stash_units;   // grid (fused: per CTA, recycled per tile) or n_tiles (forward/backward pair)
  int slab_stride;         // floats per gradient slab (n_params + loss tail, padded)
  size_t smem;
};
TrainPlan plan_train(const Layout& lo, long long B, int sm_count, bool per_tile_stash);

struct TrainIO {
  const float* packed = nullptr;
  const float* x = nullptr;
  long long x_batches = 0;   // tensor-core path with step_dev: x is a resident set of this many batches (0: one batch)
  int x_shuffle = 0;         // resident set: rows are picked through the per-epoch permutation resident_row (else in storage order)
  unsigned long long x_shuffle_seed = 0;
  const float* start = nullptr;
  const float* eps = nullptr;
  float* stash = nullptr;
  float* slabs = nullptr;
  float* recon = nullptr; float* mu = nullptr; float* logvar = nullptr; float* hc = nullptr;
  const float* g_recon = nullptr; const float* g_mu = nullptr; const float* g_logvar = nullptr; const float* g_hc = nullptr;
  unsigned long long seed = 0, sample_offset = 0, step = 0;
  const long long* step_dev = nullptr;  // tensor-core path: step index in device memory (graph-capturable)
  long long B = 0;
  float w_recon = 0.f, w_kld = 0.f, w_start = 0.f, w_time = 0.f, inv_batch = 0.f;
};
// mode: 0 fused forward+loss+backward, 1 forward only, 2 backward only
cudaError_t launch_train(const Layout& lo, const TrainPlan& plan, int mode, const TrainIO& io, cudaStream_t stream);

// The first WS_HEADER_FLOATS floats of every training workspace are the library's persistent state: the per-tile
// progress counters of train_tc_fused_kernel (reset by the reduction kernel that follows it) and the finished-block
// counter of that reduction (self-resetting).  Zero at allocation (the caller's one duty), never touched by any other
// kernel family, at the same place whatever the batch size or kernel selection - so that no step needs a memset.
constexpr size_t WS_HEADER_FLOATS = 128;
constexpr int WS_DONE_SLOT = 124;   // finished-block counter; the tile counters start at 0 (at most SMs / 4 <= 120 tiles)

// Tensor-core training pass (dmvae_train_tc.cu): chain kernel -> weight-gradient kernel -> reduction.
struct TrainTcPlan {
  long long n_tiles;          // 128-row tiles
  int chain_grid, chain_stages;
  size_t chain_smem;
  int wgrad_grid;
  int role_begin[3], role_count[3];    // CTAs of each role
  int unit_tiles[3], unit_count[3], unit_begin[3];  // tiles per unit, units (= partial slabs) per role, first slab
  int n_slabs;
  int slab_stride;
  size_t stash_floats, slab_floats, loss_floats;   // workspace = [header][stash][slabs][loss partials]
  bool overlap;               // chain and weight-gradient CTAs side by side in one launch (at most SMs / 2 tiles)
  bool streamed;              // ... with fewer weight-gradient CTAs than tiles x roles: each follows several tiles
};
bool train_tc_supported(const Layout& lo);
void set_chain_trace(long long* device_buffer, int tile = 0);  // development aid (256 int64), null = off; which of CTA 0's tiles
TrainTcPlan plan_train_tc(const Layout& lo, long long B, int sm_count, int overlap = -1);  // -1: the current setting
cudaError_t launch_chain(const Layout& lo, const TrainTcPlan& plan, const TrainIO& io, float* stash, float* loss_part,
                         cudaStream_t stream);
cudaError_t launch_wgrad(const Layout& lo, const TrainTcPlan& plan, const float* stash, float* slabs, cudaStream_t stream);
cudaError_t launch_chain_wgrad_fused(const Layout& lo, const TrainTcPlan& plan, const TrainIO& io, float* stash, float* slabs,
                                     float* loss_part, int* flags, cudaStream_t stream);
void set_train_tc_overlap(bool on);  // false: always the two-launch sequence (measurement / debugging)
cudaError_t launch_reduce_tc(const Layout& lo, const TrainTcPlan& plan, const float* slabs, const float* loss_part,
                             const float w[4], float* grads, const DmvaeAdam* adam, float* p, float* m, float* v,
                             const long long* step_dev, float* packed, long long* step_inc, unsigned int* done,
                             const DmvaeDpPeers* dp, int* tile_flags, cudaStream_t stream);
inline int dp_exchange_stride(const Layout& lo) { return round_up(lo.n_params + 5, 4); }
// inbox of a rank: [source | sum][step parity][stride] 8-byte words {step : value}, then 16 bytes whose first word is
// the rank's status (0, or 0x80000000 | step once a thread timed out waiting for a peer)
inline size_t dp_status_offset(const Layout& lo, int world) { return (size_t)(world + 1) * 2 * dp_exchange_stride(lo) * 8; }
inline size_t dp_inbox_bytes(const Layout& lo, int world) { return dp_status_offset(lo, world) + 16; }

// grads = fixed-order sum of the slabs (+ five loss terms); with `adam` also the update.
cudaError_t launch_reduce(const Layout& lo, const float* slabs, int n_slabs, int slab_stride, const float w[4],
                          float* grads, const DmvaeAdam* adam, float* p, float* m, float* v, cudaStream_t stream);
cudaError_t launch_adam(const Layout& lo, float* p, const float* g, float* m, float* v, const DmvaeAdam& a,
                        cudaStream_t stream, const long long* step_dev = nullptr);

cudaError_t launch_loss(const Layout& lo, long long B, const float* recon, const float* x, const float* mu,
                        const float* logvar, const float w[4], float* losses, cudaStream_t stream);
cudaError_t launch_loss_grad(const Layout& lo, long long B, const float* recon, const float* x, const float* mu,
                             const float* logvar, const float w[4], const float* g_out, float* g_recon, float* g_mu,
                             float* g_logvar, cudaStream_t stream);

// validation metrics over waypoint trajectories (dmvae_metrics.cu); layout 0 = [t, x, y], 1 = [x, y, t]
cudaError_t launch_speeds(const float* traj, long long n, int T, int layout, float* vel, float* minmax, cudaStream_t stream);
cudaError_t launch_histogram(const float* values, long long m, const double* edges, int nb, unsigned long long* counts, int sm_count,
                             cudaStream_t stream);
cudaError_t launch_cells(const float* traj, long long n, int T, int layout, double x0, double xstep, int nx, double y0, double ystep,
                         int ny, unsigned long long* counts, int sm_count, cudaStream_t stream);


// one Linear (+ ReLU) layer from the state_dict arena (dmvae_dense.cu): W (out, in) row-major, x (B, in), y (B, out)
cudaError_t launch_dense(const float* W, const float* b, const float* x, float* y, long long B, int in, int out, int relu,
                         cudaStream_t stream);

// batched MPC path tracker (dmvae_mpc.cu)
using MpcCfg = DmvaeMpcCfg;
constexpr int MPC_MAX_WAY = DMVAE_MPC_MAX_WAY, MPC_MAX_HOR = DMVAE_MPC_MAX_HORIZON, MPC_MAX_BLK = DMVAE_MPC_MAX_HORIZON;
size_t mpc_workspace_bytes(const MpcCfg& c, long long n);
cudaError_t launch_mpc_prepare(const MpcCfg& c, const void* way, const double* init, double* ws, long long n, double* state, int* status,
                               double* profile, cudaStream_t stream);
cudaError_t launch_mpc_track(const MpcCfg& c, double* ws, long long n, double dt, const int* n_steps, const int* status, int step_begin,
                             int step_count, double* state, double* states_out, double* controls_out, long long out_rows, int* iters_out,
                             cudaStream_t stream);
cudaError_t launch_mpc_windows(const MpcCfg& c, const double* ws, long long n, double dt, const double* times, int n_times, const int* status,
                               double* out, cudaStream_t stream);

}  // namespace dmvae
