/*
 * dmvae.h - C ABI of the B200-native trajectory-VAE hot path (libdmvae.so).
 *
 * The reference (yslf2035/Defensive-Model-VAE) has no FFI/plugin interface for
 * this path: every call below replaces a chain of PyTorch-CPU ops that the
 * reference issues from Python.  Each entry point cites the reference lines it
 * stands in for (paths relative to the reference root).  The Python drop-in
 * modules (Training_VAE.py, Tools.py at the repo root) bind these symbols with
 * ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned memory (in practice
 *     torch-owned), fp32, contiguous, 16-byte aligned unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     calls are asynchronous with respect to the host and ordered on `stream`;
 *   - return value 0 = success, negative = error (dmvae_last_error() describes
 *     the last error of the calling thread); no exception crosses the ABI;
 *   - the library allocates no persistent device memory: the caller sizes the
 *     workspaces with the *_bytes()/ *_count() queries;
 *   - sm_100 (B200) only.  On any other device every compute call returns
 *     DMVAE_ERR_DEVICE.  There is no CPU path.
 *
 * Shape envelope: dim == 3, hidden_dim == 128,
 * 1 <= latent_dim <= 64, 2 <= seq_len <= 400.  Inside it the tensor-core kernels cover all of
 * generation (latent_dim <= 56 with a per-row start point) and every seq_len with latent_dim <= 32 for
 * training (beyond 3*seq_len = 64 the first encoder / last decoder layer walk the trajectory in chunks of 128
 * features); the rest runs on the FP32 FFMA kernels behind the same entry points.
 */
#ifndef DMVAE_H_
#define DMVAE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMVAE_ABI_VERSION 3

#define DMVAE_OK 0
#define DMVAE_ERR_SHAPE (-1)   /* configuration outside the supported envelope */
#define DMVAE_ERR_ARG (-2)     /* null / misaligned pointer, bad size */
#define DMVAE_ERR_DEVICE (-3)  /* no sm_100 device / wrong device */
#define DMVAE_ERR_CUDA (-4)    /* a CUDA runtime call failed */
#define DMVAE_ERR_TIMEOUT (-5) /* a data-parallel peer did not deliver its gradients in time (dmvae_dp_status) */

/* ConditionalTrajectoryVAE(seq_len, dim, latent_dim, hidden_dim=128)
 * (Training_VAE.py:124-129). */
typedef struct DmvaeCfg {
  int32_t seq_len;
  int32_t dim;
  int32_t latent_dim;
  int32_t hidden_dim;
} DmvaeCfg;

/* conditional_vae_loss weights (Training_VAE.py:229, :300-306). */
typedef struct DmvaeLossWeights {
  float recon;
  float kld;
  float start;
  float time;
} DmvaeLossWeights;

/* torch.optim.Adam hyper-parameters (Training_VAE.py:332; torch optim/adam.py). */
/* Doubles, like the Python floats torch computes its step scalars from. */
typedef struct DmvaeAdam {
  double lr;
  double beta1;
  double beta2;
  double eps;
  int64_t step; /* 1-based index of the update being applied */
} DmvaeAdam;

/* Data-parallel peers of one training job: one process per GPU of a node, every GPU mapping the others'
 * inbox (CUDA peer access over NVLink; torch.distributed's symmetric memory provides the mapping, cudaIpc
 * would do as well).  inbox[p] is rank p's inbox as addressed from THIS device: dmvae_dp_inbox_bytes()
 * bytes, zero before the first step (every rank zeroes its own, then a barrier).
 *   owned_from  smallest world size that exchanges through element owners (0 = default 3; 2 forces the owner
 *               scheme on two ranks, DMVAE_MAX_PEERS + 1 the all-to-all scheme everywhere).  Part of the
 *               peers description because every rank must use the same value.
 *   timeout_ms  how long a thread polls for a peer's word before it gives up (0 = default 2000): the rank
 *               then stores NaN for that element, raises the status word at the end of its own inbox and
 *               finishes the step, so a dead or lagging peer cannot hang the GPU; dmvae_dp_status reports it. */
#define DMVAE_MAX_PEERS 8
typedef struct DmvaeDpPeers {
  int32_t world;
  int32_t rank;
  int32_t owned_from;
  int32_t timeout_ms;
  void* inbox[DMVAE_MAX_PEERS];
} DmvaeDpPeers;

int dmvae_abi_version(void);
const char* dmvae_last_error(void);
/* Number of SMs of the current device, or a negative error. */
int dmvae_device_sm_count(void);

/* ---- layouts ------------------------------------------------------------- */
/* Number of fp32 parameters in state_dict order (Training_VAE.py:132-167):
 * 128942 for seq_len 10, latent 8. */
int64_t dmvae_param_count(const DmvaeCfg* cfg);
/* Offset (in floats) of the i-th state_dict tensor (0..23) in the flat
 * parameter arena; i == 24 returns the total.  Negative on error. */
int64_t dmvae_param_offset(const DmvaeCfg* cfg, int index);
/* Size in floats of the kernel-layout (transposed, padded) weight arena. */
int64_t dmvae_packed_count(const DmvaeCfg* cfg);
/* params (torch layout, dmvae_param_count floats) -> packed (dmvae_packed_count). */
int dmvae_pack_weights(const DmvaeCfg* cfg, const float* params, float* packed, void* stream);

/* ---- generation ------------------------------------------------------------
 * Replaces Tools.py:44-63 (one trajectory) and Tools.py:898-912 (a batch):
 * h = condition_encoder(start); rel = decode(z, h); out = rel with
 * fp32(start) added to the x, y columns when add_start != 0.
 *   packed          kernel-layout weights from dmvae_pack_weights
 *   z               (B, L) latents, or NULL to draw them in-kernel with
 *                   Philox4x32-10 keyed by `seed`, counter = sample_offset + row
 *   start           (B, 2) start points, or (1, 2) when start_is_shared != 0
 *   out             (B, T, 3) [t, x, y]
 *   z_out           optional (B, L): receives the latents actually used */
int dmvae_decode(const DmvaeCfg* cfg, const float* packed, const float* z, uint64_t seed,
                 uint64_t sample_offset, const float* start, int start_is_shared, float* out,
                 float* z_out, int64_t B, int add_start, void* stream);

/* Which kernel dmvae_decode launches (per host thread, like dmvae_set_train_impl): 0 (default) = decode_tc_kernel, the dense layers on the
 * tcgen05 tensor cores as error-compensated 3xTF32 with activations resident in tensor
 * memory; 1 = decode_kernel, everything in FP32 FFMA.  Both hold the 1e-5 tolerance; the
 * switch exists so that bench.py can report the two side by side. */
int dmvae_set_decode_impl(int impl);
/* model.condition_encoder(c) on its own (Training_VAE.py:132-137; called directly
 * at Tools.py:55, :898): start (B,2) -> h_c (B,128). */
int dmvae_cond_encode(const DmvaeCfg* cfg, const float* packed, const float* start, float* h_c, int64_t B,
                      void* stream);
/* model.decode(z, condition) on its own (Training_VAE.py:208-215; called directly
 * at Tools.py:58, :904): z (B,L), h_c (B,128) -> relative trajectories (B,T,3). */
int dmvae_decode_from_condition(const DmvaeCfg* cfg, const float* packed, const float* z, const float* h_c,
                                float* out, int64_t B, void* stream);

/* ---- training ---------------------------------------------------------------
 * dmvae_train_fwd_bwd: one fused pass over a batch - relative-offset transform
 * (Training_VAE.py:345-348), forward (:217-226), five-term loss (:229-268) and
 * the full backward (:362) - followed by the fixed-order reduction of the
 * per-CTA gradient slabs.
 *   packed          kernel-layout weights (dmvae_pack_weights)
 *   x               (B, T, 3) absolute trajectories
 *   eps             (B, L) reparameterisation noise, or NULL for Philox
 *                   (key `seed`, counter = sample_offset + row, stream = step + 1)
 *   inv_batch       1 / (global batch size): the loss is a mean over the
 *                   GLOBAL batch, so data-parallel ranks pass the global size
 *   workspace       dmvae_train_workspace_bytes(cfg, B) bytes, 16-byte aligned, ZERO-FILLED once when it is allocated:
 *                   its first 512 bytes are counters that the kernels reset themselves (no step issues a memset);
 *                   the caller never writes to it afterwards and uses one workspace per stream
 *   grads           out: dmvae_grad_count(cfg) = param_count + 5 floats: the
 *                   gradient of the total loss in state_dict order, then
 *                   [total, recon, kld, start, time] (this rank's share of the
 *                   global means; a SUM all-reduce over ranks completes both)
 * dmvae_train_step: the same followed by the Adam update in the reduction
 * kernel and a refresh of `packed`: the whole single-GPU step
 * (Training_VAE.py:351-363) in three launches, no host synchronisation. */
int64_t dmvae_grad_count(const DmvaeCfg* cfg);
int64_t dmvae_train_workspace_bytes(const DmvaeCfg* cfg, int64_t B);
/* Which kernels the fused training pass launches (a per-THREAD setting: it applies to the calls the calling
 * host thread makes afterwards, so threads driving different streams do not disturb each other):
 *   0 (default)  tensor cores (tcgen05, 3xTF32: chain_kernel + wgrad_kernel + reduce_tc_kernel) inside their
 *                envelope for batches of more than 128 rows; the FP32 FFMA kernels otherwise - the
 *                reference's own batch sizes (16..135 rows, Training_VAE.py:278) are latency-bound either
 *                way and keep the FFMA kernels' 1e-6-class gradients;
 *   1            always the FP32 FFMA kernels (train_kernel + reduce_kernel);
 *   2            tensor cores at any batch size, always as two launches;
 *   3            tensor cores at any batch size.
 * With 0 and 3, batches of at most (SMs / 4) * 128 rows run the chain and the weight-gradient CTAs side by
 * side in ONE launch (train_tc_fused_kernel): a weight-gradient CTA starts on an image of its tile as soon
 * as the chain has written it.  The results of 2 and 3 are bit-identical.  The entry points with a
 * device-side step counter or data-parallel peers always run the tensor-core kernels.  All hold the
 * tolerances of tests/test_train_gpu.py; dmvae_train_workspace_bytes covers every setting. */
int dmvae_set_train_impl(int impl);
int dmvae_train_fwd_bwd(const DmvaeCfg* cfg, const float* packed, const float* x, const float* eps,
                        uint64_t seed, uint64_t sample_offset, uint64_t step, const DmvaeLossWeights* w,
                        float inv_batch, int64_t B, void* workspace, float* grads, void* stream);
int dmvae_train_step(const DmvaeCfg* cfg, float* params, float* packed, float* m, float* v,
                     const float* x, const float* eps, uint64_t seed, uint64_t sample_offset,
                     const DmvaeLossWeights* w, float inv_batch, int64_t B, const DmvaeAdam* adam,
                     void* workspace, float* grads, void* stream);
/* dmvae_train_step with the Adam step index in device memory, so that every launch parameter is
 * step-independent and the whole call can be captured in a CUDA graph and replayed (no per-step host
 * work: Training_VAE.py:351-363 as one graph launch).  *step_dev = number of updates applied so far;
 * the call applies update *step_dev + 1 (adam->step is ignored; the bias corrections are derived on the
 * device in double) and its last kernel increments the counter.  The Philox stream of the step is
 * *step_dev + 2, as in dmvae_train_step.  Tensor-core path only: DMVAE_ERR_SHAPE outside its envelope. */
int dmvae_train_step_dev(const DmvaeCfg* cfg, float* params, float* packed, float* m, float* v,
                         const float* x, const float* eps, uint64_t seed, uint64_t sample_offset,
                         const DmvaeLossWeights* w, float inv_batch, int64_t B, const DmvaeAdam* adam,
                         int64_t* step_dev, void* workspace, float* grads, void* stream);
/* dmvae_train_step_dev over a data set that is resident in device memory (the reference's
 * DataLoader hands out one batch per step on the host, reshuffled every epoch, Training_VAE.py:327): x_set holds
 * n_batches batches of B rows back to back; update t belongs to epoch e = (t - 1) / n_batches and is batch
 * b = (t - 1) mod n_batches of it, both derived in the kernel from the device-side step counter - an epoch is
 * n_batches replays of one captured graph with no per-step copy or host work.
 *   shuffle == 0   batch b is rows [b B, (b + 1) B) of the set, every epoch;
 *   shuffle != 0   row r of batch b is row pi(b B + r) of the set, pi = the permutation of [0, n_batches B) keyed
 *                  by (shuffle_seed, e): every row exactly once per epoch, a new order every epoch, no permutation
 *                  array anywhere (dmvae_resident_row evaluates the same function on the host).
 * peers: NULL, or the data-parallel peers of dmvae_train_step_dp (every rank walks its own resident shard with the
 * same function of (shuffle_seed, epoch, position): the order does not depend on the rank or the world size).
 * Tensor-core path only. */
int64_t dmvae_resident_row(uint64_t shuffle_seed, int64_t epoch, int64_t pos, int64_t n_rows);
int dmvae_train_step_resident(const DmvaeCfg* cfg, float* params, float* packed, float* m, float* v,
                              const float* x_set, int64_t n_batches, int shuffle, uint64_t shuffle_seed,
                              uint64_t seed, uint64_t sample_offset,
                              const DmvaeLossWeights* w, float inv_batch, int64_t B, const DmvaeAdam* adam,
                              int64_t* step_dev, void* workspace, float* grads, const DmvaeDpPeers* peers,
                              void* stream);
/* The two halves of the data-parallel step in the same graph-capturable form: forward + loss + backward
 * with the Philox stream taken from *step_dev + 2 (the counter is only read), and the Adam update for
 * step *step_dev + 1 followed by the repack, whose kernel increments the counter.  Between them the
 * caller all-reduces `grads` (NCCL collectives can be captured in the same graph). */
int dmvae_train_fwd_bwd_dev(const DmvaeCfg* cfg, const float* packed, const float* x, const float* eps,
                            uint64_t seed, uint64_t sample_offset, const int64_t* step_dev,
                            const DmvaeLossWeights* w, float inv_batch, int64_t B, void* workspace,
                            float* grads, void* stream);
int dmvae_adam_step_dev(const DmvaeCfg* cfg, float* params, const float* grads, float* m, float* v,
                        const DmvaeAdam* adam, int64_t* step_dev, float* packed, void* stream);
/* The whole data-parallel step without a library collective (replaces loss.backward() + the gradient
 * all-reduce the north star adds + optimizer.step(), Training_VAE.py:351-363): the fused pass on this rank's
 * rows (inv_batch = 1 / global batch), then ONE kernel in which every thread sums the partial slabs of its
 * gradient, writes {value, step index} as one 8-byte word into its peers' inboxes over NVLink, polls its own
 * inbox for the words of the same step, adds the ranks in rank order (up to 2 ranks: everybody adds; more:
 * the element's owner adds and pushes the sum to all - bit-identical replicas either way), applies Adam
 * and refreshes `packed`.  No fence, flag, grid-wide or job-wide barrier.  grads
 * receives the global gradient and the five loss terms of the global batch.  step_dev may be NULL
 * (host-driven: the step is adam->step) or the device-side counter of dmvae_train_step_dev
 * (graph-capturable).  The inbox is double-buffered by step parity and every word carries its step, so
 * steps need no reset; all ranks must call with the same step.  Tensor-core path only. */
int64_t dmvae_dp_inbox_bytes(const DmvaeCfg* cfg, int world);
/* Reads the status word of this rank's inbox (synchronises `stream`): DMVAE_OK, or DMVAE_ERR_TIMEOUT when a
 * thread of an earlier step gave up waiting for a peer (the parameters of this rank then hold NaN). */
int dmvae_dp_status(const DmvaeCfg* cfg, const DmvaeDpPeers* peers, void* stream);
int dmvae_train_step_dp(const DmvaeCfg* cfg, float* params, float* packed, float* m, float* v,
                        const float* x, const float* eps, uint64_t seed, uint64_t sample_offset,
                        const DmvaeLossWeights* w, float inv_batch, int64_t B, const DmvaeAdam* adam,
                        int64_t* step_dev, void* workspace, float* grads, const DmvaeDpPeers* peers,
                        void* stream);
/* optimizer.step() of torch.optim.Adam (Training_VAE.py:363; torch
 * optim/adam.py::_single_tensor_adam): updates params, m, v in place from
 * grads (dmvae_param_count floats each) and refreshes `packed` (may be NULL).
 * Data-parallel training calls this after the gradient all-reduce. */
int dmvae_adam_step(const DmvaeCfg* cfg, float* params, const float* grads, float* m, float* v,
                    const DmvaeAdam* adam, float* packed, void* stream);

/* ---- unfused pieces behind the nn.Module / autograd surface ---------------
 * dmvae_forward: model.forward (Training_VAE.py:217-226) on already-relative
 * trajectories: x_rel (B,T,3), start (B,2), eps (B,L) -> recon (B,T,3), mu,
 * logvar (B,L), h_c (B,128).  Activations needed by dmvae_backward are kept
 * in `stash` (dmvae_stash_bytes(cfg, B) bytes). */
int64_t dmvae_stash_bytes(const DmvaeCfg* cfg, int64_t B);
int dmvae_forward(const DmvaeCfg* cfg, const float* packed, const float* x_rel, const float* start,
                  const float* eps, float* recon, float* mu, float* logvar, float* h_c, void* stash,
                  int64_t B, void* stream);
/* dmvae_backward: what autograd computes for the module's parameters given the
 * upstream gradients g_recon (B,T,3), g_mu, g_logvar (B,L), g_hc (B,128) (any
 * may be NULL = zero).  workspace: dmvae_train_workspace_bytes(cfg, B) bytes.
 * grads: dmvae_grad_count floats (the 5-float tail is zero). */
int dmvae_backward(const DmvaeCfg* cfg, const float* packed, const float* g_recon, const float* g_mu,
                   const float* g_logvar, const float* g_hc, const void* stash, void* workspace,
                   float* grads, int64_t B, void* stream);
/* conditional_vae_loss (Training_VAE.py:229-268): losses[5] = total, recon,
 * kld, start, time (a term whose weight is <= 0 is reported as 0 and left out
 * of the total, as the reference does). */
int dmvae_loss(const DmvaeCfg* cfg, const float* recon, const float* x, const float* mu,
               const float* logvar, const DmvaeLossWeights* w, int64_t B, float* losses, void* stream);
/* its backward for upstream gradients g_out[5] of the five outputs (device
 * pointer, NULL = d(total) = 1): g_recon (B,T,3), g_mu, g_logvar (B,L). */
int dmvae_loss_backward(const DmvaeCfg* cfg, const float* recon, const float* x, const float* mu,
                        const float* logvar, const DmvaeLossWeights* w, int64_t B, const float* g_out,
                        float* g_recon, float* g_mu, float* g_logvar, void* stream);

/* ---- validation metrics over generated waypoint trajectories -------------------
 * The reference compares generated with human trajectories through two distributions: waypoint
 * speeds (Distribution.py:248-296) summarised as a Jensen-Shannon divergence over 50 common histogram edges (:309-331),
 * and the number of trajectories that visit each cell of a scenario grid (Spatial_Distribution.py:387-431) summarised as
 * an RMSE (:434-493).  These are the per-trajectory passes for the 10^6 trajectories the generation kernel produces; the
 * final arithmetic on 49 counts / one count map is the caller's (dmvae/validation.py).
 *   traj      (n, seq_len, 3) fp32, layout 0 = [t, x, y] (what dmvae_decode writes), 1 = [x, y, t] (the reference's order)
 * dmvae_waypoint_speeds: speeds (n * seq_len): per trajectory the speed of every step, the last point repeating the last
 *   step; a step whose time difference is not above 1e-6 repeats the value before it in the flattened array (0 at the very
 *   beginning) - the reference's loop, evaluated without its sequential dependency.  minmax (2 floats, device): the
 *   smallest and largest speed.  fp32 arithmetic as the reference's NumPy scalars (within one unit in the last place).
 * dmvae_histogram: np.histogram(values, bins=edges) for n_bins <= 256 increasing edges (host array of n_bins + 1
 *   doubles); counts (device, n_bins x uint64) is overwritten.
 * dmvae_trajectories_per_cell: edges x0 + i * x_step (i < nx_edges, np.arange) and likewise in y; counts (device,
 *   (ny_edges - 1) x (nx_edges - 1) uint64, row = y cell) is overwritten with the number of trajectories that have at
 *   least one point in the cell; points outside the grid fall into its border cells (np.clip, as the reference). */
int dmvae_waypoint_speeds(const float* traj, int64_t n, int32_t seq_len, int32_t layout, float* speeds, float* minmax,
                          void* stream);
int dmvae_histogram(const float* values, int64_t m, const double* edges, int32_t n_bins, uint64_t* counts, void* stream);
int dmvae_trajectories_per_cell(const float* traj, int64_t n, int32_t seq_len, int32_t layout, double x0, double x_step,
                                int32_t nx_edges, double y0, double y_step, int32_t ny_edges, uint64_t* counts, void* stream);

/* ---- sub-modules called on their own -------------------------------------------
 * The reference's sub-modules are plain nn.Sequential / nn.Linear objects (Training_VAE.py:141-167), so
 * model.encoder(x), model.decoder(zc), model.fc_mu(h), model.fc_logvar(h) are callable.  No reference caller does that
 * (encode / decode / forward run in the fused kernels); for completeness of the module surface one layer at a time:
 * y (B, out) = act(x (B, in) * weight^T + bias), weight (out, in) row-major as nn.Linear stores it (a view into the
 * parameter arena), relu != 0 applies max(., 0).  FP32 FFMA.  Forward only. */
int dmvae_dense(const float* weight, const float* bias, const float* x, float* y, int64_t B, int32_t in_features,
                int32_t out_features, int32_t relu, void* stream);

/* ---- batched MPC path tracker ------------------------------------------------
 * The reference turns each generated waypoint set into a driven trajectory with PathTracker (MPC/MPC_Tracking.py:418-523,
 * called from Distribution.py:91-105): PathInterpolator (:89-277) -> per time step a reference window (:464-478), one
 * solve of the controller's optimal control problem (MPCController.solve_mpc, :311-415) and one Euler step of the bicycle
 * model with the first control (:484-486).  These entry points do the same for n independent trajectories, one GPU thread
 * each, in float64.  The optimisation problem is the reference's (weights Q = diag(20, 5) on heading and speed over
 * horizon + 1 rows, R = diag(1, 50) on the control increments over `blocks` rows, the last row held for the rest of the
 * horizon, the effective bounds of :390-398 with the inequality constraint of :376-387); it is solved to convergence by a
 * control-limited DDP / Newton method where the reference runs SLSQP to ftol = 1e-6, so results agree with the
 * reference's to its early-stopping noise, not bit for bit (statistical parity).
 *
 *   waypoints      (n, n_way, 3) [x, y, t], float32 (way_f32 = 1: the VAE's output; the knot arithmetic that the reference
 *                  does in float32 is done in float32) or float64; 2 <= n_way <= 64 (cubic interpolants from four waypoints on,
 *                  one parabola for three, a line for two: MPC_Tracking.py:126-137)
 *   initial_state  (n, 5) float64 [x, y, theta, vx, vy] (Distribution.py:80)
 *   workspace      dmvae_mpc_workspace_bytes(cfg, n) bytes: interpolants, previous control and previous solution of every
 *                  trajectory; written by dmvae_mpc_prepare, carried from one dmvae_mpc_track call to the next
 *   state          (n, 4) float64 [x, y, theta, v]: written by prepare (PathTracker.__init__, :435-441), advanced by track
 *   status         (n) int32, written by prepare: 0 ok, 1 waypoint times do not increase strictly (the reference raises
 *                  ValueError, :118-119) or are not finite / beyond 1e6 s; such trajectories are skipped by track
 *   profile        optional (n, 5) float64 [start_theta, end_vx, end_vy, end_theta, t_end] (:194-221), or NULL
 * dmvae_mpc_track runs steps [step_begin, step_begin + step_count) of every trajectory j that has them (s < n_steps[j],
 * n_steps = int(total_time / dt) per trajectory, :505).  states_out (n, out_rows, 4) float64 or NULL: row s + 1 = state
 * after step s, row 0 = the initial state (written when step_begin = 0); controls_out (n, out_rows - 1, 2) or NULL: the
 * applied (a, delta) of step s; rows of steps a trajectory does not have are left untouched.  iters_out (n) int32 or NULL:
 * solver iterations summed over the trajectory's steps.
 * dmvae_mpc_windows: the reference windows [theta_ref, v_ref] that step() would build at the given current times (host
 * array of n_times doubles): out (n, n_times, horizon + 1, 2) float64. */
#define DMVAE_MPC_MAX_WAY 64
#define DMVAE_MPC_MAX_HORIZON 40
typedef struct DmvaeMpcCfg {
  int32_t n_way;
  int32_t way_f32;
  int32_t horizon;   /* prediction_horizon (Distribution.py:98: 30), <= DMVAE_MPC_MAX_HORIZON */
  int32_t blocks;    /* control_horizon (:99: 20), <= horizon */
  int32_t max_iter;  /* solver iterations per controller call at most (50) */
  int32_t reserved;
  double wheelbase, max_steer, max_accel;     /* 2.8, 0.5, 7.0 (MPC_Tracking.py:26) */
  double q_theta, q_v, r_accel, r_steer;      /* 20, 5, 1, 50 (:304-306) */
  double tol;                                 /* stop when no control changed by more than this in the last iteration (1e-6; quadratic convergence: the controls are then good to ~1e-11) */
} DmvaeMpcCfg;
int64_t dmvae_mpc_workspace_bytes(const DmvaeMpcCfg* cfg, int64_t n);
int dmvae_mpc_prepare(const DmvaeMpcCfg* cfg, const void* waypoints, const double* initial_state, int64_t n, void* workspace,
                      double* state, int32_t* status, double* profile, void* stream);
int dmvae_mpc_track(const DmvaeMpcCfg* cfg, void* workspace, int64_t n, double dt, const int32_t* n_steps, const int32_t* status,
                    int32_t step_begin, int32_t step_count, double* state, double* states_out, double* controls_out,
                    int64_t out_rows, int32_t* iters_out, void* stream);
int dmvae_mpc_windows(const DmvaeMpcCfg* cfg, const void* workspace, int64_t n, double dt, const double* times, int32_t n_times,
                      const int32_t* status, double* out, void* stream);

/* ---- instrumentation -------------------------------------------------------
 * Nothing in the reference corresponds to these: they let bench.py report what the
 * library launched and how long the dominant kernel ran inside the timed region.
 * Kernel ids: 0 pack, 1 decode, 2 train (fused), 3 train (forward), 4 train
 * (backward), 5 reduce, 6 reduce+Adam, 7 Adam, 8 loss, 9 loss backward, 10 FFMA probe,
 * 11 decode (tensor cores), 12 train chain (tensor cores), 13 weight gradients (tensor
 * cores), 14 partial-slab reduction (+ Adam), 15 chain + weight gradients in one launch
 * (small batches), 16 waypoint speeds, 17 histogram, 18 trajectories per grid cell, 19 MPC tracker
 * set-up, 20 MPC tracker steps, 21 one Linear layer on its own. */
#define DMVAE_KERNEL_COUNT 22
const char* dmvae_kernel_name(int kernel);
/* Kernels launched by this process since the library was loaded (kernel < 0: all). */
int64_t dmvae_launch_count(int kernel);
/* Between begin and end every launch is bracketed by a cudaEvent pair on its stream;
 * end synchronises on them and returns, per kernel id, the summed duration in
 * milliseconds and the number of launches (arrays of n <= DMVAE_KERNEL_COUNT). */
int dmvae_profile_begin(void);
int dmvae_profile_end(double* ms_by_kernel, int64_t* launches_by_kernel, int n);
/* FP32 roofline probe: a pure FFMA kernel over every SM (16 independent chains per
 * thread, 2 x 1024 threads per SM); *flop_out receives the FLOPs it executes.  Timed
 * by the caller (or through the profile calls) it yields the FFMA peak the fused
 * kernels are measured against.  sink: any device float. */
int dmvae_ffma_probe(int64_t iters, float* sink, double* flop_out, void* stream);
/* Tensor-core roofline probe: every SM issues iters x 16 dense tcgen05.mma kind::tf32 products (M = 128, K = 8 per
 * instruction) on resident operands; *flop_out receives the FLOPs.  mode 0: both operands in shared memory, N = 256;
 * mode 1: A in tensor memory, N = 128 (the shape the training chain issues); modes 2 / 3: one M = 128, N = 128 product
 * per PAIR of CTAs (cta_group::2, 64 rows per SM) with A in tensor / shared memory.  Timed by the caller it yields the
 * MEASURED dense TF32 rate; the fp32-equivalent ceiling of the 3xTF32 kernels is one third of it. */
int dmvae_tf32_probe(int64_t iters, int mode, float* sink, double* flop_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DMVAE_H_ */
