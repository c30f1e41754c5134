// dmvae_prof.h - launch accounting (see dmvae_prof.cu).
#pragma once
#include <cuda_runtime.h>

namespace dmvae {

enum KernelId {
  K_PACK = 0, K_DECODE, K_TRAIN_FUSED, K_TRAIN_FWD, K_TRAIN_BWD, K_REDUCE, K_REDUCE_ADAM, K_ADAM, K_LOSS, K_LOSS_GRAD,
  K_FFMA_PROBE, K_DECODE_TC, K_CHAIN, K_WGRAD, K_REDUCE_TC, K_TRAIN_TC_FUSED, K_SPEEDS, K_HISTOGRAM, K_CELLS, K_MPC_PREPARE, K_MPC_TRACK, K_DENSE, K_COUNT
};

// Brackets one kernel launch: counts it and, when profiling is on, records an event pair
// on the launching stream around it.
class ProfScope {
 public:
  ProfScope(int kernel, cudaStream_t stream);
  ~ProfScope();
  ProfScope(const ProfScope&) = delete;
  ProfScope& operator=(const ProfScope&) = delete;

 private:
  int kernel_;
  cudaStream_t stream_;
  int slot_;
};

const char* kernel_name(int id);
void profile_enable(int on);
cudaError_t profile_collect(double* ms, long long* launches, int n);
long long launch_count(int kernel);
cudaError_t launch_ffma_probe(long long iters, float* sink, int sm_count, double* flop, cudaStream_t stream);
// dense tcgen05.mma kind::tf32 issue rate (dmvae_probe_tc.cu); mode 0: operands in shared memory, N = 256; 1: A in tensor memory, N = 128
cudaError_t launch_tf32_probe(long long iters, int mode, float* sink, int sm_count, double* flop, cudaStream_t stream);

}  // namespace dmvae
