// dmvae_launch.h - host-side launchers implemented next to each kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dmvae_common.cuh"

namespace dmvae {

cudaError_t launch_pack(const Layout& lo, const float* params, float* packed, cudaStream_t stream);

// mode: 0 per-row start, 1 shared start, 2 decode from a supplied h_c, 3 condition encoder only
cudaError_t launch_decode(const Layout& lo, int mode, const float* packed, const float* z, uint64_t seed,
                          uint64_t sample_offset, const float* start, const float* hc_in, float* hc_out, float* out,
                          float* z_out, long long B, int add_start, int sm_count, cudaStream_t stream);

}  // namespace dmvae
