/*
 * dmvae_debug.h - development aids of libdmvae.so.  NOT part of the drop-in boundary (include/dmvae.h):
 * nothing in the reference corresponds to these, no product code calls them; scripts/trace_*.py and
 * tests/dev use them to time the inside of the persistent kernels on a B200.
 * Process-global, not thread-safe: set them while no training / generation call is in flight.
 */
#ifndef DMVAE_DEBUG_H_
#define DMVAE_DEBUG_H_

#ifdef __cplusplus
extern "C" {
#endif

/* Device buffer of 128 int64 that CTA 0 of decode_tc_kernel fills with clock64 stamps per layer step
 * (NULL = off, the default). */
int dmvae_debug_decode_trace(void* device_int64x128);
/* Same for chain_kernel (first tile of CTA 0): 256 int64; [4 o + 0 / 1] = MMA warp starts / has issued
 * op o, [128 + 2 e + 0 / 1] = epilogue e starts (accumulator complete) / has released the A operand;
 * from 176: %globaltimer stamps of the launch, the weight-gradient roles and the reduction kernel. */
int dmvae_debug_train_trace(void* device_int64x256);
/* The same for the tile-th tile that CTA 0 walks (0 = its first): steady-state timing of a persistent CTA. */
int dmvae_debug_train_trace_tile(void* device_int64x256, int tile);

#ifdef __cplusplus
}
#endif
#endif /* DMVAE_DEBUG_H_ */
