// k1_packed.h -- launchers of the K1 kernels that are built with the packed complex product (k1_packed.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace lb {
// SF7: k1_sf7_warp_kernel<12,2>; SF9: k1_group_kernel<9,3,2>.  Returns 0 or a cudaError_t value with a message in err.
int k1_packed_launch(int sf, int device, int n_sms, const float2 *iq, const float2 *chirp, const float2 *tw, size_t n_symbols,
                     uint32_t *bins, float *mags, cudaStream_t st, char *err, size_t err_cap);
}  // namespace lb
