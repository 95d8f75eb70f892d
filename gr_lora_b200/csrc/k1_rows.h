// k1_rows.h -- launcher of the SF11 / SF12 rows kernel (k1_rows.cuh), compiled in its own translation unit.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace lb {
// Enqueues K1 for n_symbols aligned SF11 / SF12 symbols on `st`.  SF11 writes bins / mags directly.  SF12 merges the two
// CTAs' argmax keys into packed[n_symbols] (must be zeroed on `st` before; finalised by the caller).  tw_host: the host
// copy of the W_sps table (per-q2 constants are taken from it once per device).  Returns 0 or a cudaError_t value with
// a message in err.
int k1_rows_launch(int sf, int device, int n_sms, const float2 *iq, const float2 *chirp, const float2 *tw, const float2 *tw_host,
                   size_t n_symbols, uint32_t *bins, float *mags, unsigned long long *packed, cudaStream_t st, char *err,
                   size_t err_cap);
}  // namespace lb
