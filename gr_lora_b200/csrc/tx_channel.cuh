// tx_channel.cuh -- synthetic transmitter and channel on the device.
//
// The reference is a receiver only (its README points at a hardware transmitter), so this has no reference counterpart; it
// is what tests and the benchmark need to put known LoRa traffic into HBM without a pass over PCIe:
//   * tx_symbols_kernel   aligned data symbols: out[s][n] = up[(n + decim * value[s]) mod sps] * e^{j 2 pi cfo[s] n / fs}
//                         + sigma (N(0,1) + j N(0,1)) -- the cyclic-shift modulator of gr_lora_b200/tx.py::modulate_shifts
//                         (bit-identical to it when given the same chirp table and no noise / CFO);
//   * tx_expand_kernel    a batch of concurrent channels from K base captures: out[s] = base[s mod K] + the stream's own
//                         noise (64 channels that replay one capture would correlate perfectly).
// Noise: Philox4x32-10 keyed by the seed, counter = (sample pair, row), Box-Muller on the four 32-bit outputs -- a
// counter-based generator, so the result does not depend on the launch geometry.  Both kernels are HBM-write bound.
#pragma once
#include "lora_common.cuh"

namespace lb {
#ifdef __CUDACC__
LB_HD uint32_t philox_mulhi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((unsigned long long)a * b) >> 32);
#endif
}
LB_HD void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = philox_mulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = philox_mulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"): c <- ten rounds under key (k0, k1)
LB_HD void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
// four normal deviates for (row, pair index i) under `seed`
LB_D float4 philox_normal4(unsigned long long seed, unsigned long long row, unsigned long long i) {
    uint32_t c[4] = {(uint32_t)i, (uint32_t)(i >> 32), (uint32_t)row, (uint32_t)(row >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float s = 2.3283064365386963e-10f;                    // 2^-32
    const float u0 = ((float)c[0] + 0.5f) * s, u1 = (float)c[1] * s, u2 = ((float)c[2] + 0.5f) * s, u3 = (float)c[3] * s;
    const float r0 = sqrtf(-2.0f * __logf(fminf(u0, 0.99999994f))), r1 = sqrtf(-2.0f * __logf(fminf(u2, 0.99999994f)));
    float s0, c0, s1, c1;
    sincospif(2.0f * u1, &s0, &c0);
    sincospif(2.0f * u3, &s1, &c1);
    return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

// one thread = two consecutive samples of one symbol
__global__ void tx_symbols_kernel(const float2 *__restrict__ up, uint32_t sps, uint32_t decim, const uint32_t *__restrict__ values,
                                  const float *__restrict__ cfo_hz, double inv_fs, float sigma, unsigned long long seed,
                                  size_t n_symbols, float2 *__restrict__ out) {
    const size_t pairs = (size_t)sps / 2, total = n_symbols * pairs;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const size_t s = g / pairs;
        const uint32_t n = (uint32_t)(g - s * pairs) * 2u;
        const uint32_t sh = values[s] * decim;
        float2 a = up[(n + sh) % sps], b = up[(n + 1u + sh) % sps];
        if (cfo_hz) {
            const double rev = (double)cfo_hz[s] * inv_fs;        // revolutions per sample
            double t0 = rev * (double)n, t1 = rev * (double)(n + 1u);
            t0 -= floor(t0); t1 -= floor(t1);
            float s0, c0, s1, c1;
            sincospif(2.0f * (float)t0, &s0, &c0);
            sincospif(2.0f * (float)t1, &s1, &c1);
            a = make_float2(a.x * c0 - a.y * s0, a.x * s0 + a.y * c0);
            b = make_float2(b.x * c1 - b.y * s1, b.x * s1 + b.y * c1);
        }
        if (sigma != 0.0f) {
            const float4 z = philox_normal4(seed, s, n / 2u);
            a.x = fmaf(sigma, z.x, a.x); a.y = fmaf(sigma, z.y, a.y);
            b.x = fmaf(sigma, z.z, b.x); b.y = fmaf(sigma, z.w, b.y);
        }
        __stcs(reinterpret_cast<float4 *>(out + s * sps + n), make_float4(a.x, a.y, b.x, b.y));
    }
}

// out[s][i] = base[s % k][i] + noise(seed, s, i); n_items even
__global__ void tx_expand_kernel(const float2 *__restrict__ base, uint32_t k, size_t n_items, float sigma, unsigned long long seed,
                                 size_t n_streams, float2 *__restrict__ out) {
    const size_t pairs = n_items / 2, total = n_streams * pairs;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const size_t s = g / pairs, i = (g - s * pairs) * 2;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(base + (s % k) * n_items + i));
        float4 o = v;
        if (sigma != 0.0f) {
            const float4 z = philox_normal4(seed, s, i / 2);
            o = make_float4(fmaf(sigma, z.x, v.x), fmaf(sigma, z.y, v.y), fmaf(sigma, z.z, v.z), fmaf(sigma, z.w, v.w));
        }
        __stcs(reinterpret_cast<float4 *>(out + s * n_items + i), o);
    }
}
#endif
}  // namespace lb
