// k1_fft.cuh -- K1: dechirp + pruned FFT + argmax (the north-star kernel).
//
// Computes, for each aligned symbol window x[0..sps), exactly what the reference's
// get_shift_fft does (lib/decoder_impl.cc:430-464):
//     m[n]   = x[n] * down[n]                                   (:436-438)
//     F      = forward sps-point DFT of m                        (:443)
//     tmp[k] = F[k] (k < N/2), F[sps - N + k] (k >= N/2), tmp[N/2] += F[N/2]   (:447-450)
//     bin    = first argmax |tmp|                                (:452-463)
// without ever forming the 7N unused bins:  with n = 8*n1 + r (sps = 8N at fs/bw = 8)
//     F[k'] = sum_r W_sps^{k' r} G_r[k' mod N],   G_r = N-point DFT over n1 of m[8 n1 + r]
// i.e. 8 N-point FFTs (one per polyphase branch) and one 8-term twiddled sum per kept bin.
// For SF11/12 (N > 1024) a radix-S decimation-in-frequency step (S = 2, 4) is folded into
// the load:  F[S q + s] = DFT_{sps/S}( y_s )[q],  y_s[n] = W_sps^{s n} sum_j W_S^{s j} m[n + j sps/S],
// giving S independent sub-problems of N' = N/S = 1024 bins whose partial argmaxes are
// merged with a 64-bit atomicMax.
//
// Data flow per CTA (256 threads) and batch of G = 1024/N' symbols:
//   pass 0  coalesced float4 loads straight into registers (each thread: 16 rows x 2 branches),
//           dechirp, radix-16 DIF in registers, inter-pass twiddle, store to shared memory
//   pass i  in-place radix-8/16/4 passes in shared memory (padded, conflict-free layout)
//   combine Horner evaluation of the 8-branch twiddled sum, |.|^2, argmax
// The phase functions are __host__ __device__ so tests/test_k1_emulation.py can run the
// very same index arithmetic on the CPU.
#pragma once
#include "lora_common.cuh"

namespace lb {

constexpr int K1_THREADS = 256;

template <int SF>
struct K1Cfg {
    static constexpr int N = 1 << SF;                 // bins
    static constexpr int SPS = 8 * N;                 // samples per symbol (fs/bw = 8)
    static constexpr int S = SF <= 10 ? 1 : (1 << (SF - 10));   // DIF split factor
    static constexpr int NP = N / S;                  // bins per sub-problem (128..1024)
    static constexpr int SPS_SUB = 8 * NP;
    static constexpr int G = 1024 / NP;               // symbols per CTA batch
    static constexpr int M0 = NP / 16;                // columns after the radix-16 pass 0
    static constexpr int SB0 = NP + NP / 16;
    static constexpr int SB = SB0 + ((2 - SB0 % 16) + 16) % 16;   // branch stride, == 2 (mod 16)
    static constexpr int SYM_STRIDE = 8 * SB;
    static constexpr int SMEM_ELEMS = G * SYM_STRIDE; // float2 elements
    static constexpr int TPS = K1_THREADS / G;        // threads per symbol in the combine phase
    // passes after pass 0 on blocks of M0 points: (R1, SIG1) then (R2, 1)
    static constexpr int R1 = M0 == 8 ? 8 : M0 == 16 ? 16 : 8;
    static constexpr int SIG1 = M0 / R1;              // 1, 1, 4, 8
    static constexpr int R2 = SIG1;                   // 1 (none), 1, 4, 8
    static_assert(SF >= 7 && SF <= 12, "K1 supports SF7..SF12");
};

LB_HD int k1_pad(int i) { return i + (i >> 4); }

struct K1Args {
    const float2 *x;        // n_symbols * SPS samples
    const float2 *chirp;    // down-chirp table, SPS entries
    const float2 *tw;       // W_sps^j = exp(-2 pi i j / sps), SPS entries
    size_t n_symbols;
};

// AL16: the window starts on a 16-byte boundary (batch path); the stream state machine
// hands windows at arbitrary sample offsets (8-byte aligned only)
template <bool AL16>
LB_HD float4 k1_ld_stream(const float2 *p) {
#ifdef __CUDA_ARCH__
    if (AL16) return __ldcs(reinterpret_cast<const float4 *>(p));
    const float2 a = __ldcs(p), b = __ldcs(p + 1);
    return make_float4(a.x, a.y, b.x, b.y);
#else
    return make_float4(p[0].x, p[0].y, p[1].x, p[1].y);
#endif
}
LB_HD float4 k1_ld_table4(const float2 *p) {
#ifdef __CUDA_ARCH__
    return __ldg(reinterpret_cast<const float4 *>(p));
#else
    return make_float4(p[0].x, p[0].y, p[1].x, p[1].y);
#endif
}
LB_HD float2 k1_ld_table(const float2 *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// ---- pass 0: global -> registers -> radix-16 -> shared ---------------------------------
template <int SF, bool AL16 = true>
LB_HD void k1_pass0(const K1Args &a, size_t batch, int s, int tid, float2 *buf) {
    using C = K1Cfg<SF>;
    const int g = tid / (4 * C::M0);
    const int rem = tid % (4 * C::M0);
    const int m = rem >> 2, b = rem & 3;
    const size_t sym = batch * C::G + g;
    const bool valid = sym < a.n_symbols;
    const float2 *xs = a.x + (valid ? sym : 0) * (size_t)C::SPS;
    float2 v0[16], v1[16];
#pragma unroll
    for (int c = 0; c < 16; c++) {
        const int n = 8 * (c * C::M0 + m) + 2 * b;       // index inside the sub-problem
        if (C::S == 1) {
            const float4 xv = k1_ld_stream<AL16>(xs + n);
            const float4 dv = k1_ld_table4(a.chirp + n);
            v0[c] = cmul(make_float2(xv.x, xv.y), make_float2(dv.x, dv.y));
            v1[c] = cmul(make_float2(xv.z, xv.w), make_float2(dv.z, dv.w));
        } else {
            float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < C::S; j++) {
                const int nf = n + j * C::SPS_SUB;
                const float4 xv = k1_ld_stream<AL16>(xs + nf);
                const float4 dv = k1_ld_table4(a.chirp + nf);
                const float2 w = k1_ld_table(a.tw + ((s * j * (C::SPS / C::S)) & (C::SPS - 1)));   // W_S^{s j}
                acc0 = cfma(cmul(make_float2(xv.x, xv.y), make_float2(dv.x, dv.y)), w, acc0);
                acc1 = cfma(cmul(make_float2(xv.z, xv.w), make_float2(dv.z, dv.w)), w, acc1);
            }
            v0[c] = cmul(acc0, k1_ld_table(a.tw + s * n));          // W_sps^{s n}
            v1[c] = cmul(acc1, k1_ld_table(a.tw + s * (n + 1)));
        }
        if (!valid) { v0[c] = make_float2(0.f, 0.f); v1[c] = make_float2(0.f, 0.f); }
    }
    dft_dif<16>(v0);
    dft_dif<16>(v1);
    float2 *b0 = buf + g * C::SYM_STRIDE + (2 * b) * C::SB;
    float2 *b1 = b0 + C::SB;
    // inter-pass twiddle W_{N'}^{m kc} = u^kc with the lane-invariant base u = W_{N'}^m: one table
    // load and a running product instead of 15 scattered loads (they were half of the L1 wavefronts
    // in the first profile).  15 fp32 products in a row: relative error < 2e-6.
    const float2 u = k1_ld_table(a.tw + m * 8 * C::S);
    float2 t = make_float2(1.0f, 0.0f);
#pragma unroll
    for (int kc = 0; kc < 16; kc++) {
        const int br = bitrev<16>(kc);
        const int pos = k1_pad(kc * C::M0 + m);
        if (kc == 0) {
            b0[pos] = v0[br];
            b1[pos] = v1[br];
        } else {
            t = kc == 1 ? u : cmul(t, u);
            b0[pos] = cmul(v0[br], t);
            b1[pos] = cmul(v1[br], t);
        }
    }
}

// ---- in-place shared-memory pass of radix R, stride SIG (in n1 units) -------------------
template <int SF, int R, int SIG>
LB_HD void k1_pass(const K1Args &a, int tid, float2 *buf) {
    using C = K1Cfg<SF>;
    constexpr int PER_BRANCH = C::NP / R;
    constexpr int ITEMS = C::G * 8 * PER_BRANCH;
    for (int it = tid; it < ITEMS; it += K1_THREADS) {
        const int j = it % PER_BRANCH;
        const int gr = it / PER_BRANCH;                  // g * 8 + r
        const int lo = j % SIG;
        const int base = (j / SIG) * (R * SIG) + lo;
        float2 *p = buf + gr * C::SB;
        float2 v[R];
#pragma unroll
        for (int c = 0; c < R; c++) v[c] = p[k1_pad(base + SIG * c)];
        dft_dif<R>(v);
        float2 u = make_float2(1.0f, 0.0f), t = u;
        if (SIG > 1) u = k1_ld_table(a.tw + lo * (C::SPS / (R * SIG)));      // W_{R SIG}^{lo}; t runs through its powers
#pragma unroll
        for (int kc = 0; kc < R; kc++) {
            float2 o = v[bitrev<R>(kc)];
            if (SIG > 1 && kc > 0) {
                t = kc == 1 ? u : cmul(t, u);
                o = cmul(o, t);
            }
            p[k1_pad(base + SIG * kc)] = o;
        }
    }
}

// position p (inside one branch) -> sub-problem bin q held there after all passes
template <int SF>
LB_HD int k1_pos_to_bin(int p) {
    using C = K1Cfg<SF>;
    const int d0 = p / C::M0, dr = p % C::M0;
    int rest;
    if (C::SIG1 == 1) rest = dr;                          // one more pass: digit d1 = dr
    else rest = (dr / C::SIG1) + C::R1 * (dr % C::SIG1);  // d1 + R1 * d2
    return d0 + 16 * rest;
}

// ---- combine: 8-branch twiddled sum, |.|^2, per-thread argmax over its 4 positions -------
// lane-invariant combine twiddles W_{sps'}^{qs} for the thread's NP/TPS positions (hoisted out of
// the persistent loop: they were 4 scattered loads per thread and batch)
template <int SF>
LB_HD void k1_combine_twiddles(const K1Args &a, int tid, float2 *w) {
    using C = K1Cfg<SF>;
    const int lt = tid % C::TPS;
    for (int i = 0; i < C::NP / C::TPS; i++) {
        const int q = k1_pos_to_bin<SF>(lt + C::TPS * i);
        const int qs = q < C::NP / 2 ? q : q - C::NP;
        w[i] = k1_ld_table(a.tw + ((qs * C::S) & (C::SPS - 1)));
    }
}

template <int SF>
LB_HD unsigned long long k1_combine(const K1Args &a, int s, int tid, const float2 *buf, const float2 *wtab) {
    using C = K1Cfg<SF>;
    const int g = tid / C::TPS, lt = tid % C::TPS;
    const float2 *bs = buf + g * C::SYM_STRIDE;
    unsigned long long best = 0ull;
#pragma unroll
    for (int i = 0; i < C::NP / C::TPS; i++) {
        const int p = lt + C::TPS * i;
        const int q = k1_pos_to_bin<SF>(p);
        const int qs = q < C::NP / 2 ? q : q - C::NP;      // signed bin of the sub-problem
        const float2 w = wtab[i];                          // W_{sps'}^{qs}
        const int pp = k1_pad(p);
        float2 gv[8];
#pragma unroll
        for (int r = 0; r < 8; r++) gv[r] = bs[r * C::SB + pp];
        float2 acc = gv[7];
#pragma unroll
        for (int r = 6; r >= 0; r--) acc = cfma(acc, w, gv[r]);
        if (s == 0 && q == C::NP / 2) {                    // tmp[N/2] += F[N/2]  (:450)
            const float2 wc = cconj(w);
            float2 acc2 = gv[7];
#pragma unroll
            for (int r = 6; r >= 0; r--) acc2 = cfma(acc2, wc, gv[r]);
            acc = cadd(acc, acc2);
        }
        const int kp = C::S * qs + s;
        const uint32_t idx = (uint32_t)(kp >= 0 ? kp : C::N + kp);
        const unsigned long long key = pack_key(cnorm2(acc), idx);
        best = key > best ? key : best;
    }
    return best;
}

#ifdef __CUDACC__
// ---- the kernel: persistent CTAs over (batch, s) work items ------------------------------
template <int SF>
__global__ void __launch_bounds__(K1_THREADS, 2)
k1_fft_kernel(K1Args a, uint32_t *__restrict__ bins, float *__restrict__ mags,
              unsigned long long *__restrict__ packed /* S > 1 only */) {
    using C = K1Cfg<SF>;
    extern __shared__ float2 k1_smem[];
    __shared__ unsigned long long warp_best[K1_THREADS / 32];
    float2 *buf = k1_smem;
    const int tid = threadIdx.x;
    const size_t n_batches = (a.n_symbols + C::G - 1) / C::G;
    const size_t n_work = n_batches * C::S;
    float2 wtab[C::NP / C::TPS];
    k1_combine_twiddles<SF>(a, tid, wtab);
    for (size_t w = blockIdx.x; w < n_work; w += gridDim.x) {
        const size_t batch = w / C::S;
        const int s = (int)(w % C::S);
        k1_pass0<SF>(a, batch, s, tid, buf);
        __syncthreads();
        k1_pass<SF, C::R1, C::SIG1>(a, tid, buf);
        __syncthreads();
        if (C::R2 > 1) {
            k1_pass<SF, (C::R2 > 1 ? C::R2 : 2), 1>(a, tid, buf);
            __syncthreads();
        }
        unsigned long long best = k1_combine<SF>(a, s, tid, buf, wtab);
        // warp argmax, then across the warps of one symbol
        best = warp_max_key(best);
        if ((tid & 31) == 0) warp_best[tid >> 5] = best;
        __syncthreads();
        if (tid < C::G) {
            constexpr int WPS = C::TPS / 32;              // warps per symbol
            unsigned long long bb = 0ull;
#pragma unroll
            for (int k = 0; k < WPS; k++) {
                const unsigned long long o = warp_best[tid * WPS + k];
                bb = o > bb ? o : bb;
            }
            const size_t sym = batch * C::G + tid;
            if (sym < a.n_symbols) {
                if (C::S == 1) {
                    bins[sym] = key_idx(bb);
                    if (mags) mags[sym] = sqrtf(key_mag2(bb));
                } else {
                    atomicMax(packed + sym, bb);
                }
            }
        }
        // warp_best and buf are rewritten only after the next pass0 + __syncthreads
    }
}

static __global__ void k1_finalize_kernel(const unsigned long long *__restrict__ packed, size_t n,
                                   uint32_t *__restrict__ bins, float *__restrict__ mags) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) {
        const unsigned long long k = packed[i];
        bins[i] = key_idx(k);
        if (mags) mags[i] = sqrtf(key_mag2(k));
    }
}
#endif  // __CUDACC__

// ---- CPU emulation of the kernel (same phase functions, threads run one after another) ---
template <int SF>
inline void k1_emulate(const K1Args &a, uint32_t *bins, float *mags) {
    using C = K1Cfg<SF>;
    float2 *buf = new float2[C::SMEM_ELEMS];
    const size_t n_batches = (a.n_symbols + C::G - 1) / C::G;
    unsigned long long *packed = new unsigned long long[n_batches * C::G]();
    for (size_t batch = 0; batch < n_batches; batch++) {
        for (int s = 0; s < C::S; s++) {
            for (int i = 0; i < C::SMEM_ELEMS; i++) buf[i] = make_float2(NAN, NAN);   // catch unwritten reads
            for (int t = 0; t < K1_THREADS; t++) k1_pass0<SF>(a, batch, s, t, buf);
            for (int t = 0; t < K1_THREADS; t++) k1_pass<SF, C::R1, C::SIG1>(a, t, buf);
            if (C::R2 > 1)
                for (int t = 0; t < K1_THREADS; t++) k1_pass<SF, (C::R2 > 1 ? C::R2 : 2), 1>(a, t, buf);
            for (int t = 0; t < K1_THREADS; t++) {
                float2 wtab[C::NP / C::TPS];
                k1_combine_twiddles<SF>(a, t, wtab);
                const unsigned long long k = k1_combine<SF>(a, s, t, buf, wtab);
                const size_t sym = batch * C::G + t / C::TPS;
                if (k > packed[sym]) packed[sym] = k;
            }
        }
    }
    for (size_t i = 0; i < a.n_symbols; i++) {
        bins[i] = key_idx(packed[i]);
        if (mags) mags[i] = sqrtf(key_mag2(packed[i]));
    }
    delete[] buf;
    delete[] packed;
}

}  // namespace lb
