// k1_cluster.cuh -- K1 for SF11 / SF12: one thread-block CLUSTER (2 / 4 CTAs) per symbol.
//
// A 2^SF-bin symbol (128 / 256 KiB of IQ) does not fit one SM.  The radix-16 pass 0 over the 16
// "rows" of the symbol is distributed: CTA `rank` of the cluster loads the columns
// m in [rank*M0/CL, (rank+1)*M0/CL) (coalesced float4, every sample read once), dechirps, runs the
// radix-16 DIF in registers and scatters output column kc to the CTA that owns it
// (kc / (16/CL)) through distributed shared memory (st.shared::cluster via map_shared_rank).
// After one cluster barrier every CTA holds, for its 16/CL columns kc, the 8 polyphase branches of
// M0 = N/16 points each = 1024 points per branch (64 KiB): exactly the shape the in-place passes of
// k1_fft.cuh work on.  The remaining FFT_M0 (16x8 at SF11, 16x16 at SF12), the 8-branch twiddled sum
// and the argmax run locally; the per-CTA partial argmaxes merge with a 64-bit atomicMax.
// Compared with the first version (radix-S DIF folded into the load, every CTA re-reading the whole
// symbol S times) this reads each sample once and dechirps it once.
#pragma once
#include <cooperative_groups.h>
#include "k1_fft.cuh"

namespace lb {

template <int SF>
struct KCfg {
    static constexpr int CL = 1 << (SF - 10);             // CTAs per cluster: 2, 4
    static constexpr int N = 1 << SF, SPS = 8 * N;
    static constexpr int M0 = N / 16;                     // 128, 256: points of the second FFT
    static constexpr int KPC = 16 / CL;                   // columns kc owned per CTA: 8, 4
    static constexpr int MPC = M0 / CL;                   // columns m loaded per CTA: 64
    static constexpr int R1 = 16, SIG1 = M0 / 16, R2 = SIG1;   // FFT_M0 = 16 x (8 | 16)
    static_assert(SF == 11 || SF == 12, "cluster kernel: SF11, SF12");
    static_assert(KPC * M0 == K1Cfg<SF>::NP, "per-CTA branch length must match the shared-memory layout");
};

// position inside one branch of this CTA's buffer -> bin q of the full symbol
template <int SF>
LB_HD int kc_pos_to_bin(int p, int rank) {
    using K = KCfg<SF>;
    const int kc = rank * K::KPC + p / K::M0;
    const int dr = p % K::M0;
    return kc + 16 * ((dr / K::SIG1) + K::R1 * (dr % K::SIG1));
}

// pass 0 of CTA `rank`, thread `tid`: load + dechirp + radix-16, scatter column kc to peer[kc / KPC]
template <int SF>
LB_HD void kc_pass0(const K1Args &a, size_t sym, int rank, int tid, float2 u, float2 *v0, float2 *v1) {
    using K = KCfg<SF>;
    const int m = rank * K::MPC + (tid >> 2), b = tid & 3;
    const float2 *xs = a.x + sym * (size_t)K::SPS;
#pragma unroll
    for (int c = 0; c < 16; c++) {
        const int n = 8 * (c * K::M0 + m) + 2 * b;
        const float4 xv = k1_ld_stream<true>(xs + n);
        const float4 dv = k1_ld_table4(a.chirp + n);
        v0[c] = cmul(make_float2(xv.x, xv.y), make_float2(dv.x, dv.y));
        v1[c] = cmul(make_float2(xv.z, xv.w), make_float2(dv.z, dv.w));
    }
    dft_dif<16>(v0);
    dft_dif<16>(v1);
    (void)u;
}

template <int SF>
LB_HD void kc_scatter(int rank, int tid, float2 u, const float2 *v0, const float2 *v1, float2 *const *peer) {
    using C = K1Cfg<SF>;
    using K = KCfg<SF>;
    const int m = rank * K::MPC + (tid >> 2), b = tid & 3;
    float2 t = make_float2(1.0f, 0.0f);
#pragma unroll
    for (int kc = 0; kc < 16; kc++) {
        const int br = bitrev<16>(kc);
        float2 *dst = peer[kc / K::KPC] + (2 * b) * C::SB + k1_pad((kc % K::KPC) * K::M0 + m);
        if (kc == 0) {
            dst[0] = v0[br];
            dst[C::SB] = v1[br];
        } else {
            t = kc == 1 ? u : cmul(t, u);
            dst[0] = cmul(v0[br], t);
            dst[C::SB] = cmul(v1[br], t);
        }
    }
}

template <int SF>
LB_HD void kc_twiddles(const K1Args &a, int rank, int tid, float2 *wtab) {
    using C = K1Cfg<SF>;
    using K = KCfg<SF>;
    for (int i = 0; i < C::NP / K1_THREADS; i++) {
        const int q = kc_pos_to_bin<SF>(tid + K1_THREADS * i, rank);
        wtab[i] = k1_ld_table(a.tw + ((q < K::N / 2 ? q : q - K::N) & (K::SPS - 1)));
    }
}

template <int SF>
LB_HD unsigned long long kc_combine(int rank, int tid, const float2 *buf, const float2 *wtab) {
    using C = K1Cfg<SF>;
    using K = KCfg<SF>;
    unsigned long long best = 0ull;
#pragma unroll
    for (int i = 0; i < C::NP / K1_THREADS; i++) {
        const int p = tid + K1_THREADS * i;
        const int q = kc_pos_to_bin<SF>(p, rank);
        const int pp = k1_pad(p);
        float2 gv[8];
#pragma unroll
        for (int r = 0; r < 8; r++) gv[r] = buf[r * C::SB + pp];
        const float2 w = wtab[i];
        float2 acc = gv[7];
#pragma unroll
        for (int r = 6; r >= 0; r--) acc = cfma(acc, w, gv[r]);
        if (q == K::N / 2) {                 // tmp[N/2] += F[N/2]  (:450)
            const float2 wc = cconj(w);
            float2 acc2 = gv[7];
#pragma unroll
            for (int r = 6; r >= 0; r--) acc2 = cfma(acc2, wc, gv[r]);
            acc = cadd(acc, acc2);
        }
        const unsigned long long key = pack_key(cnorm2(acc), (uint32_t)q);
        best = key > best ? key : best;
    }
    return best;
}

// CPU emulation: the CL CTAs of the cluster run one after another on CL buffers
template <int SF>
inline void kc_emulate(const K1Args &a, uint32_t *bins, float *mags) {
    using C = K1Cfg<SF>;
    using K = KCfg<SF>;
    float2 *bufs[K::CL];
    for (int q = 0; q < K::CL; q++) bufs[q] = new float2[C::SMEM_ELEMS];
    for (size_t sym = 0; sym < a.n_symbols; sym++) {
        for (int q = 0; q < K::CL; q++)
            for (int i = 0; i < C::SMEM_ELEMS; i++) bufs[q][i] = make_float2(NAN, NAN);
        for (int rank = 0; rank < K::CL; rank++)
            for (int t = 0; t < K1_THREADS; t++) {
                float2 v0[16], v1[16];
                const float2 u = k1_ld_table(a.tw + (rank * K::MPC + (t >> 2)) * 8);
                kc_pass0<SF>(a, sym, rank, t, u, v0, v1);
                kc_scatter<SF>(rank, t, u, v0, v1, bufs);
            }
        unsigned long long best = 0ull;
        for (int rank = 0; rank < K::CL; rank++) {
            for (int t = 0; t < K1_THREADS; t++) k1_pass<SF, K::R1, K::SIG1>(a, t, bufs[rank]);
            for (int t = 0; t < K1_THREADS; t++) k1_pass<SF, K::R2, 1>(a, t, bufs[rank]);
            for (int t = 0; t < K1_THREADS; t++) {
                float2 wtab[C::NP / K1_THREADS];
                kc_twiddles<SF>(a, rank, t, wtab);
                const unsigned long long k = kc_combine<SF>(rank, t, bufs[rank], wtab);
                best = k > best ? k : best;
            }
        }
        bins[sym] = key_idx(best);
        if (mags) mags[sym] = sqrtf(key_mag2(best));
    }
    for (int q = 0; q < K::CL; q++) delete[] bufs[q];
}

#ifdef __CUDACC__
template <int SF>
__global__ void __launch_bounds__(K1_THREADS, 2)
k1_cluster_kernel(K1Args a, unsigned long long *__restrict__ packed) {
    namespace cg = cooperative_groups;
    using C = K1Cfg<SF>;
    using K = KCfg<SF>;
    extern __shared__ float2 k1c_smem[];
    __shared__ unsigned long long warp_best[K1_THREADS / 32];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x;
    const size_t cid = blockIdx.x / K::CL, n_clusters = gridDim.x / K::CL;
    float2 *buf = k1c_smem;
    float2 *peer[K::CL];
#pragma unroll
    for (int q = 0; q < K::CL; q++) peer[q] = cluster.map_shared_rank(buf, q);

    // lane-invariant pieces
    const float2 u = k1_ld_table(a.tw + (rank * K::MPC + (tid >> 2)) * 8);   // W_N^m (inter-pass twiddle base)
    float2 wtab[C::NP / K1_THREADS];
    kc_twiddles<SF>(a, rank, tid, wtab);

    for (size_t sym = cid; sym < a.n_symbols; sym += n_clusters) {
        // ---- pass 0, distributed over the cluster -------------------------------------------
        float2 v0[16], v1[16];
        kc_pass0<SF>(a, sym, rank, tid, u, v0, v1);
        cluster.sync();                      // every CTA has finished reading its buffer (previous symbol)
        kc_scatter<SF>(rank, tid, u, v0, v1, peer);
        cluster.sync();                      // all remote stores have landed
        // ---- local: FFT_M0 in place, combine, argmax ------------------------------------------
        k1_pass<SF, K::R1, K::SIG1>(a, tid, buf);
        __syncthreads();
        k1_pass<SF, K::R2, 1>(a, tid, buf);
        __syncthreads();
        unsigned long long best = kc_combine<SF>(rank, tid, buf, wtab);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, off);
            best = o > best ? o : best;
        }
        if ((tid & 31) == 0) warp_best[tid >> 5] = best;
        __syncthreads();
        if (tid == 0) {
            unsigned long long bb = warp_best[0];
#pragma unroll
            for (int k = 1; k < K1_THREADS / 32; k++) bb = warp_best[k] > bb ? warp_best[k] : bb;
            atomicMax(packed + sym, bb);
        }
    }
    cluster.sync();                          // nobody exits while a peer may still write into its shared memory
}
#endif  // __CUDACC__

}  // namespace lb
