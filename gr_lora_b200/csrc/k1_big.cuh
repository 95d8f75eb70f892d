// k1_big.cuh -- K1 for SF11 / SF12: a thread-block CLUSTER of CL = 2 / 4 CTAs per symbol, every CTA a
// TMA-fed 256-thread group as in k1_sf10.cuh.
//
// N = 32 * (32 CL) bins.  The symbol (128 / 256 KiB) is split by COLUMN: with 32 rows c of ROWLEN =
// 256 CL samples, CTA `rank` owns samples [c*ROWLEN + 256 rank, +256) of every row, i.e. columns
// a = 32 rank + (t >> 3), branch r = t & 7, and receives them as 32 bulk copies of 2 KiB per symbol into
// a 64 KiB slot (2-slot ring, one mbarrier per slot).  Per symbol:
//   pass 0    dechirp, radix-32 over the rows (registers), inter-pass twiddle W_N^{a kc}
//   exch 1    scatter over DISTRIBUTED shared memory: output column kc goes to CTA kc / KPC,
//             thread ((kc % KPC) * 8 + r) * CL + (a % CL), position a / CL            [cluster barrier x2]
//   pass 1    radix-32 over a' = a / CL for the residue j = a % CL            -> F_j[ka']
//   transpose the CL lanes of an item swap residues through shared memory so that lane j_r holds
//             F_j[CL i + j_r] for all j;  radix-CL DIT across j with W_32^{j i} W_{32CL}^{j j_r}
//             -> G_r[kc + 32 ka], ka = CL i + j_r + 32 m
//   exch 2    [local bin][branch] swizzled, local
//   combine   Horner over the 8 branches with one lane-invariant twiddle per bin, |.|^2, argmax;
//             the CTAs' partial argmaxes merge with a 64-bit atomicMax.
// Every sample is read once from HBM and dechirped once; same arithmetic as get_shift_fft
// (lib/decoder_impl.cc:430-464).
#pragma once
#include <cooperative_groups.h>
#include "k1_sf10.cuh"

namespace lb {

template <int SF>
struct BCfg {
    static constexpr int CL = 1 << (SF - 10);            // CTAs per cluster (2, 4)
    static constexpr int N = 1024 * CL, SPS = 8 * N;
    static constexpr int ROWLEN = 256 * CL;              // samples per row
    static constexpr int KPC = 32 / CL;                  // output columns kc per CTA
    static constexpr int KA = 32 * CL;                   // length of the second FFT
    static constexpr int IPL = 32 / CL;                  // i values per lane after the transpose
    static_assert(SF == 11 || SF == 12, "k1_big: SF11, SF12");
};

template <int SF>
struct BConsts {
    float2 tl[4], th[8];     // W_N^{a j}, W_N^{4 a j}
    float2 wj[4];            // W_{32 CL}^{j j_r}, j = 0..CL-1
    float2 wq[4];            // W_sps^{q'} for the thread's 4 bins
};

LB_HD int b_swz2(int q) { return (((q >> 1) & 1) << 1) | ((q >> 2) & 1); }
LB_HD int b_pos2(int q, int r) { return q * 8 + ((((r >> 1) ^ b_swz2(q)) << 1) | (r & 1)); }

// local bin index q_l = kc_l * KA + ka  ->  bin of the whole symbol
template <int SF>
LB_HD int b_global_bin(int q_l, int rank) {
    using B = BCfg<SF>;
    return (B::KPC * rank + q_l / B::KA) + 32 * (q_l % B::KA);
}

template <int SF>
LB_HD void b_consts(int t, int rank, const float2 *tw, BConsts<SF> &c) {
    using B = BCfg<SF>;
    const int a = 32 * rank + (t >> 3);
    const int jr = t & (B::CL - 1);
    for (int j = 0; j < 4; j++) c.tl[j] = k1_ld_table(tw + ((a * j * 8) & (B::SPS - 1)));            // W_N = W_sps^8
    for (int j = 0; j < 8; j++) c.th[j] = k1_ld_table(tw + ((a * 4 * j * 8) & (B::SPS - 1)));
    for (int j = 0; j < 4; j++) c.wj[j] = k1_ld_table(tw + (((j * jr) * (B::SPS / B::KA)) & (B::SPS - 1)));
    for (int i = 0; i < 4; i++) {
        const int q = b_global_bin<SF>(t + 256 * i, rank);
        c.wq[i] = k1_ld_table(tw + ((q < B::N / 2 ? q : q - B::N) & (B::SPS - 1)));
    }
}

template <int SF>
LB_HD void b_pass0(int t, const float2 *slot, const float2 *chirp, const BConsts<SF> &c, float2 *v) {
#pragma unroll
    for (int r = 0; r < 32; r++) v[r] = cmul(slot[r * 256 + t], chirp[r * 256 + t]);
    dft_dif<32>(v);
#pragma unroll
    for (int kc = 1; kc < 32; kc++) {
        const int br = bitrev<32>(kc);
        float2 w;
        if ((kc & 3) == 0) w = c.th[kc >> 2];
        else if ((kc >> 2) == 0) w = c.tl[kc & 3];
        else w = cmul(c.th[kc >> 2], c.tl[kc & 3]);
        v[br] = cmul(v[br], w);
    }
}

// exchange 1: to CTA kc / KPC, thread ((kc % KPC)*8 + r)*CL + (a % CL), slot position (a / CL)*256 + thread
template <int SF>
LB_HD void b_scatter(int t, int rank, const float2 *v, float2 *const *peer_slot) {
    using B = BCfg<SF>;
    const int a = 32 * rank + (t >> 3), r = t & 7;
    const int j = a & (B::CL - 1), ap = a / B::CL;
#pragma unroll
    for (int kc = 0; kc < 32; kc++) {
        const int dst_t = ((kc % B::KPC) * 8 + r) * B::CL + j;
        peer_slot[kc / B::KPC][ap * 256 + dst_t] = v[bitrev<32>(kc)];
    }
}

template <int SF>
LB_HD void b_pass1(int t, const float2 *slot, float2 *f) {
#pragma unroll
    for (int ap = 0; ap < 32; ap++) f[ap] = slot[ap * 256 + t];
    dft_dif<32>(f);                                  // f[bitrev(ka')] = F_j[ka'], j = t % CL
}

// lane transpose through shared memory: row (ka'/CL)*CL + j, column item*CL + (ka' % CL + j) % CL
template <int SF>
LB_HD void b_store_t(int t, float2 *slot, const float2 *f) {
    using B = BCfg<SF>;
    const int item = t / B::CL, j = t & (B::CL - 1);
#pragma unroll
    for (int kap = 0; kap < 32; kap++)
        slot[((kap / B::CL) * B::CL + j) * 256 + item * B::CL + ((kap % B::CL + j) & (B::CL - 1))] = f[bitrev<32>(kap)];
}

// after the transpose: g[CL*i + m ... ] -- lane j_r computes, for i < IPL, ka' = CL i + j_r:
//   out[ka' + 32 m] = sum_j W_CL^{j m} * ( W_32^{j i} W_{32CL}^{j j_r} F_j[ka'] ),  m < CL
// stored as g[i * CL + m]
template <int SF>
LB_HD void b_load_t_radix(int t, const float2 *slot, const BConsts<SF> &c, float2 *g) {
    using B = BCfg<SF>;
    const int item = t / B::CL, jr = t & (B::CL - 1);
#pragma unroll
    for (int i = 0; i < B::IPL; i++) {
        float2 z[B::CL];
#pragma unroll
        for (int j = 0; j < B::CL; j++) {
            float2 x = slot[(i * B::CL + j) * 256 + item * B::CL + ((jr + j) & (B::CL - 1))];
            if (j > 0) {
                const int e = (j * i) & 31;                 // W_32^{j i}; W_32^{e} = -W_32^{e-16} for e >= 16
                x = mul_w32(x, e & 15);
                if (e >= 16) x = make_float2(-x.x, -x.y);
                x = cmul(x, c.wj[j]);
            }
            z[j] = x;
        }
        dft_dif<B::CL>(z);                              // z[bitrev(m)] = sum_j W_CL^{j m} z_j
#pragma unroll
        for (int m = 0; m < B::CL; m++) g[i * B::CL + m] = z[bitrev<B::CL>(m)];
    }
}

// exchange 2: local bin q_l = kc_l * KA + ka, ka = CL i + j_r + 32 m
template <int SF>
LB_HD void b_store2(int t, float2 *slot, const float2 *g) {
    using B = BCfg<SF>;
    const int item = t / B::CL, jr = t & (B::CL - 1);
    const int kcl = item >> 3, r = item & 7;
#pragma unroll
    for (int i = 0; i < B::IPL; i++)
#pragma unroll
        for (int m = 0; m < B::CL; m++) slot[b_pos2(kcl * B::KA + (B::CL * i + jr + 32 * m), r)] = g[i * B::CL + m];
}

template <int SF>
LB_HD unsigned long long b_combine(int t, int rank, const float2 *slot, const BConsts<SF> &c) {
    using B = BCfg<SF>;
    unsigned long long best = 0ull;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int ql = t + 256 * i;
        const int q = b_global_bin<SF>(ql, rank);
        float2 gv[8];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const float4 v = *reinterpret_cast<const float4 *>(slot + ql * 8 + ((u ^ b_swz2(ql)) << 1));
            gv[2 * u] = make_float2(v.x, v.y);
            gv[2 * u + 1] = make_float2(v.z, v.w);
        }
        const float2 w = c.wq[i];
        float2 acc = gv[7];
#pragma unroll
        for (int r = 6; r >= 0; r--) acc = cfma(acc, w, gv[r]);
        if (q == B::N / 2) {                             // tmp[N/2] += F[N/2]  (:450)
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

#ifdef __CUDACC__
template <int NSLOT>
struct BSmem {
    float2 chirp[8192];
    float2 slots[NSLOT][8192];
    uint64_t bars[NSLOT];
    unsigned long long keys[8];
};

template <int SF, int NSLOT>
__global__ void __launch_bounds__(256, 1)
k1_big_kernel(K1Args a, unsigned long long *__restrict__ packed) {
    namespace cg = cooperative_groups;
    using B = BCfg<SF>;
    extern __shared__ __align__(128) unsigned char b_raw[];
    BSmem<NSLOT> &sm = *reinterpret_cast<BSmem<NSLOT> *>(b_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const size_t cid = blockIdx.x / B::CL, n_clusters = gridDim.x / B::CL;

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < NSLOT; s++) mbar_init(&sm.bars[s], 1);
        fence_mbar_init();
    }
    for (int i = t; i < 8192; i += 256) sm.chirp[i] = k1_ld_table(a.chirp + (i >> 8) * B::ROWLEN + 256 * rank + (i & 255));
    __syncthreads();
    auto issue = [&](int s, size_t sym) {               // warp 0: 32 bulk copies of one row piece each
        const float2 *src = a.x + sym * (size_t)B::SPS + 256 * rank;
        if (lane == 0) mbar_expect_tx(&sm.bars[s], 65536u);
        __syncwarp();
        bulk_g2s(sm.slots[s] + lane * 256, src + (size_t)lane * B::ROWLEN, 2048u, &sm.bars[s]);
    };
    if (warp == 0) {
#pragma unroll
        for (int s = 0; s < NSLOT; s++) {
            const size_t sym = cid + (size_t)s * n_clusters;
            if (sym < a.n_symbols) issue(s, sym);
        }
    }
    BConsts<SF> c;
    b_consts<SF>(t, rank, a.tw, c);
    float2 *peer[NSLOT][B::CL];
#pragma unroll
    for (int s = 0; s < NSLOT; s++)
#pragma unroll
        for (int q = 0; q < B::CL; q++) peer[s][q] = cluster.map_shared_rank(sm.slots[s], q);

    uint32_t it = 0;
    for (size_t sym = cid; sym < a.n_symbols; sym += n_clusters, it++) {
        const int s = it % NSLOT;
        float2 *slot = sm.slots[s];
        mbar_wait(&sm.bars[s], (it / NSLOT) & 1u);
        float2 v[32];
        b_pass0<SF>(t, slot, sm.chirp, c, v);
        cluster.sync();                                  // every CTA of the cluster has consumed its slot
        b_scatter<SF>(t, rank, v, peer[s]);
        cluster.sync();                                  // all DSMEM stores have landed
        b_pass1<SF>(t, slot, v);
        __syncthreads();
        b_store_t<SF>(t, slot, v);
        __syncthreads();
        b_load_t_radix<SF>(t, slot, c, v);
        __syncthreads();
        b_store2<SF>(t, slot, v);
        __syncthreads();
        unsigned long long best = b_combine<SF>(t, rank, slot, c);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, off);
            best = o > best ? o : best;
        }
        if (lane == 0) sm.keys[warp] = best;
        __syncthreads();
        if (warp == 0) {
            const size_t nxt = sym + (size_t)NSLOT * n_clusters;
            if (nxt < a.n_symbols) {
                fence_proxy_async();
                issue(s, nxt);
            }
            if (lane == 0) {
                unsigned long long bb = sm.keys[0];
#pragma unroll
                for (int k = 1; k < 8; k++) bb = sm.keys[k] > bb ? sm.keys[k] : bb;
                atomicMax(packed + sym, bb);
            }
        }
    }
    cluster.sync();                                      // no CTA exits while a peer may still store into it
}
#endif

// CPU emulation: the CL CTAs run one after another, each on its own slot
template <int SF>
inline void b_emulate(const K1Args &a, uint32_t *bins, float *mags) {
    using B = BCfg<SF>;
    float2 *slot[B::CL], *chirp[B::CL];
    BConsts<SF> *c[B::CL];
    auto v = new float2[B::CL][256][32];
    for (int q = 0; q < B::CL; q++) {
        slot[q] = new float2[8192];
        chirp[q] = new float2[8192];
        c[q] = new BConsts<SF>[256];
        for (int i = 0; i < 8192; i++) chirp[q][i] = a.chirp[(i >> 8) * B::ROWLEN + 256 * q + (i & 255)];
        for (int t = 0; t < 256; t++) b_consts<SF>(t, q, a.tw, c[q][t]);
    }
    for (size_t sym = 0; sym < a.n_symbols; sym++) {
        const float2 *x = a.x + sym * (size_t)B::SPS;
        for (int q = 0; q < B::CL; q++) {
            for (int i = 0; i < 8192; i++) slot[q][i] = x[(i >> 8) * B::ROWLEN + 256 * q + (i & 255)];
            for (int t = 0; t < 256; t++) b_pass0<SF>(t, slot[q], chirp[q], c[q][t], v[q][t]);
        }
        for (int q = 0; q < B::CL; q++)
            for (int i = 0; i < 8192; i++) slot[q][i] = make_float2(NAN, NAN);
        for (int q = 0; q < B::CL; q++)
            for (int t = 0; t < 256; t++) b_scatter<SF>(t, q, v[q][t], slot);
        unsigned long long best = 0ull;
        for (int q = 0; q < B::CL; q++) {
            for (int t = 0; t < 256; t++) b_pass1<SF>(t, slot[q], v[q][t]);
            for (int i = 0; i < 8192; i++) slot[q][i] = make_float2(NAN, NAN);
            for (int t = 0; t < 256; t++) b_store_t<SF>(t, slot[q], v[q][t]);
            for (int t = 0; t < 256; t++) b_load_t_radix<SF>(t, slot[q], c[q][t], v[q][t]);
            for (int i = 0; i < 8192; i++) slot[q][i] = make_float2(NAN, NAN);
            for (int t = 0; t < 256; t++) b_store2<SF>(t, slot[q], v[q][t]);
            for (int t = 0; t < 256; t++) {
                const unsigned long long k = b_combine<SF>(t, q, slot[q], c[q][t]);
                best = k > best ? k : best;
            }
        }
        bins[sym] = key_idx(best);
        if (mags) mags[sym] = sqrtf(key_mag2(best));
    }
    for (int q = 0; q < B::CL; q++) { delete[] slot[q]; delete[] chirp[q]; delete[] c[q]; }
    delete[] v;
}

}  // namespace lb
