// k1_group.cuh -- K1 for SF8 / SF9: a GROUP of W = 2^(SF-7) warps per symbol.
//
// The SF7 warp kernel (k1_warp.cuh) generalised: the T = 32 W threads of a group own a ring of TMA-fed
// shared-memory slots (one symbol = 8*sps bytes each) and run, per symbol,
//   pass 0   float4 #t of each of the 16 rows -> dechirp -> two radix-16 DIF FFTs in registers,
//            inter-pass twiddle u^kc (u = W_N^a lane invariant, powers kept in registers)
//   exch 1   XOR-swizzled 128-bit exchange through the consumed slot: thread (kc, h) receives its
//            NR = 4/W polyphase branches of output column kc, all M0 = 8 W points
//   pass 1   NR radix-M0 FFTs in registers (radix 16 at SF8, radix 32 at SF9)
//   exch 2   second swizzled exchange ([bin][branch]) so that each thread holds all 8 branches of 4 bins
//   combine  Horner over the 8 branches with ONE lane-invariant twiddle per bin, |.|^2, group argmax
// Same arithmetic as get_shift_fft (lib/decoder_impl.cc:430-464); see k1_fft.cuh for the derivation.
// Shared-memory traffic per symbol: 5 x 8*sps bytes (slot read, two exchanges).  The 16 float4 of the down-chirp a thread
// multiplies with are symbol-invariant and live in tensor memory (tmem.cuh; 64 columns per thread, 12 warps = 3 per lane
// quadrant = 192 columns), like k1_sf10's: the shared-memory chirp table of round 1 was a sixth of the traffic of a kernel
// whose L1 / shared pipe was the busiest unit (ncu profiles/r2_k1_sf9.txt: l1tex 68 %).  -DLB_GROUP_CHIRP_SMEM builds the
// old form for A/B runs.  (The 16 KiB this frees at SF8 hold a seventh group, but 14 warps leave 128 registers per thread:
// measured 0.775 against 0.881 with six groups, profiles/r2_group_chirp_tmem_ab.jsonl.)
#pragma once
#include "k1_warp.cuh"
#include "tmem.cuh"
#include "k1_group_consts.h"

namespace lb {

template <int SF>
struct GCfg {
    static constexpr int W = 1 << (SF - 7);          // warps per group
    static constexpr int T = 32 * W;                 // threads per group
    static constexpr int N = 1 << SF, SPS = 8 * N;
    static constexpr int M0 = 8 * W;                 // points of the second FFT (columns)
    static constexpr int NR = 4 / W;                 // branches per thread in pass 1: 4, 2, 1
    static constexpr int LPK = 2 * W;                // lanes per output column kc
    static constexpr int SLOT_F4 = 16 * T;           // float4 per slot
    static constexpr uint32_t SLOT_BYTES = 8u * SPS;
    static_assert(SF >= 7 && SF <= 9, "group kernel: SF7..SF9");
};

template <int SF> LB_HD int g_swz1(int kc) { return SF == 7 ? ((kc & 1) | ((kc & 2) << 1)) : ((kc & 1) << 2); }
template <int SF> LB_HD int g_signed_bin(int q) { return q < GCfg<SF>::N / 2 ? q : q - GCfg<SF>::N; }

template <int SF>
struct GConsts {
    float2 twk[16];          // W_N^{a kc}, a = t >> 2 (pass-0 output twiddle), twk[0] = 1
    float2 wq[4];            // W_sps^{q'} for the thread's bins q = t + T i
    float2 wq2[4];           // wq^2 (Horner over branch PAIRS)
    float2 wb;               // pair-twiddle base of this lane: W_sps^{kc} (times the upper-half factor at SF9)
};

template <int SF>
LB_HD void g_consts(int t, const float2 *tw, GConsts<SF> &c) {
    using C = GCfg<SF>;
    const int a = t >> 2;
    for (int kc = 0; kc < 16; kc++) c.twk[kc] = k1_ld_table(tw + ((a * kc * 8) & (C::SPS - 1)));     // W_N = W_sps^8
    for (int i = 0; i < 4; i++) c.wq[i] = k1_ld_table(tw + (g_signed_bin<SF>(t + C::T * i) & (C::SPS - 1)));
    for (int i = 0; i < 4; i++) c.wq2[i] = k1_ld_table(tw + ((2 * g_signed_bin<SF>(t + C::T * i)) & (C::SPS - 1)));
    {   // w[ka] = W_sps^{q'} with q = kc + 16 ka = wb * g_cc[ka]; at SF9 the odd lane of a pair owns ka >= M0/2:
        // g_cc[ka] = g_cc[ka - M0/2] * W_sps^{16 M0/2 - N}, folded into its wb
        const int kc = t / C::LPK, h = t % C::LPK;
        int e = kc;
        if (C::NR == 1 && (h & 1)) e += 16 * (C::M0 / 2) - C::N;
        c.wb = k1_ld_table(tw + (e & (C::SPS - 1)));
    }
}

template <int SF>
LB_HD float2 g_cc(int ka) {
#ifdef __CUDA_ARCH__
    return SF == 8 ? g_cc8_dev[ka & 15] : g_cc9_dev[ka & 31];
#else
    return SF == 8 ? g_cc8_host[ka & 15] : g_cc9_host[ka & 31];
#endif
}

// pass 0 + exchange-1 write.  slot/chirp: natural sample order as float4 pairs.
template <int SF>
LB_HD void g_pass0(int t, const float4 *slot, const float4 *chirp, const GConsts<SF> &c, float2 *v0, float2 *v1) {
    using C = GCfg<SF>;
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const float4 xv = slot[r * C::T + t];
        const float4 dv = chirp[r * C::T + t];
        v0[r] = cmul(make_float2(xv.x, xv.y), make_float2(dv.x, dv.y));
        v1[r] = cmul(make_float2(xv.z, xv.w), make_float2(dv.z, dv.w));
    }
    dft_dif<16>(v0);
    dft_dif<16>(v1);
#pragma unroll
    for (int kc = 1; kc < 16; kc++) {
        const int br = bitrev<16>(kc);
        v0[br] = cmul(v0[br], c.twk[kc]);
        v1[br] = cmul(v1[br], c.twk[kc]);
    }
}

#if defined(__CUDACC__)
// pass 0 with the thread's chirp samples in tensor memory (tm: lane and first column of this thread)
template <int SF>
LB_D void g_pass0_tm(int t, const float4 *slot, uint32_t tm, const GConsts<SF> &c, float2 *v0, float2 *v1) {
    using C = GCfg<SF>;
    float2 ch[2][8];
    tm_ld16(tm, ch[0]);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        tm_wait_ld();
        if (q < 3) tm_ld16(tm + 16u * (uint32_t)(q + 1), ch[(q + 1) & 1]);      // in flight under this chunk's products
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int r = 4 * q + j;
            const float4 xv = slot[r * C::T + t];
            v0[r] = cmul(make_float2(xv.x, xv.y), ch[q & 1][2 * j]);
            v1[r] = cmul(make_float2(xv.z, xv.w), ch[q & 1][2 * j + 1]);
        }
    }
    dft_dif<16>(v0);
    dft_dif<16>(v1);
#pragma unroll
    for (int kc = 1; kc < 16; kc++) {
        const int br = bitrev<16>(kc);
        v0[br] = cmul(v0[br], c.twk[kc]);
        v1[br] = cmul(v1[br], c.twk[kc]);
    }
}
#endif

template <int SF>
LB_HD void g_store1(int t, float4 *slot, const float2 *v0, const float2 *v1) {
    using C = GCfg<SF>;
#pragma unroll
    for (int kc = 0; kc < 16; kc++) {
        const int br = bitrev<16>(kc);
        slot[kc * C::T + (t ^ g_swz1<SF>(kc))] = make_float4(v0[br].x, v0[br].y, v1[br].x, v1[br].y);
    }
}

// pass 1: thread (kc = t / LPK, h = t % LPK) loads branches r = h*NR .. h*NR+NR-1, all M0 columns,
// and runs NR radix-M0 FFTs.  g[i][bitrev(ka)] = G_{h NR + i}[kc + 16 ka].
template <int SF>
LB_HD void g_pass1(int t, const float4 *slot, float2 (*g)[GCfg<SF>::M0]) {
    using C = GCfg<SF>;
    const int kc = t / C::LPK, h = t % C::LPK;
    const int sw = g_swz1<SF>(kc);
    const float4 *row = slot + kc * C::T;
#pragma unroll
    for (int a = 0; a < C::M0; a++) {
        if constexpr (C::NR == 4) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const float4 u = row[(4 * a + 2 * h + e) ^ sw];
                g[2 * e][a] = make_float2(u.x, u.y);
                g[2 * e + 1][a] = make_float2(u.z, u.w);
            }
        } else if constexpr (C::NR == 2) {
            const float4 u = row[(4 * a + h) ^ sw];
            g[0][a] = make_float2(u.x, u.y);
            g[C::NR - 1][a] = make_float2(u.z, u.w);
        } else {
            const float2 *p2 = reinterpret_cast<const float2 *>(row + ((4 * a + (h >> 1)) ^ sw));
            g[0][a] = p2[h & 1];
        }
    }
#pragma unroll
    for (int i = 0; i < C::NR; i++) dft_dif<C::M0>(g[i]);
}

// exchange 2: [bin q][16-byte unit u = r/2], unit position XOR ((q >> 1) & 3); 4 units (64 B) per bin
LB_HD int g_unit2(int q, int u) { return q * 4 + (u ^ ((q >> 1) & 3)); }

template <int SF>
LB_HD void g_store2(int t, float4 *slot, float2 (*g)[GCfg<SF>::M0]) {
    using C = GCfg<SF>;
    const int kc = t / C::LPK, h = t % C::LPK;
#pragma unroll
    for (int ka = 0; ka < C::M0; ka++) {
        const int br = bitrev<C::M0>(ka);
        const int q = kc + 16 * ka;
        if constexpr (C::NR == 4) {
            slot[g_unit2(q, 2 * h)] = make_float4(g[0][br].x, g[0][br].y, g[1][br].x, g[1][br].y);
            slot[g_unit2(q, 2 * h + 1)] = make_float4(g[2][br].x, g[2][br].y, g[3][br].x, g[3][br].y);
        } else if constexpr (C::NR == 2) {
            slot[g_unit2(q, h)] = make_float4(g[0][br].x, g[0][br].y, g[C::NR - 1][br].x, g[C::NR - 1][br].y);
        } else {
            float2 *p2 = reinterpret_cast<float2 *>(slot + g_unit2(q, h >> 1));
            p2[h & 1] = g[0][br];
        }
    }
}

// combine: bins q = t + T i, i = 0..3: all 8 branches, Horner with w = W_sps^{q'}
template <int SF>
LB_HD unsigned long long g_combine(int t, const float4 *slot, const GConsts<SF> &c) {
    using C = GCfg<SF>;
    unsigned long long best = 0ull;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int q = t + C::T * i;
        float2 gv[8];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const float4 v = slot[g_unit2(q, u)];
            gv[2 * u] = make_float2(v.x, v.y);
            gv[2 * u + 1] = make_float2(v.z, v.w);
        }
        const float2 w = c.wq[i];
        float2 acc = gv[7];
#pragma unroll
        for (int r = 6; r >= 0; r--) acc = cfma(acc, w, gv[r]);
        if (q == C::N / 2) {                             // tmp[N/2] += F[N/2]  (:450); thread 0, i = 2
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

// ---- pair variant (SF8, SF9): combine branch pairs BEFORE the second exchange ---------------------
// P_pr[q] = G_{2pr}[q] + w[q] G_{2pr+1}[q].  SF8 (NR = 2): both branches are in the thread.  SF9 (NR = 1):
// the partner lane t^1 holds the other branch; the even lane finishes ka < M0/2, the odd lane ka >= M0/2
// (own = this lane's G, bit-reversed; other = what the partner sent for this lane's half).
// Returns NP = M0 (SF8) or M0/2 (SF9) values P[j] for ka = j (+ M0/2 on odd SF9 lanes).
// Pq = the same with the conjugate twiddle for the quirk bin q = N/2 (kc = 0, ka = M0/2), else 0.
template <int SF> struct GPair { static constexpr int NP = GCfg<SF>::NR == 2 ? GCfg<SF>::M0 : GCfg<SF>::M0 / 2; };

template <int SF>
LB_HD void g_pair_sf8(int t, float2 (*g)[GCfg<SF>::M0], const GConsts<SF> &c, float2 *P, float2 &Pq) {
    using C = GCfg<SF>;
    const int kc = t / C::LPK;
    Pq = make_float2(0.f, 0.f);
#pragma unroll
    for (int ka = 0; ka < C::M0; ka++) {
        const int br = bitrev<C::M0>(ka);
        const float2 od = cmul(g[C::NR - 1][br], g_cc<SF>(ka));
        P[ka] = cfma(od, c.wb, g[0][br]);
        if (ka == C::M0 / 2 && kc == 0) Pq = cfma(cmul(g[C::NR - 1][br], cconj(g_cc<SF>(ka))), cconj(c.wb), g[0][br]);
    }
}

// SF9: keep[j]/other[j] for ka = j + (odd ? M0/2 : 0); even_part/odd_part by lane parity
template <int SF>
LB_HD void g_pair_sf9(int t, const float2 *keep, const float2 *recv, const GConsts<SF> &c, float2 *P, float2 &Pq) {
    using C = GCfg<SF>;
    const int kc = t / C::LPK, odd = t & 1;
    Pq = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < C::M0 / 2; j++) {
        const float2 ev = odd ? recv[j] : keep[j];
        const float2 od = odd ? keep[j] : recv[j];
        const float2 odc = cmul(od, g_cc<SF>(j));
        P[j] = cfma(odc, c.wb, ev);
        if (j == 0 && kc == 0 && odd) Pq = cfma(cmul(od, cconj(g_cc<SF>(j))), cconj(c.wb), ev);   // ka = M0/2
    }
}

// exchange 2 (pair variant): [bin q][pair pr] float2, 4 pairs = 32 B per bin, 16-byte unit XOR (q >> 2) & 1
// (rows of the upper half, written by the odd lanes at SF9, take the other 64-byte half of the bank window)
LB_HD int g_row2p(int q) { return q ^ (((q >> 8) & 1) << 1); }
LB_HD int g_pos2p(int q, int pr) { return g_row2p(q) * 4 + ((((pr >> 1) ^ ((q >> 2) & 1)) << 1) | (pr & 1)); }

template <int SF>
LB_HD void g_store2p(int t, float2 *slot2, const float2 *P, float2 Pq, float2 *quirk) {
    using C = GCfg<SF>;
    const int kc = t / C::LPK, h = t % C::LPK;
    const int pr = C::NR == 2 ? h : (h >> 1);
    const int ka0 = (C::NR == 1 && (h & 1)) ? C::M0 / 2 : 0;
#pragma unroll
    for (int j = 0; j < GPair<SF>::NP; j++) slot2[g_pos2p(kc + 16 * (ka0 + j), pr)] = P[j];
    if (kc == 0 && (C::NR == 2 || (h & 1))) quirk[pr] = Pq;
}

template <int SF>
LB_HD unsigned long long g_combine_p(int t, const float2 *slot2, const float2 *quirk, const GConsts<SF> &c) {
    using C = GCfg<SF>;
    unsigned long long best = 0ull;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int q = t + C::T * i;
        float2 pv[4];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const float4 v = *reinterpret_cast<const float4 *>(slot2 + g_row2p(q) * 4 + ((u ^ ((q >> 2) & 1)) << 1));
            pv[2 * u] = make_float2(v.x, v.y);
            pv[2 * u + 1] = make_float2(v.z, v.w);
        }
        const float2 w2 = c.wq2[i];
        float2 acc = cfma(pv[3], w2, pv[2]);
        acc = cfma(acc, w2, pv[1]);
        acc = cfma(acc, w2, pv[0]);
        if (q == C::N / 2) {                             // tmp[N/2] += F[N/2]  (:450)
            const float2 wc = cconj(w2);
            float2 a2 = cfma(quirk[3], wc, quirk[2]);
            a2 = cfma(a2, wc, quirk[1]);
            a2 = cfma(a2, wc, quirk[0]);
            acc = cadd(acc, a2);
        }
        const unsigned long long key = pack_key(cnorm2(acc), (uint32_t)q);
        best = key > best ? key : best;
    }
    return best;
}

#ifdef __CUDACC__
template <int SF, int NGROUPS, int NSLOT>
struct GSmem {
#ifdef LB_GROUP_CHIRP_SMEM
    float4 chirp[GCfg<SF>::SLOT_F4];
#endif
    float4 slots[NGROUPS][NSLOT][GCfg<SF>::SLOT_F4];
    uint64_t bars[NGROUPS][NSLOT];
    unsigned long long keys[NGROUPS][GCfg<SF>::W];
    float2 quirk[NGROUPS][4];
    uint32_t tm_base;
};
constexpr int G_TM_COLS = 256;                          // 64 columns per thread, up to four warps per lane quadrant

LB_D void group_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

template <int SF, int NGROUPS, int NSLOT>
__global__ void __launch_bounds__(NGROUPS * GCfg<SF>::T, 1)
k1_group_kernel(K1Args a, uint32_t *__restrict__ bins, float *__restrict__ mags) {
    using C = GCfg<SF>;
    extern __shared__ __align__(128) unsigned char g_raw[];
    GSmem<SF, NGROUPS, NSLOT> &sm = *reinterpret_cast<GSmem<SF, NGROUPS, NSLOT> *>(g_raw);
    const int grp = threadIdx.x / C::T, t = threadIdx.x % C::T;
    const int lane = t & 31, wig = t >> 5;              // warp in group
    const int bar_id = 1 + grp;                         // named barrier of this group (0 = __syncthreads)
    const size_t gg = (size_t)blockIdx.x * NGROUPS + grp, g_total = (size_t)gridDim.x * NGROUPS;

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < NSLOT; s++) mbar_init(&sm.bars[grp][s], 1);
        fence_mbar_init();
    }
#ifdef LB_GROUP_CHIRP_SMEM
    for (int i = threadIdx.x; i < C::SLOT_F4; i += NGROUPS * C::T) sm.chirp[i] = k1_ld_table4(a.chirp + 2 * i);
    __syncthreads();
#else
    static_assert(NGROUPS * GCfg<SF>::W <= 16, "four warps per lane quadrant at most");
    const int wcta = threadIdx.x >> 5;                  // warp of the CTA
    if (wcta == 0) tm_alloc<G_TM_COLS>(&sm.tm_base);
    tm_fence_before();
    __syncthreads();
    tm_fence_after();
    const uint32_t tm = sm.tm_base + ((uint32_t)(32 * (wcta & 3)) << 16) + (uint32_t)(64 * (wcta >> 2));
#pragma unroll
    for (int q = 0; q < 4; q++) {                       // chirp float4 #(r * T + t), r = 4 q .. 4 q + 3 -> columns 16 q .. 16 q + 15
        float2 buf[8];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float4 dv = k1_ld_table4(a.chirp + 2 * ((4 * q + j) * C::T + t));
            buf[2 * j] = make_float2(dv.x, dv.y);
            buf[2 * j + 1] = make_float2(dv.z, dv.w);
        }
        tm_st16(tm + 16u * (uint32_t)q, buf);
    }
    tm_wait_st();
#endif
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < NSLOT; s++) {
            const size_t sym = gg + (size_t)s * g_total;
            if (sym < a.n_symbols) {
                mbar_expect_tx(&sm.bars[grp][s], C::SLOT_BYTES);
                bulk_g2s(sm.slots[grp][s], a.x + sym * C::SPS, C::SLOT_BYTES, &sm.bars[grp][s]);
            }
        }
    }
    GConsts<SF> c;
    g_consts<SF>(t, a.tw, c);

    uint32_t it = 0;
    for (size_t sym = gg; sym < a.n_symbols; sym += g_total, it++) {
        const int s = it % NSLOT;
        const uint32_t parity = (it / NSLOT) & 1u;
        float4 *slot = sm.slots[grp][s];
        mbar_wait(&sm.bars[grp][s], parity);
        {
            float2 v0[16], v1[16];
#ifdef LB_GROUP_CHIRP_SMEM
            g_pass0<SF>(t, slot, sm.chirp, c, v0, v1);
#else
            g_pass0_tm<SF>(t, slot, tm, c, v0, v1);
#endif
            group_bar(bar_id, C::T);                    // everyone has read the slot
            g_store1<SF>(t, slot, v0, v1);
        }
        group_bar(bar_id, C::T);
        float2 g[C::NR][C::M0];
        g_pass1<SF>(t, slot, g);
        unsigned long long best;
        if constexpr (C::NR == 4) {
            group_bar(bar_id, C::T);                    // exchange-1 reads done
            g_store2<SF>(t, slot, g);
            group_bar(bar_id, C::T);
            best = g_combine<SF>(t, slot, c);
        } else {
            float2 P[GPair<SF>::NP], Pq;
            if constexpr (C::NR == 2) {
                g_pair_sf8<SF>(t, g, c, P, Pq);
            } else {
                const int odd = t & 1;
                float2 keep[C::M0 / 2], recv[C::M0 / 2];
#pragma unroll
                for (int j = 0; j < C::M0 / 2; j++) {
                    const float2 lo = g[0][bitrev<C::M0>(j)], hi = g[0][bitrev<C::M0>(j + C::M0 / 2)];
                    keep[j] = odd ? hi : lo;
                    const float2 send = odd ? lo : hi;
                    recv[j].x = __shfl_xor_sync(0xffffffffu, send.x, 1);
                    recv[j].y = __shfl_xor_sync(0xffffffffu, send.y, 1);
                }
                g_pair_sf9<SF>(t, keep, recv, c, P, Pq);
            }
            group_bar(bar_id, C::T);                    // exchange-1 reads done
            g_store2p<SF>(t, reinterpret_cast<float2 *>(slot), P, Pq, sm.quirk[grp]);
            group_bar(bar_id, C::T);
            best = g_combine_p<SF>(t, reinterpret_cast<const float2 *>(slot), sm.quirk[grp], c);
        }
        best = warp_max_key(best);
        if (lane == 0) sm.keys[grp][wig] = best;
        group_bar(bar_id, C::T);                        // exchange-2 reads done + keys visible
        if (t == 0) {
            const size_t nxt = sym + (size_t)NSLOT * g_total;
            if (nxt < a.n_symbols) {
                fence_proxy_async();
                mbar_expect_tx(&sm.bars[grp][s], C::SLOT_BYTES);
                bulk_g2s(slot, a.x + nxt * C::SPS, C::SLOT_BYTES, &sm.bars[grp][s]);
            }
            unsigned long long bb = sm.keys[grp][0];
#pragma unroll
            for (int k = 1; k < C::W; k++) bb = sm.keys[grp][k] > bb ? sm.keys[grp][k] : bb;
            bins[sym] = key_idx(bb);
            if (mags) mags[sym] = sqrtf(key_mag2(bb));
        }
        // keys[] is rewritten only after four more group barriers: no hazard with thread 0's read
    }
#ifndef LB_GROUP_CHIRP_SMEM
    tm_fence_before();
    __syncthreads();                                    // every group has finished its symbols
    tm_fence_after();
    if (wcta == 0) tm_dealloc<G_TM_COLS>(sm.tm_base);
#endif
}
#endif  // __CUDACC__

// ---- CPU emulation ---------------------------------------------------------------------------
template <int SF>
inline void g_emulate(const K1Args &a, uint32_t *bins, float *mags) {
    using C = GCfg<SF>;
    float4 *slot = new float4[C::SLOT_F4];
    float4 *chirp = new float4[C::SLOT_F4];
    GConsts<SF> *c = new GConsts<SF>[C::T];
    for (int t = 0; t < C::T; t++) g_consts<SF>(t, a.tw, c[t]);
    for (int i = 0; i < C::SLOT_F4; i++) chirp[i] = make_float4(a.chirp[2 * i].x, a.chirp[2 * i].y, a.chirp[2 * i + 1].x, a.chirp[2 * i + 1].y);
    auto v0 = new float2[C::T][16];
    auto v1 = new float2[C::T][16];
    auto g = new float2[C::T][C::NR][C::M0];
    for (size_t sym = 0; sym < a.n_symbols; sym++) {
        const float2 *x = a.x + sym * C::SPS;
        for (int i = 0; i < C::SLOT_F4; i++) slot[i] = make_float4(x[2 * i].x, x[2 * i].y, x[2 * i + 1].x, x[2 * i + 1].y);
        for (int t = 0; t < C::T; t++) g_pass0<SF>(t, slot, chirp, c[t], v0[t], v1[t]);
        for (int i = 0; i < C::SLOT_F4; i++) slot[i] = make_float4(NAN, NAN, NAN, NAN);
        for (int t = 0; t < C::T; t++) g_store1<SF>(t, slot, v0[t], v1[t]);
        for (int t = 0; t < C::T; t++) g_pass1<SF>(t, slot, g[t]);
        for (int i = 0; i < C::SLOT_F4; i++) slot[i] = make_float4(NAN, NAN, NAN, NAN);
        unsigned long long best = 0ull;
        if constexpr (C::NR == 4) {
            for (int t = 0; t < C::T; t++) g_store2<SF>(t, slot, g[t]);
            for (int t = 0; t < C::T; t++) {
                const unsigned long long k = g_combine<SF>(t, slot, c[t]);
                best = k > best ? k : best;
            }
        } else {
            float2 quirk[4] = {};
            float2 *slot2 = reinterpret_cast<float2 *>(slot);
            for (int t = 0; t < C::T; t++) {
                float2 P[GPair<SF>::NP], Pq;
                if constexpr (C::NR == 2) {
                    g_pair_sf8<SF>(t, g[t], c[t], P, Pq);
                } else {
                    const int odd = t & 1;
                    float2 keep[C::M0 / 2], recv[C::M0 / 2];
                    for (int j = 0; j < C::M0 / 2; j++) {
                        const int mine = bitrev<C::M0>(j + (odd ? C::M0 / 2 : 0));
                        keep[j] = g[t][0][mine];
                        recv[j] = g[t ^ 1][0][mine];          // what the partner sends: its value at MY ka
                    }
                    g_pair_sf9<SF>(t, keep, recv, c[t], P, Pq);
                }
                g_store2p<SF>(t, slot2, P, Pq, quirk);
            }
            for (int t = 0; t < C::T; t++) {
                const unsigned long long k = g_combine_p<SF>(t, slot2, quirk, c[t]);
                best = k > best ? k : best;
            }
        }
        bins[sym] = key_idx(best);
        if (mags) mags[sym] = sqrtf(key_mag2(best));
    }
    delete[] slot; delete[] chirp; delete[] c; delete[] v0; delete[] v1; delete[] g;
}

}  // namespace lb
