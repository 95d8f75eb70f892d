// k1_xchg.cuh -- K1 for SF10 / SF11 / SF12: a TEAM of CL sub-CTAs (on CL different SMs) per symbol, the symbol split by
// columns, the all-to-all between pass 0 and the rest through an L2-resident scratch in global memory.  Default for SF12.
//
// History of this file (profiles/r1_k1_xchg_sf12.md has the table).  The first cluster kernel (k1_big.cuh) read every
// sample once and was lean in instructions but used 30 % of the issue slots: one 256-thread group per SM in lock step,
// two cluster-wide barriers per symbol, 8-byte remote stores.  The versions tried here, same arithmetic every time:
//   1. hardware clusters, pass-0 outputs moved by cp.async.bulk smem -> peer smem with mbarrier complete_tx at the
//      receiver and credit mbarriers: no cluster barrier, but the copy engine moves only ~10 GB/s per SM;
//   2. no hardware cluster: TMA stores into an L2-resident scratch, flags in global memory, TMA loads back;
//   3. one 512-thread CTA per SM = two sub-CTAs sharing the chirp columns, pass 0 two symbols ahead, outputs stored
//      straight from registers, double-buffered TMA slots (the consumed one is the step's scratch), group barriers
//      only, control by the last warp to arrive  <- what is below.
// All of them land at 0.21-0.28 of the HBM peak; switching the cross-SM waits off changes < 10 %.
//
// Index algebra (get_shift_fft, lib/decoder_impl.cc:430-464; pruned DFT as in k1_fft.cuh).  sample n = 8 n1 + r,
// n1 = c L + a (row c < 16, a < L = N/16), kept bin k = kc + 16 ka (kc < 16, ka < L):
//   G_r[k] = sum_a W_L^{a ka} ( W_N^{a kc} sum_c W_16^{c kc} y_r[c L + a] )            pass 0: radix 16 over c
// CTA `rank` loads columns a in [32 rank, 32 rank + 32) of every row and afterwards owns kc in
// [KPC rank, KPC rank + KPC).  With a = j + M2 a_hi (M2 = L/16 = 2 CL), ka = ka_lo + 16 ka_hi:
//   G_r[k] = sum_j W_M2^{j ka_hi} ( W_L^{j ka_lo} sum_{a_hi} W_16^{a_hi ka_lo} U[j + M2 a_hi] )
//                                                           pass 1: radix 16 over a_hi;  pass 2: radix M2 over j
//   F[k'] = sum_r W_sps^{k' r} G_r[k],  k' = k < N/2 ? k : k - N;   tmp[N/2] += F[N/2]   (:447-450)
#pragma once
#include "k1_sf10.cuh"

namespace lb {

template <int SF, int TH>
struct XCfg {
    static_assert(SF >= 10 && SF <= 12, "k1_xchg: SF10..SF12");
    static_assert(TH == 128 || TH == 256, "k1_xchg: 128 or 256 threads per CTA");
    static constexpr int T = TH;                         // threads per CTA = samples per row piece
    static constexpr int N = 1 << SF, SPS = 8 * N;
    static constexpr int L = N / 16;                     // length of the per-kc FFT
    static constexpr int ROWLEN = 8 * L;                 // samples per row; 16 rows per symbol
    static constexpr int CL = ROWLEN / TH;               // sub-CTAs per team: TH=256: 2, 4, 8;  TH=128: 4, 8, 16
    static constexpr int AW = TH / 8;                    // columns a per CTA
    static constexpr int M2 = L / 16;                    // 4, 8, 16
    static constexpr int KPC = 16 / CL;                  // output columns kc per CTA
    static constexpr int NI = 16 / M2;                   // ka_lo values per thread in pass 2: 4, 2, 1
    static constexpr int WPK = M2 / 4;                   // warps per kc_l in the receiver mapping: 1, 2, 4
    static constexpr int F2 = 16 * TH;                   // float2 per buffer (32 / 16 KiB)
    static constexpr uint32_t BYTES = 128u * TH, RUN = 8u * TH;
    static_assert(KPC >= 1 && KPC * WPK * 32 == TH, "thread mapping");
};

template <int SF, int TH>
struct XConsts {
    float2 tl[4], th[4];     // pass 0: W_N^{a kl}, W_N^{4 a kh}            (sender role, a = 32 rank + (t >> 3))
    float2 ul[4], uh[4];     // pass 1: W_L^{j l},  W_L^{4 j h}             (receiver role, j)
    float2 wq[2];            // combine: W_sps^{k'} of the thread's two bins
};

// receiver-role decode of a thread index: lane = j_lo * 8 + r, warp = kc_l * WPK + j_hi
template <int SF, int TH> LB_HD int xg_r(int t) { return t & 7; }
template <int SF, int TH> LB_HD int xg_j(int t) { return 4 * ((t >> 5) % XCfg<SF, TH>::WPK) + ((t >> 3) & 3); }
template <int SF, int TH> LB_HD int xg_kcl(int t) { return (t >> 5) / XCfg<SF, TH>::WPK; }

// the two local bins thread t combines: they belong to its own warp group (kc_l), GT = 32 WPK = L / 2 threads
template <int SF, int TH>
LB_HD int xg_my_bin(int t, int i) {
    using X = XCfg<SF, TH>;
    return xg_kcl<SF, TH>(t) * X::L + (t % (32 * X::WPK)) + (32 * X::WPK) * i;
}

// local bin (kc_l, ka) -> bin of the whole symbol
template <int SF, int TH>
LB_HD int xg_global_bin(int bl, int rank) {
    using X = XCfg<SF, TH>;
    return (X::KPC * rank + bl / X::L) + 16 * (bl % X::L);
}

template <int SF, int TH>
LB_HD void xg_consts(int t, int rank, const float2 *tw, XConsts<SF, TH> &c) {
    using X = XCfg<SF, TH>;
    const int a = X::AW * rank + (t >> 3);
    const int j = xg_j<SF, TH>(t);
    for (int i = 0; i < 4; i++) {
        c.tl[i] = k1_ld_table(tw + ((a * i * 8) & (X::SPS - 1)));                 // W_N = W_sps^8
        c.th[i] = k1_ld_table(tw + ((a * 4 * i * 8) & (X::SPS - 1)));
        c.ul[i] = k1_ld_table(tw + ((j * i * 128) & (X::SPS - 1)));               // W_L = W_sps^128
        c.uh[i] = k1_ld_table(tw + ((j * 4 * i * 128) & (X::SPS - 1)));
    }
    for (int i = 0; i < 2; i++) {
        const int q = xg_global_bin<SF, TH>(xg_my_bin<SF, TH>(t, i), rank);
        c.wq[i] = k1_ld_table(tw + ((q < X::N / 2 ? q : q - X::N) & (X::SPS - 1)));
    }
}

// v[bitrev(k)] *= lo[k & 3] * hi[k >> 2]   (k = 1..15; index-0 factors are 1 and skipped)
LB_HD void xg_twiddle16(float2 *v, const float2 *lo, const float2 *hi) {
#pragma unroll
    for (int k = 1; k < 16; k++) {
        const int br = bitrev<16>(k);
        float2 w;
        if ((k & 3) == 0) w = hi[k >> 2];
        else if ((k >> 2) == 0) w = lo[k & 3];
        else w = cmul(hi[k >> 2], lo[k & 3]);
        v[br] = cmul(v[br], w);
    }
}

// where output column kc of piece `rank` lands: exchange image of CTA kc / KPC, float2 offset ((kc % KPC) CL + rank) * TH,
// so that everything warp group kc_l = kc % KPC needs sits in its own block [kc_l CL TH, (kc_l + 1) CL TH)
template <int SF, int TH> LB_HD int xg_dst_cta(int kc) { return kc / XCfg<SF, TH>::KPC; }
template <int SF, int TH> LB_HD int xg_dst_off(int kc, int rank) { return ((kc % XCfg<SF, TH>::KPC) * XCfg<SF, TH>::CL + rank) * TH; }

// pass 0, part 1: slot[c][t] (row c, column t of this CTA's piece) x chirp -> registers (the slot is free afterwards)
template <int SF, int TH>
LB_HD void xg_pass0_load(int t, const float2 *slot, const float2 *chirp, float2 *v) {
    using X = XCfg<SF, TH>;
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = cmul(slot[r * X::T + t], chirp[r * X::T + t]);
}
// part 2: radix 16 over the rows -> twiddle W_N^{a kc}
template <int SF, int TH>
LB_HD void xg_pass0_fft(const XConsts<SF, TH> &c, float2 *v) {
    dft_dif<16>(v);
    xg_twiddle16(v, c.tl, c.th);
}
// part 3: output column kc of piece `rank` -> exchange image of CTA kc / KPC (img = the CL images of one symbol).
// A warp writes 32 consecutive float2 (256 B) per kc.
template <int SF, int TH>
LB_HD void xg_put(int t, int rank, float2 *img, const float2 *v) {
    using X = XCfg<SF, TH>;
#pragma unroll
    for (int kc = 0; kc < 16; kc++)
        img[(size_t)xg_dst_cta<SF, TH>(kc) * X::F2 + xg_dst_off<SF, TH>(kc, rank) + t] = v[bitrev<16>(kc)];
}



// pass 1: thread (kc_l, j, r) gathers a = j + M2 a_hi, radix 16 over a_hi, twiddle W_L^{j ka_lo}
template <int SF, int TH>
LB_HD void xg_pass1(int t, const float2 *rx, const XConsts<SF, TH> &c, float2 *f) {
    using X = XCfg<SF, TH>;
    const int r = xg_r<SF, TH>(t), j = xg_j<SF, TH>(t), kcl = xg_kcl<SF, TH>(t);
#pragma unroll
    for (int ah = 0; ah < 16; ah++) {
        const int a = j + X::M2 * ah;
        f[ah] = rx[(kcl * X::CL + a / X::AW) * X::T + (a % X::AW) * 8 + r];
    }
    dft_dif<16>(f);
    xg_twiddle16(f, c.ul, c.uh);
}

// transpose write: item (kc_l, r), logical offset ka_lo M2 + j, low 4 bits XOR 2r (bank spread, keeps pairs)
template <int SF, int TH>
LB_HD void xg_store_t(int t, float2 *s1, const float2 *f) {
    using X = XCfg<SF, TH>;
    const int r = xg_r<SF, TH>(t), j = xg_j<SF, TH>(t), kcl = xg_kcl<SF, TH>(t);
    float2 *item = s1 + (kcl * 8 + r) * X::L;
#pragma unroll
    for (int kl = 0; kl < 16; kl++) item[(kl * X::M2 + j) ^ (2 * r)] = f[bitrev<16>(kl)];
}

// pass 2: thread (kc_l, j', r) takes ka_lo = j' + M2 i, reads all j (128-bit), radix M2 -> g[i M2 + ka_hi]
template <int SF, int TH>
LB_HD void xg_pass2(int t, const float2 *s1, float2 *g) {
    using X = XCfg<SF, TH>;
    const int r = xg_r<SF, TH>(t), j = xg_j<SF, TH>(t), kcl = xg_kcl<SF, TH>(t);
    const float2 *item = s1 + (kcl * 8 + r) * X::L;
#pragma unroll
    for (int i = 0; i < X::NI; i++) {
        const int kl = j + X::M2 * i;
        float2 z[X::M2];
#pragma unroll
        for (int p = 0; p < X::M2 / 2; p++) {
            const float4 u = *reinterpret_cast<const float4 *>(item + ((kl * X::M2 + 2 * p) ^ (2 * r)));
            z[2 * p] = make_float2(u.x, u.y);
            z[2 * p + 1] = make_float2(u.z, u.w);
        }
        dft_dif<X::M2>(z);
#pragma unroll
        for (int kh = 0; kh < X::M2; kh++) g[i * X::M2 + kh] = z[bitrev<X::M2>(kh)];
    }
}

// exchange 2: [local bin][branch], the 16-byte unit of the row XOR-swizzled by (bin >> 1) & 3
LB_HD int xg_pos2(int bl, int r) { return bl * 8 + ((((r >> 1) ^ ((bl >> 1) & 3)) << 1) | (r & 1)); }

template <int SF, int TH>
LB_HD void xg_store2(int t, float2 *s2, const float2 *g) {
    using X = XCfg<SF, TH>;
    const int r = xg_r<SF, TH>(t), j = xg_j<SF, TH>(t), kcl = xg_kcl<SF, TH>(t);
#pragma unroll
    for (int i = 0; i < X::NI; i++)
#pragma unroll
        for (int kh = 0; kh < X::M2; kh++) s2[xg_pos2(kcl * X::L + (j + X::M2 * i) + 16 * kh, r)] = g[i * X::M2 + kh];
}

template <int SF, int TH>
LB_HD unsigned long long xg_combine(int t, int rank, const float2 *s2, const XConsts<SF, TH> &c) {
    using X = XCfg<SF, TH>;
    unsigned long long best = 0ull;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int bl = xg_my_bin<SF, TH>(t, i);
        const int q = xg_global_bin<SF, TH>(bl, rank);
        float2 gv[8];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const float4 v = *reinterpret_cast<const float4 *>(s2 + bl * 8 + ((u ^ ((bl >> 1) & 3)) << 1));
            gv[2 * u] = make_float2(v.x, v.y);
            gv[2 * u + 1] = make_float2(v.z, v.w);
        }
        const float2 w = c.wq[i];
        float2 acc = gv[7];
#pragma unroll
        for (int r = 6; r >= 0; r--) acc = cfma(acc, w, gv[r]);
        if (q == X::N / 2) {                             // tmp[N/2] += F[N/2]  (:450)
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
// ---- flag primitives --------------------------------------------------------------------------
LB_D void xg_fence_proxy_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
LB_D void xg_flag_add(uint32_t *flag) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(flag) : "memory"); }
LB_D uint32_t xg_flag_ld(const uint32_t *flag) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    return v;
}
LB_D void xg_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int XG_D = 2;            // pass 0 runs XG_D symbols ahead of the rest
constexpr int XG_NB = 2 * XG_D;    // exchange buffers per team in global memory (see the flag protocol below)

// watchdog (tools/k1_ab.py, LORA_B200_XG_WATCHDOG=1): a spin that lasts too long leaves one record in a host-mapped
// buffer, dbg[0] = count, dbg[1..] = site:8 | block:12 | sub-CTA:2 | warp:4 | symbol:16 | value:22
LB_D void xg_dbg(unsigned long long *dbg, unsigned site, unsigned half, unsigned warp, unsigned sym, unsigned val) {
    if (!dbg) return;
    const unsigned idx = atomicAdd(reinterpret_cast<unsigned *>(dbg), 1u);
    if (idx < 255u) {
        dbg[1 + idx] = ((unsigned long long)(site & 0xffu) << 56) | ((unsigned long long)(blockIdx.x & 0xfffu) << 44) |
                       ((unsigned long long)(half & 3u) << 42) | ((unsigned long long)(warp & 15u) << 38) |
                       ((unsigned long long)(sym & 0xffffu) << 22) | (unsigned long long)(val & 0x3fffffu);
        __threadfence_system();
    }
}
LB_D void xg_wait(uint64_t *bar, uint32_t parity, unsigned long long *dbg, unsigned site, unsigned half, unsigned warp, unsigned sym) {
    uint32_t ok, spins = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!ok && ++spins == (1u << 20) && (threadIdx.x & 31) == 0) xg_dbg(dbg, site, half, warp, sym, parity);
    } while (!ok);
}

template <int TH>
struct XHalf {
    float2 slot[2][16 * TH]; // TMA destinations: symbol m -> slot[m & 1].  Once pass 0 has read a slot into registers it
                             // serves as the scratch of that step's two local exchanges and is refilled at the end of the
                             // step, a whole step before it is needed again (the other slot feeds the next step)
    float2 rx[16 * TH];      // this CTA's output columns of all CL pieces (free again once pass 1 has gathered it)
};
template <int TH>
struct XSmem {
    float2 chirp[16 * TH];   // this CTA's columns of the down-chirp, [row][TH], shared by the sub-CTAs
    XHalf<TH> h[512 / TH];
    uint64_t slot_full[512 / TH][2];
    uint64_t rx_full[512 / TH];
    uint64_t slot_free[512 / TH];      // one arrival per warp: pass 0 has read the slot (it may be used as scratch)
    uint32_t cnt_done[512 / TH];       // warps that have finished a step (monotonic; the last one refills the slot)
    uint32_t cnt_loaded[512 / TH];     // warps that have read the slot (monotonic; the last one refills it)
    uint32_t cnt_stored[512 / TH][2];  // warps that have written their pass-0 outputs of an even / odd symbol (the last
                                       // one raises the flag; two counters because a warp may finish symbol m + 1
                                       // before a slow one has finished symbol m)
    uint32_t cnt_gathered[512 / TH];   // warps that have gathered rx (the last one fetches the next image)
};

// One CTA of 512 threads per SM = 512 / TH independent sub-CTAs of TH threads with the same `rank` (they share the
// chirp columns) that belong to different teams.  A team = the CL sub-CTAs (one per rank, on CL different SMs) that
// split one symbol by columns.  Per symbol a sub-CTA does
//   A: TMA-load its 16 row pieces, dechirp, radix 16 over the rows, and write output column kc straight from
//      registers into the exchange image of rank kc / KPC (global memory, L2 resident, coalesced 256 B per warp);
//   B: TMA-load its own image (32 / 16 KiB contiguous), pass 1, pass 2, combine, argmax.
// A runs XG_D symbols ahead of B, so the store -> flag -> load round trip of the exchange is never waited for; both
// TMA destinations are released as soon as their contents are in registers, so the next symbol's data streams in
// during the whole computation; and no barrier spans more than the WPK warps of one kc_l group: control work
// (refills, flags, fetches) is done by whichever warp arrives LAST at the point that enables it.
// Flag protocol: flags[team][b] counts the ranks whose outputs of the symbols m = b (mod XG_NB) are written; a rank
// adds its count for symbol it + XG_D at the end of step `it`, i.e. after it has seen its own image of symbol `it` arrive; a rank writes symbol
// m + XG_NB only after polling the flag of symbol m + XG_NB - XG_D, which therefore implies that every rank has
// fetched symbol m + XG_NB - 2 XG_D = m: the buffer is free.
template <int SF, int TH>
__global__ void __launch_bounds__(512, 1)
k1_xchg_kernel(K1Args a, float2 *__restrict__ xs, uint32_t *__restrict__ flags, unsigned long long *__restrict__ packed,
               unsigned long long *dbg, int nosync) {
    using X = XCfg<SF, TH>;
    constexpr int NH = 512 / TH, NW = TH / 32, GT = 32 * X::WPK;
    extern __shared__ __align__(128) unsigned char xg_raw[];
    XSmem<TH> &sm = *reinterpret_cast<XSmem<TH> *>(xg_raw);
    const int rank = (int)(blockIdx.x % X::CL);
    const int half = threadIdx.x / TH, t = threadIdx.x % TH, lane = t & 31, warp = t >> 5;
    const size_t team = (blockIdx.x / X::CL) * NH + half, n_teams = (gridDim.x / X::CL) * NH;
    const size_t n_it = team < a.n_symbols ? (a.n_symbols - team + n_teams - 1) / n_teams : 0;
    float2 *xs_team = xs + team * (size_t)(XG_NB * X::SPS);
    uint32_t *fl = flags + XG_NB * team;
    XHalf<TH> &hb = sm.h[half];
    uint64_t *slot_full = sm.slot_full[half], *rx_full = &sm.rx_full[half], *slot_free = &sm.slot_free[half];

    if (t == 0) {
        mbar_init(&slot_full[0], 1);
        mbar_init(&slot_full[1], 1);
        mbar_init(rx_full, 1);
        mbar_init(slot_free, NW);
        sm.cnt_done[half] = 0;
        sm.cnt_loaded[half] = 0;
        sm.cnt_stored[half][0] = 0;
        sm.cnt_stored[half][1] = 0;
        sm.cnt_gathered[half] = 0;
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < X::F2; i += 512) sm.chirp[i] = k1_ld_table(a.chirp + (i / TH) * X::ROWLEN + TH * rank + (i % TH));
    __syncthreads();
    if (n_it == 0) return;

    auto group_sync = [&]() {                            // the WPK warps that share (sub-CTA, kc_l); literal ids
        if (X::WPK == 1) { __syncwarp(); return; }
        switch (half * X::KPC + warp / X::WPK) {
        case 0: asm volatile("bar.sync 1, %0;" ::"n"(GT) : "memory"); break;
        case 1: asm volatile("bar.sync 2, %0;" ::"n"(GT) : "memory"); break;
        case 2: asm volatile("bar.sync 3, %0;" ::"n"(GT) : "memory"); break;
        case 3: asm volatile("bar.sync 4, %0;" ::"n"(GT) : "memory"); break;
        case 4: asm volatile("bar.sync 5, %0;" ::"n"(GT) : "memory"); break;
        case 5: asm volatile("bar.sync 6, %0;" ::"n"(GT) : "memory"); break;
        case 6: asm volatile("bar.sync 7, %0;" ::"n"(GT) : "memory"); break;
        default: asm volatile("bar.sync 8, %0;" ::"n"(GT) : "memory"); break;
        }
    };
    // warp-uniform: true for the warp whose arrival is number `target`; what the other warps did before arriving is
    // visible to it (release: fence + atomic at cta scope; acquire: atomic + fence in the last warp).  gpu: the arrivals
    // publish global-memory writes, so the LAST warp's fence has gpu scope -- causality order composes across scopes
    // (PTX memory model), which keeps the expensive MEMBAR.GPU (21 % of the stall samples when every warp issued
    // two of them) to one per sub-CTA and symbol.
    auto arrive_last = [&](uint32_t *cnt, uint32_t target, bool gpu) -> bool {
        __syncwarp();
        int last = 0;
        if (lane == 0) {
            __threadfence_block();
            last = atomicAdd(cnt, 1u) == target;
            if (last) { if (gpu) __threadfence(); else __threadfence_block(); }
        }
        return __shfl_sync(0xffffffffu, last, 0) != 0;
    };
    auto load = [&](size_t it) {                         // whole warp: lanes 0..15 fetch one row piece each
        const float2 *src = a.x + (team + it * n_teams) * (size_t)X::SPS + TH * rank;
        if (lane == 0) mbar_expect_tx(&slot_full[it & 1], X::BYTES);
        __syncwarp();
        if (lane < 16) bulk_g2s(hb.slot[it & 1] + lane * TH, src + (size_t)lane * X::ROWLEN, X::RUN, &slot_full[it & 1]);
    };
    auto fetch = [&](size_t it) {                        // lane 0: wait until all CL pieces of symbol `it` are written
        if (lane == 0) {
            const uint32_t need = (uint32_t)(X::CL * (it / XG_NB + 1));
            uint32_t spins = 0, got;
            while (!nosync && (got = xg_flag_ld(fl + (it % XG_NB))) < need)    // nosync: experiment only (wrong results)
                if (++spins == (1u << 18)) xg_dbg(dbg, 4, half, warp, (unsigned)it, got);
            xg_fence_proxy_all();                         // peers' generic writes -> my async-proxy read
            mbar_expect_tx(rx_full, X::BYTES);
            bulk_g2s(hb.rx, xs_team + (it % XG_NB) * (size_t)X::SPS + (size_t)rank * X::F2, X::BYTES, rx_full);
        }
    };
    XConsts<SF, TH> c;
    xg_consts<SF, TH>(t, rank, a.tw, c);
    float2 v[16];
    uint32_t n_a = 0;                                    // A phases done so far by this warp

    // A(m): pass 0 of symbol m.  Its outputs are published (publish(m)) one phase later, when the stores have long
    // been acknowledged and the gpu-scope fence costs nothing (fencing right after the stores: 18 % membar stalls).
    // early: prologue, nothing uses the slot as scratch -> the last warp to read it refills it at once.
    auto phase_a = [&](size_t m, bool early) {
        xg_wait(&slot_full[m & 1], (uint32_t)(m >> 1) & 1u, dbg, 1, half, warp, (unsigned)m);
        xg_pass0_load<SF, TH>(t, hb.slot[m & 1], sm.chirp, v);
        if (early) {
            n_a++;
            if (arrive_last(&sm.cnt_loaded[half], NW * n_a - 1, false) && m + 2 < n_it) load(m + 2);
        } else {
            __syncwarp();
            if (lane == 0) xg_mbar_arrive(slot_free);
        }
        xg_pass0_fft<SF, TH>(c, v);
        xg_put<SF, TH>(t, rank, xs_team + (m % XG_NB) * (size_t)X::SPS, v);
    };
    auto publish = [&](size_t m) {
        if (arrive_last(&sm.cnt_stored[half][m & 1], (uint32_t)(NW * (m / 2 + 1) - 1), true) && lane == 0) xg_flag_add(fl + (m % XG_NB));
    };

    static_assert(XG_D == 2, "the slot schedule assumes pass 0 runs two symbols ahead");
    if (warp == 0) {
        load(0);
        if (n_it > 1) load(1);
    }
    for (size_t m = 0; m < XG_D && m < n_it; m++) {
        phase_a(m, true);
        publish(m);
        if (m == 0 && warp == 0) fetch(0);               // (polls until every rank of the team has written symbol 0)
    }

    for (size_t it = 0; it < n_it; it++) {
        const uint32_t ph = (uint32_t)it & 1u;
        const bool has_a = it + XG_D < n_it;
        float2 *sc = hb.slot[it & 1];                     // the slot A reads in this step, then the scratch of B
        if (has_a) phase_a(it + XG_D, false);
        xg_wait(rx_full, ph, dbg, 3, half, warp, (unsigned)it);      // B: symbol it
        xg_pass1<SF, TH>(t, hb.rx, c, v);
        if (arrive_last(&sm.cnt_gathered[half], (uint32_t)(NW * (it + 1) - 1), false) && it + 1 < n_it) fetch(it + 1);
        if (has_a) xg_wait(slot_free, ph, dbg, 5, half, warp, (unsigned)it);     // every warp has read the slot
        xg_store_t<SF, TH>(t, sc, v);
        group_sync();
        xg_pass2<SF, TH>(t, sc, v);
        group_sync();
        xg_store2<SF, TH>(t, sc, v);
        group_sync();
        unsigned long long best = xg_combine<SF, TH>(t, rank, sc, c);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, off);
            best = o > best ? o : best;
        }
        if (has_a) publish(it + XG_D);                    // (after the image of symbol `it` has landed here: see above)
        fence_proxy_async();                              // generic writes to the scratch before the TMA refill of that slot
        if (arrive_last(&sm.cnt_done[half], (uint32_t)(NW * (it + 1) - 1), false) && it + 2 * XG_D < n_it) load(it + 2 * XG_D);
        if (lane == 0) atomicMax(packed + team + it * n_teams, best);
    }
}
#endif

// CPU emulation: the CL sub-CTAs of a team run one after another, phase by phase
template <int SF, int TH>
inline void xg_emulate(const K1Args &a, uint32_t *bins, float *mags) {
    using X = XCfg<SF, TH>;
    float2 *slot[X::CL], *rx[X::CL], *chirp[X::CL];
    XConsts<SF, TH> *c[X::CL];
    auto v = new float2[X::CL][TH][16];
    float2 *img = new float2[X::SPS];
    for (int q = 0; q < X::CL; q++) {
        slot[q] = new float2[X::F2];
        rx[q] = new float2[X::F2];
        chirp[q] = new float2[X::F2];
        c[q] = new XConsts<SF, TH>[TH];
        for (int i = 0; i < X::F2; i++) chirp[q][i] = a.chirp[(i / TH) * X::ROWLEN + TH * q + (i % TH)];
        for (int t = 0; t < TH; t++) xg_consts<SF, TH>(t, q, a.tw, c[q][t]);
    }
    for (size_t sym = 0; sym < a.n_symbols; sym++) {
        const float2 *x = a.x + sym * (size_t)X::SPS;
        for (int q = 0; q < X::CL; q++) {
            for (int i = 0; i < X::F2; i++) slot[q][i] = x[(i / TH) * X::ROWLEN + TH * q + (i % TH)];
            for (int i = 0; i < X::F2; i++) rx[q][i] = make_float2(NAN, NAN);
            for (int t = 0; t < TH; t++) {
                xg_pass0_load<SF, TH>(t, slot[q], chirp[q], v[q][t]);
                xg_pass0_fft<SF, TH>(c[q][t], v[q][t]);
            }
        }
        for (int q = 0; q < X::CL; q++)                  // every piece writes its columns into the CL exchange images
            for (int t = 0; t < TH; t++) xg_put<SF, TH>(t, q, img, v[q][t]);
        for (int q = 0; q < X::CL; q++)                  // the TMA fetch of every rank's image
            for (int i = 0; i < X::F2; i++) rx[q][i] = img[(size_t)q * X::F2 + i];
        unsigned long long best = 0ull;
        for (int q = 0; q < X::CL; q++) {
            for (int t = 0; t < TH; t++) xg_pass1<SF, TH>(t, rx[q], c[q][t], v[q][t]);
            for (int i = 0; i < X::F2; i++) rx[q][i] = make_float2(NAN, NAN);
            for (int t = 0; t < TH; t++) xg_store_t<SF, TH>(t, rx[q], v[q][t]);
            for (int t = 0; t < TH; t++) xg_pass2<SF, TH>(t, rx[q], v[q][t]);
            for (int i = 0; i < X::F2; i++) rx[q][i] = make_float2(NAN, NAN);
            for (int t = 0; t < TH; t++) xg_store2<SF, TH>(t, rx[q], v[q][t]);
            for (int t = 0; t < TH; t++) {
                const unsigned long long k = xg_combine<SF, TH>(t, q, rx[q], c[q][t]);
                best = k > best ? k : best;
            }
        }
        bins[sym] = key_idx(best);
        if (mags) mags[sym] = sqrtf(key_mag2(best));
    }
    for (int q = 0; q < X::CL; q++) { delete[] slot[q]; delete[] rx[q]; delete[] chirp[q]; delete[] c[q]; }
    delete[] v;
    delete[] img;
}

}  // namespace lb
