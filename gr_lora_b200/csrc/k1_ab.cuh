// k1_ab.cuh -- K1 for SF10 / SF11 / SF12 as a producer / consumer pair inside ONE persistent kernel, the exchange in L2.
//
// What the other large-SF kernels taught (profiles/r1_k1_xchg_sf12.md): the arithmetic is cheap enough, what costs is
// every point where many warps have to meet; the fraction of the HBM peak falls 0.90 / 0.81 / 0.64 / 0.58 / 0.30 / 0.25
// as 1 / 2 / 4 / 8 / 16 / 64 warps share a symbol.  Here no warp ever waits for another warp of its SM:
//
//   role A (n_a CTAs): one WARP per (symbol, block of 32 columns).  With n = c*1024 + n' (row c < R = N/128, column
//     n' < 1024) a lane owns one column: R coalesced loads (256 B per warp and row), dechirp, radix-R DIF over the rows
//     in registers, twiddle W_sps^{kc n'}, and R coalesced stores of output row kc into a ring of exchange buffers in
//     global memory that stays L2 resident.  No shared memory, no barrier.
//   role B (the other CTAs): one WARP per (symbol, kc).  Row kc of the exchange buffer IS an SF7-shaped problem: 1024
//     samples z[n'] = 8 branches x 128 points, and F[kc + R q] of the symbol = the SF7 pipeline of k1_warp.cuh applied
//     to z (chirp = 1, same polyphase sum: W_sps^{k' r} = W_sps^{kc r} W_1024^{q' r}, the first factor already inside
//     the twiddle of role A).  TMA-fed 2-slot ring per warp, register FFTs, one swizzled exchange: the kernel that runs
//     SF7 at 0.90.  The N/2 quirk (lib/decoder_impl.cc:450) belongs to kc = 0 only; bins merge by 64-bit atomicMax.
//
// Flags (global memory, monotonic): ready[slot] counts the 32 column blocks of the symbol in ring slot `slot` that role
// A has stored (B waits for 32 * (s / K + 1) before it lets the TMA read row kc); done[slot] counts the rows role B has
// pulled into shared memory (A waits for R * (s / K) before it overwrites the slot).  Every CTA of the grid is resident
// (one per SM, sized by the shared memory of role B) and the ring is longer than the symbols a producer can hold
// unpublished (launch_k1_ab), so the waits cannot deadlock: the oldest unfinished symbol never depends on a younger one.
// tests/test_exchange_protocols.py runs the protocol as a model under random schedules (liveness at the bound, no torn
// or overwritten buffer for any ring).
#pragma once
#include "k1_xchg.cuh"

namespace lb {

template <int SF>
struct ACfg {
    static_assert(SF >= 10 && SF <= 12, "k1_ab: SF10..SF12");
    static constexpr int N = 1 << SF, SPS = 8 * N;
    static constexpr int R = N / 128;                    // rows = radix of role A = sub-problems per symbol: 8, 16, 32
    static constexpr int COLS = 1024;                    // samples per row = samples of one sub-problem
    static constexpr int NBLK = COLS / 32;               // column blocks (warp items of role A) per symbol
    static constexpr int RH = R / 4;                     // kc = 4 h + l
};

// lane-invariant twiddles of role A: W_sps^{l n'} (l < 4) and W_sps^{4 h n'} (h < R/4), n' = the lane's column
template <int SF>
struct AConsts {
    float2 tl[4], th[ACfg<SF>::RH];
};

template <int SF>
LB_HD void ab_consts(int col, const float2 *tw, AConsts<SF> &c) {
    using A = ACfg<SF>;
    for (int l = 0; l < 4; l++) c.tl[l] = k1_ld_table(tw + ((l * col) & (A::SPS - 1)));
    for (int h = 0; h < A::RH; h++) c.th[h] = k1_ld_table(tw + ((4 * h * col) & (A::SPS - 1)));
}

// role A, one column: x = symbol base + col, chirp = table + col (row stride 1024); z = exchange image of the symbol
// ([kc][1024]) + col
template <int SF>
LB_HD void ab_column_load(const float2 *x, const float2 *chirp, float2 *v) {
    using A = ACfg<SF>;
#pragma unroll
    for (int r = 0; r < A::R; r++) v[r] = cmul(x[r * A::COLS], k1_ld_table(chirp + r * A::COLS));
}
template <int SF>
LB_HD void ab_column_fft(const AConsts<SF> &c, float2 *v) {
    using A = ACfg<SF>;
    dft_dif<A::R>(v);
#pragma unroll
    for (int kc = 1; kc < A::R; kc++) {
        const int br = bitrev<A::R>(kc);
        float2 w;
        if ((kc & 3) == 0) w = c.th[kc >> 2];
        else if ((kc >> 2) == 0) w = c.tl[kc & 3];
        else w = cmul(c.th[kc >> 2], c.tl[kc & 3]);
        v[br] = cmul(v[br], w);
    }
}
template <int SF>
LB_HD void ab_column_store(float2 *z, const float2 *v) {
    using A = ACfg<SF>;
#pragma unroll
    for (int kc = 0; kc < A::R; kc++) z[kc * A::COLS] = v[bitrev<A::R>(kc)];
}

// the lane-invariant twiddles of the SF7 pipeline (w7_consts) from the symbol's table: W_1024^j = W_sps^{j R}
LB_HD void ab_w7_consts(int lane, const float2 *tw, int stride, W7Consts &c) {
    const int kc = lane >> 1, h = lane & 1;
    for (int a = 0; a < 8; a++) c.tw1[a] = k1_ld_table(tw + ((a * kc * 8) & 1023) * stride);
    for (int ka = 0; ka < 8; ka++) c.wq[ka] = k1_ld_table(tw + (w7_signed_bin(kc + 16 * ka) & 1023) * stride);
    for (int j = 0; j < 4; j++) c.w4[j] = k1_ld_table(tw + ((4 * w7_signed_bin(kc + 16 * (4 * h + j))) & 1023) * stride);
}

// role B, last step of the SF7 pipeline (w7_final of k1_warp.cuh) for sub-problem kc of a symbol with R rows:
// bin of the symbol = kc + R q, quirk term only where the symbol has it (kc = 0, q = 64)
LB_HD unsigned long long ab_final(int lane, const W7Consts &c, const float2 *own, const float2 *other, float2 own_q, float2 other_q,
                                  int kc_sym, int rows) {
    const int kc = lane >> 1, h = lane & 1;
    unsigned long long best = 0ull;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int ka = 4 * h + j;
        const float2 p0 = h ? other[j] : own[j];
        const float2 p1 = h ? own[j] : other[j];
        float2 f = cfma(p1, c.w4[j], p0);
        const int q = kc + 16 * ka;
        if (q == 64 && kc_sym == 0) {
            const float2 q0 = h ? other_q : own_q, q1 = h ? own_q : other_q;
            f = cadd(f, cfma(q1, cconj(c.w4[j]), q0));
        }
        const unsigned long long key = pack_key(cnorm2(f), (uint32_t)(kc_sym + rows * q));
        best = key > best ? key : best;
    }
    return best;
}

#ifdef __CUDACC__
constexpr int AB_WARPS = 12, AB_NSLOT = 2;
constexpr int AB_FSTRIDE = 32;                           // uint32 per flag: one 128-byte line each (hot-spot relief)

struct ABSmem {                                          // role B only
    float4 ones[W7_SLOT_F4];                             // "chirp" of the sub-problems: (1, 0)
    float4 slots[AB_WARPS][AB_NSLOT][W7_SLOT_F4];
    uint64_t bars[AB_WARPS][AB_NSLOT];
};

// Polls are RELAXED loads (an acquire load invalidates the L1 that holds role A's chirp and twiddle tables, and costs a
// gpu-scope fence per 8 KiB item in role B).  What follows the poll is either a TMA read or plain stores, both go to L2
// -- the point of coherence where the producer's data was performed before its releasing flag update -- and neither is
// issued before the loop exits.
LB_D uint32_t ab_flag_ld(const uint32_t *flag) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    return v;
}
LB_D void ab_flag_add_relaxed(uint32_t *flag) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(flag) : "memory"); }
// The first captures counted ~100 polls per consumer item, 1.7e10 polls/s from ~1800 warps on counters that sat in six
// 128-byte lines: every flag now has a line of its own (AB_FSTRIDE) and a waiting warp backs off.
LB_D void ab_spin(const uint32_t *flag, uint32_t need, unsigned long long *dbg, unsigned site, unsigned warp, unsigned sym) {
    uint32_t spins = 0, got, ns = 32;
    while ((got = ab_flag_ld(flag)) < need) {
        __nanosleep(ns);
        if (ns < 1024) ns *= 2;
        if (++spins == (1u << 14)) xg_dbg(dbg, site, 0, warp, sym, got);
    }
}

LB_D void ab_bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
LB_D void ab_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
LB_D void ab_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
LB_D void ab_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Role A, version 2 (prod == 2): a GROUP of 4 warps owns 128 columns.  An item = 32 row pieces of 1 KiB (U symbols x R
// rows) in one 32 KiB slot of the group's 2-slot TMA ring; thread t of the group owns column t: it reads its 32 values,
// dechirps, runs the radix-R row pass and writes the outputs back IN PLACE (it overwrites only what it read), then one
// thread hands the 32 rows to the copy engine (32 bulk stores of 1 KiB into the exchange ring) and refills the slot as
// soon as those stores have read it.  The ready flags of an item are raised one item later, after
// cp.async.bulk.wait_group has confirmed its stores: no gpu-scope fence and no scattered LSU stores on the critical
// path, one named barrier per item.  (Version 1's warp spent 20 % of its time issuing 32 x 256 B copies lane by lane,
// 27 % on polls and 256 B stores, 13 % in fences, 8 % on arithmetic: profiles/r1_k1_ab_sf12.md.)
template <int SF>
LB_D void ab_producer_groups(const K1Args &a, float2 *scratch, uint32_t *ready, const uint32_t *done, uint32_t ring, uint32_t n_b,
                             unsigned char *smem_raw, unsigned long long *dbg) {
    using A = ACfg<SF>;
    constexpr int U = 32 / A::R, GT = 128, TILES = A::COLS / GT;           // 8 column tiles per symbol
    constexpr uint32_t SLOT_BYTES = 32u * GT * 8u;                          // 32 KiB
    const int grp = threadIdx.x / GT, t = threadIdx.x % GT, lane = threadIdx.x & 31;
    float2 *slots = reinterpret_cast<float2 *>(smem_raw + sizeof(float4) * W7_SLOT_F4) + (size_t)grp * 2 * (SLOT_BYTES / 8);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + sizeof(float4) * W7_SLOT_F4 + (size_t)AB_WARPS * AB_NSLOT * W7_SLOT_F4 * sizeof(float4))
                     + grp * 8;                           // the mbarrier words of the group's first warp
    const size_t n_ag = (size_t)(gridDim.x - n_b) * 3;    // groups; a multiple of TILES (host side)
    const size_t g = (size_t)(blockIdx.x - n_b) * 3 + grp;
    const int colbase = (int)(g % TILES) * GT, col = colbase + t;
    const size_t step = n_ag / TILES, first = g / TILES;  // this group's symbols: first, first + step, ...
    const size_t n_mine = first < a.n_symbols ? (a.n_symbols - first + step - 1) / step : 0;
    const size_t n_items = (n_mine + U - 1) / U;
    auto group_sync = [&]() {
        if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else if (grp == 1) asm volatile("bar.sync 2, 128;" ::: "memory");
        else asm volatile("bar.sync 3, 128;" ::: "memory");
    };
    if (t == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    group_sync();
    auto sym_of = [&](size_t item, int u) { return first + (item * U + u) * step; };
    auto issue_loads = [&](size_t item, int si) {        // one thread: 32 row pieces of 1 KiB
        uint32_t bytes = 0;
#pragma unroll
        for (int u = 0; u < U; u++) if (sym_of(item, u) < a.n_symbols) bytes += A::R * 1024u;
        mbar_expect_tx(&bars[si], bytes);
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t s = sym_of(item, u);
            if (s >= a.n_symbols) break;
#pragma unroll
            for (int r = 0; r < A::R; r++)
                bulk_g2s(slots + (size_t)si * (SLOT_BYTES / 8) + (u * A::R + r) * GT, a.x + s * (size_t)A::SPS + r * A::COLS + colbase, 1024u, &bars[si]);
        }
    };
    AConsts<SF> c;
    ab_consts<SF>(col, a.tw, c);
    if (t == 0)
        for (int si = 0; si < 2; si++) if ((size_t)si < n_items) issue_loads(si, si);
    bool stores_pending = false;                          // thread 0: the previous item's stores are not yet published
    size_t pend_item = 0;
    for (size_t item = 0; item < n_items; item++) {
        const int si = (int)(item & 1);
        float2 *slot = slots + (size_t)si * (SLOT_BYTES / 8);
        float2 v[U][A::R];
#pragma unroll
        for (int r = 0; r < A::R; r++) v[0][r] = k1_ld_table(a.chirp + col + r * A::COLS);      // L2 hits, before the wait
        xg_wait(&bars[si], (uint32_t)(item >> 1) & 1u, dbg, 9, 0, (unsigned)(threadIdx.x >> 5), (unsigned)item);
#pragma unroll
        for (int u = U - 1; u >= 0; u--)
#pragma unroll
            for (int r = 0; r < A::R; r++) v[u][r] = cmul(slot[(u * A::R + r) * GT + t], v[0][r]);
#pragma unroll
        for (int u = 0; u < U; u++) {
            ab_column_fft<SF>(c, v[u]);
#pragma unroll
            for (int kc = 0; kc < A::R; kc++) slot[(u * A::R + kc) * GT + t] = v[u][bitrev<A::R>(kc)];   // in place
        }
        fence_proxy_async();                              // my generic writes -> read by the copy engine
        group_sync();
        if (t == 0) {
            if (stores_pending) {                         // publish the previous item: its stores have had a whole item
                ab_bulk_wait_all();
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const size_t s = sym_of(pend_item, u);
                    if (s < a.n_symbols) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ready + AB_FSTRIDE * (uint32_t)(s % ring)), "r"(GT / 32) : "memory");
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const size_t s = sym_of(item, u);
                if (s >= a.n_symbols) break;
                const uint32_t rs = (uint32_t)(s % ring);
                ab_spin(done + AB_FSTRIDE * rs, (uint32_t)(A::R * (s / ring)), dbg, 6, (unsigned)(threadIdx.x >> 5), (unsigned)s);
#pragma unroll
                for (int kc = 0; kc < A::R; kc++)
                    ab_bulk_s2g(scratch + ((size_t)rs * A::R + kc) * A::COLS + colbase, slot + (u * A::R + kc) * GT, 1024u);
            }
            ab_bulk_commit();
            stores_pending = true;
            pend_item = item;
            if (item + 2 < n_items) {                     // refill the slot once the stores have read it
                ab_bulk_wait_read();
                issue_loads(item + 2, si);
            }
        }
    }
    if (t == 0 && stores_pending) {
        ab_bulk_wait_all();
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t s = sym_of(pend_item, u);
            if (s < a.n_symbols) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ready + AB_FSTRIDE * (uint32_t)(s % ring)), "r"(GT / 32) : "memory");
        }
    }
    (void)lane;
}

// scratch: [ring][R][1024] float2; ready / done: [ring] counters (zeroed before the launch)
template <int SF>
__global__ void __launch_bounds__(AB_WARPS * 32, 1)
k1_ab_kernel(K1Args a, float2 *__restrict__ scratch, uint32_t *__restrict__ ready, uint32_t *__restrict__ done, uint32_t ring,
             uint32_t n_b, unsigned long long *__restrict__ packed, unsigned long long *dbg, int prod) {
    using A = ACfg<SF>;
    extern __shared__ __align__(128) unsigned char ab_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (blockIdx.x >= n_b && prod == 2) {
        ab_producer_groups<SF>(a, scratch, ready, done, ring, n_b, ab_raw, dbg);
        return;
    }
    if (blockIdx.x >= n_b) {
        // ---- role A (version 1): warp items = U symbols x one block of 32 columns; the block of a warp never changes ---------
        // The 32 row pieces of an item (256 B each, one per lane) stream through the warp's own 2-slot TMA ring, so the
        // next item loads while this one is computed; 12 warps x 16 KiB in flight per SM.
        ABSmem &sa = *reinterpret_cast<ABSmem *>(ab_raw);
        constexpr int U = 32 / A::R;                      // symbols per item: 32 rows for every SF
        const size_t n_aw = (size_t)(gridDim.x - n_b) * AB_WARPS;           // a multiple of NBLK (host side)
        const size_t g = (size_t)(blockIdx.x - n_b) * AB_WARPS + warp;
        const int col = (int)(g % A::NBLK) * 32 + lane;
        const size_t step = n_aw / A::NBLK;               // symbols between two consecutive symbols of this warp
        const size_t first = g / A::NBLK;
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < AB_NSLOT; s++) mbar_init(&sa.bars[warp][s], 1);
            fence_mbar_init();
        }
        __syncwarp();
        // lane l fetches row l % R of symbol (item * U + l / R) of this warp's sequence
        auto a_issue = [&](size_t item, int slot_i) {
            const size_t sym = first + (item * U + lane / A::R) * step;
            const bool valid = sym < a.n_symbols;
            const unsigned n_valid = __popc(__ballot_sync(0xffffffffu, valid));
            if (lane == 0) mbar_expect_tx(&sa.bars[warp][slot_i], 256u * n_valid);
            __syncwarp();
            if (valid) bulk_g2s(reinterpret_cast<float2 *>(sa.slots[warp][slot_i]) + lane * 32,
                                a.x + sym * (size_t)A::SPS + (lane % A::R) * A::COLS + (col - lane), 256u, &sa.bars[warp][slot_i]);
        };
        const size_t n_mine = first < a.n_symbols ? (a.n_symbols - first + step - 1) / step : 0;   // symbols of this warp
        const size_t n_items = (n_mine + U - 1) / U;
        AConsts<SF> c;
        ab_consts<SF>(col, a.tw, c);
        for (int s = 0; s < AB_NSLOT; s++)
            if ((size_t)s < n_items) a_issue(s, s);
        constexpr int PUB = 1;                            // publish every PUB items (one gpu-scope fence each); what is stored but
                                                          // unpublished counts against the ring (see launch_k1_ab)
        uint32_t pend[PUB * U];                           // ring slots stored but not yet published
        int n_pend = 0;
        for (size_t item = 0; item < n_items; item++) {
            const int si = (int)(item % AB_NSLOT);
            const float2 *slot = reinterpret_cast<const float2 *>(sa.slots[warp][si]);
            // everything that does not depend on the samples goes out before the wait: the chirp column (it does not
            // fit the L1 next to 204 KiB of shared memory, so these are L2 hits) and the throttle counters
            float2 v[U][A::R];
#pragma unroll
            for (int r = 0; r < A::R; r++) v[0][r] = k1_ld_table(a.chirp + col + r * A::COLS);
            uint32_t seen[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const size_t s = first + (item * U + u) * step;
                seen[u] = (lane == 0 && s < a.n_symbols) ? ab_flag_ld(done + AB_FSTRIDE * (uint32_t)(s % ring)) : 0u;
            }
            xg_wait(&sa.bars[warp][si], (uint32_t)(item / AB_NSLOT) & 1u, dbg, 9, 0, (unsigned)warp, (unsigned)item);
#pragma unroll
            for (int u = U - 1; u >= 0; u--)
#pragma unroll
                for (int r = 0; r < A::R; r++) v[u][r] = cmul(slot[(u * A::R + r) * 32 + lane], v[0][r]);
            __syncwarp();
            if (item + AB_NSLOT < n_items) {              // every lane has read the slot: refill it
                fence_proxy_async();
                a_issue(item + AB_NSLOT, si);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const size_t s = first + (item * U + u) * step;
                if (s >= a.n_symbols) break;
                ab_column_fft<SF>(c, v[u]);
                const uint32_t rs = (uint32_t)(s % ring);
                const uint32_t need = (uint32_t)(A::R * (s / ring));
                if (lane == 0 && seen[u] < need) ab_spin(done + AB_FSTRIDE * rs, need, dbg, 6, (unsigned)warp, (unsigned)s);
                __syncwarp();
                ab_column_store<SF>(scratch + (size_t)rs * A::SPS + col, v[u]);
                pend[n_pend++] = rs;
            }
            if (n_pend == PUB * U) {
                __syncwarp();
                if (lane == 0) {
                    __threadfence();
                    for (int u = 0; u < n_pend; u++) ab_flag_add_relaxed(ready + AB_FSTRIDE * pend[u]);
                }
                n_pend = 0;
            }
        }
        __syncwarp();
        if (lane == 0 && n_pend) {
            __threadfence();
            for (int u = 0; u < n_pend; u++) ab_flag_add_relaxed(ready + AB_FSTRIDE * pend[u]);
        }
        return;
    }

    // ---- role B: warp items u = symbol * R + kc, the SF7 pipeline on row kc of the symbol's exchange image ------------
    ABSmem &sm = *reinterpret_cast<ABSmem *>(ab_raw);
    const size_t n_items = a.n_symbols * (size_t)A::R;
    const size_t gw = (size_t)blockIdx.x * AB_WARPS + warp, w_total = (size_t)n_b * AB_WARPS;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < AB_NSLOT; s++) mbar_init(&sm.bars[warp][s], 1);
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < W7_SLOT_F4; i += AB_WARPS * 32) sm.ones[i] = make_float4(1.f, 0.f, 1.f, 0.f);
    __syncthreads();

    auto issue = [&](size_t u, int s) {                  // lane 0: wait for the 32 column blocks of the symbol, then load row kc
        const size_t sym = u / A::R;
        const uint32_t slot = (uint32_t)(sym % ring);
        ab_spin(ready + AB_FSTRIDE * slot, (uint32_t)(A::NBLK * (sym / ring + 1)), dbg, 7, (unsigned)warp, (unsigned)sym);
        asm volatile("fence.proxy.async.global;" ::: "memory");     // role A's generic stores -> my async-proxy read
        mbar_expect_tx(&sm.bars[warp][s], 8192);
        bulk_g2s(sm.slots[warp][s], scratch + ((size_t)slot * A::R + (u % A::R)) * A::COLS, 8192, &sm.bars[warp][s]);
    };
    W7Consts c;
    ab_w7_consts(lane, a.tw, A::R, c);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < AB_NSLOT; s++) {
            const size_t u = gw + (size_t)s * w_total;
            if (u < n_items) issue(u, s);
        }
    }
    uint32_t it = 0;
    for (size_t u = gw; u < n_items; u += w_total, it++) {
        const int s = it % AB_NSLOT;
        float4 *slot = sm.slots[warp][s];
        xg_wait(&sm.bars[warp][s], (it / AB_NSLOT) & 1u, dbg, 8, 0, (unsigned)warp, (unsigned)(u / A::R));
        if (lane == 0) ab_flag_add_relaxed(done + AB_FSTRIDE * (uint32_t)((u / A::R) % ring));   // the row is in shared memory: role A may reuse it
        float2 v0[16], v1[16];
        w7_pass0(lane, slot, sm.ones, v0, v1);
        __syncwarp();
        w7_store(lane, slot, v0, v1);
        __syncwarp();
        float2 P[8], Pq;
        w7_pass1(lane, slot, c, P, Pq);
        __syncwarp();
        if (lane == 0) {
            const size_t nxt = u + (size_t)AB_NSLOT * w_total;
            if (nxt < n_items) {
                fence_proxy_async();
                issue(nxt, s);
            }
        }
        const int h = lane & 1;
        float2 own[4], other[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float2 send = h ? P[j] : P[4 + j];
            own[j] = h ? P[4 + j] : P[j];
            other[j].x = __shfl_xor_sync(0xffffffffu, send.x, 1);
            other[j].y = __shfl_xor_sync(0xffffffffu, send.y, 1);
        }
        float2 other_q;
        other_q.x = __shfl_xor_sync(0xffffffffu, Pq.x, 1);
        other_q.y = __shfl_xor_sync(0xffffffffu, Pq.y, 1);
        unsigned long long best = ab_final(lane, c, own, other, Pq, other_q, (int)(u % A::R), A::R);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, off);
            best = o > best ? o : best;
        }
        if (lane == 0) atomicMax(packed + u / A::R, best);
    }
}
#endif

// CPU emulation: role A for every column, then the SF7 pipeline (lanes one after another) for every row kc
template <int SF>
inline void ab_emulate(const K1Args &a, uint32_t *bins, float *mags) {
    using A = ACfg<SF>;
    float2 *img = new float2[A::SPS];
    float4 *slot = new float4[W7_SLOT_F4];
    float4 *ones = new float4[W7_SLOT_F4];
    for (int i = 0; i < W7_SLOT_F4; i++) ones[i] = make_float4(1.f, 0.f, 1.f, 0.f);
    W7Consts c[32];
    for (int l = 0; l < 32; l++) ab_w7_consts(l, a.tw, A::R, c[l]);
    for (size_t sym = 0; sym < a.n_symbols; sym++) {
        for (int col = 0; col < A::COLS; col++) {
            AConsts<SF> ac;
            ab_consts<SF>(col, a.tw, ac);
            float2 v[A::R];
            ab_column_load<SF>(a.x + sym * (size_t)A::SPS + col, a.chirp + col, v);
            ab_column_fft<SF>(ac, v);
            ab_column_store<SF>(img + col, v);
        }
        unsigned long long best = 0ull;
        for (int kc = 0; kc < A::R; kc++) {
            const float2 *z = img + (size_t)kc * A::COLS;
            for (int i = 0; i < W7_SLOT_F4; i++) slot[i] = make_float4(z[2 * i].x, z[2 * i].y, z[2 * i + 1].x, z[2 * i + 1].y);
            float2 v0[32][16], v1[32][16];
            for (int l = 0; l < 32; l++) w7_pass0(l, slot, ones, v0[l], v1[l]);
            for (int i = 0; i < W7_SLOT_F4; i++) slot[i] = make_float4(NAN, NAN, NAN, NAN);
            for (int l = 0; l < 32; l++) w7_store(l, slot, v0[l], v1[l]);
            float2 P[32][8], Pq[32];
            for (int l = 0; l < 32; l++) w7_pass1(l, slot, c[l], P[l], Pq[l]);
            for (int l = 0; l < 32; l++) {
                const int h = l & 1;
                float2 own[4], other[4];
                for (int j = 0; j < 4; j++) {
                    own[j] = h ? P[l][4 + j] : P[l][j];
                    other[j] = h ? P[l ^ 1][4 + j] : P[l ^ 1][j];
                }
                const unsigned long long k = ab_final(l, c[l], own, other, Pq[l], Pq[l ^ 1], kc, A::R);
                best = k > best ? k : best;
            }
        }
        bins[sym] = key_idx(best);
        if (mags) mags[sym] = sqrtf(key_mag2(best));
    }
    delete[] img; delete[] slot; delete[] ones;
}

}  // namespace lb
