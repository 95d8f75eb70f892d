// tmem.cuh -- tensor memory (tcgen05.alloc / st / ld) as a per-thread constant store.
//
// Nothing on this path is a matrix product, so the 256 KiB of tensor memory per SM are otherwise idle.  The K1 kernels keep
// thread-invariant tables there (the dechirp samples a thread multiplies with, inter-pass twiddles): written once per CTA,
// read once per symbol with 16-column loads, they cost neither registers nor shared-memory bandwidth.
// Addressing: (lane << 16) | column; warp w of a CTA may only touch lanes 32 (w & 3) .. 32 (w & 3) + 31, so warps w and
// w + 4 share a lane quadrant and use different columns.
#pragma once
#include "k1_warp.cuh"

namespace lb {
#ifdef __CUDACC__
template <int COLS>
LB_D void tm_alloc(uint32_t *smem_dst) {      // one warp; COLS a power of two >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
LB_D void tm_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
LB_D void tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
LB_D void tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
LB_D void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
LB_D void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 16 consecutive columns of the thread's lane <-> 8 complex values
LB_D void tm_st16(uint32_t taddr, const float2 *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "f"(v[0].x), "f"(v[0].y), "f"(v[1].x), "f"(v[1].y), "f"(v[2].x), "f"(v[2].y), "f"(v[3].x), "f"(v[3].y),
                 "f"(v[4].x), "f"(v[4].y), "f"(v[5].x), "f"(v[5].y), "f"(v[6].x), "f"(v[6].y), "f"(v[7].x), "f"(v[7].y) : "memory");
}
LB_D void tm_ld16(uint32_t taddr, float2 *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y), "=f"(v[2].x), "=f"(v[2].y), "=f"(v[3].x), "=f"(v[3].y),
                   "=f"(v[4].x), "=f"(v[4].y), "=f"(v[5].x), "=f"(v[5].y), "=f"(v[6].x), "=f"(v[6].y), "=f"(v[7].x), "=f"(v[7].y)
                 : "r"(taddr) : "memory");
}
#endif  // __CUDACC__
}  // namespace lb
