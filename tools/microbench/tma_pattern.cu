// Microbenchmark: DRAM -> shared memory throughput of cp.async.bulk for the access patterns of the K1 kernels.
// Every CTA (one warp) streams "units" of UNIT bytes through a ring of NSLOT shared-memory slots; a unit is either one
// contiguous copy or PIECES copies of UNIT / PIECES bytes at a fixed stride (the column slice of a symbol that a
// k1_xchg sub-CTA loads: 16 row pieces of 2 KiB, stride = 8 * ROWLEN bytes).  No compute: the number is the ceiling
// the load pattern itself allows.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_pattern tma_pattern.cu && ./tma_pattern
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// unit u of the grid-wide sequence -> (symbol, slice): symbol = u / slices, slice = u % slices;  slices = stride / piece
template <int NSLOT>
__global__ void stream(const char *x, size_t n_units, uint32_t unit, int pieces, size_t stride, int slices, size_t sym_bytes,
                       unsigned long long *sink) {
    extern __shared__ __align__(128) unsigned char raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(raw);
    unsigned char *slots = raw + 128;
    const int lane = threadIdx.x;
    if (lane == 0) {
        for (int s = 0; s < NSLOT; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const uint32_t piece = unit / pieces;
    auto issue = [&](size_t u, int s) {
        const size_t sym = u / slices, sl = u % slices;
        const char *src = x + sym * sym_bytes + sl * (size_t)piece;
        if (lane == 0) mbar_expect(&bars[s], unit);
        __syncwarp();
        for (int p = lane; p < pieces; p += 32) bulk(slots + (size_t)s * unit + (size_t)p * piece, src + (size_t)p * stride, piece, &bars[s]);
    };
    size_t u = blockIdx.x;
    for (int s = 0; s < NSLOT && u + (size_t)s * gridDim.x < n_units; s++) issue(u + (size_t)s * gridDim.x, s);
    unsigned long long acc = 0;
    uint32_t it = 0;
    for (; u < n_units; u += gridDim.x, it++) {
        const int s = it % NSLOT;
        mbar_wait(&bars[s], (it / NSLOT) & 1u);
        acc += *reinterpret_cast<const unsigned long long *>(slots + (size_t)s * unit + lane * 8);
        __syncwarp();
        const size_t nxt = u + (size_t)NSLOT * gridDim.x;
        if (nxt < n_units) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(nxt, s);
        }
    }
    if (acc == 0x1234567ull) sink[0] = acc;
}

int main() {
    const size_t total = 4ull << 30;
    char *x;
    unsigned long long *sink;
    cudaMalloc(&x, total);
    cudaMalloc(&sink, 8);
    cudaMemset(x, 1, total);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    struct Case { const char *name; uint32_t unit; int pieces; size_t stride; int ctas_per_sm; };
    const Case cases[] = {
        {"contiguous 64 KiB (k1_sf10)", 65536, 1, 65536, 1},
        {"contiguous 32 KiB", 32768, 1, 32768, 2},
        {"contiguous 16 KiB", 16384, 1, 16384, 4},
        {"contiguous  8 KiB (k1_sf7: per warp)", 8192, 1, 8192, 8},
        {"16 x 2 KiB, stride  4 KiB (SF10 xchg)", 32768, 16, 4096, 2},
        {"16 x 2 KiB, stride  8 KiB (SF11 xchg)", 32768, 16, 8192, 2},
        {"16 x 2 KiB, stride 16 KiB (SF12 xchg)", 32768, 16, 16384, 2},
        {"16 x 1 KiB, stride 16 KiB (SF12 xchg, 128 thr)", 16384, 16, 16384, 4},
        {" 8 x 4 KiB, stride 32 KiB", 32768, 8, 32768, 2},
        {" 4 x 8 KiB, stride 64 KiB", 32768, 4, 65536, 2},
        {"32 x 2 KiB, stride  8 KiB (k1_big SF11)", 65536, 32, 8192, 1},
        {"32 x 256 B, stride  8 KiB (k1_ab role A), 12 warps/SM", 8192, 32, 8192, 12},
        {"32 x 256 B, stride  8 KiB (k1_ab role A),  8 warps/SM", 8192, 32, 8192, 8},
        {"32 x 512 B, stride  8 KiB, 6 warps/SM", 16384, 32, 8192, 6},
        {"32 x 1 KiB, stride  8 KiB, 3 warps/SM", 32768, 32, 8192, 3},
    };
    constexpr int NSLOT = 3;
    for (const Case &c : cases) {
        const int slices = (int)(c.stride / (c.unit / c.pieces));
        const size_t sym_bytes = c.stride * c.pieces;                 // one "symbol" = pieces rows of `stride` bytes
        const size_t n_units = total / c.unit;
        const size_t smem = 128 + (size_t)NSLOT * c.unit;
        if (smem * c.ctas_per_sm > 227 * 1024) { printf("%-48s skipped (smem)\n", c.name); continue; }
        cudaFuncSetAttribute(stream<NSLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        const int grid = sms * c.ctas_per_sm;
        stream<NSLOT><<<grid, 32, smem>>>(x, n_units, c.unit, c.pieces, c.stride, slices, sym_bytes, sink);
        cudaEventRecord(e0);
        for (int r = 0; r < 3; r++) stream<NSLOT><<<grid, 32, smem>>>(x, n_units, c.unit, c.pieces, c.stride, slices, sym_bytes, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const cudaError_t err = cudaGetLastError();
        printf("%-48s %8.1f GB/s  (%d CTAs/SM, %d slots in flight per CTA)%s\n", c.name, 3.0 * total / (ms * 1e-3) / 1e9, c.ctas_per_sm, NSLOT,
               err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
    return 0;
}
