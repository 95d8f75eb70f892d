// Microbenchmark: issue/throughput of scalar FFMA/FADD vs packed FFMA2/FADD2 (sm_100a f32x2 PTX).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_tput f32x2_tput.cu && ./f32x2_tput
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { float2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*(u64*)&r) : "l"(*(u64*)&a), "l"(*(u64*)&b), "l"(*(u64*)&c)); return r; }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { float2 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(*(u64*)&r) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return r; }
constexpr int CH = 8, ITERS = 4096;
template <int MODE> __global__ void k(float2 *p, long long *cyc) {
    float2 a[CH], b = p[threadIdx.x], c = p[threadIdx.x + 32];
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = p[threadIdx.x + 64 + i];
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }       // 2 scalar FFMA
            if (MODE == 1) { a[i] = fma2(a[i], b, c); }                                                // 1 FFMA2
            if (MODE == 2) { a[i].x = a[i].x + b.x; a[i].y = a[i].y + b.y; }                           // 2 scalar FADD
            if (MODE == 3) { a[i] = add2(a[i], b); }                                                   // 1 FADD2
            if (MODE == 4) { a[i] = fma2(a[i], b, c); a[i].x = a[i].x + b.x; }                         // FFMA2 + FADD mix
            if (MODE == 5) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = a[i].y + b.y; }                 // FFMA + FADD
        }
    }
    long long t1 = clock64();
    float2 s = a[0];
#pragma unroll
    for (int i = 1; i < CH; i++) { s.x += a[i].x; s.y += a[i].y; }
    p[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float2 *p; long long *c; cudaMalloc(&p, 1 << 24); cudaMemset(p, 0, 1 << 24); cudaMallocManaged(&c, 8);
    const char *names[] = {"2xFFMA (scalar)", "1xFFMA2", "2xFADD (scalar)", "1xFADD2", "FFMA2+FADD", "FFMA+FADD"};
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int m = 0; m < 6; m++) {
            for (int rep = 0; rep < 2; rep++) {
                switch (m) { case 0: k<0><<<148, warps * 32>>>(p, c); break; case 1: k<1><<<148, warps * 32>>>(p, c); break;
                    case 2: k<2><<<148, warps * 32>>>(p, c); break; case 3: k<3><<<148, warps * 32>>>(p, c); break;
                    case 4: k<4><<<148, warps * 32>>>(p, c); break; case 5: k<5><<<148, warps * 32>>>(p, c); break; }
                cudaDeviceSynchronize();
            }
            double per = (double)*c / (ITERS * CH);      // cycles per unrolled body per warp
            printf("warps/SM=%2d %-18s cycles per 2-flop-pair-op per warp = %.3f  -> SM-level: %.3f cycles per warp-op (x%d warps)\n",
                   warps, names[m], per, per / warps, warps);
        }
    }
    return 0;
}
