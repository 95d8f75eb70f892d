// k1_packed.cu -- the SF7 warp kernel and the SF9 group kernel compiled with the packed complex product
// (LB_PACKED_CMUL: cmul / cfma as FMUL2 + FFMA2 with the swap and the half negation folded into operand modifiers,
// lora_common.cuh).  Measured on B200 against the scalar-product build of the same sources (profiles/r2_packed_cmul_ab.jsonl):
// SF7 0.903 -> 0.921 of the HBM roofline, SF9 0.615 -> 0.643, but SF8 0.796 -> 0.729 (0.881 -> 0.866 after its chirp table
// moved to tensor memory) and SF10 0.571 -> 0.566 (0.608 -> 0.608 after the same move) -- so the choice is
// per kernel, which is why these two live in their own translation unit (the inline product is a per-TU definition).
#define LB_PACKED_CMUL 1
#include "k1_group.cuh"
#include "k1_packed.h"

#include <cstdio>

namespace lb {
namespace {

#define PCU(call)                                                                     \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            snprintf(err, err_cap, "%s: %s", #call, cudaGetErrorString(e_));          \
            return (int)e_;                                                           \
        }                                                                             \
    } while (0)

}  // namespace

int k1_packed_launch(int sf, int device, int n_sms, const float2 *iq, const float2 *chirp, const float2 *tw, size_t n_symbols,
                     uint32_t *bins, float *mags, cudaStream_t st, char *err, size_t err_cap) {
    K1Args a{iq, chirp, tw, n_symbols};
    if (n_symbols == 0) return 0;
    if (sf == 7) {
        constexpr int NW = 12, NS = 2;
        static bool attr_set[64] = {};
        const size_t smem = sizeof(W7Smem<NW, NS>);
        if (!attr_set[device & 63]) {
            PCU(cudaFuncSetAttribute(k1_sf7_warp_kernel<NW, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set[device & 63] = true;
        }
        const int grid = (int)(((n_symbols + NW - 1) / NW) < (size_t)n_sms ? ((n_symbols + NW - 1) / NW) : (size_t)n_sms);
        k1_sf7_warp_kernel<NW, NS><<<grid, NW * 32, smem, st>>>(a, bins, mags);
        PCU(cudaGetLastError());
        return 0;
    }
    if (sf == 9) {
        constexpr int NG = 3, NS = 2;
        static bool attr_set[64] = {};
        const size_t smem = sizeof(GSmem<9, NG, NS>);
        if (!attr_set[device & 63]) {
            PCU(cudaFuncSetAttribute(k1_group_kernel<9, NG, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set[device & 63] = true;
        }
        const int grid = (int)(((n_symbols + NG - 1) / NG) < (size_t)n_sms ? ((n_symbols + NG - 1) / NG) : (size_t)n_sms);
        k1_group_kernel<9, NG, NS><<<grid, NG * GCfg<9>::T, smem, st>>>(a, bins, mags);
        PCU(cudaGetLastError());
        return 0;
    }
    snprintf(err, err_cap, "k1_packed: SF7 / SF9 only");
    return (int)cudaErrorInvalidValue;
}

}  // namespace lb
