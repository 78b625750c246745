// nms_common.cuh — device helpers shared by nms.cu and the short-segment detection tail (detect_tail.cu): the sort key,
// the shared-memory bitonic network and the suppression test with torchvision's arithmetic.
#pragma once
#include "common.cuh"

namespace g3d {

__device__ __forceinline__ uint64_t make_key(float score, uint32_t idx) {
    if (score == 0.0f) score = 0.0f;  // -0.0 and +0.0 compare equal in the reference's sort
    uint32_t u = __float_as_uint(score);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone float -> uint
    return ((uint64_t)(~u) << 32) | idx;             // descending score, ascending index on ties
}

__device__ __forceinline__ void bitonic_smem(uint64_t* sk, int P, int k_from, int k_to, int j_from) {
    // runs stages k = k_from..k_to (doubling); within the first stage starts at j = j_from (0 -> k/2)
    for (int k = k_from; k <= k_to; k <<= 1) {
        for (int j = (k == k_from && j_from > 0) ? j_from : (k >> 1); j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const uint64_t a = sk[i], b = sk[l];
                const bool up = (i & k) == 0;
                if ((a > b) == up) { sk[i] = b; sk[l] = a; }
            }
            __syncthreads();
        }
    }
}

// suppression test "IoU(a,b) > thr" with torchvision's arithmetic.  For thr >= 0 a pair whose clamped intersection is
// zero can never pass (IoU is 0, -0 or NaN), so the union and the IEEE division are skipped for disjoint pairs.
__device__ __forceinline__ bool suppresses(const float4& a, float area_a, const float4& b, float area_b, float thr,
                                           bool thr_nonneg) {
    const float w = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
    const float h = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
    if (thr_nonneg && !(w > 0.0f && h > 0.0f)) return false;
    const float inter = __fmul_rn(fmaxf(w, 0.0f), fmaxf(h, 0.0f));
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter)) > thr;
}

}  // namespace g3d
