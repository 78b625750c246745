// box_decode.cuh — where a candidate's NMS box comes from (shared by filter.cu and detect_tail.cu).
#pragma once
#include "decode_row.cuh"

namespace g3d {

// Where a candidate's NMS box comes from: a decoded-box tensor (boxes != null), or decoded on the fly from the
// regression output and the anchors (reg != null) - bit-identical to the corresponding BBoxTransform row - so that the
// detection tail never has to materialise the [B, A, 20] decoded tensor (80 bytes per anchor) for the ~1 % of rows
// that survive the score filter.
struct BoxDecode {
    const float4* anchors;   // [A] (or [B][A] if per_image_anchors)
    const float* reg;        // [B][A][12] (variant 3D) or [B][A][4] (variant 2D); null = not used
    int variant, per_image_anchors, clip;
    float cw, ch;
    float4 mean, stdv;
};

__device__ __forceinline__ float4 decoded_nms_box(const BoxDecode& d, int64_t o, int64_t N, int64_t e) {
    const float4 an = __ldg(d.anchors + (d.per_image_anchors ? o * N + e : e));
    if (d.variant == G3D_VARIANT_3D) {
        const float4 r8 = __ldg(reinterpret_cast<const float4*>(d.reg + (o * N + e) * 12) + 2);
        return decode3d_box(r8, anchor_geom(an));
    }
    const float4 dl = __ldg(reinterpret_cast<const float4*>(d.reg + (o * N + e) * 4));
    return decode2d_row(an, dl, d.mean, d.stdv, d.clip, d.cw, d.ch);
}

}  // namespace g3d

static inline int make_box_decode(g3d::BoxDecode& d, const float* anchors, int64_t Ba, int64_t B, const float* reg, int variant,
                           const float* mean_host, const float* std_host, int clip, float clip_w, float clip_h) {
    G3D_REQUIRE(variant == G3D_VARIANT_2D || variant == G3D_VARIANT_3D, "unknown variant");
    G3D_REQUIRE(anchors && reg, "null pointer");
    G3D_REQUIRE(Ba == 1 || Ba == B, "anchors batch must be 1 or B");
    G3D_REQUIRE(((uintptr_t)anchors % 16) == 0 && ((uintptr_t)reg % 16) == 0, "anchors / regression must be 16-byte aligned");
    G3D_REQUIRE(variant == G3D_VARIANT_3D || (mean_host && std_host), "2D decode needs mean / std");
    d.anchors = (const float4*)anchors; d.reg = reg; d.variant = variant;
    d.per_image_anchors = (Ba == B && B > 1) ? 1 : 0;
    d.clip = clip; d.cw = clip_w; d.ch = clip_h;
    d.mean = mean_host ? make_float4(mean_host[0], mean_host[1], mean_host[2], mean_host[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    d.stdv = std_host ? make_float4(std_host[0], std_host[1], std_host[2], std_host[3]) : make_float4(1.f, 1.f, 1.f, 1.f);
    return G3D_OK;
}

