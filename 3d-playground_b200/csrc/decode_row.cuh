// decode_row.cuh — per-row box decoding shared by the full-tensor decode kernels (decode.cu) and the fused detection
// tail (detect.cu), so that a row decoded on the fly is bit-identical to the same row of BBoxTransform's output.
// Arithmetic is the eager reference's, op by op: left-to-right +/- chains, then a separately rounded mul and add.
#pragma once
#include "common.cuh"

namespace g3d {

struct AnchorGeom {
    float w, h, cx, cy;
};
// anchor widths / heights / centres as both BBoxTransform copies form them (3D utils.py:104-107, 2D utils.py:104-107)
__device__ __forceinline__ AnchorGeom anchor_geom(const float4& an) {
    AnchorGeom g;
    g.w = __fsub_rn(an.z, an.x);
    g.h = __fsub_rn(an.w, an.y);
    g.cx = __fadd_rn(an.x, __fmul_rn(0.5f, g.w));
    g.cy = __fadd_rn(an.y, __fmul_rn(0.5f, g.h));
    return g;
}

// a7: 3D BBoxTransform row, pytorch_retinanet_detector_directional/retinanet/utils.py:114-135: r[12] -> p[20]
__device__ __forceinline__ void decode3d_row(const float* r, const AnchorGeom& g, float* p) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        // corner k = c -/+ L -/+ W +/- H  (utils.py:114-130)
        const bool lp = k & 2, wp = k & 1, hp = !(k & 4);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float v = lp ? __fadd_rn(r[c], r[2 + c]) : __fsub_rn(r[c], r[2 + c]);
            v = wp ? __fadd_rn(v, r[4 + c]) : __fsub_rn(v, r[4 + c]);
            v = hp ? __fadd_rn(v, r[6 + c]) : __fsub_rn(v, r[6 + c]);
            p[2 * k + c] = v;
        }
    }
    p[16] = r[8]; p[17] = r[9]; p[18] = r[10]; p[19] = r[11];
#pragma unroll
    for (int i = 0; i < 20; i += 2) {  // utils.py:134-135
        p[i] = __fadd_rn(__fmul_rn(p[i], g.w), g.cx);
        p[i + 1] = __fadd_rn(__fmul_rn(p[i + 1], g.h), g.cy);
    }
}

// columns 16..19 of the same row (the 2D box NMS runs on, 3D model.py:383): regression columns 8..11
__device__ __forceinline__ float4 decode3d_box(const float4& r8, const AnchorGeom& g) {
    return make_float4(__fadd_rn(__fmul_rn(r8.x, g.w), g.cx), __fadd_rn(__fmul_rn(r8.y, g.h), g.cy),
                       __fadd_rn(__fmul_rn(r8.z, g.w), g.cx), __fadd_rn(__fmul_rn(r8.w, g.h), g.cy));
}

// a8 (+a9): 2D BBoxTransform row, retinanet/utils.py:102-126 (+ ClipBoxes :134-144)
__device__ __forceinline__ float4 decode2d_row(const float4& an, const float4& d, const float4& mean, const float4& stdv,
                                               int clip, float cw, float ch) {
    const AnchorGeom g = anchor_geom(an);
    const float dx = __fadd_rn(__fmul_rn(d.x, stdv.x), mean.x), dy = __fadd_rn(__fmul_rn(d.y, stdv.y), mean.y);
    const float dw = __fadd_rn(__fmul_rn(d.z, stdv.z), mean.z), dh = __fadd_rn(__fmul_rn(d.w, stdv.w), mean.w);
    const float pcx = __fadd_rn(g.cx, __fmul_rn(dx, g.w)), pcy = __fadd_rn(g.cy, __fmul_rn(dy, g.h));
    const float pw = __fmul_rn(expf(dw), g.w), ph = __fmul_rn(expf(dh), g.h);
    float4 o;
    o.x = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
    o.y = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
    o.z = __fadd_rn(pcx, __fmul_rn(0.5f, pw));
    o.w = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
    if (clip) {
        o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f);
        o.z = fminf(o.z, cw);   o.w = fminf(o.w, ch);
    }
    return o;
}

}  // namespace g3d
