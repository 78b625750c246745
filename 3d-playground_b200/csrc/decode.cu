// decode.cu — a7 3D BBoxTransform (12 -> 20), a8 2D BBoxTransform (+ fused a9 clip), a9 ClipBoxes.
#include "decode_row.cuh"

namespace g3d {

constexpr int kDecTile = 128;  // anchors per CTA iteration == threads per CTA

// a7: pytorch_retinanet_detector_directional/retinanet/utils.py:102-149.
// HBM-bound (48 B in, 80 B out per anchor-image).  Regression rows are 48 B and output rows 80 B, so a
// thread-per-anchor mapping would issue 16-byte accesses at a 48/80-byte lane stride; instead both tiles are staged
// through shared memory so that every global access of a warp is one contiguous 512-byte run.
// (smem access pattern: 16-byte accesses at 12- and 20-word lane strides are bank-conflict free per quarter warp.)
// Arithmetic is the eager reference's, op by op: left-to-right +/- chains, then a separately rounded mul and add.
__global__ void __launch_bounds__(kDecTile) decode3d_kernel(const float4* __restrict__ anchors,
                                                            const float4* __restrict__ reg, int B, int A,
                                                            float4* __restrict__ out) {
    __shared__ float4 s_in[kDecTile * 3];
    __shared__ float4 s_out[kDecTile * 5];
    const int tid = threadIdx.x;
    const int a0 = blockIdx.x * kDecTile;
    const int nvalid = min(kDecTile, A - a0);
    const int a = a0 + tid;
    AnchorGeom ag;
    ag.w = ag.h = ag.cx = ag.cy = 0.f;
    if (tid < nvalid) ag = anchor_geom(__ldg(anchors + a));
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const float4* src = reg + ((int64_t)b * A + a0) * 3;
        float4* dst = out + ((int64_t)b * A + a0) * 5;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int i = tid + j * kDecTile;
            if (i < nvalid * 3) s_in[i] = ld_stream(src + i);
        }
        __syncthreads();
        if (tid < nvalid) {
            const float4 q0 = s_in[3 * tid], q1 = s_in[3 * tid + 1], q2 = s_in[3 * tid + 2];
            const float r[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
            float p[20];
            decode3d_row(r, ag, p);
#pragma unroll
            for (int j = 0; j < 5; ++j)
                s_out[5 * tid + j] = make_float4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int i = tid + j * kDecTile;
            if (i < nvalid * 5) st_stream(dst + i, s_out[i]);
        }
    }
}

// a8 (+a9): retinanet/utils.py:102-126 (+ :134-144).  One float4 in, one float4 out per anchor-image: already coalesced.
__global__ void __launch_bounds__(256) decode2d_kernel(const float4* __restrict__ anchors, int per_image_anchors,
                                                       const float4* __restrict__ deltas, int64_t BA, int A,
                                                       float4 mean, float4 stdv, int clip, float cw, float ch,
                                                       float4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < BA; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 an = __ldg(anchors + (per_image_anchors ? i : (i % A)));
        const float4 d = ld_stream(deltas + i);
        const float4 o = decode2d_row(an, d, mean, stdv, clip, cw, ch);
        st_stream(out + i, o);
    }
}

// a9: ClipBoxes in place on rows of K >= 4 floats.
__global__ void __launch_bounds__(256) clip_kernel(float* __restrict__ boxes, int64_t N, int K, float cw, float ch) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        float* row = boxes + i * K;
        if ((K & 3) == 0) {
            float4 v = *reinterpret_cast<float4*>(row);
            v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fminf(v.z, cw); v.w = fminf(v.w, ch);
            *reinterpret_cast<float4*>(row) = v;
        } else {
            row[0] = fmaxf(row[0], 0.0f); row[1] = fmaxf(row[1], 0.0f);
            row[2] = fminf(row[2], cw);   row[3] = fminf(row[3], ch);
        }
    }
}

}  // namespace g3d

using namespace g3d;

extern "C" int g3d_decode3d(const float* anchors, const float* reg, int64_t B, int64_t A, float* out, int device,
                            void* stream) {
    G3D_REQUIRE(B >= 0 && A >= 0, "negative size");
    G3D_REQUIRE(A < ((int64_t)1 << 31) - kDecTile, "A out of range");
    if (B == 0 || A == 0) return G3D_OK;
    G3D_REQUIRE(anchors && reg && out, "null pointer");
    G3D_REQUIRE(((uintptr_t)anchors % 16) == 0 && ((uintptr_t)reg % 16) == 0 && ((uintptr_t)out % 16) == 0,
                "pointers must be 16-byte aligned");
    G3D_GUARD(device);
    // one CTA per anchor tile and image group: anchors are read once per CTA and reused over its images
    const int64_t tiles = ceil_div(A, kDecTile);
    int64_t gy = B;
    while (gy > 1 && tiles * gy > (int64_t)sm_count(device) * 14 * 8) gy = (gy + 1) / 2;
    dim3 grid((unsigned)tiles, (unsigned)gy);
    decode3d_kernel<<<grid, kDecTile, 0, (cudaStream_t)stream>>>((const float4*)anchors, (const float4*)reg, (int)B,
                                                                 (int)A, (float4*)out);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int g3d_decode2d(const float* anchors, int64_t Ba, const float* deltas, int64_t B, int64_t A,
                            const float* mean_host, const float* std_host, int clip, float clip_w, float clip_h,
                            float* out, int device, void* stream) {
    G3D_REQUIRE(B >= 0 && A >= 0, "negative size");
    G3D_REQUIRE(Ba == 1 || Ba == B, "anchors batch must be 1 or B");
    G3D_REQUIRE(A < ((int64_t)1 << 31), "A out of range");
    if (B == 0 || A == 0) return G3D_OK;
    G3D_REQUIRE(anchors && deltas && out && mean_host && std_host, "null pointer");
    G3D_REQUIRE(((uintptr_t)anchors % 16) == 0 && ((uintptr_t)deltas % 16) == 0 && ((uintptr_t)out % 16) == 0,
                "pointers must be 16-byte aligned");
    G3D_GUARD(device);
    const int64_t BA = B * A;
    const int64_t blocks = ceil_div(BA, 256);
    const int grid = (int)(blocks < (int64_t)sm_count(device) * 32 ? blocks : (int64_t)sm_count(device) * 32);
    decode2d_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const float4*)anchors, (Ba == B && B > 1) ? 1 : 0, (const float4*)deltas, BA, (int)A,
        make_float4(mean_host[0], mean_host[1], mean_host[2], mean_host[3]),
        make_float4(std_host[0], std_host[1], std_host[2], std_host[3]), clip, clip_w, clip_h, (float4*)out);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int g3d_clip_boxes(float* boxes, int64_t N, int64_t K, float width, float height, int device, void* stream) {
    G3D_REQUIRE(N >= 0 && K >= 4 && K < (1 << 20), "need K >= 4");
    if (N == 0) return G3D_OK;
    G3D_REQUIRE(boxes, "null pointer");
    G3D_REQUIRE((K & 3) != 0 || ((uintptr_t)boxes % 16) == 0, "boxes must be 16-byte aligned");
    G3D_GUARD(device);
    const int64_t blocks = ceil_div(N, 256);
    const int grid = (int)(blocks < (int64_t)sm_count(device) * 32 ? blocks : (int64_t)sm_count(device) * 32);
    clip_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(boxes, N, (int)K, width, height);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}
