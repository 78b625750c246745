// iou_assign.cu — a1 calc_iou, the GT prologue of a6, and a2 the IoU-argmax assignment.
#include "assign_tile.cuh"

namespace g3d {

// ------------------------------------------------------------------------------------------------------ a1 calc_iou
// losses.py:5-22.  One thread per output element, G fastest (coalesced stores; a row of `a` is a warp broadcast).
__global__ void __launch_bounds__(256) calc_iou_kernel(const float4* __restrict__ a, int64_t A,
                                                       const float4* __restrict__ b, int64_t G,
                                                       float* __restrict__ out) {
    const int64_t total = A * G;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t ia = i / G, ig = i - ia * G;
        const float4 av = __ldg(a + ia), bv = __ldg(b + ig);
        out[i] = iou_retinanet(av, box_area_rn(av.x, av.y, av.z, av.w), bv, box_area_rn(bv.x, bv.y, bv.z, bv.w));
    }
}

// ------------------------------------------------------------------------------------------------- GT prologue (a6)
// One CTA per image: drop rows whose class column is -1 (3D losses.py:54, 2D retinanet/losses.py:46), keep order,
// and form the 2D box used for assignment (3D: min/max over the 8 projected corners, losses.py:93-107).
// Optionally zero-fills `zero_n` int32 counters (the loss kernels' tickets) in the same launch.
__global__ void __launch_bounds__(kTile) gt_prepare_kernel(const float* __restrict__ ann, int Gmax, int W, int variant,
                                                           float4* __restrict__ gt_box, int32_t* __restrict__ gt_row,
                                                           int32_t* __restrict__ gt_count, int32_t* __restrict__ zero_ptr,
                                                           int zero_n) {
    __shared__ int wcount[kWarps];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = b * kTile + threadIdx.x; i < zero_n; i += gridDim.x * kTile) zero_ptr[i] = 0;
    const float* img = ann + (int64_t)b * Gmax * W;
    const int cls_col = (variant == G3D_VARIANT_3D) ? 20 : 4;
    int count = 0;
    for (int base = 0; base < Gmax; base += kTile) {
        const int g = base + threadIdx.x;
        bool keep = false;
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < Gmax) {
            const float* row = img + (int64_t)g * W;
            keep = (row[cls_col] != -1.0f);
            if (keep) {
                if (variant == G3D_VARIANT_3D) {
                    float xmin = row[0], xmax = row[0], ymin = row[1], ymax = row[1];
#pragma unroll
                    for (int k = 1; k < 8; ++k) {
                        xmin = fminf(xmin, row[2 * k]); xmax = fmaxf(xmax, row[2 * k]);
                        ymin = fminf(ymin, row[2 * k + 1]); ymax = fmaxf(ymax, row[2 * k + 1]);
                    }
                    box = make_float4(xmin, ymin, xmax, ymax);
                } else {
                    box = make_float4(row[0], row[1], row[2], row[3]);
                }
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (base > 0) __syncthreads();
        if (lane == 0) wcount[warp] = __popc(bal);
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const int c = wcount[w];
            off += (w < warp) ? c : 0;
            total += c;
        }
        if (keep) {
            const int pos = count + off + __popc(bal & ((1u << lane) - 1u));
            gt_box[(int64_t)b * Gmax + pos] = box;
            gt_row[(int64_t)b * Gmax + pos] = g;
        }
        count += total;
    }
    if (threadIdx.x == 0) gt_count[b] = count;
}

// ------------------------------------------------------------------------------------------------------ a2 assign
__global__ void __launch_bounds__(kTile) assign_kernel(const float4* __restrict__ anchors, int A,
                                                       const float4* __restrict__ gt_box,
                                                       const int32_t* __restrict__ gt_row,
                                                       const int32_t* __restrict__ gt_count, int Gmax,
                                                       float* __restrict__ iou_max, int64_t* __restrict__ iou_argmax,
                                                       int32_t* __restrict__ assign, int32_t* __restrict__ num_pos) {
    __shared__ TileSmem sm;
    __shared__ int s_npos[kWarps];
    const int b = blockIdx.y;
    const int a = blockIdx.x * kTile + threadIdx.x;
    const bool valid = a < A;
    float4 an = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) an = __ldg(anchors + a);
    const float area_a = box_area_rn(an.x, an.y, an.z, an.w);
    const float4 bb = tile_bbox(an, valid, sm);
    const int G = gt_count[b];
    float best = 0.0f;
    int besti = 0;
    tile_argmax(an, area_a, bb, gt_box + (int64_t)b * Gmax, G, sm, best, besti);
    int code = G3D_ASSIGN_NEGATIVE;
    if (G > 0) code = assign_code(best, besti, gt_row + (int64_t)b * Gmax);
    if (valid) {
        const int64_t o = (int64_t)b * A + a;
        if (iou_max) iou_max[o] = best;
        if (iou_argmax) iou_argmax[o] = besti;
        if (assign) assign[o] = code;
    }
    if (num_pos) {
        int np = warp_sum((valid && code >= 0) ? 1 : 0);
        if ((threadIdx.x & 31) == 0) s_npos[threadIdx.x >> 5] = np;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) t += s_npos[w];
            if (t) atomicAdd(num_pos + b, t);
        }
    }
}

}  // namespace g3d

using namespace g3d;

extern "C" int g3d_calc_iou(const float* a, int64_t A, const float* b, int64_t G, float* out, int device, void* stream) {
    G3D_REQUIRE(A >= 0 && G >= 0, "negative size");
    if (A == 0 || G == 0) return G3D_OK;
    G3D_REQUIRE(a && b && out, "null pointer");
    G3D_GUARD(device);
    const int64_t total = A * G;
    const int64_t blocks = ceil_div(total, 256);
    const int grid = (int)(blocks < (int64_t)sm_count(device) * 64 ? blocks : (int64_t)sm_count(device) * 64);
    calc_iou_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)a, A, (const float4*)b, G, out);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

namespace g3d {
int gt_prepare_launch(const float* ann, int64_t B, int64_t Gmax, int64_t W, int variant, float* gt_box, int32_t* gt_row,
                      int32_t* gt_count, int32_t* zero_ptr, int64_t zero_n, int device, void* stream);
}

extern "C" int g3d_gt_prepare(const float* ann, int64_t B, int64_t Gmax, int64_t W, int variant, float* gt_box,
                              int32_t* gt_row, int32_t* gt_count, int device, void* stream) {
    return g3d::gt_prepare_launch(ann, B, Gmax, W, variant, gt_box, gt_row, gt_count, nullptr, 0, device, stream);
}

int g3d::gt_prepare_launch(const float* ann, int64_t B, int64_t Gmax, int64_t W, int variant, float* gt_box,
                           int32_t* gt_row, int32_t* gt_count, int32_t* zero_ptr, int64_t zero_n, int device,
                           void* stream) {
    G3D_REQUIRE(B >= 0 && Gmax >= 0, "negative size");
    G3D_REQUIRE(variant == G3D_VARIANT_2D || variant == G3D_VARIANT_3D, "unknown variant");
    G3D_REQUIRE(variant == G3D_VARIANT_3D ? W >= 21 : W >= 5, "annotation rows too narrow for this variant");
    G3D_REQUIRE(Gmax < (1 << 30) && W < (1 << 20) && B < (1 << 30), "size out of range");
    if (B == 0) return G3D_OK;
    G3D_REQUIRE(gt_count, "null pointer");
    G3D_REQUIRE(Gmax == 0 || (ann && gt_box && gt_row), "null pointer");
    G3D_GUARD(device);
    G3D_REQUIRE(zero_n >= 0 && zero_n < (1 << 30) && (zero_n == 0 || zero_ptr), "bad counter range");
    gt_prepare_kernel<<<(int)B, kTile, 0, (cudaStream_t)stream>>>(ann, (int)Gmax, (int)W, variant, (float4*)gt_box,
                                                                  gt_row, gt_count, zero_ptr, (int)zero_n);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int g3d_assign(const float* anchors, int64_t A, const float* gt_box, const int32_t* gt_row,
                          const int32_t* gt_count, int64_t B, int64_t Gmax, float* iou_max, int64_t* iou_argmax,
                          int32_t* assign, int32_t* num_pos, int device, void* stream) {
    G3D_REQUIRE(A >= 0 && B >= 0 && Gmax >= 0, "negative size");
    G3D_REQUIRE(A < (int64_t)1 << 31 && B <= 65535, "size out of range (A < 2^31, B <= 65535)");
    if (A == 0 || B == 0) return G3D_OK;
    G3D_REQUIRE(anchors && gt_count, "null pointer");
    G3D_REQUIRE(Gmax == 0 || (gt_box && gt_row), "null pointer");
    G3D_GUARD(device);
    cudaStream_t st = (cudaStream_t)stream;
    if (num_pos) G3D_CUDA(cudaMemsetAsync(num_pos, 0, sizeof(int32_t) * B, st));
    dim3 grid((unsigned)ceil_div(A, kTile), (unsigned)B);
    assign_kernel<<<grid, kTile, 0, st>>>((const float4*)anchors, (int)A, (const float4*)gt_box, gt_row, gt_count,
                                          (int)Gmax, iou_max, iou_argmax, assign, num_pos);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}
