// pairwise.cu — a20 md_iou: FP64 IoU of axis-aligned footprints (MC3D_crop_tracker.py:1030-1049).
//
// The callers broadcast both operands to [n,m,4] float64 with .repeat() before calling md_iou (:278-280, :687-689,
// :1013); the un-broadcast kernel below forms out[i,j] = IoU(first[i], second[j]) directly, reading 16 B per box and
// writing the [n,m] float64 matrix once.  Arithmetic order is the reference's, in FP64, no epsilon (0/0 -> NaN).
#include "common.cuh"

namespace g3d {

__device__ __forceinline__ double iou_f64(double ax0, double ay0, double ax1, double ay1, double bx0, double by0,
                                          double bx1, double by1, double eps) {
    const double area_a = __dmul_rn(__dsub_rn(ax1, ax0), __dsub_rn(ay1, ay0));
    const double area_b = __dmul_rn(__dsub_rn(bx1, bx0), __dsub_rn(by1, by0));
    const double minx = fmax(ax0, bx0), maxx = fmin(ax1, bx1);
    const double miny = fmax(ay0, by0), maxy = fmin(ay1, by1);
    const double inter = __dmul_rn(fmax(0.0, __dsub_rn(maxx, minx)), fmax(0.0, __dsub_rn(maxy, miny)));
    double uni = __dsub_rn(__dadd_rn(area_a, area_b), inter);
    if (eps != 0.0) uni = __dadd_rn(uni, eps);
    return __ddiv_rn(inter, uni);
}

__global__ void __launch_bounds__(256) pairwise_iou_kernel(const float4* __restrict__ first, int64_t n,
                                                           const float4* __restrict__ second, int64_t m, double eps,
                                                           int one_minus, double* __restrict__ out) {
    const int64_t total = n * m;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / m, c = i - r * m;
        const float4 a = __ldg(first + r), b = __ldg(second + c);
        const double v = iou_f64(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, eps);
        out[i] = one_minus ? __dsub_rn(1.0, v) : v;
    }
}

__global__ void __launch_bounds__(256) md_iou_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                     int64_t n, double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double* p = a + 4 * i;
        const double* q = b + 4 * i;
        out[i] = iou_f64(p[0], p[1], p[2], p[3], q[0], q[1], q[2], q[3], 0.0);
    }
}

// estimate_ts_bias's pair mining (MC3D_crop_tracker.py:277-289): all (i, j), i < j, seen by different cameras, whose
// footprints overlap with IoU > threshold, in the row-major order of the reference's double loop (the bias update that
// consumes them is sequential, so the order is part of the result).  One warp per row i; lanes stride over j; the
// d x d float64 matrix of the reference is never formed.  WRITE = false counts per row, WRITE = true stores the pairs at
// the row's offset (ballot-ordered).
template <bool WRITE>
__global__ void __launch_bounds__(256) cross_camera_pairs_kernel(const float4* __restrict__ boxes,
                                                                 const int32_t* __restrict__ cams, int d, double thr,
                                                                 int32_t* __restrict__ row_count,
                                                                 const int32_t* __restrict__ row_offsets,
                                                                 int2* __restrict__ pairs, int64_t capacity) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < d; i += warps) {
        const float4 a = __ldg(boxes + i);
        const int ca = __ldg(cams + i);
        int64_t base = WRITE ? (int64_t)row_offsets[i] : 0;
        int count = 0;
        for (int j0 = i + 1; j0 < d; j0 += 32) {
            const int j = j0 + lane;
            bool hit = false;
            if (j < d && __ldg(cams + j) != ca) {
                const float4 b = __ldg(boxes + j);
                // iou[i, j] = md_iou(boxes[j], boxes[i]) (:278-280): the formula is symmetric bit for bit
                hit = iou_f64(b.x, b.y, b.z, b.w, a.x, a.y, a.z, a.w, 0.0) > thr;
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, hit);
            if (WRITE) {
                const int64_t pos = base + __popc(ballot & ((1u << lane) - 1u));
                if (hit && pos < capacity) pairs[pos] = make_int2(i, j);
                base += __popc(ballot);
            } else {
                count += __popc(ballot);
            }
        }
        if (!WRITE && lane == 0) row_count[i] = count;
    }
}

}  // namespace g3d

using namespace g3d;

extern "C" int g3d_cross_camera_pairs(const float* boxes, const int32_t* cams, int64_t d, double threshold,
                                      int32_t* row_count, const int32_t* row_offsets, int32_t* pairs, int64_t capacity,
                                      int device, void* stream) {
    G3D_REQUIRE(d >= 0 && d < ((int64_t)1 << 31) && capacity >= 0, "bad size");
    if (d == 0) return G3D_OK;
    G3D_REQUIRE(boxes && cams, "null pointer");
    G3D_REQUIRE(((uintptr_t)boxes % 16) == 0, "boxes must be 16-byte aligned");
    G3D_REQUIRE(row_offsets ? (pairs != nullptr || capacity == 0) : (row_count != nullptr),
                "pass row_count (count pass) or row_offsets + pairs (write pass)");
    G3D_GUARD(device);
    const int64_t blocks = ceil_div(d, 8);       // 8 warps per CTA, one row per warp
    const int grid = (int)(blocks < (int64_t)sm_count(device) * 8 ? blocks : (int64_t)sm_count(device) * 8);
    if (row_offsets)
        cross_camera_pairs_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(
            (const float4*)boxes, cams, (int)d, threshold, nullptr, row_offsets, (int2*)pairs, capacity);
    else
        cross_camera_pairs_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(
            (const float4*)boxes, cams, (int)d, threshold, row_count, nullptr, nullptr, 0);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int g3d_pairwise_iou_f64(const float* first, int64_t n, const float* second, int64_t m, double eps,
                                    int one_minus, double* out, int device, void* stream) {
    G3D_REQUIRE(n >= 0 && m >= 0, "negative size");
    if (n == 0 || m == 0) return G3D_OK;
    G3D_REQUIRE(first && second && out, "null pointer");
    G3D_REQUIRE(((uintptr_t)first % 16) == 0 && ((uintptr_t)second % 16) == 0, "boxes must be 16-byte aligned");
    G3D_GUARD(device);
    const int64_t blocks = ceil_div(n * m, 256);
    const int grid = (int)(blocks < (int64_t)sm_count(device) * 16 ? blocks : (int64_t)sm_count(device) * 16);
    pairwise_iou_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)first, n, (const float4*)second, m, eps,
                                                                one_minus, out);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int g3d_md_iou(const double* a, const double* b, int64_t n, double* out, int device, void* stream) {
    G3D_REQUIRE(n >= 0, "negative size");
    if (n == 0) return G3D_OK;
    G3D_REQUIRE(a && b && out, "null pointer");
    G3D_GUARD(device);
    const int64_t blocks = ceil_div(n, 256);
    const int grid = (int)(blocks < (int64_t)sm_count(device) * 16 ? blocks : (int64_t)sm_count(device) * 16);
    md_iou_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, n, out);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}
