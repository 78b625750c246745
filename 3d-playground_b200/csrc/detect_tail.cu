// detect_tail.cu — the whole detection tail up to the one host decision, as ONE library call.
//
// ResNet.forward's post-processing (3D model.py:346-397, 2D retinanet/model.py:270-311) is, on this side: score filter ->
// candidate gather with on-the-fly decode -> segmented sort + NMS -> offsets of the kept rows.  Each step is an entry
// point of its own (filter.cu, nms.cu); issued from Python they cost ~10 us of host time apiece, which is what the
// ~10 short launches of this latency-bound chain wait for.  g3d_detect_tail issues them back to back from C++ on the
// caller's stream, carving its temporaries out of one caller-owned workspace, and leaves the two integers the host needs
// (number of detections, largest candidate count) in `summary`.
#include "common.cuh"

namespace g3d {

__global__ void __launch_bounds__(256) tail_summary_kernel(const int32_t* __restrict__ out_offsets,
                                                           const int32_t* __restrict__ count, int S,
                                                           int32_t* __restrict__ summary) {
    __shared__ int red[8];
    int m = 0;
    for (int i = threadIdx.x; i < S; i += blockDim.x) m = max(m, count[i]);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = max(m, red[w]);
        summary[0] = out_offsets[S];
        summary[1] = m;
    }
}

struct TailWorkspace {
    int32_t* idx;       // [S][cap] filter_compact's arrival-order indices
    float* cand_boxes;  // [S*cap][4]
    void* nms;          // g3d_nms_workspace_bytes(S*cap, S, cap)
    int64_t nms_bytes, bytes;
};
static TailWorkspace carve_tail(void* base, int64_t S, int64_t cap) {
    TailWorkspace w;
    char* p = (char*)base;
    int64_t off = 0;
    w.idx = (int32_t*)(p + off);      off += align_up(S * cap * 4, 256);
    w.cand_boxes = (float*)(p + off); off += align_up(S * cap * 16, 256);
    w.nms = (void*)(p + off);
    w.nms_bytes = g3d_nms_workspace_bytes(S * cap, S, cap);
    off += align_up(w.nms_bytes, 256);
    w.bytes = off;
    return w;
}

}  // namespace g3d

using namespace g3d;

extern "C" int64_t g3d_detect_tail_workspace_bytes(int64_t S, int64_t cap) {
    if (S < 0 || cap < 0) return G3D_ERR_INVALID;
    return carve_tail(nullptr, S, cap).bytes;
}

extern "C" int g3d_detect_tail(const float* scores, int64_t outer, int64_t inner, int64_t N, int64_t outer_pitch,
                               const float* thr, int64_t cap, const float* anchors, int64_t Ba, const float* reg,
                               int variant, const float* mean_host, const float* std_host, int clip, float clip_w,
                               float clip_h, double iou_threshold, int32_t* count, int32_t* seg_offsets,
                               float* cand_scores, int32_t* cand_src, int64_t* keep, int32_t* keep_count,
                               int32_t* out_offsets, int32_t* summary, void* workspace, int64_t workspace_bytes,
                               int device, void* stream) {
    G3D_REQUIRE(outer >= 1 && inner >= 1 && cap >= 1, "bad size");
    G3D_REQUIRE(count && seg_offsets && cand_scores && cand_src && keep && keep_count && out_offsets && summary && workspace,
                "null pointer");
    G3D_REQUIRE(((uintptr_t)workspace % 256) == 0, "workspace must be 256-byte aligned");
    const int64_t S = outer * inner;
    TailWorkspace w = carve_tail(workspace, S, cap);
    G3D_REQUIRE(workspace_bytes >= w.bytes, "workspace too small (see g3d_detect_tail_workspace_bytes)");
    int rc = g3d_filter_compact(scores, outer, inner, N, outer_pitch, thr, cap, w.idx, count, device, stream);
    if (rc) return rc;
    rc = g3d_gather_candidates_decoded(scores, outer, inner, N, outer_pitch, anchors, Ba, reg, variant, mean_host, std_host,
                                       clip, clip_w, clip_h, w.idx, count, cap, seg_offsets, cand_scores, w.cand_boxes,
                                       cand_src, device, stream);
    if (rc) return rc;
    rc = g3d_nms_segmented(w.cand_boxes, 4, 0, cand_scores, S * cap, seg_offsets, S, cap, iou_threshold, 0, keep, keep_count,
                           w.nms, w.nms_bytes, device, stream);
    if (rc) return rc;
    rc = g3d_exclusive_scan_i32(keep_count, S, out_offsets, device, stream);
    if (rc) return rc;
    G3D_GUARD(device);
    tail_summary_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(out_offsets, count, (int)S, summary);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}
