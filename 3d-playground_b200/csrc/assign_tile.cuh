// assign_tile.cuh — the tiled anchor x GT IoU / argmax core shared by g3d_assign and the fused loss kernel.
//
// One CTA owns a tile of kTile consecutive anchors (one per thread).  Anchors arrive in the reference's spatial
// order (level -> row -> col -> 9 shapes, anchors.py:109-129) so a tile has a tight bounding box; GT boxes that
// cannot touch that bounding box are culled while being staged into shared memory.  Culling is exact: a culled
// pair has iw or ih <= 0 -> clamped to 0 -> IoU == +0.0 exactly (union is clamped >= 1e-8 > 0), and the running
// (max, argmax) starts at (0.0, 0) and only moves on a strict '>' while the survivors are visited in ascending GT
// index, which is torch.max(IoU, dim=1)'s first-maximal-index rule (losses.py:110) for every anchor.
#pragma once
#include "common.cuh"

namespace g3d {

constexpr int kTile = 256;           // anchors per CTA == threads per CTA
constexpr int kWarps = kTile / 32;

struct TileSmem {
    float4 box[kTile];   // surviving GT boxes of the current chunk, ascending GT index
    float area[kTile];
    int idx[kTile];      // compacted-GT index of each survivor
    float red[4][kWarps];
    int wcount[kWarps];
};

// Bounding box of the tile's anchors (min x1, min y1, max x2, max y2); invalid threads pass neutral values.
__device__ __forceinline__ float4 tile_bbox(const float4& an, bool valid, TileSmem& sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mnx = valid ? an.x : INFINITY, mny = valid ? an.y : INFINITY;
    float mxx = valid ? an.z : -INFINITY, mxy = valid ? an.w : -INFINITY;
    mnx = warp_min(mnx); mny = warp_min(mny); mxx = warp_max(mxx); mxy = warp_max(mxy);
    if (lane == 0) { sm.red[0][warp] = mnx; sm.red[1][warp] = mny; sm.red[2][warp] = mxx; sm.red[3][warp] = mxy; }
    __syncthreads();
    float4 bb = make_float4(sm.red[0][0], sm.red[1][0], sm.red[2][0], sm.red[3][0]);
#pragma unroll
    for (int w = 1; w < kWarps; ++w) {
        bb.x = fminf(bb.x, sm.red[0][w]); bb.y = fminf(bb.y, sm.red[1][w]);
        bb.z = fmaxf(bb.z, sm.red[2][w]); bb.w = fmaxf(bb.w, sm.red[3][w]);
    }
    return bb;
}

// max / first-argmax of IoU(anchor, gt[g]) over g in [0, G).  All kTile threads of the CTA must call this.
// gt points at the image's compacted GT boxes.  best/besti must be initialised to (0.0f, 0) by the caller.
__device__ __forceinline__ void tile_argmax(const float4& an, float area_a, const float4 bb,
                                            const float4* __restrict__ gt, int G, TileSmem& sm,
                                            float& best, int& besti) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < G; base += kTile) {
        const int g = base + threadIdx.x;
        bool hit = false;
        float4 gb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < G) {
            gb = __ldg(gt + g);
            // keep unless provably disjoint from every anchor of the tile (NaN coordinates are never culled)
            hit = !(gb.z <= bb.x || gb.x >= bb.z || gb.w <= bb.y || gb.y >= bb.w);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (base > 0) __syncthreads();  // previous chunk fully consumed before its slots are reused
        if (lane == 0) sm.wcount[warp] = __popc(bal);
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const int c = sm.wcount[w];
            off += (w < warp) ? c : 0;
            total += c;
        }
        if (hit) {
            const int pos = off + __popc(bal & ((1u << lane) - 1u));
            sm.box[pos] = gb;
            sm.area[pos] = box_area_rn(gb.x, gb.y, gb.z, gb.w);
            sm.idx[pos] = g;
        }
        __syncthreads();
#pragma unroll 2
        for (int k = 0; k < total; ++k) {
            const float4 gk = sm.box[k];
            const float iw = __fsub_rn(fminf(an.z, gk.z), fmaxf(an.x, gk.x));
            const float ih = __fsub_rn(fminf(an.w, gk.w), fmaxf(an.y, gk.y));
            // iw <= 0 or ih <= 0 -> the clamped intersection is 0 -> IoU == +0.0 exactly, which can never beat `best`
            // under the strict '>' rule: skip the union / IEEE division (most survivors of the tile-level cull are
            // still disjoint from most anchors of the tile)
            if (iw > 0.0f && ih > 0.0f) {
                const float inter = __fmul_rn(iw, ih);
                const float ua = fmaxf(__fsub_rn(__fadd_rn(area_a, sm.area[k]), inter), 1e-8f);
                const float v = __fdiv_rn(inter, ua);
                if (v > best) { best = v; besti = sm.idx[k]; }
            }
        }
    }
}

// classification of the assignment (losses.py:121-131): IoU_max < 0.4 -> negative, >= 0.5 -> positive, else ignore
__device__ __forceinline__ int assign_code(float best, int besti, const int32_t* __restrict__ gt_row_img) {
    if (best >= 0.5f) return __ldg(gt_row_img + besti);
    return (best < 0.4f) ? G3D_ASSIGN_NEGATIVE : G3D_ASSIGN_IGNORE;
}

}  // namespace g3d
