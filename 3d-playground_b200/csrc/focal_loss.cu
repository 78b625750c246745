// focal_loss.cu — a2..a6: FocalLoss.forward / backward of both retinanet copies.
//
//   3D: pytorch_retinanet_detector_directional/retinanet/losses.py:27-362
//   2D: retinanet/losses.py:27-177
//
// Two launches per call (plus the tiny GT prologue of iou_assign.cu):
//
//   1. assign_codes_kernel  (compute only, anchors and GT boxes are L2 resident)
//        grid (anchor tiles, image groups).  A CTA owns 256 consecutive anchors and kImgPerCta images.  GT boxes are
//        culled against the tile (bounding box + "can this box reach IoU 0.4 with any anchor of the tile at all"),
//        compacted into shared memory in ascending GT order, refined per warp, and only pairs whose IoU can still
//        reach 0.4 pay the IEEE division.  Output: one int32 assignment code per (image, anchor) and the number of
//        positives per image (integer atomics - deterministic).
//   2. focal_stream_kernel  (HBM bound - the dominant kernel)
//        grid (anchor tiles, images), one warp per 32 consecutive (image, anchor) rows.  Streams the classification
//        rows (fully coalesced 16-byte loads: lane l takes float4 l and l+32 of the warp's 1 KB - the focal term of a
//        negative anchor does not depend on which row an element belongs to), evaluates the focal terms AND their
//        gradients from the same -log(1-p) and streams the gradient rows out.  (The zero-fill of the regression
//        gradient - 41 % of the bytes this pass would otherwise move - is done by launch 1, which is issue-bound and
//        leaves HBM idle.)
//        The per-image normaliser 1/num_pos is already known from launch 1, so forward and backward of the whole
//        loss are ONE pass over the data: cls is read once, -log once, dcls / dreg are written once.
//        Partial sums: FP32 inside a warp (<= 256 terms), FP64 across warps / tiles, fixed order; the last CTA of an
//        image (atomic ticket) reduces that image's partials, the last image forms the batch means - deterministic,
//        no second launch, no host synchronisation.
//
// The classification gradient written by launch 2 assumes the upstream gradient of the classification loss that the
// host announces (1 for `(cls + reg + vp).backward()`, 1/world under dist.py).  g3d_focal_loss_bwd checks that
// assumption ON THE DEVICE: if the real upstream gradient is the announced one, dcls is already right and the kernel
// only scans the assignment codes and writes the regression-gradient rows of the positive anchors (for whatever the
// real upstream gradients of the regression / direction losses are); otherwise it also recomputes dcls.  No host
// synchronisation either way.
#include <stdlib.h>
#include "assign_tile.cuh"

namespace g3d {
int gt_prepare_launch(const float* ann, int64_t B, int64_t Gmax, int64_t W, int variant, float* gt_box, int32_t* gt_row,
                      int32_t* gt_count, int32_t* zero_ptr, int64_t zero_n, int device, void* stream);
}

namespace g3d {

constexpr int kImgPerCta = 4;  // images processed per CTA of assign_codes_kernel (anchors / statistics loaded once)
constexpr int kStageGroup = 4; // images staged behind one pair of barriers

// clamp bounds of losses.py:56: torch.clamp(classification, 1e-4, 1.0 - 1e-4) - python doubles cast to f32
#define G3D_PMIN ((float)1e-4)
#define G3D_PMAX ((float)(1.0 - 1e-4))
#define G3D_BETA ((float)(1.0 / 9.0))        // smooth-L1 switch point (losses.py:346)
#define G3D_HALF_BETA ((float)(0.5 / 9.0))   // losses.py:348

// one focal term (losses.py:138-150): alpha_t * (1 - p_t)^2 * bce, target t in {0,1}
__device__ __forceinline__ float focal_term(float p_raw, bool t) {
    const float p = fminf(fmaxf(p_raw, G3D_PMIN), G3D_PMAX);
    const float u = 1.0f - p;
    const float fw = t ? u : p;
    const float x = t ? p : u;
    const float w = (t ? 0.25f : 0.75f) * (fw * fw);
    return w * (-logf(x));
}

// -log(u) for u = fl(1 - p) in (0, 1).  Negative anchors dominate and their probabilities are small, so the common
// case avoids the ~22-instruction logf: with pe = 1 - u (exact, Sterbenz) and z = pe / (2 - pe) = pe / (1 + u),
//     -log(1 - pe) = 2 atanh(z) = 2 z (1 + z^2/3 + z^4/5 + z^6/7 + z^8/9 + ...),
// truncated after z^6 (|z| < 1/7 for pe < 0.25: relative truncation error < 1.9e-8).  The series is evaluated on the
// SAME rounded u the reference takes the log of, so it tracks torch.log(1.0 - classification) to ~2e-7 relative.
// 1/x for x in a benign range (here [0.75, 2] and [1e-4, 1]): a single MUFU.RCP (<= 1 ulp).  __fdividef would add four
// instructions of denormal-range scaling per quotient.
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float neg_log_u_series(float u) {   // valid for 1 - u < 0.25
    const float pe = 1.0f - u;
    const float z = pe * rcp_fast(1.0f + u);
    const float z2 = z * z;
    // 2 (1 + z^2/3 + z^4/5 + z^6/7): the first dropped term, z^8/9, is < 1.9e-8 relative for |z| < 1/7 - below FP32 epsilon
    float s = fmaf(z2, 2.0f / 7.0f, 0.4f);
    s = fmaf(z2, s, 2.0f / 3.0f);
    s = fmaf(z2, s, 2.0f);
    return z * s;
}
__device__ __noinline__ float neg_log_full(float u) { return -logf(u); }   // rare: kept out of line (code size)
__device__ __forceinline__ float neg_log_u(float u) {
    return (1.0f - u < 0.25f) ? neg_log_u_series(u) : -logf(u);
}

// N elements of NEGATIVE anchors (target 0): sum of 0.75 p^2 * -log(1-p) and, if GRAD, d/dp of each term times
// `scale` (zero outside the clamp range: torch.clamp's backward passes min <= x <= max).  Straight-line code - the
// series of all N elements first, so the dependency chains interleave - with a rare, separate fix-up for
// probabilities >= 0.25 (which need the full logf).  The quotient p^2/u uses the 1-ulp MUFU reciprocal: gradients are
// compared at 1e-5 relative, nothing here is index-critical.
template <int N, bool GRAD>
__device__ __forceinline__ float focal_neg(const float* pv, float scale, float* g) {
    float p[N], nl[N];
    bool big = false;
#pragma unroll
    for (int c = 0; c < N; ++c) {
        p[c] = fminf(fmaxf(pv[c], G3D_PMIN), G3D_PMAX);
        const float u = 1.0f - p[c];
        nl[c] = neg_log_u_series(u);
        big |= !(1.0f - u < 0.25f);
    }
    if (big) {
#pragma unroll
        for (int c = 0; c < N; ++c)
            if (!(1.0f - (1.0f - p[c]) < 0.25f)) nl[c] = neg_log_full(1.0f - p[c]);
    }
    // value 0.75 p^2 nl, gradient 0.75 (2 p nl + p^2 / u): the common factor is applied once (sum) / folded into `scale`
    float acc = 0.0f;
    const float scale75 = 0.75f * scale;
#pragma unroll
    for (int c = 0; c < N; ++c) {
        const float a = p[c] * nl[c];
        acc = fmaf(p[c], a, acc);
        if (GRAD) {
            const float pr = p[c] * rcp_fast(1.0f - p[c]);
            const float t = fmaf(2.0f, a, p[c] * pr);
            g[c] = (p[c] == pv[c]) ? scale75 * t : 0.0f;   // p == p_raw  <=>  p_raw inside [min, max]
        }
    }
    return 0.75f * acc;
}

// d(focal term)/dp, zero outside the clamp range.
__device__ __forceinline__ float focal_term_grad_neg(float p) {
    if (!(p >= G3D_PMIN && p <= G3D_PMAX)) return 0.0f;
    const float u = 1.0f - p;
    return 1.5f * p * neg_log_u(u) + (0.75f * (p * p)) * rcp_fast(u);
}
__device__ __forceinline__ float focal_term_grad(float p, bool t) {
    if (!t) return focal_term_grad_neg(p);
    if (!(p >= G3D_PMIN && p <= G3D_PMAX)) return 0.0f;
    const float u = 1.0f - p;
    return 0.5f * u * logf(p) - (0.25f * (u * u)) * rcp_fast(p);
}

__device__ __forceinline__ float smooth_l1(float d) {
    return (d <= G3D_BETA) ? 4.5f * (d * d) : d - G3D_HALF_BETA;
}

__device__ __forceinline__ float cos_loss(float rx, float ry, float tx, float ty) {
    const float rn = sqrtf(rx * rx + ry * ry), tn = sqrtf(tx * tx + ty * ty);
    return 1.0f - (rx * tx + ry * ty) / (rn * tn);
}
// gradient of cos_loss w.r.t. (rx, ry)
__device__ __forceinline__ void cos_loss_grad(float rx, float ry, float tx, float ty, float& gx, float& gy) {
    // d(1 - r.t/(|r||t|))/dr = -(t^ - cos * r^)/|r|.  In the plane t^ - cos*r^ = (r_perp^ . t^) r_perp^, which gives the
    // cancellation-free form  g = cross * (ry, -rx) / (|r|^3 |t|),  cross = rx*ty - ry*tx  (evaluated with an exact
    // product residual).  The textbook form subtracts two terms of size 1/|r| and loses digits when r is nearly
    // parallel to t or very short; this one stays within a few ulp of the exact gradient.
    const float r2 = rx * rx + ry * ry;
    const float rn = sqrtf(r2), tn = sqrtf(tx * tx + ty * ty);
    const float p = ry * tx, e = fmaf(ry, tx, -p);
    const float cross = fmaf(rx, ty, -p) - e;
    const float k = cross / ((r2 * rn) * tn);
    gx = k * ry;
    gy = -k * rx;
}

// corner sign table of losses.py:311-327 / utils.py:114-130: corner k = c + sl*L + sw*W + sh*H
__device__ __forceinline__ float sgn_l(int k) { return (k & 2) ? 1.0f : -1.0f; }
__device__ __forceinline__ float sgn_w(int k) { return (k & 1) ? 1.0f : -1.0f; }
__device__ __forceinline__ float sgn_h(int k) { return (k & 4) ? -1.0f : 1.0f; }

// the three GT direction vectors (losses.py:222-223, 252-253, 281-282) from the raw 16 corner coordinates
__device__ __forceinline__ void gt_directions(const float* t, float* tv /*6*/) {
    tv[0] = ((t[4] + t[6] + t[12] + t[14]) - (t[0] + t[2] + t[8] + t[10])) / 4.0f;
    tv[1] = ((t[5] + t[7] + t[13] + t[15]) - (t[1] + t[3] + t[9] + t[11])) / 4.0f;
    tv[2] = ((t[2] + t[6] + t[10] + t[14]) - (t[0] + t[4] + t[8] + t[12])) / 4.0f;
    tv[3] = ((t[3] + t[7] + t[11] + t[15]) - (t[1] + t[5] + t[9] + t[13])) / 4.0f;
    tv[4] = ((t[0] + t[2] + t[4] + t[6]) - (t[8] + t[10] + t[12] + t[14])) / 4.0f;
    tv[5] = ((t[1] + t[3] + t[5] + t[7]) - (t[9] + t[11] + t[13] + t[15])) / 4.0f;
}

__device__ __forceinline__ void pred_corners(const float* r, float* p /*20*/) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        p[2 * k] = ((r[0] + sgn_l(k) * r[2]) + sgn_w(k) * r[4]) + sgn_h(k) * r[6];
        p[2 * k + 1] = ((r[1] + sgn_l(k) * r[3]) + sgn_w(k) * r[5]) + sgn_h(k) * r[7];
    }
    p[16] = r[8]; p[17] = r[9]; p[18] = r[10]; p[19] = r[11];
}

// 2D targets (retinanet/losses.py:137-157)
__device__ __forceinline__ void targets_2d(const float* __restrict__ grow, const float4& an, float* t /*4*/) {
    const float aw = an.z - an.x, ah = an.w - an.y;
    const float acx = an.x + 0.5f * aw, acy = an.y + 0.5f * ah;
    float gw = grow[2] - grow[0], gh = grow[3] - grow[1];
    const float gcx = grow[0] + 0.5f * gw, gcy = grow[1] + 0.5f * gh;
    gw = fmaxf(gw, 1.0f);
    gh = fmaxf(gh, 1.0f);
    t[0] = ((gcx - acx) / aw) / 0.1f;
    t[1] = ((gcy - acy) / ah) / 0.1f;
    t[2] = logf(gw / aw) / 0.2f;
    t[3] = logf(gh / ah) / 0.2f;
}

// One POSITIVE anchor (rare path, out of line so that the streaming main path stays small in registers and code):
// the regression loss terms of the row (3D: 20 smooth-L1 terms + mean of the three cosine losses, losses.py:156-350;
// 2D: 4 smooth-L1 terms, retinanet/losses.py:129-173) and, if drow != null, the row's regression gradient for the
// upstream gradients (g_reg, g_vp).
template <int VARIANT>
__device__ __noinline__ void positive_row(const float* __restrict__ rrow, const float* __restrict__ grow, const float4 an,
                                           float s_reg, float s_vp, float* __restrict__ drow, float& reg_sum,
                                           float& vp_term) {
    if (VARIANT == G3D_VARIANT_3D) {
        float r[12], t[20], pr[20], tv[6];
#pragma unroll
        for (int i = 0; i < 12; ++i) r[i] = rrow[i];
#pragma unroll
        for (int i = 0; i < 20; ++i) t[i] = grow[i];
        gt_directions(t, tv);
        vp_term = (cos_loss(r[2], r[3], tv[0], tv[1]) + cos_loss(r[4], r[5], tv[2], tv[3]) +
                   cos_loss(r[6], r[7], tv[4], tv[5])) / 3.0f;
        pred_corners(r, pr);
        const float aw = an.z - an.x, ah = an.w - an.y;
        const float acx = an.x + 0.5f * aw, acy = an.y + 0.5f * ah;
        float s = 0.0f, g[20];
#pragma unroll
        for (int i = 0; i < 20; ++i) {
            const float tn = (i & 1) ? (t[i] - acy) / ah : (t[i] - acx) / aw;   // losses.py:330-331
            const float diff = tn - pr[i];
            const float w = (i >= 8 && i < 16) ? 0.5f : 1.0f;                   // top_weighting, losses.py:343
            const float d = fabsf(diff) * w;
            s += smooth_l1(d);
            const float sg = (diff > 0.0f) ? 1.0f : ((diff < 0.0f) ? -1.0f : 0.0f);
            // d smooth_l1 / d pred = slope(d) * w * d|diff|/dpred = slope * w * (-sign(diff))
            g[i] = -s_reg * ((d <= G3D_BETA) ? 9.0f * d : 1.0f) * w * sg;
        }
        reg_sum = s;
        if (drow) {
            float dr[12];
#pragma unroll
            for (int i = 0; i < 8; ++i) dr[i] = 0.0f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                dr[0] += g[2 * k];            dr[1] += g[2 * k + 1];
                dr[2] += sgn_l(k) * g[2 * k]; dr[3] += sgn_l(k) * g[2 * k + 1];
                dr[4] += sgn_w(k) * g[2 * k]; dr[5] += sgn_w(k) * g[2 * k + 1];
                dr[6] += sgn_h(k) * g[2 * k]; dr[7] += sgn_h(k) * g[2 * k + 1];
            }
            dr[8] = g[16]; dr[9] = g[17]; dr[10] = g[18]; dr[11] = g[19];
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                float gx, gy;
                cos_loss_grad(r[2 + 2 * v], r[3 + 2 * v], tv[2 * v], tv[2 * v + 1], gx, gy);
                dr[2 + 2 * v] += s_vp * gx;
                dr[3 + 2 * v] += s_vp * gy;
            }
#pragma unroll
            for (int i = 0; i < 12; ++i) drow[i] = dr[i];
        }
    } else {
        float t[4];
        targets_2d(grow, an, t);
        float s = 0.0f;
        vp_term = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float diff = t[i] - rrow[i];
            const float d = fabsf(diff);
            s += smooth_l1(d);
            if (drow) {
                const float sg = (diff > 0.0f) ? 1.0f : ((diff < 0.0f) ? -1.0f : 0.0f);
                drow[i] = -s_reg * ((d <= G3D_BETA) ? 9.0f * d : 1.0f) * sg;
            }
        }
        reg_sum = s;
    }
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// =====================================================================================================================
// launch 1: assignment codes
// =====================================================================================================================
// zero-fill `nrows` consecutive rows of dreg (R floats each; 16-byte aligned because R is 4 or 12) with coalesced
// 16-byte streaming stores
template <int R>
__device__ __forceinline__ void zero_rows(float* base, int nrows, int lane) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* b4 = reinterpret_cast<float4*>(base);
    if (nrows == 32) {
#pragma unroll
        for (int k = 0; k < R / 4; ++k) st_stream(b4 + lane + 32 * k, z);
    } else {
        for (int i = lane; i < nrows * (R / 4); i += 32) st_stream(b4 + i, z);
    }
}

struct AssignCodesArgs {
    const float4* anchors;
    const float4* gt_box;
    const int32_t* gt_row;
    const int32_t* gt_count;
    int32_t* assign;     // [B][A]
    int32_t* npos;       // [B], zero on entry
    int32_t* pos_list;   // [B][A]: anchor indices of the positives of each image, in arrival order (first npos[b] valid)
    float* dreg;         // [B][A][R] or null: zero-filled here (see the kernel)
    int B, A, Gmax, R;
};

// One-lane atomics as single instructions.  `if (lane == 0) atomicAdd(...)` makes nvcc wrap the call in its warp-aggregation
// pattern (vote, leader election, popc, shuffle: ~15 instructions around one ATOMS), which at one ticket per 32-anchor
// unit was 15 % of this issue-bound kernel's instructions.
// The address is made lane-dependent in a way the compiler cannot fold (offset zero for lane 0, the one lane that executes
// it; %laneid read through asm): ptxas then emits the bare ATOMS / ATOMG instead of its aggregation sequence.
__device__ __forceinline__ int lane0_atomic_add_shared(int* addr, int v, int lane) {
    int old;
    unsigned l2;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l2));          // opaque to the front end: not folded under `lane == 0`
    (void)lane;
    const unsigned a = (unsigned)__cvta_generic_to_shared(addr) + (l2 ? 4u : 0u);
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ int lane0_atomic_add_global(int* addr, int v) {
    int old;
    unsigned l2;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l2));
    asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(addr + (l2 ? 1 : 0)), "r"(v) : "memory");
    return old;
}

struct StageSmem {
    float4 box[kImgPerCta][kTile];
    int idx[kImgPerCta][kTile];
    float4 an[kTile];          // the tile's anchors: any warp can process any 32-anchor slice
    float4 wbb[kWarps];        // bounding box of each slice
    int next_unit;             // dynamic (slice, image) work units of the item
    float red[8][kWarps];
    float tile[8];
    int tile_wf;
    int wcount[kImgPerCta][kWarps];
    int total[kImgPerCta];
};

// Size statistics of a group of anchors (a warp or a tile), for the "can the IoU reach 0.4 at all" cull:
// for well-formed boxes  IoU = inter / union,  inter <= min(wa, wg) * min(ha, hg)  and  inter <= min(area_a, area_g),
// union >= max(area_a, area_g).  So a GT box whose best case over the group stays below 0.38 (margin for the FP32
// roundings of the real thing, which are ~1e-7 relative) has IoU < 0.4 with every anchor of the group: whatever
// its exact IoU, it cannot move an anchor out of the `negative` class, and it can be skipped.
struct GroupStats {
    float4 bb;          // min x1, min y1, max x2, max y2
    float wmax, hmax;   // largest width / height
    float amin, amax;   // smallest / largest area
    bool wellformed;    // every anchor has positive width and height
};

__device__ __forceinline__ bool can_touch(const float4& g, const GroupStats& s) {
    // keep unless provably disjoint from every anchor of the group (NaN coordinates are never culled here)
    if (g.z <= s.bb.x || g.x >= s.bb.z || g.w <= s.bb.y || g.y >= s.bb.w) return false;
    const float gw = g.z - g.x, gh = g.w - g.y;
    if (s.wellformed && gw > 0.0f && gh > 0.0f) {
        const float ga = gw * gh;
        const float best_inter = fminf(fminf(s.wmax, gw) * fminf(s.hmax, gh), fminf(s.amax, ga));
        if (best_inter < 0.38f * fmaxf(s.amin, ga)) return false;
    }
    return true;
}

// warp maximum through the integer redux unit: IEEE floats order like sign-magnitude integers, so flipping the low 31
// bits of negative values gives two's-complement keys with the same order (NaN keys sort beyond +-inf: a NaN anchor
// only disables culling, it never matches anything)
__device__ __forceinline__ float warp_max_redux(float x) {
    int k = __float_as_int(x);
    k ^= (k >> 31) & 0x7fffffff;
    k = __reduce_max_sync(0xffffffffu, k);
    k ^= (k >> 31) & 0x7fffffff;
    return __int_as_float(k);
}

// One work item of the assignment: anchor tile `tile` (kTile consecutive anchors) x image group `group` (kImgPerCta
// images).  All kTile threads of the CTA take part (barriers inside).
__device__ __forceinline__ void assign_item(const AssignCodesArgs& p, StageSmem& sm, int tile, int group) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int a = tile * kTile + tid;
    const bool valid = a < p.A;
    float4 an = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) an = __ldg(p.anchors + a);
    const float area_a = box_area_rn(an.x, an.y, an.z, an.w);
    const int b0 = group * kImgPerCta;
    const int nimg = min(kImgPerCta, p.B - b0);

    // ---- statistics of the warp's and of the tile's anchors (one barrier)
    GroupStats ws, ts;
    {
        const float w = an.z - an.x, h = an.w - an.y;
        float v[8];
        v[0] = valid ? -an.x : -INFINITY; v[1] = valid ? -an.y : -INFINITY;    // maxima of negated values = minima
        v[2] = valid ? an.z : -INFINITY;  v[3] = valid ? an.w : -INFINITY;
        v[4] = valid ? w : -INFINITY;     v[5] = valid ? h : -INFINITY;
        v[6] = valid ? -area_a : -INFINITY; v[7] = valid ? area_a : -INFINITY;
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = warp_max_redux(v[k]);
        const bool wf = __all_sync(0xffffffffu, !valid || (w > 0.0f && h > 0.0f));
        ws.bb = make_float4(-v[0], -v[1], v[2], v[3]);
        ws.wmax = v[4]; ws.hmax = v[5]; ws.amin = -v[6]; ws.amax = v[7];
        ws.wellformed = wf;
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) sm.red[k][warp] = v[k];
            sm.wcount[0][warp] = wf ? 1 : 0;
            sm.wbb[warp] = ws.bb;
        }
        sm.an[tid] = an;
        if (tid < kImgPerCta) sm.total[tid] = 0;
        if (tid == 0) sm.next_unit = 0;
        __syncthreads();
        if (warp == 0) {     // tile statistics: lane k < 8 reduces statistic k over the warps
            if (lane < 8) {
                float t = sm.red[lane][0];
#pragma unroll
                for (int w2 = 1; w2 < kWarps; ++w2) t = fmaxf(t, sm.red[lane][w2]);
                sm.tile[lane] = t;
            }
            if (lane == 8) {
                int twf = 1;
#pragma unroll
                for (int w2 = 0; w2 < kWarps; ++w2) twf &= sm.wcount[0][w2];
                sm.tile_wf = twf;
            }
        }
        __syncthreads();
        ts.bb = make_float4(-sm.tile[0], -sm.tile[1], sm.tile[2], sm.tile[3]);
        ts.wmax = sm.tile[4]; ts.hmax = sm.tile[5]; ts.amin = -sm.tile[6]; ts.amax = sm.tile[7];
        ts.wellformed = sm.tile_wf != 0;
        __syncthreads();   // sm.wcount[0] is reused below
    }

    // ---- stage the GT boxes of the group's images, culled against the tile.  An image with more than kTile GT rows
    // keeps its first kTile candidates here and is finished by the (rare) overflow loop further down.
#pragma unroll 1
    for (int i0 = 0; i0 < nimg; i0 += kStageGroup) {
        const int g = tid;
        float4 gb[kStageGroup];
        unsigned bal[kStageGroup];
#pragma unroll
        for (int j = 0; j < kStageGroup; ++j) {
            const int i = i0 + j;
            bool hit = false;
            gb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int Gi = (i < nimg) ? __ldg(p.gt_count + b0 + i) : 0;
            if (g < Gi) {
                gb[j] = __ldg(p.gt_box + (int64_t)(b0 + i) * p.Gmax + g);
                hit = can_touch(gb[j], ts);
            }
            bal[j] = __ballot_sync(0xffffffffu, hit);
        }
        // unordered compaction (one shared-memory atomic per warp and image): the candidate loop below breaks ties by
        // GT index explicitly, so the order of the survivors does not matter
        int base[kStageGroup];
#pragma unroll
        for (int j = 0; j < kStageGroup; ++j) {
            base[j] = 0;
            if (lane == 0 && bal[j]) base[j] = lane0_atomic_add_shared(&sm.total[i0 + j], __popc(bal[j]), lane);
        }
#pragma unroll
        for (int j = 0; j < kStageGroup; ++j) {
            if (bal[j] == 0) continue;
            const int bs = __shfl_sync(0xffffffffu, base[j], 0);
            if ((bal[j] >> lane) & 1u) {
                const int pos = bs + __popc(bal[j] & ((1u << lane) - 1u));
                sm.box[i0 + j][pos] = gb[j];
                sm.idx[i0 + j][pos] = g;
            }
        }
    }
    __syncthreads();

    // ---- (slice of 32 anchors) x (image) work units, handed out dynamically: the candidate loops of the slices differ a
    // lot in length, and a warp that is done early takes the next unit instead of waiting at the end of the item.
    // No barriers; units are image-major so that the warps walk the same staged image together.
    const int nunits = nimg * kWarps;
#pragma unroll 1
    for (;;) {
        int u = 0;
        if (lane == 0) u = lane0_atomic_add_shared(&sm.next_unit, 1, lane);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= nunits) break;
        const int i = u / kWarps, slice = u - i * kWarps;
        const int b = b0 + i;
        const int a = tile * kTile + slice * 32 + lane;
        const bool valid = a < p.A;
        const float4 an = sm.an[slice * 32 + lane];
        const float area_a = box_area_rn(an.x, an.y, an.z, an.w);
        GroupStats ws;
        ws.bb = sm.wbb[slice];
        const int Gi = __ldg(p.gt_count + b);
        // The regression gradient is zero except on the few positive rows.  This kernel is issue-bound and leaves HBM
        // idle, so the 48 bytes / row of zeros are written from here (3 coalesced 16-byte stores per lane, drained in
        // the background) instead of costing the HBM-bound streaming kernel 41 % more traffic.
        if (p.dreg) {
            const int wa0 = tile * kTile + slice * 32, nrows = min(32, p.A - wa0);
            if (nrows > 0) {
                if (p.R == 12) zero_rows<12>(p.dreg + ((int64_t)b * p.A + wa0) * 12, nrows, lane);
                else           zero_rows<4>(p.dreg + ((int64_t)b * p.A + wa0) * 4, nrows, lane);
            }
        }
        float best = 0.0f;
        int besti = 0;
        const int total = sm.total[i];
        for (int k0 = 0; k0 < total; k0 += 32) {
            // warp-level refinement: which of these (up to 32) tile survivors matter for this warp's anchors?
            bool near = false;
            if (k0 + lane < total) {   // the size test was already applied with the tile's statistics
                const float4 t = sm.box[i][k0 + lane];
                near = !(t.z <= ws.bb.x || t.x >= ws.bb.z || t.w <= ws.bb.y || t.y >= ws.bb.w);
            }
            unsigned m = __ballot_sync(0xffffffffu, near);
            while (m) {
                const int k = k0 + __ffs(m) - 1;
                m &= m - 1;
                const float4 gk = sm.box[i][k];
                const float iw = __fsub_rn(fminf(an.z, gk.z), fmaxf(an.x, gk.x));
                const float ih = __fsub_rn(fminf(an.w, gk.w), fmaxf(an.y, gk.y));
                // a disjoint pair has IoU == +0.0 exactly and can never beat `best` under the strict '>' rule
                if (iw > 0.0f && ih > 0.0f) {
                    const float inter = __fmul_rn(iw, ih);
                    const float ua0 = __fsub_rn(__fadd_rn(area_a, box_area_rn(gk.x, gk.y, gk.z, gk.w)), inter);
                    // inter <= ua0 / 2.6  =>  IoU <= 0.3847 < 0.4: cannot change the code of this anchor, skip the
                    // division (the clamp of the union only matters below 1e-8, where this test passes)
                    if (__fmul_rn(inter, 2.6f) > ua0) {
                        const float v = __fdiv_rn(inter, fmaxf(ua0, 1e-8f));
                        const int gi = sm.idx[i][k];
                        if (v > best || (v == best && gi < besti)) { best = v; besti = gi; }   // first maximal index
                    }
                }
            }
        }
        // overflow: GT rows beyond the first kTile of this image, straight from global memory (warp-uniform loop)
        for (int g = kTile; g < Gi; ++g) {
            const float4 gk = __ldg(p.gt_box + (int64_t)b * p.Gmax + g);
            const float iw = __fsub_rn(fminf(an.z, gk.z), fmaxf(an.x, gk.x));
            const float ih = __fsub_rn(fminf(an.w, gk.w), fmaxf(an.y, gk.y));
            if (iw > 0.0f && ih > 0.0f) {
                const float inter = __fmul_rn(iw, ih);
                const float ua = fmaxf(__fsub_rn(__fadd_rn(area_a, box_area_rn(gk.x, gk.y, gk.z, gk.w)), inter), 1e-8f);
                const float v = __fdiv_rn(inter, ua);
                if (v > best) { best = v; besti = g; }
            }
        }
        // `best` is exact whenever it is >= 0.4 (every pair that can reach 0.3847 was evaluated exactly; ties go to the
        // lower GT index = torch.max's first-maximal-index); below that only "< 0.4" is used.
        int code = G3D_ASSIGN_NEGATIVE;
        if (Gi > 0) code = assign_code(best, besti, p.gt_row + (int64_t)b * p.Gmax);
        if (valid) p.assign[(int64_t)b * p.A + a] = code;
        // positives are appended to the image's list (integer atomics: the COUNT is deterministic, the order is not -
        // everything downstream is order independent: per-row gradients, and loss sums in exact fixed point)
        const bool is_pos = valid && code >= 0;
        const unsigned posmask = __ballot_sync(0xffffffffu, is_pos);
        if (posmask) {
            int base = 0;
            if (lane == 0) base = lane0_atomic_add_global(p.npos + b, __popc(posmask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (is_pos) p.pos_list[(int64_t)b * p.A + base + __popc(posmask & ((1u << lane) - 1u))] = a;
        }
    }
}

__global__ void __launch_bounds__(kTile, 6) assign_codes_kernel(const AssignCodesArgs p) {
    __shared__ StageSmem sm;
    assign_item(p, sm, blockIdx.x, blockIdx.y);
}

// =====================================================================================================================
// launch 1, GT-centric form (anchors known to be the regular pyramid of Anchors.forward, Gmax <= 256)
// =====================================================================================================================
// assign_codes_kernel above walks every (32-anchor slice, image) unit - 389 k of them at cfg2 - although 98 % of the
// anchors are negatives: it is instruction-bound on per-unit overhead.  When the anchor table is the pyramid (level ->
// cell -> shape, anchors.py:21-40) the loop can be inverted.  One warp per (image, GT row): for every level and shape the
// cells whose anchor can reach IoU 0.385 with this box form a small window - IoU >= t needs inter >= t/(1+t) (Aa + Ag),
// and inter <= iw * min(ah, gh), so iw >= that / min(ah, gh), which bounds the anchor centre to
// [gx1 + iw_min - aw/2, gx2 - iw_min + aw/2] (same in y; FP64, widened by 0.01 px) - and only those pairs (~110 per GT
// row, 1.3 x the pairs that really reach 0.385) are evaluated, with the exact arithmetic of the kernel above on the
// anchor values read from the table.  A pair with IoU >= 0.4 goes into a per-(image, anchor) key with atomicMax:
// key = (IoU bits - bits(0.4f) + 1) << 8 | (255 - GT index): larger IoU wins, then the lower index = torch.max's
// first-maximal rule (a fire-and-forget RED: a first version that used the returned old value to build a list of touched
// anchors spent 70 % of its time waiting on that round trip).  One coalesced pass over the keys (4 B / anchor) then
// writes the codes of the anchors that have one and builds the positives list.  Everything else is a pure fill
// (codes = -1, keys = 0, dreg = 0) at HBM speed.
// Measured at cfg2 (B = 32, 1080p, 200 GT rows / image): keys + codes fill 17 us, pairs 26 us, resolve 28 us = 71 us
// against 150+ us for the anchor-centric kernel - but the 0.6 GB zero-fill of dreg, which that issue-bound kernel hides
// behind its instruction stream, costs 97 us here (6.2 TB/s), and running it on a second stream saturates the memory
// system and slows this latency-bound chain by as much as it saves.  So the callers choose: GT-centric when no gradient
// buffers are requested (forward-only), anchor-centric for the training step.
constexpr int kPyrLevels = 8, kPyrShapes = 16;
constexpr unsigned kIou04Bits = 0x3ECCCCCDu;   // 0.4f
struct Pyramid {
    float aw[kPyrLevels * kPyrShapes], ah[kPyrLevels * kPyrShapes];
    float inv_stride[kPyrLevels];
    int first[kPyrLevels], rows[kPyrLevels], cols[kPyrLevels];
    int L, S;
};

__global__ void __launch_bounds__(256) assign_fill_kernel(int32_t* __restrict__ codes, uint32_t* __restrict__ keys,
                                                          float* __restrict__ dreg, long long n_rows, int R) {
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    const float4 zf = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dreg) {
        float4* d4 = reinterpret_cast<float4*>(dreg);
        const long long n4 = n_rows * R / 4;                 // R is 4 or 12
        for (long long i = tid; i < n4; i += nthr) st_stream(d4 + i, zf);
    }
    if (!codes) return;
    const long long q = n_rows >> 2;
    const float4 neg = make_float4(__int_as_float(-1), __int_as_float(-1), __int_as_float(-1), __int_as_float(-1));
    for (long long i = tid; i < q; i += nthr) {
        st_stream(reinterpret_cast<float4*>(codes) + i, neg);
        st_stream(reinterpret_cast<float4*>(keys) + i, zf);
    }
    for (long long i = (q << 2) + tid; i < n_rows; i += nthr) { codes[i] = G3D_ASSIGN_NEGATIVE; keys[i] = 0u; }
}

struct PairArgs {
    const float4* anchors;
    const float4* gt_box;      // [B][Gmax] compacted valid rows
    const int32_t* gt_count;   // [B]
    uint32_t* keys;            // [B][A]
    int B, A, Gmax;
};

__global__ void __launch_bounds__(256) assign_pairs_kernel(const PairArgs p, const __grid_constant__ Pyramid pyr) {
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= p.B * p.Gmax) return;
    const int b = wid / p.Gmax, g = wid - b * p.Gmax;
    if (g >= __ldg(p.gt_count + b)) return;
    const float4 gk = __ldg(p.gt_box + (int64_t)b * p.Gmax + g);
    const float gw = gk.z - gk.x, gh = gk.w - gk.y;
    if (!(gw > 0.0f && gh > 0.0f)) return;                    // no overlap with anything is possible
    const float Ag = gw * gh;
    const float area_g = box_area_rn(gk.x, gk.y, gk.z, gk.w);
    uint32_t* keys = p.keys + (int64_t)b * p.A;
    // The window only has to be a superset: FP32 with approximate reciprocals (relative error ~1e-6) against a threshold
    // 3.9 % below the 0.4 that matters, plus 0.05 px of slack on the centre range.
    // The (level, shape) combinations are spread over the lanes - each lane derives the window of its own combination
    // (all of them in one or two rounds instead of L*S uniform iterations) - and the warp then walks the non-empty
    // windows together, 32 cells at a time.
    constexpr float kq = 0.385f / 1.385f, eps = 0.05f;
    const int ncombo = pyr.L * pyr.S;
    for (int k0 = 0; k0 < ncombo; k0 += 32) {
        const int k = k0 + lane;
        int c0 = 0, r0 = 0, wc = 0, ncell = 0, l = 0;
        if (k < ncombo) {
            l = k / pyr.S;
            const float aw = pyr.aw[k], ah = pyr.ah[k];
            const float imin = kq * fmaf(aw, ah, Ag);
            const float mw = fminf(aw, gw), mh = fminf(ah, gh);
            if (mw * mh >= imin) {
                const float inv_stride = pyr.inv_stride[l];
                const float iw_min = 0.9999f * __fdividef(imin, mh), ih_min = 0.9999f * __fdividef(imin, mw);
                const float lo_x = gk.x + iw_min - 0.5f * aw - eps, hi_x = gk.z - iw_min + 0.5f * aw + eps;
                const float lo_y = gk.y + ih_min - 0.5f * ah - eps, hi_y = gk.w - ih_min + 0.5f * ah + eps;
                c0 = max(0, (int)ceilf(lo_x * inv_stride - 0.5f));
                r0 = max(0, (int)ceilf(lo_y * inv_stride - 0.5f));
                const int c1 = min(pyr.cols[l] - 1, (int)floorf(hi_x * inv_stride - 0.5f));
                const int r1 = min(pyr.rows[l] - 1, (int)floorf(hi_y * inv_stride - 0.5f));
                if (c1 >= c0 && r1 >= r0) { wc = c1 - c0 + 1; ncell = wc * (r1 - r0 + 1); }
            }
        }
        unsigned live = __ballot_sync(0xffffffffu, ncell > 0);
        while (live) {
            const int src = __ffs(live) - 1;
            live &= live - 1;
            const int kc = k0 + src;
            const int lc = __shfl_sync(0xffffffffu, l, src), sc = kc - lc * pyr.S;
            const int c0c = __shfl_sync(0xffffffffu, c0, src), r0c = __shfl_sync(0xffffffffu, r0, src);
            const int wcc = __shfl_sync(0xffffffffu, wc, src), nc = __shfl_sync(0xffffffffu, ncell, src);
            const int cols = pyr.cols[lc], first = pyr.first[lc];
            for (int tcell = lane; tcell < nc; tcell += 32) {
                const int r = r0c + tcell / wcc, c = c0c + tcell % wcc;
                const int a = first + (r * cols + c) * pyr.S + sc;
                const float4 an = __ldg(p.anchors + a);
                const float iw = __fsub_rn(fminf(an.z, gk.z), fmaxf(an.x, gk.x));
                const float ih = __fsub_rn(fminf(an.w, gk.w), fmaxf(an.y, gk.y));
                if (!(iw > 0.0f && ih > 0.0f)) continue;
                const float inter = __fmul_rn(iw, ih);
                const float ua0 = __fsub_rn(__fadd_rn(box_area_rn(an.x, an.y, an.z, an.w), area_g), inter);
                if (!(__fmul_rn(inter, 2.6f) > ua0)) continue;                 // IoU <= 0.3847: cannot matter
                const float v = __fdiv_rn(inter, fmaxf(ua0, 1e-8f));
                if (!(v >= 0.4f)) continue;
                const unsigned key = ((__float_as_uint(v) - kIou04Bits + 1u) << 8) | (unsigned)(255 - g);
                atomicMax(keys + a, key);      // result unused: a fire-and-forget RED, nothing waits on it
            }
        }
    }
}

struct ResolveArgs {
    const uint32_t* keys;
    const int32_t* gt_row;     // [B][Gmax] original annotation row of each compacted row
    int32_t* assign;
    int32_t* pos_list;
    int32_t* npos;
    int A, Gmax;
};

// grid (x, B): one coalesced pass over the keys of image blockIdx.y, four per thread (16-byte loads when the row
// allows); an anchor with a key gets its code (original GT row for IoU >= 0.5, else IGNORE); positives are appended to
// the image's list with one atomic per warp and round (order is irrelevant downstream: exact fixed-point sums, per-row
// gradients)
__global__ void __launch_bounds__(256) assign_resolve_kernel(const ResolveArgs p) {
    constexpr int U = 4;                                        // 16-byte loads in flight per thread
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const uint32_t* __restrict__ keys = p.keys + (int64_t)b * p.A;
    const bool vec = (((int64_t)b * p.A) & 3) == 0;             // the image's key row starts 16-byte aligned
    const int nq = (p.A + 3) >> 2;
    for (int q0 = blockIdx.x * blockDim.x * U; q0 < nq; q0 += gridDim.x * blockDim.x * U) {   // warp-uniform trip count
        unsigned k[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int q = q0 + u * blockDim.x + threadIdx.x, a0 = q << 2;
            k[u][0] = k[u][1] = k[u][2] = k[u][3] = 0u;
            if (q < nq) {
                if (vec && a0 + 3 < p.A) {
                    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(keys) + q);
                    k[u][0] = v.x; k[u][1] = v.y; k[u][2] = v.z; k[u][3] = v.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) if (a0 + e < p.A) k[u][e] = __ldcs(keys + a0 + e);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!__any_sync(0xffffffffu, (k[u][0] | k[u][1] | k[u][2] | k[u][3]) != 0u)) continue;
            const int a0 = (q0 + u * blockDim.x + threadIdx.x) << 2;
            int npos_mine = 0, code[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                code[e] = G3D_ASSIGN_NEGATIVE;
                if (k[u][e]) {
                    const float v = __uint_as_float((k[u][e] >> 8) - 1u + kIou04Bits);
                    const int g = 255 - (int)(k[u][e] & 255u);
                    code[e] = (v >= 0.5f) ? __ldg(p.gt_row + (int64_t)b * p.Gmax + g) : G3D_ASSIGN_IGNORE;
                    p.assign[(int64_t)b * p.A + a0 + e] = code[e];
                    npos_mine += code[e] >= 0;
                }
            }
            // exclusive prefix of the per-lane positive counts, one atomic for the warp
            int incl = npos_mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (total == 0) continue;
            int base = 0;
            if (lane == 0) base = lane0_atomic_add_global(p.npos + b, total);
            base = __shfl_sync(0xffffffffu, base, 0) + incl - npos_mine;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (code[e] >= 0) p.pos_list[(int64_t)b * p.A + base++] = a0 + e;
        }
    }
}

// =====================================================================================================================
// launch 2: the positive anchors (dense: one thread per positive, from the lists launch 1 built)
// =====================================================================================================================
// Loss sums of the positives are accumulated in exact fixed point - two int64 limbs per sum, units 2^-20 and 2^-52 -
// with integer atomics: integer addition is associative, so the result does not depend on the (non-deterministic)
// order of the lists.  Range 2^43 per sum, absolute resolution 2^-52 per term.
struct PosArgs {
    const float* reg;
    const float4* anchors;
    const float* ann;
    const int32_t* assign;
    const int32_t* pos_list;
    const int32_t* npos;
    long long* acc;            // [B][4]: reg_hi, reg_lo, vp_hi, vp_lo (forward)
    int32_t* nonfinite;        // [B]: set if a term was NaN / Inf / out of range (forward)
    const float* grad_out;     // [3] device (backward)
    const float* grad_scale;   // [3] device or null (backward): multiplies grad_out (dist.py: local -> global means)
    const float* losses;       // [4] (backward: losses[3] = number of images with >= 1 GT row)
    float* dreg;               // (backward)
    int B, A, R, Gmax, W;
};

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void fixed_split(float t, long long& hi, long long& lo, bool& bad) {
    bad = !(fabsf(t) < 1.0e12f);            // NaN, Inf, or beyond the 2^43 range of the high limb
    const double td = bad ? 0.0 : (double)t;
    const double h = floor(td * 1048576.0);                              // 2^20
    hi = (long long)h;
    lo = (long long)((td - h * (1.0 / 1048576.0)) * 4503599627370496.0);  // 2^52: in [0, 2^32)
}

template <int VARIANT, bool BWD>
__device__ __forceinline__ void positives_body(const PosArgs& p) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int n = min(__ldg(p.npos + b), p.A);
    const float npos = (float)n;
    float s_reg = 0.0f, s_vp = 0.0f;
    if (BWD) {
        const float per_pos = (VARIANT == G3D_VARIANT_3D) ? 20.0f : 4.0f;
        const float go1 = __ldg(p.grad_out + 1) * (p.grad_scale ? __ldg(p.grad_scale + 1) : 1.0f);
        s_reg = go1 / ((float)p.B * per_pos * npos);
        if (VARIANT == G3D_VARIANT_3D) {
            const float go2 = __ldg(p.grad_out + 2) * (p.grad_scale ? __ldg(p.grad_scale + 2) : 1.0f);
            s_vp = go2 / (__ldg(p.losses + 3) * npos * 3.0f);
        }
    }
    const int n_up = (n + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_up; i += gridDim.x * blockDim.x) {
        float reg_sum = 0.0f, vp_term = 0.0f;
        if (i < n) {
            const int a = __ldg(p.pos_list + (int64_t)b * p.A + i);
            const int64_t row = (int64_t)b * p.A + a;
            const int code = __ldg(p.assign + row);
            const float* grow = p.ann + ((int64_t)b * p.Gmax + code) * p.W;
            positive_row<VARIANT>(p.reg + row * p.R, grow, __ldg(p.anchors + a), s_reg, s_vp,
                                  BWD ? p.dreg + row * p.R : nullptr, reg_sum, vp_term);
        }
        if (!BWD) {
            long long rh, rl, vh, vl;
            bool bad_r, bad_v;
            fixed_split(reg_sum, rh, rl, bad_r);
            fixed_split(vp_term, vh, vl, bad_v);
            rh = warp_sum_ll(rh); rl = warp_sum_ll(rl);
            if (VARIANT == G3D_VARIANT_3D) { vh = warp_sum_ll(vh); vl = warp_sum_ll(vl); }
            const bool any_bad = __any_sync(0xffffffffu, bad_r || bad_v);
            if (lane == 0) {
                unsigned long long* acc = reinterpret_cast<unsigned long long*>(p.acc + 4 * b);
                atomicAdd(acc + 0, (unsigned long long)rh);
                atomicAdd(acc + 1, (unsigned long long)rl);
                if (VARIANT == G3D_VARIANT_3D) {
                    atomicAdd(acc + 2, (unsigned long long)vh);
                    atomicAdd(acc + 3, (unsigned long long)vl);
                }
                if (any_bad) atomicOr(p.nonfinite + b, 1);
            }
        }
    }
}

template <int VARIANT, bool BWD>
__global__ void __launch_bounds__(128) positives_kernel(const PosArgs p) { positives_body<VARIANT, BWD>(p); }

// =====================================================================================================================
// launch 3: the streaming classification pass (loss terms + gradient) - the dominant, HBM-bound kernel
// =====================================================================================================================
constexpr int kChunksPerWarp = 8;                            // 32-row chunks handled by one warp
constexpr int kRowsPerCta = kWarps * 32 * kChunksPerWarp;    // 1024 (image, anchor) rows per CTA

struct StreamArgs {
    const float* cls;
    const float* ann;
    const int32_t* assign;     // [B][A]
    const int32_t* npos;       // [B]
    const int32_t* gt_count;   // [B]
    const long long* acc;      // [B][4] fixed-point sums of the positives (launch 2)
    const int32_t* nonfinite;  // [B]
    int32_t* gt_count_out;     // [B] or null: copy of gt_count for the caller
    double* partials;          // [B][T]: classification partial sums, one per CTA
    int32_t* counters;         // [B] image tickets + [1] batch ticket, zero on entry
    float* losses;             // [4] : cls, reg, vp, number of non-empty images
    float* per_image;          // [B][4]
    double* shard_stats;       // [5] or null: sum cls_j, sum reg_j, sum vp_j (images with GT), B, #images with GT
    float* dcls;               // [B][A][C]  (GRAD only)
    float* dreg;               // [B][A][R]  (GRAD only; zero-filled by launch 1)
    float g0;                  // upstream gradient of the classification loss that dcls is formed for
    int B, A, C, R, Gmax, W, T;
    int cpw;                   // 32-row chunks per warp and work item (rows per item = kWarps * 32 * cpw; T = items per image)
};

// Executed by ONE warp - the last CTA of image b: reduce the image's T partials in a fixed order (lane-strided
// accumulation + shuffle tree), then (last image of the batch) the batch means.
template <int VARIANT>
__device__ __forceinline__ void finalize_image(const StreamArgs& p, int b) {
    const int lane = threadIdx.x & 31;
    __threadfence();
    double tc = 0.0;
    const double* src = p.partials + (int64_t)b * p.T;
#pragma unroll 4
    for (int t = lane; t < p.T; t += 32) tc += __ldcg(src + t);
    tc = warp_sum(tc);
    int last = 0;
    if (lane == 0) {
        const double tn = (double)__ldcg(p.npos + b);
        const double per_pos = (VARIANT == G3D_VARIANT_3D) ? 20.0 : 4.0;
        const long long* acc = p.acc + 4 * b;
        const double lo_unit = 1.0 / 4503599627370496.0, hi_unit = 1.0 / 1048576.0;
        double tr = (double)__ldcg(acc + 0) * hi_unit + (double)__ldcg(acc + 1) * lo_unit;
        double tv = (double)__ldcg(acc + 2) * hi_unit + (double)__ldcg(acc + 3) * lo_unit;
        if (__ldcg(p.nonfinite + b)) tr = tv = __longlong_as_double(0x7ff8000000000000LL);   // NaN
        float4 o;
        o.x = (float)(tc / fmax(tn, 1.0));                      // losses.py:152 (and :70 for empty images)
        o.y = tn > 0.0 ? (float)(tr / (tn * per_pos)) : 0.0f;  // .mean() over P x 20 (:350) / P x 4
        o.z = tn > 0.0 ? (float)(tv / tn) : 0.0f;              // vp_loss.mean() (:304)
        o.w = (float)tn;
        __stcg(reinterpret_cast<float4*>(p.per_image) + b, o);
        if (p.gt_count_out) p.gt_count_out[b] = __ldg(p.gt_count + b);
        __threadfence();
        last = (atomicAdd(p.counters + p.B, 1) == p.B - 1);
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
        // last image: batch means (losses.py:362).  vp: only images with >= 1 GT row contribute (:304,:353-358)
        __threadfence();
        double sc = 0.0, sr = 0.0, sv = 0.0, ne = 0.0;
        for (int j = lane; j < p.B; j += 32) {
            const float4 o = __ldcg(reinterpret_cast<const float4*>(p.per_image) + j);
            sc += o.x; sr += o.y;
            if (__ldg(p.gt_count + j) > 0) { sv += o.z; ne += 1.0; }
        }
        sc = warp_sum(sc); sr = warp_sum(sr); sv = warp_sum(sv); ne = warp_sum(ne);
        if (lane == 0) {
            p.losses[0] = (float)(sc / p.B);
            p.losses[1] = (float)(sr / p.B);
            p.losses[2] = (VARIANT == G3D_VARIANT_3D) ? (float)(sv / ne) : 0.0f;  // 0/0 -> NaN when all empty
            p.losses[3] = (float)ne;
            if (p.shard_stats) {
                p.shard_stats[0] = sc; p.shard_stats[1] = sr; p.shard_stats[2] = (VARIANT == G3D_VARIANT_3D) ? sv : 0.0;
                p.shard_stats[3] = (double)p.B; p.shard_stats[4] = ne;
            }
        }
    }
}

// Fix-up of one float4 (4 classes of row r) of a POSITIVE anchor whose assigned class is element e of this float4: the
// other three elements keep their negative-anchor terms, element e trades its target-0 term for the target-1 term.
// Rare and divergent: out of line, full logf; arguments and result by value so that nothing of the hot path is forced
// into local memory.
struct QuadOut {
    float4 g;
    float acc;
};
template <bool GRAD>
__device__ __noinline__ QuadOut positive_fix(float4 v, int e, float s_cls, float acc, float4 g) {
    const float pe = (e == 0) ? v.x : (e == 1) ? v.y : (e == 2) ? v.z : v.w;
    QuadOut o;
    o.acc = acc + (focal_term(pe, true) - focal_term(pe, false));
    o.g = g;
    if (GRAD) {
        const float ge = s_cls * focal_term_grad(pe, true);
        if (e == 0) o.g.x = ge; else if (e == 1) o.g.y = ge; else if (e == 2) o.g.z = ge; else o.g.w = ge;
    }
    return o;
}

// The data one lane holds of a 32-row chunk of the C == 8 path: lane l owns float4 l and l + 32 of the chunk's 64
// (32 rows x 2 float4), i.e. classes (l & 1) * 4 .. + 3 of rows l >> 1 and 16 + (l >> 1) - fully coalesced in both
// directions - plus the assignment code of row l.
struct Chunk8 {
    float4 v0, v1;
    int code;
};

// all 32 rows of the chunk exist (the caller routes an image's ragged last chunk to stream_chunk_any)
// COHERENT: the codes were written earlier in the SAME launch (fused kernel) - they must not come through the
// non-coherent read-only path
template <bool COHERENT>
__device__ __forceinline__ Chunk8 load_chunk8(const StreamArgs& p, int64_t row0, int lane) {
    Chunk8 c;
    const float4* cp = reinterpret_cast<const float4*>(p.cls + row0 * 8);
    c.code = COHERENT ? __ldcg(p.assign + row0 + lane) : __ldg(p.assign + row0 + lane);
    c.v0 = ld_stream(cp + lane);
    c.v1 = ld_stream(cp + 32 + lane);
    return c;
}

// Focal terms (+ gradient) of one loaded chunk; returns the lane's share of the sum.
template <int VARIANT, bool GRAD>
__device__ __forceinline__ float process_chunk8(const StreamArgs& p, int b, int64_t row0, int lane, float s_cls,
                                                const Chunk8& c) {
    // every element as if its anchor were negative (the overwhelmingly common case) ...
    const float pa[4] = {c.v0.x, c.v0.y, c.v0.z, c.v0.w}, pb[4] = {c.v1.x, c.v1.y, c.v1.z, c.v1.w};
    float ga[4], gb[4];
    float acc0 = focal_neg<4, GRAD>(pa, s_cls, ga);
    float acc1 = focal_neg<4, GRAD>(pb, s_cls, gb);
    float4 g0 = make_float4(ga[0], ga[1], ga[2], ga[3]), g1 = make_float4(gb[0], gb[1], gb[2], gb[3]);
    // ... then redo the float4s that belong to positive / ignored anchors
    if (__any_sync(0xffffffffu, c.code != G3D_ASSIGN_NEGATIVE)) {
        const int cls_col = (VARIANT == G3D_VARIANT_3D) ? 20 : 4;
        int pos_cls = -1;
        if (c.code >= 0) pos_cls = (int)(long long)p.ann[((int64_t)b * p.Gmax + c.code) * p.W + cls_col];
        const int r0 = lane >> 1, r1 = 16 + (lane >> 1), c0 = (lane & 1) * 4;
        const int code0 = __shfl_sync(0xffffffffu, c.code, r0), pc0 = __shfl_sync(0xffffffffu, pos_cls, r0);
        const int code1 = __shfl_sync(0xffffffffu, c.code, r1), pc1 = __shfl_sync(0xffffffffu, pos_cls, r1);
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (code0 == G3D_ASSIGN_IGNORE) { acc0 = 0.0f; g0 = zero4; }          // ignored anchor: no loss, no gradient
        else if (code0 >= 0 && pc0 >= c0 && pc0 < c0 + 4) {
            const QuadOut o = positive_fix<GRAD>(c.v0, pc0 - c0, s_cls, acc0, g0);
            acc0 = o.acc; g0 = o.g;
        }
        if (code1 == G3D_ASSIGN_IGNORE) { acc1 = 0.0f; g1 = zero4; }
        else if (code1 >= 0 && pc1 >= c0 && pc1 < c0 + 4) {
            const QuadOut o = positive_fix<GRAD>(c.v1, pc1 - c0, s_cls, acc1, g1);
            acc1 = o.acc; g1 = o.g;
        }
    }
    if (GRAD) {
        float4* dp = reinterpret_cast<float4*>(p.dcls + row0 * 8);
        st_stream(dp + lane, g0);
        st_stream(dp + 32 + lane, g1);
    }
    return acc0 + acc1;
}

// generic class count: one thread per row, scalar accesses
template <int VARIANT, bool GRAD, bool COHERENT>
__device__ __forceinline__ float stream_chunk_any(const StreamArgs& p, int b, int a0, int lane, float s_cls) {
    const int nrows = min(32, p.A - a0);
    if (nrows <= 0) return 0.0f;
    const int64_t row0 = (int64_t)b * p.A + a0;
    float acc = 0.0f;
    if (lane < nrows) {
        const int code = COHERENT ? __ldcg(p.assign + row0 + lane) : __ldg(p.assign + row0 + lane);
        const int cls_col = (VARIANT == G3D_VARIANT_3D) ? 20 : 4;
        int pos_cls = -1;
        if (code >= 0) pos_cls = (int)(long long)p.ann[((int64_t)b * p.Gmax + code) * p.W + cls_col];
        const int C = p.C;
        const float* cp = p.cls + (row0 + lane) * C;
        float* dp = p.dcls + (row0 + lane) * C;
        const bool ign = (code == G3D_ASSIGN_IGNORE);
#pragma unroll 1
        for (int c = 0; c < C; ++c) {
            const float pr = __ldg(cp + c);
            if (!ign) acc += focal_term(pr, c == pos_cls);
            if (GRAD) dp[c] = ign ? 0.0f : s_cls * focal_term_grad(pr, c == pos_cls);
        }
    }
    return acc;
}

struct StreamSmem {
    double dred[kWarps];
    int arrive;
};

// One work item of the streaming pass: kWarps * 32 * cpw consecutive rows of image b, cpw chunks of 32 rows per warp.
// Returns this warp's share of the focal sum (valid in every lane).
template <int VARIANT, int CS, bool GRAD, bool COHERENT>
__device__ __forceinline__ float stream_item(const StreamArgs& p, int tile, int b) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float npos = (float)(COHERENT ? __ldcg(p.npos + b) : __ldg(p.npos + b));
    const float s_cls = GRAD ? p.g0 / ((float)p.B * fmaxf(npos, 1.0f)) : 0.0f;
    const int wa0 = (tile * kWarps + warp) * (32 * p.cpw);   // first anchor of this warp
    float cls_acc = 0.0f;
    int c = 0;
    if (CS == 8) {
        // full 32-row chunks: software pipeline, the loads of chunk c + 1 are in flight while chunk c is evaluated
        const int nfull = max(0, min(p.cpw, (p.A - wa0) >> 5));
        if (nfull > 0) {
            int64_t row0 = (int64_t)b * p.A + wa0;
            Chunk8 cur = load_chunk8<COHERENT>(p, row0, lane);
#pragma unroll 1
            for (; c < nfull; ++c, row0 += 32) {
                Chunk8 nxt = cur;
                if (c + 1 < nfull) nxt = load_chunk8<COHERENT>(p, row0 + 32, lane);
                cls_acc += process_chunk8<VARIANT, GRAD>(p, b, row0, lane, s_cls, cur);
                cur = nxt;
            }
        }
    }
    // generic class count, and the ragged last chunk of an image: one thread per row
#pragma unroll 1
    for (; c < p.cpw; ++c)
        cls_acc += stream_chunk_any<VARIANT, GRAD, COHERENT>(p, b, wa0 + 32 * c, lane, s_cls);
    return warp_sum_f(cls_acc);   // FP32 inside the warp (<= 1024 terms), FP64 from here on
}

template <int VARIANT, int CS, bool GRAD>
__global__ void __launch_bounds__(kTile, 5) focal_stream_kernel(const StreamArgs p) {
    __shared__ StreamSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    if (tid == 0) sm.arrive = 0;
    __syncthreads();   // the ticket must be zero before the first warp finishes (all warps are still at the start: cheap)
    const float cs = stream_item<VARIANT, CS, GRAD, false>(p, blockIdx.x, b);
    // ---- the last warp of the CTA to get here (shared-memory ticket, no block barrier: finished warps retire
    // immediately) combines the 8 warp partials in warp order; the last CTA of the image (global ticket) reduces it.
    int arrived = 0;
    if (lane == 0) {
        sm.dred[warp] = (double)cs;
        __threadfence_block();
        arrived = atomicAdd(&sm.arrive, 1);
    }
    arrived = __shfl_sync(0xffffffffu, arrived, 0);
    if (arrived != kWarps - 1) return;
    __threadfence_block();
    int is_last = 0;
    if (lane == 0) {
        const volatile double* dr = sm.dred;
        double tc = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) tc += dr[w];
        __stcg(p.partials + (int64_t)b * p.T + blockIdx.x, tc);
        __threadfence();
        is_last = (atomicAdd(p.counters + b, 1) == p.T - 1);
    }
    is_last = __shfl_sync(0xffffffffu, is_last, 0);
    if (is_last) finalize_image<VARIANT>(p, b);
}

// =====================================================================================================================
// fused launch (C == 8): assignment and streaming pass in ONE persistent kernel
// =====================================================================================================================
// The assignment is issue-bound and leaves HBM almost idle; the streaming pass is HBM / latency bound and leaves issue
// slots idle.  Run back to back they cost the sum; here persistent CTAs (as many as fit on the GPU) pull work items
// from two queues - assignment items (tile x image group, group-major) and streaming items (row tile x image,
// image-major) - so both kinds are resident on every SM at once and the issue-bound work fills the stalls of the
// memory-bound work.  A streaming item of image b may only start when every assignment item of b's group has finished
// (its normaliser num_pos and its codes are complete): `group_done` counts them; writers publish with
// __threadfence + atomicAdd, readers observe the count, fence, and read the codes through L2 (ld.cg).
// CTAs in odd / even launch slots prefer different queues so that the mix is there from the start; a CTA whose
// preferred queue is empty (or not ready) takes from the other; only when no assignment item is left does a CTA wait
// for the streaming queue - the items it waits for are held by running CTAs, so the wait always ends (no co-residency
// assumption).  Per-item partial sums go to fixed slots; loss_finalize_kernel reduces them in fixed order.
struct FusedArgs {
    AssignCodesArgs q;
    StreamArgs p;
    int32_t* ctr;             // work counters, zero on entry, one per 128-byte line (kCtrPitch ints apart):
                              //   [0] next assignment item, [1 + g] next streaming item of image group g,
                              //   [1 + n_groups + g] finished assignment items of group g
    int n_tiles_assign;       // ceil(A / kTile)
    int n_groups;             // ceil(B / kImgPerCta)
    int n_assign;             // assignment items = n_tiles_assign * n_groups
    int sms;                  // SM count (consecutive CTA indices land on different SMs)
    int mix;                  // 0: every CTA prefers assignment items; k > 0: launch slots with (slot % k) == k-1 prefer streaming
};

__device__ __forceinline__ int ld_volatile(const int32_t* p) { return *reinterpret_cast<const volatile int32_t*>(p); }

constexpr int kCtrPitch = 32;
__device__ __forceinline__ int32_t* ctr_assign_next(const FusedArgs& f) { return f.ctr; }
__device__ __forceinline__ int32_t* ctr_stream_next(const FusedArgs& f, int g) { return f.ctr + kCtrPitch * (1 + g); }
__device__ __forceinline__ int32_t* ctr_group_done(const FusedArgs& f, int g) { return f.ctr + kCtrPitch * (1 + f.n_groups + g); }

// thread 0 of a CTA: try to take a streaming item of the first image group that is completely assigned and still has
// items (g_lo: first group not known to be exhausted).  Tickets are taken with atomicAdd per group - an over-claim on
// an exhausted group is harmless, nobody ever holds an item that is not ready, nothing retries under contention.
// Returns 0 with `item` set, 1 if the next group with items is still being assigned, 2 if every group is exhausted.
__device__ __forceinline__ int claim_stream_item(const FusedArgs& f, int& g_lo, int& item) {
    for (int g = g_lo; g < f.n_groups; ++g) {
        if (ld_volatile(ctr_group_done(f, g)) < f.n_tiles_assign) return 1;
        const int n_items = min(kImgPerCta, f.p.B - g * kImgPerCta) * f.p.T;
        const int idx = atomicAdd(ctr_stream_next(f, g), 1);
        if (idx < n_items) { item = (1 << 30) | (g * kImgPerCta * f.p.T + idx); return 0; }
        g_lo = g + 1;
    }
    return 2;
}

// thread 0 of a CTA, non-blocking: the next work item as (kind << 30) | index, -1 when all work is done, -2 when the only
// work left is streaming items whose image group is still being assigned (try again later).
__device__ __forceinline__ int claim_item(const FusedArgs& f, bool prefer_stream, int& g_lo, bool& assign_left) {
    int item = -2, st = 0;      // st: streaming queue 0 not looked at, 1 blocked, 2 exhausted
    if (prefer_stream || !assign_left) {
        st = claim_stream_item(f, g_lo, item);
        if (item != -2) return item;
    }
    if (assign_left) {
        const int a = atomicAdd(ctr_assign_next(f), 1);
        if (a < f.n_assign) return a;
        assign_left = false;
    }
    if (st == 0) {
        st = claim_stream_item(f, g_lo, item);
        if (item != -2) return item;
    }
    return st == 2 ? -1 : -2;
}

template <int VARIANT, bool GRAD>
__global__ void __launch_bounds__(kTile, 5) focal_fused_kernel(const FusedArgs f) {
    __shared__ StageSmem sm;
    __shared__ double s_dred[kWarps];
    __shared__ int s_item[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool prefer_stream = f.mix > 0 && ((int)(blockIdx.x / f.sms) % f.mix) == f.mix - 1;
    // thread 0 keeps the scheduling state and always holds the NEXT item (claimed while the current one is processed,
    // so the L2 round trips of the work queue are off the critical path)
    int g_lo = 0, next = -2;
    bool assign_left = true;
    if (tid == 0) next = claim_item(f, prefer_stream, g_lo, assign_left);
    for (int it = 0;; ++it) {
        if (tid == 0) {
            int item = next;
            while (item == -2) {                 // streaming items exist but their image group is still being assigned
                __nanosleep(500);
                item = claim_item(f, prefer_stream, g_lo, assign_left);
            }
            if (item >= (1 << 30)) __threadfence();   // acquire: the group's codes / num_pos were published before the count
            s_item[it & 1] = item;
        }
        __syncthreads();
        const int item = s_item[it & 1];
        if (item == -1) break;
        if (tid == 0) next = claim_item(f, prefer_stream, g_lo, assign_left);
        if (item >> 30) {
            const int s = item & ((1 << 30) - 1);
            const int b = s / f.p.T, tile = s - b * f.p.T;
            const float cs = stream_item<VARIANT, 8, GRAD, true>(f.p, tile, b);   // codes / num_pos through L2 (ld.cg)
            if (lane == 0) s_dred[warp] = (double)cs;
            __syncthreads();
            if (tid == 0) {
                double tc = 0.0;
#pragma unroll
                for (int w = 0; w < kWarps; ++w) tc += s_dred[w];
                __stcg(f.p.partials + (int64_t)b * f.p.T + tile, tc);
            }
        } else {
            const int group = item / f.n_tiles_assign, tile = item - group * f.n_tiles_assign;
            assign_item(f.q, sm, tile, group);
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                atomicAdd(ctr_group_done(f, group), 1);
            }
        }
    }
}

// one warp per image: per-image losses from the partial sums, then (last image) the batch means
template <int VARIANT>
__global__ void __launch_bounds__(32) loss_finalize_kernel(const StreamArgs p) { finalize_image<VARIANT>(p, blockIdx.x); }

// Fused path: the positives launch also does the reduction - the last CTA of an image (per-image ticket) finalises that
// image, the last image the batch - so the forward is gt_prepare, focal_fused_kernel and this.
template <int VARIANT>
__global__ void __launch_bounds__(128) positives_finalize_kernel(const PosArgs q, const StreamArgs p) {
    positives_body<VARIANT, false>(q);
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(p.counters + blockIdx.y, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) finalize_image<VARIANT>(p, blockIdx.y);
}

// =====================================================================================================================
// backward, part 1: the classification gradient for upstream gradients other than the one launch 3 was told to expect
// =====================================================================================================================
struct ClsGradArgs {
    const float* cls;
    const float* ann;
    const float* grad_out;   // [3] device
    const float* grad_scale; // [3] device or null
    const int32_t* npos;     // [B]
    const int32_t* assign;
    float* dcls;
    float* dreg;
    float e0;                // upstream classification gradient dcls was formed for (valid if have_dcls)
    int have_dcls;           // dcls already holds the gradient for e0 and dreg is already zero-filled
    int B, A, C, R, Gmax, W, T;
};

// Persistent grid-stride kernel over (image, 256-row tile) items, one thread per row.  The usual training step
// (dcls already right) exits at once: one wave of CTAs.
template <int VARIANT, int CS>
__global__ void __launch_bounds__(256, 4) focal_cls_grad_kernel(const ClsGradArgs p) {
    const float go0 = __ldg(p.grad_out + 0) * (p.grad_scale ? __ldg(p.grad_scale + 0) : 1.0f);
    if (p.have_dcls && go0 == p.e0) return;
    const int lane = threadIdx.x & 31;
    const int C = (CS > 0) ? CS : p.C;
    const int cls_col = (VARIANT == G3D_VARIANT_3D) ? 20 : 4;
    const int64_t items = (int64_t)p.T * p.B;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int b = (int)(item / p.T);
        const int a = (int)(item - (int64_t)b * p.T) * 256 + threadIdx.x;
        const bool valid = a < p.A;
        const int64_t row = (int64_t)b * p.A + a;
        if (!p.have_dcls) {
            const int nrows = min(32, p.A - (a - lane));
            if (nrows > 0) {
                if (VARIANT == G3D_VARIANT_3D) zero_rows<12>(p.dreg + (row - lane) * 12, nrows, lane);
                else                           zero_rows<4>(p.dreg + (row - lane) * 4, nrows, lane);
            }
        }
        if (!valid) continue;
        const int code = __ldg(p.assign + row);
        const float npos = (float)__ldg(p.npos + b);
        int pos_cls = -1;
        if (code >= 0) pos_cls = (int)(long long)p.ann[((int64_t)b * p.Gmax + code) * p.W + cls_col];
        const float s_cls = go0 / ((float)p.B * fmaxf(npos, 1.0f));
        if (CS == 8) {
            const float4* cp = reinterpret_cast<const float4*>(p.cls + row * 8);
            const float4 c0 = ld_stream(cp), c1 = ld_stream(cp + 1);
            const float pv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
            float g[8];
            if (code == G3D_ASSIGN_NEGATIVE) {
                focal_neg<8, true>(pv, s_cls, g);
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    g[c] = (code == G3D_ASSIGN_IGNORE) ? 0.0f : s_cls * focal_term_grad(pv[c], c == pos_cls);
            }
            float4* dp = reinterpret_cast<float4*>(p.dcls + row * 8);
            st_stream(dp, make_float4(g[0], g[1], g[2], g[3]));
            st_stream(dp + 1, make_float4(g[4], g[5], g[6], g[7]));
        } else {
            const float* cp = p.cls + row * C;
            float* dp = p.dcls + row * C;
            for (int c = 0; c < C; ++c)
                dp[c] = (code == G3D_ASSIGN_IGNORE) ? 0.0f : s_cls * focal_term_grad(__ldg(cp + c), c == pos_cls);
        }
    }
}

struct FocalWorkspace {
    float4* gt_box;
    int32_t* gt_row;
    int32_t* gt_count;
    double* partials;    // [B][T]
    int32_t* pos_list;   // [B][A]
    uint32_t* keys;      // [B][A]  GT-centric assignment: best (IoU, GT index) key per anchor
    int32_t* counters;   // zeroed per call: [B] image tickets, [1] batch ticket, [B] npos, [B] nonfinite flags, then
                         // (8-byte aligned) [B][4] int64 fixed-point sums, then the fused kernel's work counters
    int64_t n_counters;  // number of int32 words to zero
    int64_t bytes;
};

static FocalWorkspace carve(void* base, int64_t B, int64_t A, int64_t Gmax) {
    FocalWorkspace w;
    const int64_t T = ceil_div(A, kRowsPerCta);
    int64_t off = 0;
    char* p = (char*)base;
    w.gt_box = (float4*)(p + off);   off += align_up(B * Gmax * 16, 256);
    w.gt_row = (int32_t*)(p + off);  off += align_up(B * Gmax * 4, 256);
    w.gt_count = (int32_t*)(p + off); off += align_up(B * 4, 256);
    w.partials = (double*)(p + off); off += align_up(B * T * 8, 256);
    w.pos_list = (int32_t*)(p + off); off += align_up(B * A * 4, 256);
    w.keys = (uint32_t*)(p + off); off += align_up(B * A * 4, 256);
    w.counters = (int32_t*)(p + off);
    const int64_t head = align_up(3 * B + 1, 2);          // int32 words before the int64 sums
    w.n_counters = head + 8 * B + 32 * (2 * ceil_div(B, 4) + 1);   // ... then the fused kernel's counters (kCtrPitch apart)
    off += align_up(w.n_counters * 4, 256);
    w.bytes = off;
    return w;
}
static inline int32_t* ws_npos(const FocalWorkspace& w, int64_t B) { return w.counters + B + 1; }
static inline int32_t* ws_nonfinite(const FocalWorkspace& w, int64_t B) { return w.counters + 2 * B + 1; }
static inline long long* ws_acc(const FocalWorkspace& w, int64_t B) { return (long long*)(w.counters + align_up(3 * B + 1, 2)); }
static inline int32_t* ws_fused(const FocalWorkspace& w, int64_t B) { return w.counters + align_up(3 * B + 1, 2) + 8 * B; }

// multi-GPU: the [world][5] shard statistics (all-gathered) -> global batch means and this rank's gradient scales.
// Fixed (rank) summation order: the same bits on every rank and from run to run.
__global__ void combine_shard_stats_kernel(const double* __restrict__ gathered, int world, int rank,
                                           float* __restrict__ losses, float* __restrict__ scale) {
    if (threadIdx.x != 0) return;
    double t[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int r = 0; r < world; ++r)
        for (int k = 0; k < 5; ++k) t[k] += gathered[r * 5 + k];
    losses[0] = (float)(t[0] / t[3]);
    losses[1] = (float)(t[1] / t[3]);
    losses[2] = (float)(t[2] / t[4]);                       // 0/0 -> NaN when no image of the global batch has GT
    const double bl = gathered[rank * 5 + 3], nl = gathered[rank * 5 + 4];
    // d(global mean) / d(local mean): B_l / B_g for cls and reg, NE_l / NE_g for vp (0 when the shard has no GT at all)
    scale[0] = scale[1] = (float)(bl / t[3]);
    scale[2] = nl > 0.0 ? (float)(nl / t[4]) : 0.0f;
}

}  // namespace g3d

using namespace g3d;

extern "C" int g3d_combine_shard_stats(const double* gathered, int64_t world, int64_t rank, float* losses, float* scale,
                                       int device, void* stream) {
    G3D_REQUIRE(gathered && losses && scale, "null pointer");
    G3D_REQUIRE(world >= 1 && rank >= 0 && rank < world && world < (1 << 20), "bad world / rank");
    G3D_GUARD(device);
    combine_shard_stats_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(gathered, (int)world, (int)rank, losses, scale);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int64_t g3d_focal_workspace_bytes(int64_t B, int64_t A, int64_t Gmax) {
    if (B < 0 || A < 0 || Gmax < 0) return G3D_ERR_INVALID;
    return carve(nullptr, B, A, Gmax).bytes;
}

static int check_focal_shapes(int64_t B, int64_t A, int64_t C, int64_t R, int64_t Gmax, int64_t W, int variant) {
    G3D_REQUIRE(variant == G3D_VARIANT_2D || variant == G3D_VARIANT_3D, "unknown variant");
    G3D_REQUIRE(B >= 1 && A >= 1 && C >= 1 && Gmax >= 0, "sizes must be positive");
    G3D_REQUIRE(B <= 65535 && A < ((int64_t)1 << 31) - kRowsPerCta && Gmax < (1 << 30) && C < (1 << 20), "size out of range");
    if (variant == G3D_VARIANT_3D) {
        G3D_REQUIRE(R == 12, "3D variant needs 12 regression outputs per anchor");
        G3D_REQUIRE(W >= 21, "3D variant needs >= 21 annotation columns");
    } else {
        G3D_REQUIRE(R == 4, "2D variant needs 4 regression outputs per anchor");
        G3D_REQUIRE(W >= 5, "2D variant needs >= 5 annotation columns");
    }
    return G3D_OK;
}

template <int VARIANT, int CS>
static void launch_stream(const StreamArgs& p, bool grad, dim3 grid, cudaStream_t st) {
    if (grad) focal_stream_kernel<VARIANT, CS, true><<<grid, kTile, 0, st>>>(p);
    else      focal_stream_kernel<VARIANT, CS, false><<<grid, kTile, 0, st>>>(p);
}

static dim3 positives_grid(int64_t B) { return dim3(64, (unsigned)B); }

// pyramid_host (nullable): {L, S, then L x (rows, cols, stride), then L x S x (anchor width, anchor height)} as doubles -
// the structure of Anchors.forward's table (anchors.py:21-40).  Accepted only if it accounts for exactly A anchors.
static bool load_pyramid(const double* h, int64_t A, Pyramid& pyr) {
    if (!h) return false;
    const int L = (int)h[0], S = (int)h[1];
    if (L < 1 || L > kPyrLevels || S < 1 || S > kPyrShapes) return false;
    pyr.L = L; pyr.S = S;
    int64_t first = 0;
    for (int l = 0; l < L; ++l) {
        const double rows = h[2 + 3 * l], cols = h[3 + 3 * l], stride = h[4 + 3 * l];
        if (!(rows >= 0 && cols >= 0 && stride > 0) || rows * cols * S > 2e9) return false;
        pyr.first[l] = (int)first; pyr.rows[l] = (int)rows; pyr.cols[l] = (int)cols; pyr.inv_stride[l] = (float)(1.0 / stride);
        first += (int64_t)rows * (int64_t)cols * S;
        for (int s = 0; s < S; ++s) {
            pyr.aw[l * S + s] = (float)h[2 + 3 * L + 2 * (l * S + s)];
            pyr.ah[l * S + s] = (float)h[3 + 3 * L + 2 * (l * S + s)];
            if (!(pyr.aw[l * S + s] > 0 && pyr.ah[l * S + s] > 0)) return false;
        }
    }
    return first == A;
}

// G3D_LOSS_FUSED=1 in the environment selects the experimental single persistent kernel for the assignment and the
// streaming pass (read per call: no state).  Default: two plain launches - measured equal or slightly faster (both
// kinds of work sit at ~60 % issue-slot utilisation, limited by latency, so sharing an SM buys nothing) and simpler.
static bool fused_path_enabled() {
    const char* e = getenv("G3D_LOSS_FUSED");
    return e && e[0] == '1';
}

extern "C" int g3d_focal_loss_fwd_bwd(const float* cls, const float* reg, const float* anchors, const float* ann,
                                      int64_t B, int64_t A, int64_t C, int64_t R, int64_t Gmax, int64_t W, int variant,
                                      float grad_cls_expected, float* losses, float* per_image, int32_t* assign,
                                      int32_t* gt_count_out, double* shard_stats, float* dcls, float* dreg,
                                      void* workspace, int64_t workspace_bytes, const double* pyramid_host,
                                      void* const* trace_events, int device, void* stream) {
    int rc = check_focal_shapes(B, A, C, R, Gmax, W, variant);
    if (rc != G3D_OK) return rc;
    G3D_REQUIRE(cls && reg && anchors && losses && per_image && assign && workspace, "null pointer");
    G3D_REQUIRE(Gmax == 0 || ann, "null annotations");
    G3D_REQUIRE((dcls == nullptr) == (dreg == nullptr), "dcls and dreg must both be given or both be null");
    FocalWorkspace w = carve(workspace, B, A, Gmax);
    G3D_REQUIRE(workspace_bytes >= w.bytes, "workspace too small (see g3d_focal_workspace_bytes)");
    G3D_REQUIRE(((uintptr_t)cls % 16) == 0 && ((uintptr_t)anchors % 16) == 0 && ((uintptr_t)workspace % 256) == 0 &&
                    ((uintptr_t)per_image % 16) == 0 && ((uintptr_t)dcls % 16) == 0 && ((uintptr_t)dreg % 16) == 0,
                "cls/dcls/dreg/anchors/per_image must be 16-byte aligned and the workspace 256-byte aligned");
    G3D_GUARD(device);
    cudaStream_t st = (cudaStream_t)stream;
    // launch 0: GT prologue; the same kernel zeroes the tickets, counters and fixed-point sums
    rc = gt_prepare_launch(ann, B, Gmax, W, variant, (float*)w.gt_box, w.gt_row, w.gt_count, w.counters, w.n_counters,
                           device, stream);
    if (rc != G3D_OK) return rc;
    int32_t* npos = ws_npos(w, B);

    AssignCodesArgs q;
    q.anchors = (const float4*)anchors; q.gt_box = w.gt_box; q.gt_row = w.gt_row; q.gt_count = w.gt_count;
    q.assign = assign; q.npos = npos; q.pos_list = w.pos_list; q.dreg = dreg; q.B = (int)B; q.A = (int)A; q.Gmax = (int)Gmax;
    q.R = (int)R;
    PosArgs pp;
    pp.reg = reg; pp.anchors = (const float4*)anchors; pp.ann = ann; pp.assign = assign; pp.pos_list = w.pos_list;
    pp.npos = npos; pp.acc = ws_acc(w, B); pp.nonfinite = ws_nonfinite(w, B); pp.grad_out = nullptr; pp.grad_scale = nullptr;
    pp.losses = nullptr;
    pp.dreg = nullptr; pp.B = (int)B; pp.A = (int)A; pp.R = (int)R; pp.Gmax = (int)Gmax; pp.W = (int)W;
    StreamArgs p;
    p.cls = cls; p.ann = ann; p.assign = assign; p.npos = npos; p.gt_count = w.gt_count;
    p.acc = ws_acc(w, B); p.nonfinite = ws_nonfinite(w, B); p.gt_count_out = gt_count_out;
    p.partials = w.partials; p.counters = w.counters; p.losses = losses; p.per_image = per_image;
    p.shard_stats = shard_stats;
    p.dcls = dcls; p.dreg = dreg; p.g0 = grad_cls_expected;
    p.B = (int)B; p.A = (int)A; p.C = (int)C; p.R = (int)R; p.Gmax = (int)Gmax; p.W = (int)W;
    p.cpw = kChunksPerWarp;
    p.T = (int)ceil_div(A, kRowsPerCta);
    const bool grad = dcls != nullptr;
    const dim3 agrid((unsigned)ceil_div(A, kTile), (unsigned)ceil_div(B, kImgPerCta));

    if (trace_events) G3D_CUDA(cudaEventRecord((cudaEvent_t)trace_events[0], st));
    if (C == 8 && fused_path_enabled()) {
        // ---- one persistent kernel for assignment + streaming pass, then the positives, then the reduction
        FusedArgs f;
        p.cpw = kChunksPerWarp;
        p.T = (int)ceil_div(A, kRowsPerCta);
        f.q = q; f.p = p;
        f.ctr = ws_fused(w, B);
        f.n_tiles_assign = (int)agrid.x;
        f.n_groups = (int)agrid.y;
        f.n_assign = (int)(agrid.x * agrid.y);
        int sms = 148, per_sm = 1;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        f.sms = sms;
        { const char* e = getenv("G3D_FUSED_MIX"); f.mix = e ? atoi(e) : 3; }
        const void* kern = nullptr;
        if (variant == G3D_VARIANT_3D) kern = grad ? (const void*)focal_fused_kernel<G3D_VARIANT_3D, true> : (const void*)focal_fused_kernel<G3D_VARIANT_3D, false>;
        else                           kern = grad ? (const void*)focal_fused_kernel<G3D_VARIANT_2D, true> : (const void*)focal_fused_kernel<G3D_VARIANT_2D, false>;
        G3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTile, 0));
        int64_t nctas = (int64_t)sms * (per_sm > 0 ? per_sm : 1);
        if (nctas > (int64_t)f.n_assign + p.T * B) nctas = (int64_t)f.n_assign + p.T * B;
        void* kargs[] = {(void*)&f};
        G3D_CUDA(cudaLaunchKernel(kern, dim3((unsigned)nctas), dim3(kTile), kargs, 0, st));
        if (trace_events) G3D_CUDA(cudaEventRecord((cudaEvent_t)trace_events[1], st));
        if (variant == G3D_VARIANT_3D) positives_finalize_kernel<G3D_VARIANT_3D><<<positives_grid(B), 128, 0, st>>>(pp, p);
        else                           positives_finalize_kernel<G3D_VARIANT_2D><<<positives_grid(B), 128, 0, st>>>(pp, p);
        G3D_LAUNCH_CHECK();
        if (trace_events) {
            G3D_CUDA(cudaEventRecord((cudaEvent_t)trace_events[2], st));
            G3D_CUDA(cudaEventRecord((cudaEvent_t)trace_events[3], st));
        }
        return G3D_OK;
    }
    // ---- separate launches (the default)
    Pyramid pyr;
    if (Gmax >= 1 && Gmax <= 256 && load_pyramid(pyramid_host, A, pyr)) {
        // anchors are the regular pyramid: fill, then the few (anchor, GT) pairs that can matter, then their codes
        const long long n_rows = (long long)B * A;
        if (dreg) {
            assign_fill_kernel<<<148 * 8, 256, 0, st>>>(nullptr, nullptr, dreg, n_rows, (int)R);
            G3D_LAUNCH_CHECK();
        }
        assign_fill_kernel<<<148 * 4, 256, 0, st>>>(assign, w.keys, nullptr, n_rows, (int)R);
        G3D_LAUNCH_CHECK();
        PairArgs pa;
        pa.anchors = (const float4*)anchors; pa.gt_box = w.gt_box; pa.gt_count = w.gt_count; pa.keys = w.keys;
        pa.B = (int)B; pa.A = (int)A; pa.Gmax = (int)Gmax;
        assign_pairs_kernel<<<(unsigned)ceil_div(B * Gmax, 8), 256, 0, st>>>(pa, pyr);
        G3D_LAUNCH_CHECK();
        ResolveArgs ra;
        ra.keys = w.keys; ra.gt_row = w.gt_row; ra.assign = assign; ra.pos_list = w.pos_list; ra.npos = npos;
        ra.A = (int)A; ra.Gmax = (int)Gmax;
        assign_resolve_kernel<<<dim3((unsigned)ceil_div(A, 256 * 4 * 4), (unsigned)B), 256, 0, st>>>(ra);
        G3D_LAUNCH_CHECK();
    } else {
        assign_codes_kernel<<<agrid, kTile, 0, st>>>(q);
        G3D_LAUNCH_CHECK();
    }
    if (trace_events) G3D_CUDA(cudaEventRecord((cudaEvent_t)trace_events[1], st));
    if (variant == G3D_VARIANT_3D) positives_kernel<G3D_VARIANT_3D, false><<<positives_grid(B), 128, 0, st>>>(pp);
    else                           positives_kernel<G3D_VARIANT_2D, false><<<positives_grid(B), 128, 0, st>>>(pp);
    G3D_LAUNCH_CHECK();
    if (trace_events) G3D_CUDA(cudaEventRecord((cudaEvent_t)trace_events[2], st));
    const dim3 grid((unsigned)p.T, (unsigned)B);
    if (variant == G3D_VARIANT_3D) {
        if (C == 8) launch_stream<G3D_VARIANT_3D, 8>(p, grad, grid, st);
        else        launch_stream<G3D_VARIANT_3D, 0>(p, grad, grid, st);
    } else {
        if (C == 8) launch_stream<G3D_VARIANT_2D, 8>(p, grad, grid, st);
        else        launch_stream<G3D_VARIANT_2D, 0>(p, grad, grid, st);
    }
    G3D_LAUNCH_CHECK();
    if (trace_events) G3D_CUDA(cudaEventRecord((cudaEvent_t)trace_events[3], st));
    return G3D_OK;
}

extern "C" int g3d_focal_loss_fwd(const float* cls, const float* reg, const float* anchors, const float* ann,
                                  int64_t B, int64_t A, int64_t C, int64_t R, int64_t Gmax, int64_t W, int variant,
                                  float* losses, float* per_image, int32_t* assign, int32_t* gt_count_out,
                                  void* workspace, int64_t workspace_bytes, const double* pyramid_host, int device,
                                  void* stream) {
    return g3d_focal_loss_fwd_bwd(cls, reg, anchors, ann, B, A, C, R, Gmax, W, variant, 0.0f, losses, per_image,
                                  assign, gt_count_out, nullptr, nullptr, nullptr, workspace, workspace_bytes, pyramid_host,
                                  nullptr, device, stream);
}

extern "C" int g3d_focal_loss_bwd(const float* cls, const float* reg, const float* anchors, const float* ann,
                                  int64_t B, int64_t A, int64_t C, int64_t R, int64_t Gmax, int64_t W, int variant,
                                  const float* grad_out, const float* grad_scale, int have_dcls, float grad_cls_expected,
                                  const float* losses,
                                  const int32_t* assign, const void* workspace, int64_t workspace_bytes, float* dcls,
                                  float* dreg, int device, void* stream) {
    int rc = check_focal_shapes(B, A, C, R, Gmax, W, variant);
    if (rc != G3D_OK) return rc;
    G3D_REQUIRE(cls && reg && anchors && grad_out && losses && assign && workspace && dcls && dreg, "null pointer");
    G3D_REQUIRE(Gmax == 0 || ann, "null annotations");
    G3D_REQUIRE(((uintptr_t)cls % 16) == 0 && ((uintptr_t)dcls % 16) == 0 && ((uintptr_t)dreg % 16) == 0 &&
                    ((uintptr_t)anchors % 16) == 0 && ((uintptr_t)workspace % 256) == 0,
                "cls/dcls/dreg/anchors must be 16-byte aligned and the workspace 256-byte aligned");
    FocalWorkspace w = carve(const_cast<void*>(workspace), B, A, Gmax);
    G3D_REQUIRE(workspace_bytes >= w.bytes, "workspace too small (pass the forward's workspace, untouched)");
    G3D_GUARD(device);
    cudaStream_t st = (cudaStream_t)stream;
    ClsGradArgs p;
    p.cls = cls; p.ann = ann; p.grad_out = grad_out; p.grad_scale = grad_scale; p.npos = ws_npos(w, B); p.assign = assign;
    p.dcls = dcls; p.dreg = dreg; p.e0 = grad_cls_expected; p.have_dcls = have_dcls ? 1 : 0;
    p.B = (int)B; p.A = (int)A; p.C = (int)C; p.R = (int)R; p.Gmax = (int)Gmax; p.W = (int)W;
    p.T = (int)ceil_div(A, 256);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int64_t items = (int64_t)p.T * B;
    const int grid = (int)(items < (int64_t)sms * 8 ? items : (int64_t)sms * 8);
    if (variant == G3D_VARIANT_3D) {
        if (C == 8) focal_cls_grad_kernel<G3D_VARIANT_3D, 8><<<grid, 256, 0, st>>>(p);
        else        focal_cls_grad_kernel<G3D_VARIANT_3D, 0><<<grid, 256, 0, st>>>(p);
    } else {
        if (C == 8) focal_cls_grad_kernel<G3D_VARIANT_2D, 8><<<grid, 256, 0, st>>>(p);
        else        focal_cls_grad_kernel<G3D_VARIANT_2D, 0><<<grid, 256, 0, st>>>(p);
    }
    G3D_LAUNCH_CHECK();
    PosArgs pp;
    pp.reg = reg; pp.anchors = (const float4*)anchors; pp.ann = ann; pp.assign = assign; pp.pos_list = w.pos_list;
    pp.npos = ws_npos(w, B); pp.acc = nullptr; pp.nonfinite = nullptr; pp.grad_out = grad_out; pp.grad_scale = grad_scale;
    pp.losses = losses;
    pp.dreg = dreg; pp.B = (int)B; pp.A = (int)A; pp.R = (int)R; pp.Gmax = (int)Gmax; pp.W = (int)W;
    if (variant == G3D_VARIANT_3D) positives_kernel<G3D_VARIANT_3D, true><<<positives_grid(B), 128, 0, st>>>(pp);
    else                           positives_kernel<G3D_VARIANT_2D, true><<<positives_grid(B), 128, 0, st>>>(pp);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}
