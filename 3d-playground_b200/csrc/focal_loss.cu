// focal_loss.cu — a2..a6 fused: FocalLoss.forward / backward of both retinanet copies in one pass over the anchors.
//
//   3D: pytorch_retinanet_detector_directional/retinanet/losses.py:27-362
//   2D: retinanet/losses.py:27-177
//
// Forward, one launch: grid (anchor tiles, image groups).  A CTA loads its 256 anchors once, then for each image of
// its group: culls/stages the GT boxes (assign_tile.cuh), finds IoU max/argmax per anchor, classifies the anchor
// (negative / ignore / positive), evaluates the C focal terms from the streamed classification row, evaluates the
// corner smooth-L1 and the three direction-cosine terms for the (rare) positives, block-reduces in FP64 and stores
// one partial per (image, tile).  The last CTA to finish an image (atomic ticket) reduces that image's partials in a
// fixed order, and the last image finalises the batch means - so the result is deterministic and needs no second
// launch and no host synchronisation.
#include "assign_tile.cuh"

namespace g3d {

constexpr int kImgPerCta = 4;  // images processed per CTA for one anchor tile (anchors/bbox loaded once)

// clamp bounds of losses.py:56: torch.clamp(classification, 1e-4, 1.0 - 1e-4) - python doubles cast to f32
#define G3D_PMIN ((float)1e-4)
#define G3D_PMAX ((float)(1.0 - 1e-4))
#define G3D_BETA ((float)(1.0 / 9.0))        // smooth-L1 switch point (losses.py:346)
#define G3D_HALF_BETA ((float)(0.5 / 9.0))   // losses.py:348

struct FocalArgs {
    const float* cls;
    const float* reg;
    const float4* anchors;
    const float* ann;
    const float4* gt_box;
    const int32_t* gt_row;
    const int32_t* gt_count;
    double* partials;    // [B][T][4] : cls_sum, num_pos, reg_sum, vp_sum
    int32_t* counters;   // [B+1], zero on entry
    float* losses;       // [4] : cls, reg, vp, number of non-empty images
    float* per_image;    // [B][4]
    int32_t* assign;     // [B][A] or null
    int B, A, C, R, Gmax, W, T;
};

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
// truncated after z^8 (|z| < 1/7 for pe < 0.25: relative truncation error < 4e-10).  The series is evaluated on the
// SAME rounded u the reference takes the log of, so it tracks torch.log(1.0 - classification) to ~2e-7 relative.
__device__ __forceinline__ float neg_log_u_series(float u) {   // valid for 1 - u < 0.25
    const float pe = 1.0f - u;
    const float z = __fdividef(pe, 1.0f + u);
    const float z2 = z * z;
    float s = fmaf(z2, 1.0f / 9.0f, 1.0f / 7.0f);
    s = fmaf(z2, s, 0.2f);
    s = fmaf(z2, s, 1.0f / 3.0f);
    s = fmaf(z2, s, 1.0f);
    return (z + z) * s;
}
__device__ __forceinline__ float neg_log_u(float u) {
    return (1.0f - u < 0.25f) ? neg_log_u_series(u) : -logf(u);
}

// The 8 class terms of a NEGATIVE anchor, forward (sum) and backward (per-class derivative times `scale`).
// Straight-line code - the series for all 8 classes first, so the 8 dependency chains interleave - and a rare,
// separate fix-up for probabilities >= 0.25 (which need the full logf).
__device__ __forceinline__ float focal_neg_sum8(const float* pv) {
    float p[8], nl[8];
    bool big = false;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        p[c] = fminf(fmaxf(pv[c], G3D_PMIN), G3D_PMAX);
        const float u = 1.0f - p[c];
        nl[c] = neg_log_u_series(u);
        big |= (1.0f - u >= 0.25f);
    }
    if (big) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (1.0f - (1.0f - p[c]) >= 0.25f) nl[c] = -logf(1.0f - p[c]);
    }
    float acc = 0.0f;
#pragma unroll
    for (int c = 0; c < 8; ++c) acc += (0.75f * (p[c] * p[c])) * nl[c];
    return acc;
}
__device__ __forceinline__ void focal_neg_grad8(const float* pv, float scale, float* g) {
    float nl[8];
    bool big = false;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float u = 1.0f - pv[c];
        nl[c] = neg_log_u_series(u);
        big |= !(1.0f - u < 0.25f);
    }
    if (big) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (!(1.0f - (1.0f - pv[c]) < 0.25f)) nl[c] = -logf(1.0f - pv[c]);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float p = pv[c], u = 1.0f - p;
        const float d = 1.5f * p * nl[c] + __fdividef(0.75f * (p * p), u);
        g[c] = (p >= G3D_PMIN && p <= G3D_PMAX) ? scale * d : 0.0f;   // clamp backward: zero outside [min, max]
    }
}

// focal term of a negative anchor (target 0 for every class): 0.75 p^2 * -log(1-p)
__device__ __forceinline__ float focal_term_neg(float p_raw) {
    const float p = fminf(fmaxf(p_raw, G3D_PMIN), G3D_PMAX);
    return (0.75f * (p * p)) * neg_log_u(1.0f - p);
}

// d(focal term)/dp, zero outside the clamp range (torch.clamp backward passes min <= x <= max).
// The two quotients use the 2-ulp fast division: gradients are compared at 1e-5 relative, nothing is index-critical.
__device__ __forceinline__ float focal_term_grad_neg(float p) {
    if (!(p >= G3D_PMIN && p <= G3D_PMAX)) return 0.0f;
    const float u = 1.0f - p;
    return 1.5f * p * neg_log_u(u) + __fdividef(0.75f * (p * p), u);
}
__device__ __forceinline__ float focal_term_grad(float p, bool t) {
    if (!t) return focal_term_grad_neg(p);
    if (!(p >= G3D_PMIN && p <= G3D_PMAX)) return 0.0f;
    const float u = 1.0f - p;
    return 0.5f * u * logf(p) - __fdividef(0.25f * (u * u), p);
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
    const float rn = sqrtf(rx * rx + ry * ry), tn = sqrtf(tx * tx + ty * ty);
    const float dot = rx * tx + ry * ty, den = rn * tn;
    // d(-dot/den)/dr = -(t/den) + dot * (r/rn) * tn / den^2
    const float k = dot / (den * den) * tn / rn;
    gx = -tx / den + k * rx;
    gy = -ty / den + k * ry;
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

// 3D positive anchor: sum of the 20 smooth-L1 terms and the mean of the three cosine losses (losses.py:156-350)
__device__ __noinline__ void positive_terms_3d(const float* __restrict__ rrow, const float* __restrict__ grow,
                                                  const float4& an, float& reg_sum, float& vp_term) {
    float r[12], t[20], p[20], tv[6];
#pragma unroll
    for (int i = 0; i < 12; ++i) r[i] = rrow[i];
#pragma unroll
    for (int i = 0; i < 20; ++i) t[i] = grow[i];
    gt_directions(t, tv);
    vp_term = (cos_loss(r[2], r[3], tv[0], tv[1]) + cos_loss(r[4], r[5], tv[2], tv[3]) +
               cos_loss(r[6], r[7], tv[4], tv[5])) / 3.0f;
    pred_corners(r, p);
    const float aw = an.z - an.x, ah = an.w - an.y;
    const float acx = an.x + 0.5f * aw, acy = an.y + 0.5f * ah;
    const float iaw = 1.0f / aw, iah = 1.0f / ah;   // 2 divisions instead of 20 (the loss is compared at 1e-5)
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 20; ++i) {
        const float tn = (i & 1) ? (t[i] - acy) * iah : (t[i] - acx) * iaw;
        float d = fabsf(tn - p[i]);
        if (i >= 8 && i < 16) d *= 0.5f;  // top_weighting, losses.py:343
        s += smooth_l1(d);
    }
    reg_sum = s;
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

__device__ __forceinline__ double block_sum(double v, double* s) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += s[w];
    __syncthreads();
    return t;
}

// Shared state of one CTA of the forward kernel: the culled GT lists of its kImgPerCta images (staged once, behind two
// barriers); everything after that is warp-private - no barrier per image and none at the end: the last warp of the
// CTA to finish (shared-memory ticket) combines the 8 warp partials.
struct FwdSmem {
    float4 box[kImgPerCta][kTile];
    float area[kImgPerCta][kTile];
    int idx[kImgPerCta][kTile];
    float bbred[4][kWarps];
    int wcount[kImgPerCta][kWarps];
    double dred[kImgPerCta][3][kWarps];
    int nred[kImgPerCta][kWarps];
    int total[kImgPerCta];
    int arrive;
};

// Executed by ONE warp - the last tile of image b: reduce the image's T partials in a fixed order (lane-strided
// accumulation + shuffle tree), then (last image of the batch) the batch means.
template <int VARIANT>
__device__ __forceinline__ void finalize_image(const FocalArgs& p, int b) {
    const int lane = threadIdx.x & 31;
    __threadfence();
    double tc = 0.0, tn = 0.0, tr = 0.0, tv = 0.0;
    const double2* src = reinterpret_cast<const double2*>(p.partials + (int64_t)b * p.T * 4);
#pragma unroll 4
    for (int t = lane; t < p.T; t += 32) {
        const double2 u = __ldcg(src + 2 * t), v = __ldcg(src + 2 * t + 1);
        tc += u.x; tn += u.y; tr += v.x; tv += v.y;
    }
    tc = warp_sum(tc); tn = warp_sum(tn); tr = warp_sum(tr); tv = warp_sum(tv);
    int last = 0;
    if (lane == 0) {
        const double per_pos = (VARIANT == G3D_VARIANT_3D) ? 20.0 : 4.0;
        float4 o;
        o.x = (float)(tc / fmax(tn, 1.0));                      // losses.py:152 (and :70 for empty images)
        o.y = tn > 0.0 ? (float)(tr / (tn * per_pos)) : 0.0f;  // .mean() over P x 20 (:350) / P x 4
        o.z = tn > 0.0 ? (float)(tv / tn) : 0.0f;              // vp_loss.mean() (:304)
        o.w = (float)tn;
        __stcg(reinterpret_cast<float4*>(p.per_image) + b, o);
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
        }
    }
}

template <int VARIANT, int CS>
__global__ void __launch_bounds__(kTile, 4) focal_fwd_kernel(const FocalArgs p) {
    __shared__ FwdSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int a = blockIdx.x * kTile + tid;
    const bool valid = a < p.A;
    float4 an = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) an = __ldg(p.anchors + a);
    const float area_a = box_area_rn(an.x, an.y, an.z, an.w);
    const int cls_col = (VARIANT == G3D_VARIANT_3D) ? 20 : 4;
    const int C = (CS > 0) ? CS : p.C;
    const int b0 = blockIdx.y * kImgPerCta;
    const int nimg = min(kImgPerCta, p.B - b0);
    if (tid == 0) sm.arrive = 0;

    // ---- bounding boxes of the warp's and of the tile's anchors (one barrier)
    float4 wb, bb;
    {
        float mnx = valid ? an.x : INFINITY, mny = valid ? an.y : INFINITY;
        float mxx = valid ? an.z : -INFINITY, mxy = valid ? an.w : -INFINITY;
        mnx = warp_min(mnx); mny = warp_min(mny); mxx = warp_max(mxx); mxy = warp_max(mxy);
        wb = make_float4(mnx, mny, mxx, mxy);
        if (lane == 0) { sm.bbred[0][warp] = mnx; sm.bbred[1][warp] = mny; sm.bbred[2][warp] = mxx; sm.bbred[3][warp] = mxy; }
        __syncthreads();
        bb = make_float4(sm.bbred[0][0], sm.bbred[1][0], sm.bbred[2][0], sm.bbred[3][0]);
#pragma unroll
        for (int w = 1; w < kWarps; ++w) {
            bb.x = fminf(bb.x, sm.bbred[0][w]); bb.y = fminf(bb.y, sm.bbred[1][w]);
            bb.z = fmaxf(bb.z, sm.bbred[2][w]); bb.w = fmaxf(bb.w, sm.bbred[3][w]);
        }
    }

    // ---- stage the GT boxes of the group's images: cull against the tile bounding box with an ordered compaction
    // (ascending GT index), all images behind the same two barriers.  An image with more than kTile GT rows keeps its
    // first kTile candidates here and is finished by the (rare) overflow loop further down.
    {
        const int g = tid;
        float4 gb[kImgPerCta];
        unsigned bal[kImgPerCta];
#pragma unroll
        for (int i = 0; i < kImgPerCta; ++i) {
            bool hit = false;
            gb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int Gi = (i < nimg) ? __ldg(p.gt_count + b0 + i) : 0;
            if (g < Gi) {
                gb[i] = __ldg(p.gt_box + (int64_t)(b0 + i) * p.Gmax + g);
                // keep unless provably disjoint from every anchor of the tile (NaN coordinates are never culled)
                hit = !(gb[i].z <= bb.x || gb[i].x >= bb.z || gb[i].w <= bb.y || gb[i].y >= bb.w);
            }
            bal[i] = __ballot_sync(0xffffffffu, hit);
        }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < kImgPerCta; ++i) sm.wcount[i][warp] = __popc(bal[i]);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kImgPerCta; ++i) {
            int off = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const int c = sm.wcount[i][w];
                off += (w < warp) ? c : 0;
                tot += c;
            }
            if (tid == 0) sm.total[i] = tot;
            if ((bal[i] >> lane) & 1u) {
                const int pos = off + __popc(bal[i] & ((1u << lane) - 1u));
                sm.box[i][pos] = gb[i];
                sm.area[i][pos] = box_area_rn(gb[i].x, gb[i].y, gb[i].z, gb[i].w);
                sm.idx[i][pos] = g;
            }
        }
        __syncthreads();
    }

    // ---- per image (a real loop: the body is large, unrolling it thrashes the instruction cache): IoU max/argmax over
    // the staged survivors, classify the anchor, focal terms, positive terms, warp-level partial sums.  No barriers.
#pragma unroll 1
    for (int i = 0; i < nimg; ++i) {
        const int b = b0 + i;
        const int64_t row = (int64_t)b * p.A + a;
        // request the classification row first: it is in flight during the IoU search
        float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
        if (CS == 8 && valid) {
            const float4* cp = reinterpret_cast<const float4*>(p.cls + row * 8);
            c0 = ld_stream(cp);
            c1 = ld_stream(cp + 1);
        }
        const int Gi = __ldg(p.gt_count + b);
        float best = 0.0f;
        int besti = 0;
        const int total = sm.total[i];
        for (int k0 = 0; k0 < total; k0 += 32) {
            // warp-level refinement: which of these (up to 32) tile survivors can touch this warp's anchors at all?
            bool near = false;
            if (k0 + lane < total) {
                const float4 t = sm.box[i][k0 + lane];
                near = !(t.z <= wb.x || t.x >= wb.z || t.w <= wb.y || t.y >= wb.w);
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
                    const float ua = fmaxf(__fsub_rn(__fadd_rn(area_a, sm.area[i][k]), inter), 1e-8f);
                    const float v = __fdiv_rn(inter, ua);
                    if (v > best) { best = v; besti = sm.idx[i][k]; }
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
        int code = G3D_ASSIGN_NEGATIVE;
        if (Gi > 0) code = assign_code(best, besti, p.gt_row + (int64_t)b * p.Gmax);
        if (!valid) code = G3D_ASSIGN_IGNORE;
        if (p.assign && valid) p.assign[row] = code;

        float cls_acc = 0.0f, reg_acc = 0.0f, vp_acc = 0.0f;
        const float* grow = nullptr;
        int pos_cls = -1;
        if (code >= 0) {
            grow = p.ann + ((int64_t)b * p.Gmax + code) * p.W;
            pos_cls = (int)(long long)grow[cls_col];
        }
        if (code != G3D_ASSIGN_IGNORE) {
            if (CS == 8) {
                const float pv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                if (code == G3D_ASSIGN_NEGATIVE) {  // the overwhelmingly common case: no per-class selects
                    cls_acc = focal_neg_sum8(pv);
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) cls_acc += focal_term(pv[c], c == pos_cls);
                }
            } else {
                const float* cp = p.cls + row * C;
                for (int c = 0; c < C; ++c) cls_acc += focal_term(__ldg(cp + c), c == pos_cls);
            }
        }
        if (code >= 0) {
            const float* rrow = p.reg + row * p.R;
            if (VARIANT == G3D_VARIANT_3D) {
                positive_terms_3d(rrow, grow, an, reg_acc, vp_acc);
            } else {
                float t[4];
                targets_2d(grow, an, t);
#pragma unroll
                for (int j = 0; j < 4; ++j) reg_acc += smooth_l1(fabsf(t[j] - rrow[j]));
            }
        }
        const double cs = warp_sum((double)cls_acc);
        const unsigned posmask = __ballot_sync(0xffffffffu, code >= 0);
        double rs = 0.0, vs = 0.0;
        if (posmask) {
            rs = warp_sum((double)reg_acc);
            vs = warp_sum((double)vp_acc);
        }
        if (lane == 0) {
            sm.dred[i][0][warp] = cs; sm.dred[i][1][warp] = rs; sm.dred[i][2][warp] = vs;
            sm.nred[i][warp] = __popc(posmask);
        }
    }
    // ---- the last warp of the CTA to get here combines the warp partials: one partial per (image, tile); the last
    // tile of an image (global ticket) reduces that image.  No block barrier: finished warps retire immediately.
    int arrived = 0;
    if (lane == 0) {
        __threadfence_block();
        arrived = atomicAdd(&sm.arrive, 1);
    }
    arrived = __shfl_sync(0xffffffffu, arrived, 0);
    if (arrived != kWarps - 1) return;
    __threadfence_block();
    int is_last = 0;
    if (lane < nimg) {
        const int i = lane, b = b0 + i;
        const volatile double* dr = &sm.dred[i][0][0];
        const volatile int* nr = &sm.nred[i][0];
        double tc = 0.0, tr = 0.0, tv = 0.0;
        int tn = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) { tc += dr[w]; tr += dr[kWarps + w]; tv += dr[2 * kWarps + w]; tn += nr[w]; }
        double* out = p.partials + ((int64_t)b * p.T + blockIdx.x) * 4;
        __stcg(reinterpret_cast<double2*>(out), make_double2(tc, (double)tn));
        __stcg(reinterpret_cast<double2*>(out) + 1, make_double2(tr, tv));
        __threadfence();
        is_last = (atomicAdd(p.counters + b, 1) == p.T - 1);
    }
    unsigned lastmask = __ballot_sync(0xffffffffu, is_last);
    while (lastmask) {
        const int i = __ffs(lastmask) - 1;
        lastmask &= lastmask - 1;
        finalize_image<VARIANT>(p, b0 + i);
    }
}

// ----------------------------------------------------------------------------------------------------- backward
struct FocalBwdArgs {
    const float* cls;
    const float* reg;
    const float4* anchors;
    const float* ann;
    const float* grad_out;   // [3]
    const float* per_image;  // [B][4]
    const float* losses;     // [4] (losses[3] = number of non-empty images)
    const int32_t* assign;
    float* dcls;
    float* dreg;             // pre-zeroed; only positive rows are written
    int B, A, C, R, Gmax, W;
};

// regression gradient of one positive anchor (rare path, kept out of line so that the streaming main path of the
// backward kernel stays small in registers and code)
template <int VARIANT>
__device__ __noinline__ void positive_grad(const FocalBwdArgs& p, int b, int a, int code, float npos) {
    const int64_t row = (int64_t)b * p.A + a;
    const float* grow = p.ann + ((int64_t)b * p.Gmax + code) * p.W;
    const float4 an = __ldg(p.anchors + a);
    const float* rrow = p.reg + row * p.R;
    float* drow = p.dreg + row * p.R;
    if (VARIANT == G3D_VARIANT_3D) {
        const float s_reg = __ldg(p.grad_out + 1) / ((float)p.B * 20.0f * npos);
        const float s_vp = __ldg(p.grad_out + 2) / (__ldg(p.losses + 3) * npos * 3.0f);
        float r[12], t[20], pr[20], tv[6], g[20], dr[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) r[i] = rrow[i];
#pragma unroll
        for (int i = 0; i < 20; ++i) t[i] = grow[i];
        gt_directions(t, tv);
        pred_corners(r, pr);
        const float aw = an.z - an.x, ah = an.w - an.y;
        const float acx = an.x + 0.5f * aw, acy = an.y + 0.5f * ah;
#pragma unroll
        for (int i = 0; i < 20; ++i) {
            const float tn = (i & 1) ? (t[i] - acy) / ah : (t[i] - acx) / aw;
            const float diff = tn - pr[i];
            const float w = (i >= 8 && i < 16) ? 0.5f : 1.0f;
            const float d = fabsf(diff) * w;
            const float sg = (diff > 0.0f) ? 1.0f : ((diff < 0.0f) ? -1.0f : 0.0f);
            // d smooth_l1 / d pred = slope(d) * w * d|diff|/dpred = slope * w * (-sign(diff))
            g[i] = -s_reg * ((d <= G3D_BETA) ? 9.0f * d : 1.0f) * w * sg;
        }
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
    } else {
        const float s_reg = __ldg(p.grad_out + 1) / ((float)p.B * 4.0f * npos);
        float t[4];
        targets_2d(grow, an, t);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float diff = t[i] - rrow[i];
            const float d = fabsf(diff);
            const float sg = (diff > 0.0f) ? 1.0f : ((diff < 0.0f) ? -1.0f : 0.0f);
            drow[i] = -s_reg * ((d <= G3D_BETA) ? 9.0f * d : 1.0f) * sg;
        }
    }
}

// Backward: one thread per (image, anchor).  Streaming part: assignment code + classification row in, gradient row
// out; the regression-gradient tile of the warp (32 rows) is zero-filled with coalesced 16-byte stores and the rare
// positive rows are then overwritten (ordered by __syncwarp), so dreg needs no separate memset pass.
template <int VARIANT, int CS>
__global__ void __launch_bounds__(256, 5) focal_bwd_kernel(const FocalBwdArgs p) {
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int a = blockIdx.x * 256 + threadIdx.x;
    const bool valid = a < p.A;
    const int C = (CS > 0) ? CS : p.C;
    const int64_t row = (int64_t)b * p.A + a;
    int code = G3D_ASSIGN_IGNORE;
    float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
    if (valid) {
        code = __ldg(p.assign + row);
        if (CS == 8) {
            const float4* cp = reinterpret_cast<const float4*>(p.cls + row * 8);
            c0 = ld_stream(cp);
            c1 = ld_stream(cp + 1);
        }
    }
    const float npos = __ldg(p.per_image + 4 * b + 3);
    const float s_cls = __ldg(p.grad_out + 0) / ((float)p.B * fmaxf(npos, 1.0f));
    // ---- zero-fill this warp's rows of dreg (R floats each, contiguous across the warp)
    {
        const int64_t wrow0 = (int64_t)b * p.A + (a - lane);             // first row of the warp
        const int nrows = min(32, p.A - (a - lane));
        if (nrows > 0) {
            float* base = p.dreg + wrow0 * p.R;
            const int nfloat = nrows * p.R;
            if (((uintptr_t)base & 15) == 0) {
                const int nvec = nfloat >> 2;
                for (int i = lane; i < nvec; i += 32) reinterpret_cast<float4*>(base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int i = (nvec << 2) + lane; i < nfloat; i += 32) base[i] = 0.0f;
            } else {
                for (int i = lane; i < nfloat; i += 32) base[i] = 0.0f;
            }
        }
    }
    if (valid) {
        const int cls_col = (VARIANT == G3D_VARIANT_3D) ? 20 : 4;
        int pos_cls = -1;
        if (code >= 0) pos_cls = (int)(long long)p.ann[((int64_t)b * p.Gmax + code) * p.W + cls_col];
        if (CS == 8) {
            const float pv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
            float g[8];
            if (code == G3D_ASSIGN_NEGATIVE) {
                focal_neg_grad8(pv, s_cls, g);
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
    __syncwarp();   // the zero-fill above is ordered before the positive rows written below
    if (code >= 0) positive_grad<VARIANT>(p, b, a, code, npos);
}

struct FocalWorkspace {
    float4* gt_box;
    int32_t* gt_row;
    int32_t* gt_count;
    double* partials;
    int32_t* counters;
    int64_t bytes;
};

static FocalWorkspace carve(void* base, int64_t B, int64_t A, int64_t Gmax) {
    FocalWorkspace w;
    const int64_t T = ceil_div(A, kTile);
    int64_t off = 0;
    char* p = (char*)base;
    w.gt_box = (float4*)(p + off);   off += align_up(B * Gmax * 16, 256);
    w.gt_row = (int32_t*)(p + off);  off += align_up(B * Gmax * 4, 256);
    w.gt_count = (int32_t*)(p + off); off += align_up(B * 4, 256);
    w.partials = (double*)(p + off); off += align_up(B * T * 32, 256);
    w.counters = (int32_t*)(p + off); off += align_up((B + 1) * 4, 256);
    w.bytes = off;
    return w;
}

}  // namespace g3d

using namespace g3d;

extern "C" int64_t g3d_focal_workspace_bytes(int64_t B, int64_t A, int64_t Gmax) {
    if (B < 0 || A < 0 || Gmax < 0) return G3D_ERR_INVALID;
    return carve(nullptr, B, A, Gmax).bytes;
}

static int check_focal_shapes(int64_t B, int64_t A, int64_t C, int64_t R, int64_t Gmax, int64_t W, int variant) {
    G3D_REQUIRE(variant == G3D_VARIANT_2D || variant == G3D_VARIANT_3D, "unknown variant");
    G3D_REQUIRE(B >= 1 && A >= 1 && C >= 1 && Gmax >= 0, "sizes must be positive");
    G3D_REQUIRE(B <= 65535 * (int64_t)kImgPerCta && A < ((int64_t)1 << 31) - kTile && Gmax < (1 << 30) && C < (1 << 20),
                "size out of range");
    if (variant == G3D_VARIANT_3D) {
        G3D_REQUIRE(R == 12, "3D variant needs 12 regression outputs per anchor");
        G3D_REQUIRE(W >= 21, "3D variant needs >= 21 annotation columns");
    } else {
        G3D_REQUIRE(R == 4, "2D variant needs 4 regression outputs per anchor");
        G3D_REQUIRE(W >= 5, "2D variant needs >= 5 annotation columns");
    }
    return G3D_OK;
}

extern "C" int g3d_focal_loss_fwd(const float* cls, const float* reg, const float* anchors, const float* ann,
                                  int64_t B, int64_t A, int64_t C, int64_t R, int64_t Gmax, int64_t W, int variant,
                                  float* losses, float* per_image, int32_t* assign, int32_t* gt_count_out,
                                  void* workspace, int64_t workspace_bytes, int device, void* stream) {
    int rc = check_focal_shapes(B, A, C, R, Gmax, W, variant);
    if (rc != G3D_OK) return rc;
    G3D_REQUIRE(cls && reg && anchors && losses && per_image && workspace, "null pointer");
    G3D_REQUIRE(Gmax == 0 || ann, "null annotations");
    FocalWorkspace w = carve(workspace, B, A, Gmax);
    G3D_REQUIRE(workspace_bytes >= w.bytes, "workspace too small (see g3d_focal_workspace_bytes)");
    G3D_REQUIRE(((uintptr_t)cls % 16) == 0 && ((uintptr_t)anchors % 16) == 0 && ((uintptr_t)workspace % 256) == 0 &&
                    ((uintptr_t)per_image % 16) == 0,
                "cls/anchors/per_image must be 16-byte aligned and the workspace 256-byte aligned");
    G3D_GUARD(device);
    cudaStream_t st = (cudaStream_t)stream;
    G3D_CUDA(cudaMemsetAsync(w.counters, 0, sizeof(int32_t) * (B + 1), st));
    rc = g3d_gt_prepare(ann, B, Gmax, W, variant, (float*)w.gt_box, w.gt_row, w.gt_count, device, stream);
    if (rc != G3D_OK) return rc;
    FocalArgs p;
    p.cls = cls; p.reg = reg; p.anchors = (const float4*)anchors; p.ann = ann;
    p.gt_box = w.gt_box; p.gt_row = w.gt_row; p.gt_count = w.gt_count;
    p.partials = w.partials; p.counters = w.counters;
    p.losses = losses; p.per_image = per_image; p.assign = assign;
    p.B = (int)B; p.A = (int)A; p.C = (int)C; p.R = (int)R; p.Gmax = (int)Gmax; p.W = (int)W;
    p.T = (int)ceil_div(A, kTile);
    dim3 grid((unsigned)p.T, (unsigned)ceil_div(B, kImgPerCta));
    if (variant == G3D_VARIANT_3D) {
        if (C == 8) focal_fwd_kernel<G3D_VARIANT_3D, 8><<<grid, kTile, 0, st>>>(p);
        else        focal_fwd_kernel<G3D_VARIANT_3D, 0><<<grid, kTile, 0, st>>>(p);
    } else {
        if (C == 8) focal_fwd_kernel<G3D_VARIANT_2D, 8><<<grid, kTile, 0, st>>>(p);
        else        focal_fwd_kernel<G3D_VARIANT_2D, 0><<<grid, kTile, 0, st>>>(p);
    }
    G3D_LAUNCH_CHECK();
    if (gt_count_out)
        G3D_CUDA(cudaMemcpyAsync(gt_count_out, w.gt_count, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, st));
    return G3D_OK;
}

extern "C" int g3d_focal_loss_bwd(const float* cls, const float* reg, const float* anchors, const float* ann,
                                  int64_t B, int64_t A, int64_t C, int64_t R, int64_t Gmax, int64_t W, int variant,
                                  const float* grad_out, const float* per_image, const float* losses,
                                  const int32_t* assign, float* dcls, float* dreg, int device, void* stream) {
    int rc = check_focal_shapes(B, A, C, R, Gmax, W, variant);
    if (rc != G3D_OK) return rc;
    G3D_REQUIRE(cls && reg && anchors && grad_out && per_image && losses && assign && dcls && dreg, "null pointer");
    G3D_REQUIRE(Gmax == 0 || ann, "null annotations");
    G3D_REQUIRE(((uintptr_t)cls % 16) == 0 && ((uintptr_t)dcls % 16) == 0 && ((uintptr_t)anchors % 16) == 0,
                "cls/dcls/anchors must be 16-byte aligned");
    G3D_REQUIRE(B <= 65535, "B out of range for the backward grid");
    G3D_GUARD(device);
    cudaStream_t st = (cudaStream_t)stream;
    FocalBwdArgs p;
    p.cls = cls; p.reg = reg; p.anchors = (const float4*)anchors; p.ann = ann;
    p.grad_out = grad_out; p.per_image = per_image; p.losses = losses; p.assign = assign;
    p.dcls = dcls; p.dreg = dreg;
    p.B = (int)B; p.A = (int)A; p.C = (int)C; p.R = (int)R; p.Gmax = (int)Gmax; p.W = (int)W;
    dim3 grid((unsigned)ceil_div(A, 256), (unsigned)B);
    if (variant == G3D_VARIANT_3D) {
        if (C == 8) focal_bwd_kernel<G3D_VARIANT_3D, 8><<<grid, 256, 0, st>>>(p);
        else        focal_bwd_kernel<G3D_VARIANT_3D, 0><<<grid, 256, 0, st>>>(p);
    } else {
        if (C == 8) focal_bwd_kernel<G3D_VARIANT_2D, 8><<<grid, 256, 0, st>>>(p);
        else        focal_bwd_kernel<G3D_VARIANT_2D, 0><<<grid, 256, 0, st>>>(p);
    }
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}
