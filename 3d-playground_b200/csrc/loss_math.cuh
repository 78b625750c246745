// loss_math.cuh — arithmetic of the loss path shared by the kernels of focal_loss.cu.
//
//   focal classification term / gradient      3D losses.py:56,133-152 ; 2D retinanet/losses.py:49,106-125
//   smooth-L1, cosine direction loss          3D losses.py:217-304,343-350
//   corner sign table                         3D losses.py:311-327
//   2D regression targets                     2D retinanet/losses.py:137-157
//
// Hyper-parameters (alpha, gamma, thresholds, beta, top weighting, clamp bounds) are runtime values with the
// reference's constants as defaults (losses.py:28-30,56,121,124,343,346-348); the fast paths are taken when gamma == 2.
#pragma once
#include "common.cuh"

namespace g3d {

// Device-side copy of the hyper-parameters (filled on the host by make_hyper, focal_loss.cu).
struct LossHyper {
    float alpha, one_minus_alpha;   // losses.py:28,138-139: alpha for targets == 1, 1 - alpha otherwise (FP32 difference)
    float gamma;                    // losses.py:29
    int gamma_is_two;
    float pmin, pmax;               // losses.py:56: torch.clamp(classification, 1e-4, 1 - 1e-4)
    float pos_thr, neg_thr;         // losses.py:124 (>= 0.5 positive), :121 (< 0.4 negative)
    float cull_mul;                 // inter * cull_mul <= union  =>  IoU < neg_thr with margin: the pair cannot matter
    float group_cull;               // best-case IoU of a GT box over a group of anchors below this: skip the box
    float win_q;                    // t / (1 + t), t = 0.9625 neg_thr: window of cells whose anchor can reach IoU t
    unsigned neg_thr_bits;          // bit pattern of neg_thr (key encoding of the GT-centric assignment)
    float sl1_beta, sl1_quad, sl1_off, sl1_slope;   // smooth-L1 (losses.py:345-349): d <= beta ? quad d^2 : d - off
    float top_w;                    // losses.py:30,343: weight of regression_diff[:, 8:16]
};

// ---------------------------------------------------------------------------------------------------------------
// packed FP32 pairs (Blackwell FFMA2 / FADD2 / FMUL2: two IEEE FP32 operations per issue slot, same roundings as the
// scalar instructions)
// ---------------------------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// 1/x for x in a benign range (here [0.75, 2] and [1e-4, 1]): a single MUFU.RCP (<= 1 ulp).
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// scalar focal term and gradient (general: any alpha / gamma, full logf) - the rare paths and the generic kernels
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pow_gamma(float x, const LossHyper& h) {
    return h.gamma_is_two ? x * x : powf(x, h.gamma);    // torch.pow(x, 2.0) is x * x
}
// one focal term (losses.py:138-150): alpha_t * (1 - p_t)^gamma * bce, target t in {0,1}
__device__ __forceinline__ float focal_term(float p_raw, bool t, const LossHyper& h) {
    const float p = fminf(fmaxf(p_raw, h.pmin), h.pmax);
    const float u = 1.0f - p;
    const float fw = t ? u : p;
    const float x = t ? p : u;
    const float w = (t ? h.alpha : h.one_minus_alpha) * pow_gamma(fw, h);
    return w * (-logf(x));
}
// d(focal term)/dp, zero outside the clamp range (torch.clamp's backward passes min <= x <= max)
__device__ __forceinline__ float focal_term_grad(float p, bool t, const LossHyper& h) {
    if (!(p >= h.pmin && p <= h.pmax)) return 0.0f;
    const float u = 1.0f - p;
    if (h.gamma_is_two) {
        if (!t) return h.one_minus_alpha * (2.0f * p * (-logf(u)) + (p * p) * rcp_fast(u));
        return h.alpha * (2.0f * u * logf(p) - (u * u) * rcp_fast(p));
    }
    // general gamma: d/dp [a f^g * -log(x)], f = p (t = 0, x = 1 - p) or 1 - p (t = 1, x = p)
    if (!t) return h.one_minus_alpha * (h.gamma * powf(p, h.gamma - 1.0f) * (-logf(u)) + powf(p, h.gamma) / u);
    return h.alpha * (h.gamma * powf(u, h.gamma - 1.0f) * logf(p) - powf(u, h.gamma) / p);
}

// ---------------------------------------------------------------------------------------------------------------
// the streaming fast path: 8 elements of one NEGATIVE anchor row (target 0 everywhere), gamma == 2
// ---------------------------------------------------------------------------------------------------------------
// -log(u) for u = fl(1 - p), p < 0.25: with pe = 1 - u (exact, Sterbenz) and z = pe / (2 - pe) = pe / (1 + u),
//     -log(1 - pe) = 2 atanh(z) = 2 z (1 + z^2/3 + z^4/5 + z^6/7 + ...),
// truncated after z^6 (|z| < 1/7: relative truncation error < 1.9e-8).  The series is evaluated on the SAME rounded u
// the reference takes the log of, so it tracks torch.log(1.0 - classification) to ~2e-7 relative; one MUFU.RCP and
// three (packed) FMAs instead of ~22 instructions of logf.
//
// Returns sum_c p_c^2 * -log(1 - p_c) over the row (the caller applies 1 - alpha) and, if GRAD, g[c] = scale *
// (2 p nl + p^2 / u) for elements inside the clamp range, 0 outside; pmax_out = largest clamped probability of the
// row (>= 0.25 -> the caller redoes the row with the full logf).
template <bool GRAD>
__device__ __forceinline__ float focal_neg_row8(const float (&pv)[8], float pmin, float pmax, float scale, float (&g)[8],
                                                float& pmax_out) {
    float p[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) p[c] = fminf(fmaxf(pv[c], pmin), pmax);
    pmax_out = fmaxf(fmaxf(fmaxf(p[0], p[1]), fmaxf(p[2], p[3])), fmaxf(fmaxf(p[4], p[5]), fmaxf(p[6], p[7])));
    const f32x2 one2 = pk2(1.0f, 1.0f), neg2 = pk2(-1.0f, -1.0f), two2 = pk2(2.0f, 2.0f);
    const f32x2 c7 = pk2(2.0f / 7.0f, 2.0f / 7.0f), c5 = pk2(0.4f, 0.4f), c3 = pk2(2.0f / 3.0f, 2.0f / 3.0f);
    const f32x2 scale2 = pk2(scale, scale);
    f32x2 P[4], U[4], NL[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        P[k] = pk2(p[2 * k], p[2 * k + 1]);
        U[k] = fma2(P[k], neg2, one2);                  // u = fl(1 - p)
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const f32x2 pe = fma2(U[k], neg2, one2);        // 1 - u, exact
        const f32x2 d = add2(U[k], one2);               // 1 + u
        float d0, d1;
        upk2(d, d0, d1);
        const f32x2 r = pk2(rcp_fast(d0), rcp_fast(d1));
        const f32x2 z = mul2(pe, r);
        const f32x2 z2 = mul2(z, z);
        f32x2 s = fma2(z2, c7, c5);
        s = fma2(z2, s, c3);
        s = fma2(z2, s, two2);
        NL[k] = mul2(z, s);
    }
    f32x2 acc2 = pk2(0.0f, 0.0f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const f32x2 a = mul2(P[k], NL[k]);              // p * nl
        acc2 = fma2(P[k], a, acc2);                     // + p^2 nl
        if (GRAD) {
            float u0, u1;
            upk2(U[k], u0, u1);
            const f32x2 ru = pk2(rcp_fast(u0), rcp_fast(u1));
            const f32x2 q = mul2(P[k], mul2(P[k], ru)); // p^2 / u
            const f32x2 t = fma2(two2, a, q);           // 2 p nl + p^2 / u
            float g0, g1;
            upk2(mul2(t, scale2), g0, g1);
            g[2 * k] = (p[2 * k] == pv[2 * k]) ? g0 : 0.0f;             // p == p_raw  <=>  p_raw inside [min, max]
            g[2 * k + 1] = (p[2 * k + 1] == pv[2 * k + 1]) ? g1 : 0.0f;
        }
    }
    float a0, a1;
    upk2(acc2, a0, a1);
    return a0 + a1;
}

// ---------------------------------------------------------------------------------------------------------------
// regression / direction losses of one positive anchor
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float smooth_l1(float d, const LossHyper& h) {
    return (d <= h.sl1_beta) ? h.sl1_quad * (d * d) : d - h.sl1_off;
}
__device__ __forceinline__ float smooth_l1_slope(float d, const LossHyper& h) {
    return (d <= h.sl1_beta) ? h.sl1_slope * d : 1.0f;
}

__device__ __forceinline__ float cos_loss(float rx, float ry, float tx, float ty) {
    const float rn = sqrtf(rx * rx + ry * ry), tn = sqrtf(tx * tx + ty * ty);
    return 1.0f - (rx * tx + ry * ty) / (rn * tn);     // no epsilon: a zero vector gives NaN, as in losses.py:227
}
// gradient of cos_loss w.r.t. (rx, ry)
__device__ __forceinline__ void cos_loss_grad(float rx, float ry, float tx, float ty, float& gx, float& gy) {
    // d(1 - r.t/(|r||t|))/dr = -(t^ - cos * r^)/|r|.  In the plane t^ - cos*r^ = (r_perp^ . t^) r_perp^, which gives the
    // cancellation-free form  g = cross * (ry, -rx) / (|r|^3 |t|),  cross = rx*ty - ry*tx  (evaluated with an exact
    // product residual).  The textbook form subtracts two terms of size 1/|r| and loses digits when r is nearly
    // parallel to t or very short; this one stays within a few ulp of the exact gradient.  A zero vector gives 0/0 = NaN
    // in both components, as autograd does for the reference's expression.
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
__device__ __forceinline__ void targets_2d(const float* grow, const float4& an, float* t /*4*/) {
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

// One POSITIVE anchor: the regression loss terms of the row (3D: 20 smooth-L1 terms + mean of the three cosine losses,
// losses.py:156-350; 2D: 4 smooth-L1 terms, retinanet/losses.py:129-173) and, if drow != null, the row's regression
// gradient for the scaled upstream gradients (s_reg, s_vp).
//   3D: tab[0..19] = the GT row's 20 regression targets (annotation cols 0..19), tab[20..25] its three direction vectors
//   2D: tab[0..3]  = the GT box
template <int VARIANT>
__device__ __forceinline__ void positive_row(const float* r /*12|4*/, const float* tab, const float4 an, float s_reg,
                                             float s_vp, const LossHyper& h, float* dr /*12|4 or null*/, float& reg_sum,
                                             float& vp_term) {
    if (VARIANT == G3D_VARIANT_3D) {
        float pr[20];
        const float* tv = tab + 20;
        vp_term = (cos_loss(r[2], r[3], tv[0], tv[1]) + cos_loss(r[4], r[5], tv[2], tv[3]) +
                   cos_loss(r[6], r[7], tv[4], tv[5])) / 3.0f;
        pred_corners(r, pr);
        const float aw = an.z - an.x, ah = an.w - an.y;
        const float acx = an.x + 0.5f * aw, acy = an.y + 0.5f * ah;
        float s = 0.0f, g[20];
#pragma unroll
        for (int i = 0; i < 20; ++i) {
            const float tn = (i & 1) ? (tab[i] - acy) / ah : (tab[i] - acx) / aw;   // losses.py:330-331
            const float diff = tn - pr[i];
            const float w = (i >= 8 && i < 16) ? h.top_w : 1.0f;                    // top_weighting, losses.py:343
            const float d = fabsf(diff) * w;
            s += smooth_l1(d, h);
            const float sg = (diff > 0.0f) ? 1.0f : ((diff < 0.0f) ? -1.0f : 0.0f);
            // d smooth_l1 / d pred = slope(d) * w * d|diff|/dpred = slope * w * (-sign(diff))
            g[i] = -s_reg * smooth_l1_slope(d, h) * w * sg;
        }
        reg_sum = s;
        if (dr) {
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
        }
    } else {
        float t[4];
        targets_2d(tab, an, t);
        float s = 0.0f;
        vp_term = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float diff = t[i] - r[i];
            const float d = fabsf(diff);
            s += smooth_l1(d, h);
            if (dr) {
                const float sg = (diff > 0.0f) ? 1.0f : ((diff < 0.0f) ? -1.0f : 0.0f);
                dr[i] = -s_reg * smooth_l1_slope(d, h) * sg;
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
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------------------------
// 256-bit global accesses (one 8-class row per lane) and the bulk-copy (TMA) zero fill
// ---------------------------------------------------------------------------------------------------------------
// EVICT_FIRST: the row is marked evict-first in L2.  Measured at cfg2 (A/B in one session): it helps the read-only
// forward pass (170 -> 164 us) and hurts the pass that also writes gradients (240 -> 250 us; so do evict-first hints
// on the gradient stores and on the bulk zero fill: +4 us each), hence a template switch.
template <bool EVICT_FIRST>
__device__ __forceinline__ void ld_row8(const float* p, float (&v)[8]) {
    if (EVICT_FIRST) {
        unsigned long long policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                     : "l"(p), "l"(policy));
    } else {
        asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                     : "l"(p));
    }
}
__device__ __forceinline__ void st_row8(float* p, const float (&v)[8]) {
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]),
                 "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

constexpr int kZeroTile = 8192;   // bytes of zeroed shared memory that the bulk copies replicate into global memory

// A byte range of global memory that a kernel zero-fills on the side with bulk async copies (cp.async.bulk
// shared -> global, SASS UBLKCP): one elected thread per CTA issues them, they use no LSU issue slots and no registers.
struct FillSlice {
    char* base;
    long long bytes;    // multiple of 16; base 16-byte aligned
};

__device__ __forceinline__ void bulk_zero(char* dst, long long bytes, uint32_t tile) {
    while (bytes > 0) {
        const uint32_t n = bytes < (long long)kZeroTile ? (uint32_t)bytes : (uint32_t)kZeroTile;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(dst)),
                     "r"(tile), "r"(n)
                     : "memory");
        dst += n;
        bytes -= n;
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// the source tile must stay intact until the copies have read it: call before the CTA can exit
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// all threads of the CTA: zero the tile and make it visible to the async proxy (barrier included)
__device__ __forceinline__ void zero_tile_init(unsigned char* ztile) {
    for (int i = threadIdx.x; i < kZeroTile / 16; i += blockDim.x)
        reinterpret_cast<float4*>(ztile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
}
// one thread: part `rank` of `n` equal 1 KB-granular parts of the slice
__device__ __forceinline__ void fill_part(const FillSlice& f, long long rank, long long n, uint32_t tile) {
    if (f.bytes <= 0) return;
    const long long per = (((f.bytes + n - 1) / n) + 1023) & ~1023LL;
    const long long lo = per * rank;
    const long long hi = lo + per < f.bytes ? lo + per : f.bytes;
    if (lo < hi) bulk_zero(f.base + lo, hi - lo, tile);
}
}  // namespace g3d
