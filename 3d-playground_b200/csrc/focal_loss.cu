// focal_loss.cu — a2..a6: FocalLoss.forward / backward of both retinanet copies.
//
//   3D: pytorch_retinanet_detector_directional/retinanet/losses.py:27-362
//   2D: retinanet/losses.py:27-177
//
// One training step (forward + every gradient for the expected upstream gradients) is five launches; the backward
// call adds one that only verifies the expectation on the device.
//
//   K0 loss_prologue_kernel   one CTA per image: drop class == -1 rows, the 2D box each GT row is matched with, a
//                             128-byte table row per GT (regression targets, direction vectors), the class; zero the
//                             counters; extra CTAs zero the keys / chunk mask of the GT-centric assignment.
//   K1 assignment, one of
//      assign_pairs_kernel    (anchors = the regular pyramid of Anchors.forward, <= 256 GT rows) GT-centric: one warp per
//                             (image, GT row) evaluates only the window of cells whose anchors can reach the negative
//                             threshold and atomicMax-es a 32-bit (IoU, GT index) key per (image, anchor); a chunk mask
//                             records which 32-anchor chunks hold a key at all (a few percent);
//      assign_codes_kernel    (any anchor table) anchor-centric tiles with GT culling -> byte codes + positives lists
//                             (and the zero fill of dreg: this kernel is issue-bound and leaves HBM idle).
//   K2 assign_resolve_kernel  (GT-centric only) touched chunks -> per-image lists of positive anchors (anchor, GT index),
//                             their count (the normaliser of everything that follows).
//   K3 focal_stream8_kernel   the HBM-bound sweep: lane <-> row, one 8-class row per lane with 256-bit loads, the keys of
//                             touched chunks only; focal terms AND their gradient from one -log(1-p) (packed FP32x2
//                             arithmetic); dcls with 256-bit stores; per-CTA partial sums; and the zero fill of dreg.
//   K4 positives_kernel       launched as a programmatic dependent of K3 (griddepcontrol): one thread per positive anchor
//                             from the lists - 20 smooth-L1 terms + 3 direction cosines, exact fixed-point sums, the
//                             gradient row written into dreg - then waits for K3 and the last CTA of an image (ticket)
//                             reduces the image, the last image forms the batch means.  No host synchronisation.
//
// The regression gradient dreg is zero except on the ~1 % positive rows, but autograd needs it dense: 48 B/row of
// zeros, 41 % of all bytes the step writes.  The stream warps write them with bulk async copies (cp.async.bulk shared ->
// global) of a zeroed shared-memory tile, issued by ONE lane per warp for the warp's rows: no LSU issue slots, no
// registers.  Only chunks that hold a key are written with ordinary stores, lane by lane, and the lanes of positive
// rows write nothing - those rows belong to K4, so the two kernels never touch the same bytes and need no ordering
// (which is what lets K4 start before K3 has drained).
//
// Gradients written during the forward assume the upstream gradients the host announces (1 for `(cls + reg +
// vp).backward()`, 1/world under dist.py).  g3d_focal_loss_bwd (focal_bwd_kernel) checks the assumption ON THE DEVICE
// and recomputes what does not hold; in the usual step it exits after one wave.
#include <atomic>
#include <string.h>
#include "assign_tile.cuh"
#include "loss_math.cuh"

namespace g3d {

constexpr int kImgPerCta = 4;  // images processed per CTA of assign_codes_kernel (anchors / statistics loaded once)
constexpr int kStageGroup = 4; // images staged behind one pair of barriers
constexpr int kTabW = 32;      // floats per GT table row (128 bytes)

// byte codes handed from the assignment to the streaming pass
constexpr int kCodeNegative = 0, kCodeIgnore = 1, kCodePositive = 2;   // positive: 2 + class (class < 253, else 255)

__device__ __forceinline__ int code8_of_class(int cls, int C) { return (cls >= 0 && cls < C && cls < 253) ? kCodePositive + cls : 255; }

// Where the streaming pass finds the assignment of a row.  Anchor-centric path: one byte per (image, anchor).  GT-centric
// path: the 32-bit keys themselves, read only for the 32-anchor chunks whose bit is set in the chunk mask (a few percent
// of all chunks hold a key at all).
struct CodeSrc {
    const uint8_t* code8;      // [B][Ap] or null
    const uint32_t* keys;      // [B][Ap]
    const uint32_t* mask;      // [B][MW]: bit (a >> 5) & 31 of word a >> 10 = chunk a >> 5 holds at least one key
    const int32_t* gt_cls;     // [B][Gmax]
    int Ap, MW, Gmax, C;
    float pos_thr;
    unsigned neg_thr_bits;
};
__device__ __forceinline__ float key_iou(unsigned key, unsigned neg_thr_bits) { return __uint_as_float((key >> 8) - 1u + neg_thr_bits); }
__device__ __forceinline__ int key_gt(unsigned key) { return 255 - (int)(key & 255u); }
// non-zero key -> byte code
__device__ __forceinline__ int code_of_key(const CodeSrc& s, int b, unsigned key) {
    if (!(key_iou(key, s.neg_thr_bits) >= s.pos_thr)) return kCodeIgnore;
    return code8_of_class(__ldg(s.gt_cls + (int64_t)b * s.Gmax + key_gt(key)), s.C);
}
// any row (slow paths): byte code from whichever representation is present
__device__ __forceinline__ int code_of_row(const CodeSrc& s, int b, int a) {
    if (s.code8) return __ldg(s.code8 + (int64_t)b * s.Ap + a);
    if (!((__ldg(s.mask + (int64_t)b * s.MW + (a >> 10)) >> ((a >> 5) & 31)) & 1u)) return kCodeNegative;
    const unsigned key = __ldg(s.keys + (int64_t)b * s.Ap + a);
    return key ? code_of_key(s, b, key) : kCodeNegative;
}

// =====================================================================================================================
// tuning knobs (process-wide, set through g3d_set_tuning; read once per call)
// =====================================================================================================================
static std::atomic<int> g_force_anchor_centric{0};   // != 0: ignore pyramid_host (benchmark comparisons)
static std::atomic<int> g_pdl{1};                     // 0: launch K4 without programmatic dependent launch

// =====================================================================================================================
// K0: prologue
// =====================================================================================================================
struct PrologueArgs {
    const float* ann;
    float4* gt_box;        // [B][Gmax] compacted valid rows: the 2D box used for assignment
    int32_t* gt_row;       // [B][Gmax] original annotation row of each compacted row
    int32_t* gt_cls;       // [B][Gmax] class of each compacted row
    float* gt_tab;         // [B][Gmax][kTabW]
    int32_t* gt_count;     // [B]
    int32_t* zero_ptr;     // counters to zero
    int zero_n;
    int B, Gmax, W, variant;
    FillSlice fill;        // keys + chunk mask of the GT-centric assignment: zeroed by the first nfill CTAs
    int nfill;
    // persistent regression gradient (dreg_state == G3D_DREG_CLEAN): dreg is zero except on the rows the previous step
    // wrote - its positive anchors, still listed in pos_anchor (prev_npos[b] of them) - which the image's CTA zeroes here
    float* dreg;           // null: nothing to re-zero
    const int32_t* prev_npos;
    const int32_t* pos_anchor;
    int A, R;
};

__global__ void __launch_bounds__(kTile) loss_prologue_kernel(const PrologueArgs p) {
    __shared__ int wcount[kWarps];
    if ((int)blockIdx.x < p.nfill) {
        // keys + chunk mask: 50 MB that mostly stay in L2 - plain 16-byte stores from every thread (the per-SM bulk-copy
        // path is the slower one for a fill that is not competing with anything)
        float4* dst = reinterpret_cast<float4*>(p.fill.base);
        const long long n16 = p.fill.bytes >> 4;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (long long i = (long long)blockIdx.x * kTile + threadIdx.x; i < n16; i += (long long)p.nfill * kTile) dst[i] = z;
        return;
    }
    const int b = blockIdx.x - p.nfill, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = b * kTile + threadIdx.x; i < p.zero_n; i += p.B * kTile) p.zero_ptr[i] = 0;
    if (p.dreg) {
        const int nprev = min(max(__ldg(p.prev_npos + b), 0), p.A);
        const int q = p.R >> 2;                                         // float4 pieces per row (3 or 1)
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = threadIdx.x; i < nprev * q; i += kTile) {
            const int r = i / q;
            const int a = __ldg(p.pos_anchor + (int64_t)b * p.A + r);
            if (a >= 0 && a < p.A) reinterpret_cast<float4*>(p.dreg + ((int64_t)b * p.A + a) * p.R)[i - r * q] = z;
        }
    }
    const float* img = p.ann + (int64_t)b * p.Gmax * p.W;
    const bool three_d = p.variant == G3D_VARIANT_3D;
    const int cls_col = three_d ? 20 : 4;
    int count = 0;
    for (int base = 0; base < p.Gmax; base += kTile) {
        const int g = base + threadIdx.x;
        bool keep = false;
        const float* row = img + (int64_t)g * p.W;
        if (g < p.Gmax) keep = (row[cls_col] != -1.0f);                  // losses.py:54 / retinanet/losses.py:46
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
            const int64_t pos = (int64_t)b * p.Gmax + count + off + __popc(bal & ((1u << lane) - 1u));
            float* tab = p.gt_tab + pos * kTabW;
            float4 box;
            if (three_d) {
                float t[20];
#pragma unroll
                for (int i = 0; i < 20; ++i) t[i] = row[i];
                float xmin = t[0], xmax = t[0], ymin = t[1], ymax = t[1];     // losses.py:93-107
#pragma unroll
                for (int k = 1; k < 8; ++k) {
                    xmin = fminf(xmin, t[2 * k]); xmax = fmaxf(xmax, t[2 * k]);
                    ymin = fminf(ymin, t[2 * k + 1]); ymax = fmaxf(ymax, t[2 * k + 1]);
                }
                box = make_float4(xmin, ymin, xmax, ymax);
                float tv[6];
                gt_directions(t, tv);
#pragma unroll
                for (int i = 0; i < 20; i += 4) reinterpret_cast<float4*>(tab)[i >> 2] = make_float4(t[i], t[i + 1], t[i + 2], t[i + 3]);
                reinterpret_cast<float4*>(tab)[5] = make_float4(tv[0], tv[1], tv[2], tv[3]);
                reinterpret_cast<float4*>(tab)[6] = make_float4(tv[4], tv[5], 0.f, 0.f);
            } else {
                box = make_float4(row[0], row[1], row[2], row[3]);
                reinterpret_cast<float4*>(tab)[0] = box;
            }
            p.gt_box[pos] = box;
            p.gt_row[pos] = g;
            p.gt_cls[pos] = (int)(long long)row[cls_col];                   // .long() of the class column (losses.py:131)
        }
        count += total;
    }
    if (threadIdx.x == 0) p.gt_count[b] = count;
}

// =====================================================================================================================
// K1, anchor-centric form (any anchor table)
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
    const int32_t* gt_cls;
    const int32_t* gt_count;
    uint8_t* code8;      // [B][Ap]
    int32_t* assign;     // [B][A] or null
    int32_t* npos;       // [B], zero on entry
    int32_t* pos_anchor; // [B][A]: anchor indices of the positives of each image, in arrival order (first npos[b] valid)
    int32_t* pos_gt;     // [B][A]: compacted GT index of each
    float* dreg;         // [B][A][R] or null: zero-filled here (see the kernel)
    int B, A, Ap, Gmax, R, C;
    float pos_thr, neg_thr, cull_mul, group_cull;
};

// One-lane atomics as single instructions.  `if (lane == 0) atomicAdd(...)` makes nvcc wrap the call in its warp-aggregation
// pattern (vote, leader election, popc, shuffle: ~15 instructions around one ATOMS), which at one ticket per 32-anchor
// unit was 15 % of this issue-bound kernel's instructions.
// The address is made lane-dependent in a way the compiler cannot fold (offset zero for lane 0, the one lane that executes
// it; %laneid read through asm): ptxas then emits the bare ATOMS / ATOMG instead of its aggregation sequence.
__device__ __forceinline__ int lane0_atomic_add_shared(int* addr, int v) {
    int old;
    unsigned l2;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l2));          // opaque to the front end: not folded under `lane == 0`
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

// Size statistics of a group of anchors (a warp or a tile), for the "can the IoU reach the negative threshold at all"
// cull: for well-formed boxes  IoU = inter / union,  inter <= min(wa, wg) * min(ha, hg)  and  inter <= min(area_a,
// area_g), union >= max(area_a, area_g).  So a GT box whose best case over the group stays below 0.95 x the threshold
// (margin for the FP32 roundings of the real thing, which are ~1e-7 relative) has IoU below the threshold with every
// anchor of the group: whatever its exact IoU, it cannot move an anchor out of the `negative` class, and it can be skipped.
struct GroupStats {
    float4 bb;          // min x1, min y1, max x2, max y2
    float wmax, hmax;   // largest width / height
    float amin, amax;   // smallest / largest area
    bool wellformed;    // every anchor has positive width and height
};

__device__ __forceinline__ bool can_touch(const float4& g, const GroupStats& s, float group_cull) {
    // keep unless provably disjoint from every anchor of the group (NaN coordinates are never culled here)
    if (g.z <= s.bb.x || g.x >= s.bb.z || g.w <= s.bb.y || g.y >= s.bb.w) return false;
    const float gw = g.z - g.x, gh = g.w - g.y;
    if (s.wellformed && gw > 0.0f && gh > 0.0f) {
        const float ga = gw * gh;
        const float best_inter = fminf(fminf(s.wmax, gw) * fminf(s.hmax, gh), fminf(s.amax, ga));
        if (best_inter < group_cull * fmaxf(s.amin, ga)) return false;
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
    GroupStats ts;
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
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) sm.red[k][warp] = v[k];
            sm.wcount[0][warp] = wf ? 1 : 0;
            sm.wbb[warp] = make_float4(-v[0], -v[1], v[2], v[3]);
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
                hit = can_touch(gb[j], ts, p.group_cull);
            }
            bal[j] = __ballot_sync(0xffffffffu, hit);
        }
        // unordered compaction (one shared-memory atomic per warp and image): the candidate loop below breaks ties by
        // GT index explicitly, so the order of the survivors does not matter
        int base[kStageGroup];
#pragma unroll
        for (int j = 0; j < kStageGroup; ++j) {
            base[j] = 0;
            if (lane == 0 && bal[j]) base[j] = lane0_atomic_add_shared(&sm.total[i0 + j], __popc(bal[j]));
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
        if (lane == 0) u = lane0_atomic_add_shared(&sm.next_unit, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= nunits) break;
        const int i = u / kWarps, slice = u - i * kWarps;
        const int b = b0 + i;
        const int a = tile * kTile + slice * 32 + lane;
        const bool valid = a < p.A;
        const float4 an = sm.an[slice * 32 + lane];
        const float area_a = box_area_rn(an.x, an.y, an.z, an.w);
        const float4 wbb = sm.wbb[slice];
        const int Gi = __ldg(p.gt_count + b);
        // The regression gradient is zero except on the few positive rows.  This kernel is issue-bound and leaves HBM
        // idle, so the 48 bytes / row of zeros are written from here (3 coalesced 16-byte stores per lane, drained in
        // the background).
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
                near = !(t.z <= wbb.x || t.x >= wbb.z || t.w <= wbb.y || t.y >= wbb.w);
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
                    // inter <= ua0 / cull_mul  =>  IoU < neg_thr (0.3847 < 0.4 by default): cannot change the code of
                    // this anchor, skip the division (the clamp of the union only matters below 1e-8, where this test passes)
                    if (__fmul_rn(inter, p.cull_mul) > ua0) {
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
        // `best` is exact whenever it is >= neg_thr (every pair that can reach it was evaluated exactly; ties go to the
        // lower GT index = torch.max's first-maximal-index); below that only "< neg_thr" is used.
        int code = G3D_ASSIGN_NEGATIVE, c8 = kCodeNegative;
        if (Gi > 0) {
            if (best >= p.pos_thr) {
                code = __ldg(p.gt_row + (int64_t)b * p.Gmax + besti);
                c8 = code8_of_class(__ldg(p.gt_cls + (int64_t)b * p.Gmax + besti), p.C);
            } else if (!(best < p.neg_thr)) {
                code = G3D_ASSIGN_IGNORE;
                c8 = kCodeIgnore;
            }
        }
        if (valid) {
            p.code8[(int64_t)b * p.Ap + a] = (uint8_t)c8;
            if (p.assign) p.assign[(int64_t)b * p.A + a] = code;
        }
        // positives are appended to the image's list (integer atomics: the COUNT is deterministic, the order is not -
        // everything downstream is order independent: per-row gradients, and loss sums in exact fixed point)
        const bool is_pos = valid && code >= 0;
        const unsigned posmask = __ballot_sync(0xffffffffu, is_pos);
        if (posmask) {
            int base = 0;
            if (lane == 0) base = lane0_atomic_add_global(p.npos + b, __popc(posmask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (is_pos) {
                const int64_t slot = (int64_t)b * p.A + base + __popc(posmask & ((1u << lane) - 1u));
                p.pos_anchor[slot] = a;
                p.pos_gt[slot] = besti;
            }
        }
    }
}

__global__ void __launch_bounds__(kTile, 6) assign_codes_kernel(const AssignCodesArgs p) {
    __shared__ StageSmem sm;
    assign_item(p, sm, blockIdx.x, blockIdx.y);
}

// =====================================================================================================================
// K1 / K2, GT-centric form (anchors known to be the regular pyramid of Anchors.forward, Gmax <= 256)
// =====================================================================================================================
// assign_codes_kernel above walks every (32-anchor slice, image) unit - 389 k of them at cfg2 - although 98 % of the
// anchors are negatives: it is instruction-bound on per-unit overhead.  When the anchor table is the pyramid (level ->
// cell -> shape, anchors.py:21-40) the loop can be inverted.  One warp per (image, GT row): for every level and shape the
// cells whose anchor can reach IoU t = 0.9625 neg_thr with this box form a small window - IoU >= t needs inter >=
// t/(1+t) (Aa + Ag), and inter <= iw * min(ah, gh), so iw >= that / min(ah, gh), which bounds the anchor centre to
// [gx1 + iw_min - aw/2, gx2 - iw_min + aw/2] (same in y; FP32, widened by 0.05 px; tests/test_assign_windows.py checks
// the superset property) - and only those pairs (~110 per GT row, 1.3 x the pairs that really reach t) are evaluated,
// with the exact arithmetic of the kernel above on the anchor values read from the table.  The (level, shape) windows of
// a GT row are flattened over the lanes (prefix sum + binary search by shuffles), so the ~110 pairs take 4 rounds of 32
// instead of one round per non-empty window.  A pair with IoU >= neg_thr goes into a per-(image, anchor) key with
// atomicMax: key = (IoU bits - bits(neg_thr) + 1) << 8 | (255 - GT index): larger IoU wins, then the lower index =
// torch.max's first-maximal rule (a fire-and-forget RED).  assign_resolve_kernel then makes one coalesced pass over the
// keys (4 B / anchor) and writes the byte codes and the positives lists.
constexpr int kPyrLevels = 8, kPyrShapes = 16;
struct Pyramid {
    float aw[kPyrLevels * kPyrShapes], ah[kPyrLevels * kPyrShapes];
    float inv_stride[kPyrLevels];
    int first[kPyrLevels], rows[kPyrLevels], cols[kPyrLevels];
    int L, S;
};

struct PairArgs {
    const float4* anchors;
    const float4* gt_box;      // [B][Gmax] compacted valid rows
    const int32_t* gt_count;   // [B]
    uint32_t* keys;            // [B][Ap], zero on entry
    uint32_t* mask;            // [B][MW] chunk mask, zero on entry
    int B, A, Ap, MW, Gmax;
    float win_q, cull_mul, neg_thr;
    unsigned neg_thr_bits;
};

__global__ void __launch_bounds__(256) assign_pairs_kernel(const PairArgs p, const __grid_constant__ Pyramid pyr) {
    const int lane = threadIdx.x & 31;
    const int wid = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (wid >= p.B * p.Gmax) return;
    const int b = wid / p.Gmax, g = wid - b * p.Gmax;
    if (g >= __ldg(p.gt_count + b)) return;
    const float4 gk = __ldg(p.gt_box + (int64_t)b * p.Gmax + g);
    const float gw = gk.z - gk.x, gh = gk.w - gk.y;
    if (!(gw > 0.0f && gh > 0.0f)) return;                    // no overlap with anything is possible
    const float Ag = gw * gh;
    const float area_g = box_area_rn(gk.x, gk.y, gk.z, gk.w);
    uint32_t* keys = p.keys + (int64_t)b * p.Ap;
    uint32_t* mask = p.mask + (int64_t)b * p.MW;
    // The window only has to be a superset: FP32 with approximate reciprocals (relative error ~1e-6) against a threshold
    // 3.75 % below the one that matters, plus 0.05 px of slack on the centre range.
    const float kq = p.win_q, eps = 0.05f;
    const int ncombo = pyr.L * pyr.S;
    for (int k0 = 0; k0 < ncombo; k0 += 32) {
        // lane k0 + lane derives the window of its own (level, shape) combination
        const int k = k0 + lane;
        int c0 = 0, r0 = 0, wc = 1, ncell = 0, cols = 0, first = 0, sc = 0;
        if (k < ncombo) {
            const int l = k / pyr.S;
            sc = k - l * pyr.S;
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
                cols = pyr.cols[l];
                first = pyr.first[l];
                const int c1 = min(cols - 1, (int)floorf(hi_x * inv_stride - 0.5f));
                const int r1 = min(pyr.rows[l] - 1, (int)floorf(hi_y * inv_stride - 0.5f));
                if (c1 >= c0 && r1 >= r0) { wc = c1 - c0 + 1; ncell = wc * (r1 - r0 + 1); }
            }
        }
        // flatten the windows over the lanes: inclusive prefix sum of the cell counts, then every lane finds the window
        // its item belongs to by binary search over the (monotone) prefix sums
        int incl = ncell;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int excl = incl - ncell;
        for (int t0 = 0; t0 < total; t0 += 32) {
            const int t = t0 + lane;
            int src = 0;                                       // first lane whose inclusive sum exceeds t
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const int v = __shfl_sync(0xffffffffu, incl, src + step - 1);
                if (v <= t) src += step;
            }
            const int e = t - __shfl_sync(0xffffffffu, excl, src);
            const int c0c = __shfl_sync(0xffffffffu, c0, src), r0c = __shfl_sync(0xffffffffu, r0, src);
            const int wcc = __shfl_sync(0xffffffffu, wc, src), colsc = __shfl_sync(0xffffffffu, cols, src);
            const int firstc = __shfl_sync(0xffffffffu, first, src), scc = __shfl_sync(0xffffffffu, sc, src);
            if (t >= total) continue;
            const int rr = e / wcc;
            const int r = r0c + rr, c = c0c + (e - rr * wcc);
            const int a = firstc + (r * colsc + c) * pyr.S + scc;
            const float4 an = __ldg(p.anchors + a);
            const float iw = __fsub_rn(fminf(an.z, gk.z), fmaxf(an.x, gk.x));
            const float ih = __fsub_rn(fminf(an.w, gk.w), fmaxf(an.y, gk.y));
            if (!(iw > 0.0f && ih > 0.0f)) continue;
            const float inter = __fmul_rn(iw, ih);
            const float ua0 = __fsub_rn(__fadd_rn(box_area_rn(an.x, an.y, an.z, an.w), area_g), inter);
            if (!(__fmul_rn(inter, p.cull_mul) > ua0)) continue;                 // IoU < neg_thr with margin: cannot matter
            const float v = __fdiv_rn(inter, fmaxf(ua0, 1e-8f));
            if (!(v >= p.neg_thr)) continue;
            const unsigned key = ((__float_as_uint(v) - p.neg_thr_bits + 1u) << 8) | (unsigned)(255 - g);
            atomicMax(keys + a, key);      // result unused: a fire-and-forget RED, nothing waits on it
            atomicOr(mask + (a >> 10), 1u << ((a >> 5) & 31));
        }
    }
}

struct ResolveArgs {
    const uint32_t* keys;      // [B][Ap]
    const uint32_t* mask;      // [B][MW]
    const int32_t* gt_row;     // [B][Gmax] original annotation row of each compacted row
    int32_t* assign;           // [B][A] or null
    int32_t* pos_anchor;       // [B][A]
    int32_t* pos_gt;
    int32_t* npos;
    int B, A, Ap, MW, Gmax, tiles;
    float pos_thr;
    unsigned neg_thr_bits;
};

constexpr int kResolveTile = 4096;   // anchors per CTA: 8 warps x 16 chunks of 32

// Builds the per-image lists of positive anchors (and the optional int32 codes).  Item = (image, 4096-anchor tile); warp
// w owns 16 consecutive 32-anchor chunks = 16 bits of one chunk-mask word and loads the keys of the touched chunks only
// (one key per lane and chunk, all loads issued up front).  Positives are appended with ONE atomic per CTA (block scan).
// Almost all CTAs see an empty mask and leave after the barrier.
__global__ void __launch_bounds__(256) assign_resolve_kernel(const ResolveArgs p) {
    __shared__ int wsum[kWarps];
    __shared__ int s_base;
    const int item = (int)blockIdx.x;
    const int b = item / p.tiles, tile = item - b * p.tiles;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int a_w = tile * kResolveTile + warp * 512;           // first anchor of this warp's 16 chunks
    unsigned m16 = 0u;
    if (a_w < p.Ap) m16 = (__ldg(p.mask + (int64_t)b * p.MW + (a_w >> 10)) >> ((a_w >> 5) & 31)) & 0xffffu;
    // nothing in the whole tile (the usual case) and no dense codes wanted: done
    if (!__syncthreads_or(m16 != 0u) && !p.assign) return;
    unsigned key[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) key[c] = 0u;
    int npos_mine = 0;
    if (m16 != 0u) {               // warp-uniform: most warps hold nothing
#pragma unroll
        for (int c = 0; c < 16; ++c)
            if ((m16 >> c) & 1u) key[c] = __ldcs(p.keys + (int64_t)b * p.Ap + a_w + 32 * c + lane);   // inside Ap: the bit was set
#pragma unroll
        for (int c = 0; c < 16; ++c)
            if (((m16 >> c) & 1u) && key[c] != 0u && key_iou(key[c], p.neg_thr_bits) >= p.pos_thr) ++npos_mine;
    }
    if (p.assign) {                // dense int32 codes (tests, ops.assign-style consumers): every anchor gets one
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const int a = a_w + 32 * c + lane;
            int code = G3D_ASSIGN_NEGATIVE;
            if (key[c] != 0u)
                code = key_iou(key[c], p.neg_thr_bits) >= p.pos_thr ? __ldg(p.gt_row + (int64_t)b * p.Gmax + key_gt(key[c]))
                                                                    : G3D_ASSIGN_IGNORE;
            if (a < p.A) p.assign[(int64_t)b * p.A + a] = code;
        }
    }
    // block-level exclusive scan of the positive counts, one atomic for the CTA
    int incl = npos_mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const int c = wsum[w];
        woff += (w < warp) ? c : 0;
        total += c;
    }
    if (total == 0) return;
    if (tid == 0) s_base = atomicAdd(p.npos + b, total);
    __syncthreads();
    if (npos_mine == 0) return;
    // second pass over the keys this thread still holds: append its positives
    int64_t slot = (int64_t)b * p.A + s_base + woff + incl - npos_mine;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const int a = a_w + 32 * c + lane;
        if (key[c] != 0u && key_iou(key[c], p.neg_thr_bits) >= p.pos_thr) {
            p.pos_anchor[slot] = a;
            p.pos_gt[slot] = key_gt(key[c]);
            ++slot;
        }
    }
}

// =====================================================================================================================
// K3: the streaming classification pass (loss terms + gradient, HBM-bound) + the positive anchors + the reductions
// =====================================================================================================================
constexpr int kChunksPerWarp = 8;                            // 32-row chunks handled by one warp
constexpr int kRowsPerCta = kWarps * 32 * kChunksPerWarp;    // (image, anchor) rows per stream CTA

struct StreamArgs {
    const float* cls;
    CodeSrc src;
    const int32_t* npos;       // [B]
    double* partials;          // [B][T]: classification partial sums, one per stream CTA
    float* dcls;               // [B][A][C]  (GRAD only)
    float* dreg;               // [B][A][R]  (GRAD, GT-centric path only: zero-filled here; null otherwise)
    float g0;                  // upstream gradient of the classification loss that dcls is formed for
    int B, A, C, R, T;
    LossHyper h;
};

// Loss sums of the positives are accumulated in exact fixed point - two int64 limbs per sum, units 2^-20 and 2^-52 -
// with integer atomics: integer addition is associative, so the result does not depend on the (non-deterministic)
// order of the lists.  Range 2^43 per sum, absolute resolution 2^-52 per term.
struct PosArgs {
    const float* reg;
    const float4* anchors;
    const float* gt_tab;       // [B][Gmax][kTabW]
    const int32_t* pos_anchor;
    const int32_t* pos_gt;
    const int32_t* npos;
    const int32_t* gt_count;
    long long* acc;            // [B][4]: reg_hi, reg_lo, vp_hi, vp_lo (forward)
    int32_t* nonfinite;        // [B]: bit 0 / 1 set if a regression / direction term was NaN / Inf / out of range (forward)
    float* dreg;               // gradient rows of the positives (null: losses only)
    int32_t* prev_npos;        // (forward, nullable) [B]: receives the image's number of positives - the rows of dreg this
                               // step wrote, for the next step's re-zeroing (persistent dreg)
    float g1, g2;              // (forward) upstream gradients of the regression / direction losses the rows are formed for
    const float* grad_out;     // [3] device (backward)
    const float* grad_scale;   // [3] device or null (backward): multiplies grad_out (dist.py: local -> global means)
    float e1, e2;              // (backward) the values the forward assumed; have_rows: rows already written for them
    int have_rows;
    int B, A, R, Gmax;
    LossHyper h;
    // reduction (forward)
    const double* partials;    // [B][T]
    int32_t* counters;         // [B] image tickets + [1] batch ticket, zero on entry
    float* losses;             // [4] : cls, reg, vp, number of non-empty images
    float* per_image;          // [B][4]
    int32_t* gt_count_out;     // [B] or null: copy of gt_count for the caller
    double* shard_stats;       // [5] or null: sum cls_j, sum reg_j, sum vp_j (images with GT), B, #images with GT
    int T;
};

// The fully general row (any class count, any gamma, probabilities beyond the series range): scalar, full logf.
// Returns the row's focal sum; writes its gradient row if GRAD.
template <bool GRAD>
__device__ __forceinline__ float stream_row_general(const float* __restrict__ cp, float* __restrict__ dp, int C, int code,
                                                    float s_cls, const LossHyper& h) {
    if (code == kCodeIgnore) {
        if (GRAD) for (int c = 0; c < C; ++c) dp[c] = 0.0f;
        return 0.0f;
    }
    const int pos_cls = (code >= kCodePositive && code != 255) ? code - kCodePositive : -1;
    float acc = 0.0f;
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
        const float pr = __ldg(cp + c);
        acc += focal_term(pr, c == pos_cls, h);
        if (GRAD) dp[c] = s_cls * focal_term_grad(pr, c == pos_cls, h);
    }
    return acc;
}
// out-of-line copy for the rare rows of the fast kernel (keeps its main path small in registers and code)
template <bool GRAD>
__device__ __noinline__ float stream_row_general_ool(const float* cp, float* dp, int code, float s_cls, const LossHyper h) {
    return stream_row_general<GRAD>(cp, dp, 8, code, s_cls, h);
}

// Fix-up of a POSITIVE row processed by the fast path as if it were negative: element e trades its target-0 term for
// the target-1 term.  Rare and divergent: out of line, full logf, arguments and results by value.
struct FixOut { float dacc, ge; };
__device__ __noinline__ FixOut positive_fix(float pe, float s_cls, const LossHyper h) {
    FixOut o;
    o.dacc = focal_term(pe, true, h) - focal_term(pe, false, h);
    o.ge = s_cls * focal_term_grad(pe, true, h);
    return o;
}

__device__ __forceinline__ void fixed_split(float t, long long& hi, long long& lo, bool& bad) {
    bad = !(fabsf(t) < 1.0e12f);            // NaN, Inf, or beyond the 2^43 range of the high limb
    const double td = bad ? 0.0 : (double)t;
    const double h = floor(td * 1048576.0);                              // 2^20
    hi = (long long)h;
    lo = (long long)((td - h * (1.0 / 1048576.0)) * 4503599627370496.0);  // 2^52: in [0, 2^32)
}

// number of images with at least one GT row (all threads of the CTA call this)
__device__ __forceinline__ int count_nonempty(const int32_t* gt_count, int B) {
    int n = 0;
    for (int j0 = 0; j0 < B; j0 += blockDim.x) {
        const int j = j0 + threadIdx.x;
        n += __syncthreads_count(j < B && __ldg(gt_count + j) > 0);
    }
    return n;
}

// Executed by ONE warp - the last CTA of image b: reduce the image's T partials in a fixed order (lane-strided
// accumulation + shuffle tree), then (last image of the batch) the batch means.
template <int VARIANT>
__device__ __forceinline__ void finalize_image(const PosArgs& p, int b) {
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
        const int bad = __ldcg(p.nonfinite + b);         // bit 0: a regression term, bit 1: a direction term was NaN / Inf
        if (bad & 1) tr = __longlong_as_double(0x7ff8000000000000LL);
        if (bad & 2) tv = __longlong_as_double(0x7ff8000000000000LL);
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

template <int VARIANT>
__device__ __forceinline__ void positives_accumulate(const PosArgs& q, int b, float reg_sum, float vp_term);

// rows of the positives straight from global memory: loss terms (FWD) and / or gradient rows (s_reg, s_vp already
// include the normalisers)
template <int VARIANT, bool FWD>
__device__ __forceinline__ void positives_rows(const PosArgs& p, int b, int n, float s_reg, float s_vp, bool write_rows,
                                               int first, int stride) {
    constexpr int R = (VARIANT == G3D_VARIANT_3D) ? 12 : 4;
    const int n_up = (n + 31) & ~31;
    for (int i = first; i < n_up; i += stride) {
        float reg_sum = 0.0f, vp_term = 0.0f;
        if (i < n) {
            const int a = __ldg(p.pos_anchor + (int64_t)b * p.A + i);
            const int g = __ldg(p.pos_gt + (int64_t)b * p.A + i);
            const int64_t row = (int64_t)b * p.A + a;
            float r[R], tab[(VARIANT == G3D_VARIANT_3D) ? 28 : 4], dr[R];
            const float4* rp = reinterpret_cast<const float4*>(p.reg + row * R);
            const float4* tp = reinterpret_cast<const float4*>(p.gt_tab + ((int64_t)b * p.Gmax + g) * kTabW);
#pragma unroll
            for (int k = 0; k < R / 4; ++k) {
                const float4 v = __ldg(rp + k);
                r[4 * k] = v.x; r[4 * k + 1] = v.y; r[4 * k + 2] = v.z; r[4 * k + 3] = v.w;
            }
#pragma unroll
            for (int k = 0; k < ((VARIANT == G3D_VARIANT_3D) ? 7 : 1); ++k) {
                const float4 v = __ldg(tp + k);
                tab[4 * k] = v.x; tab[4 * k + 1] = v.y; tab[4 * k + 2] = v.z; tab[4 * k + 3] = v.w;
            }
            positive_row<VARIANT>(r, tab, __ldg(p.anchors + a), s_reg, s_vp, p.h, write_rows ? dr : nullptr, reg_sum, vp_term);
            if (write_rows) {
                float4* dp = reinterpret_cast<float4*>(p.dreg + row * R);
#pragma unroll
                for (int k = 0; k < R / 4; ++k) dp[k] = make_float4(dr[4 * k], dr[4 * k + 1], dr[4 * k + 2], dr[4 * k + 3]);
            }
        }
        if (FWD) positives_accumulate<VARIANT>(p, b, reg_sum, vp_term);
    }
}

// exact fixed-point accumulation of one warp's (reg_sum, vp_term) values
template <int VARIANT>
__device__ __forceinline__ void positives_accumulate(const PosArgs& q, int b, float reg_sum, float vp_term) {
    long long rh, rl, vh, vl;
    bool bad_r, bad_v;
    fixed_split(reg_sum, rh, rl, bad_r);
    fixed_split(vp_term, vh, vl, bad_v);
    rh = warp_sum_ll(rh); rl = warp_sum_ll(rl);
    if (VARIANT == G3D_VARIANT_3D) { vh = warp_sum_ll(vh); vl = warp_sum_ll(vl); }
    const int any_bad = (__any_sync(0xffffffffu, bad_r) ? 1 : 0) | (__any_sync(0xffffffffu, bad_v) ? 2 : 0);
    if ((threadIdx.x & 31) == 0) {
        unsigned long long* acc = reinterpret_cast<unsigned long long*>(q.acc + 4 * b);
        atomicAdd(acc + 0, (unsigned long long)rh);
        atomicAdd(acc + 1, (unsigned long long)rl);
        if (VARIANT == G3D_VARIANT_3D) {
            atomicAdd(acc + 2, (unsigned long long)vh);
            atomicAdd(acc + 3, (unsigned long long)vl);
        }
        if (any_bad) atomicOr(q.nonfinite + b, any_bad);
    }
}

// Programmatic dependent launch (PDL): K4 is launched while K3 is still running - its CTAs move into the SM slots the
// last wave of the sweep leaves free - and does everything that does not need K3's results (the rows of the positives:
// dependent loads, ~1000 instructions of arithmetic, gradient rows, fixed-point sums) before it waits for K3 to finish.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_primary() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// K4.  grid (x, B): loss sums of the positives, their gradient rows for the expected upstream gradients (g1, g2); then,
// once K3 is complete, the reductions: the last CTA of an image reduces the image, the last image the batch.
template <int VARIANT>
__global__ void __launch_bounds__(128) positives_kernel(const __grid_constant__ PosArgs p) {
    __shared__ int s_last;
    const int b = blockIdx.y;
    const int n = min(__ldg(p.npos + b), p.A);
    const float npos = (float)n;
    float s_reg = 0.0f, s_vp = 0.0f;
    if (p.prev_npos && blockIdx.x == 0 && threadIdx.x == 0) p.prev_npos[b] = p.dreg ? n : 0;
    if (p.dreg) {
        const float per_pos = (VARIANT == G3D_VARIANT_3D) ? 20.0f : 4.0f;
        s_reg = p.g1 / ((float)p.B * per_pos * npos);
        if (VARIANT == G3D_VARIANT_3D) s_vp = p.g2 / ((float)count_nonempty(p.gt_count, p.B) * npos * 3.0f);
    }
    positives_rows<VARIANT, true>(p, b, n, s_reg, s_vp, p.dreg != nullptr, blockIdx.x * blockDim.x + threadIdx.x,
                                  gridDim.x * blockDim.x);
    __threadfence();          // every warp's fixed-point atomics are ordered before the CTA's ticket
    pdl_wait_primary();       // from here on K3's partial sums are complete and visible
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(p.counters + b, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (s_last && threadIdx.x < 32) finalize_image<VARIANT>(p, b);
}

struct StreamSmem {
    double dred[kWarps];
    int arrive;
};

// zero rows of dreg in a chunk that holds keys: ordinary stores, lane by lane; the lanes of positive rows write nothing
// (the positives CTAs own those rows)
__device__ __forceinline__ void zero_row_unless_positive(float* dreg_row, int R, bool valid, int code) {
    if (!valid || code >= kCodePositive) return;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* d4 = reinterpret_cast<float4*>(dreg_row);
    if (R == 12) { st_stream(d4, z); st_stream(d4 + 1, z); st_stream(d4 + 2, z); }
    else st_stream(d4, z);
}

// C == 8, gamma == 2.  grid (T, B): CTA x streams rows [x * kRowsPerCta, ...) of image b - lane l of a warp owns row l
// of a 32-row chunk: one 256-bit load, one code, one 256-bit store per lane and chunk; the next chunk's loads are in
// flight while the current one is evaluated; KEYS: the codes come from the GT-centric keys, read only for chunks whose
// mask bit is set, and the warp zero-fills its 256 rows of dreg (bulk copies for chunks without keys); else from the byte
// codes.
template <bool GRAD, bool KEYS>
__global__ void __launch_bounds__(kTile, 4) focal_stream8_kernel(const __grid_constant__ StreamArgs p) {
    __shared__ StreamSmem sm;
    __shared__ __align__(128) unsigned char ztile[kZeroTile];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int bx = (int)blockIdx.x;
    const bool fill = GRAD && KEYS && p.dreg != nullptr;
    pdl_launch_dependents();      // K4 (positives + reductions) may move in as soon as every CTA of this grid has started
    if (tid == 0) sm.arrive = 0;
    if (fill) zero_tile_init(ztile); else __syncthreads();
    const uint32_t tile = (uint32_t)__cvta_generic_to_shared(ztile);
    const float npos = (float)__ldg(p.npos + b);
    const float s_cls = GRAD ? p.g0 / ((float)p.B * fmaxf(npos, 1.0f)) : 0.0f;
    const float scale_neg = p.h.one_minus_alpha * s_cls;
    const int wa0 = (bx * kWarps + warp) * (32 * kChunksPerWarp);   // first anchor of this warp (multiple of 256)
    const int nchunks = max(0, min(kChunksPerWarp, (p.A - wa0 + 31) >> 5));
    float cls_acc = 0.0f;
    if (nchunks > 0) {
        int a = wa0 + lane;
        const float* cp = p.cls + ((int64_t)b * p.A + a) * 8;
        float* dp = p.dcls + ((int64_t)b * p.A + a) * 8;
        const int64_t kbase = (int64_t)b * p.src.Ap;
        unsigned m8 = 0u;       // the 8 chunk-mask bits of this warp
        if (KEYS) m8 = (__ldg(p.src.mask + (int64_t)b * p.src.MW + (wa0 >> 10)) >> ((wa0 >> 5) & 31)) & 0xffu;
        char* dreg_w = nullptr;                 // this warp's rows of dreg
        const int rb = p.R * 4;
        if (fill) {
            dreg_w = reinterpret_cast<char*>(p.dreg) + ((int64_t)b * p.A + wa0) * rb;
            if (m8 == 0u && lane == 0) bulk_zero(dreg_w, (long long)min(32 * kChunksPerWarp, p.A - wa0) * rb, tile);
        }
        float cur[8], nxt[8];
        unsigned raw = 0u, nraw = 0u;           // key (KEYS) or byte code of the row
#pragma unroll
        for (int e = 0; e < 8; ++e) { cur[e] = 0.0f; nxt[e] = 0.0f; }
        if (a < p.A) {
            ld_row8<!GRAD>(cp, cur);
            if (KEYS) { if (m8 & 1u) raw = __ldg(p.src.keys + kbase + a); }
            else raw = __ldg(p.src.code8 + kbase + a);
        }
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
            const bool valid = a < p.A;
            nraw = 0u;
            if (c + 1 < nchunks && a + 32 < p.A) {
                ld_row8<!GRAD>(cp + 32 * 8, nxt);
                if (KEYS) { if ((m8 >> (c + 1)) & 1u) nraw = __ldg(p.src.keys + kbase + a + 32); }
                else nraw = __ldg(p.src.code8 + kbase + a + 32);
            }
            int code = kCodeNegative;
            if (raw != 0u) code = KEYS ? code_of_key(p.src, b, raw) : (int)raw;
            if (fill && m8 != 0u) {
                if ((m8 >> c) & 1u) zero_row_unless_positive(reinterpret_cast<float*>(dreg_w + (int64_t)(32 * c + lane) * rb), p.R, valid, code);
                else if (lane == 0) bulk_zero(dreg_w + (int64_t)(32 * c) * rb, (long long)min(32, p.A - (wa0 + 32 * c)) * rb, tile);
            }
            // every element as if its anchor were negative (the overwhelmingly common case); an ignored (or absent) row
            // gets scale 0 and contributes nothing
            float g[8], pmax;
            const bool ign = code == kCodeIgnore || !valid;
            float acc = focal_neg_row8<GRAD>(cur, p.h.pmin, p.h.pmax, ign ? 0.0f : scale_neg, g, pmax);
            acc = ign ? 0.0f : p.h.one_minus_alpha * acc;
            if (pmax >= 0.25f && valid) {
                // a probability beyond the series range: the whole row again with the full logf (handles its code too)
                acc = stream_row_general_ool<GRAD>(cp, dp, code, s_cls, p.h);
            } else {
                if (GRAD && valid) st_row8(dp, g);
                if (code >= kCodePositive && code != 255) {
                    const int e = code - kCodePositive;         // < 8 because C == 8
                    const FixOut o = positive_fix(__ldg(cp + e), s_cls, p.h);
                    acc += o.dacc;
                    if (GRAD) dp[e] = o.ge;                      // after the row store of the same thread
                }
            }
            cls_acc += acc;
#pragma unroll
            for (int e = 0; e < 8; ++e) cur[e] = nxt[e];
            raw = nraw;
            a += 32; cp += 32 * 8; dp += 32 * 8;
        }
    }
    const float cs = warp_sum_f(cls_acc);     // FP32 inside the warp (<= 2048 terms), FP64 from here on
    if (fill && lane == 0) bulk_wait_read();  // the zero tile must outlive the copies that read it
    // ---- the last warp of the CTA to get here (shared-memory ticket, no block barrier: finished warps retire
    // immediately) combines the 8 warp partials in warp order
    int arrived = 0;
    if (lane == 0) {
        sm.dred[warp] = (double)cs;
        __threadfence_block();
        arrived = atomicAdd(&sm.arrive, 1);
    }
    arrived = __shfl_sync(0xffffffffu, arrived, 0);
    if (arrived != kWarps - 1) return;
    __threadfence_block();
    if (lane == 0) {
        const volatile double* dr = sm.dred;
        double tc = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) tc += dr[w];
        __stcg(p.partials + (int64_t)b * p.T + bx, tc);
    }
}

// any class count / any gamma: one thread per row, scalar accesses; same fill
template <bool GRAD>
__global__ void __launch_bounds__(kTile, 4) focal_stream_generic_kernel(const __grid_constant__ StreamArgs p) {
    __shared__ StreamSmem sm;
    __shared__ __align__(128) unsigned char ztile[kZeroTile];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int bx = (int)blockIdx.x;
    pdl_launch_dependents();
    const bool keys = p.src.code8 == nullptr;
    const bool fill = GRAD && keys && p.dreg != nullptr;
    if (fill) zero_tile_init(ztile);
    const uint32_t tile = (uint32_t)__cvta_generic_to_shared(ztile);
    const float npos = (float)__ldg(p.npos + b);
    const float s_cls = GRAD ? p.g0 / ((float)p.B * fmaxf(npos, 1.0f)) : 0.0f;
    const int wa0 = (bx * kWarps + warp) * (32 * kChunksPerWarp);
    const int rb = p.R * 4;
    unsigned m8 = 0u;
    if (keys && wa0 < p.A) m8 = (__ldg(p.src.mask + (int64_t)b * p.src.MW + (wa0 >> 10)) >> ((wa0 >> 5) & 31)) & 0xffu;
    char* dreg_w = fill ? reinterpret_cast<char*>(p.dreg) + ((int64_t)b * p.A + wa0) * rb : nullptr;
    float cls_acc = 0.0f;
#pragma unroll 1
    for (int c = 0; c < kChunksPerWarp; ++c) {
        const int a = wa0 + 32 * c + lane;
        if (wa0 + 32 * c >= p.A) break;
        const bool valid = a < p.A;
        int code = kCodeNegative;
        if (valid) code = code_of_row(p.src, b, a);
        if (fill) {
            if ((m8 >> c) & 1u) zero_row_unless_positive(reinterpret_cast<float*>(dreg_w + (int64_t)(32 * c + lane) * rb), p.R, valid, code);
            else if (lane == 0) bulk_zero(dreg_w + (int64_t)(32 * c) * rb, (long long)min(32, p.A - (wa0 + 32 * c)) * rb, tile);
        }
        if (valid) {
            const int64_t row = (int64_t)b * p.A + a;
            cls_acc += stream_row_general<GRAD>(p.cls + row * p.C, p.dcls + row * p.C, p.C, code, s_cls, p.h);
        }
    }
    const float cs = warp_sum_f(cls_acc);
    if (lane == 0) sm.dred[warp] = (double)cs;
    if (fill && lane == 0) bulk_wait_read();
    __syncthreads();
    if (tid == 0) {
        double tc = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) tc += sm.dred[w];
        __stcg(p.partials + (int64_t)b * p.T + bx, tc);
    }
}

// =====================================================================================================================
// backward: verify the expectation; recompute what does not hold
// =====================================================================================================================
struct ClsGradArgs {
    const float* cls;
    const float* grad_out;   // [3] device
    const float* grad_scale; // [3] device or null
    const int32_t* npos;     // [B]
    CodeSrc src;
    float* dcls;
    float* dreg;
    float e0;                // upstream classification gradient dcls was formed for (valid if have_dcls)
    int have_dcls;           // dcls already holds the gradient for e0 and dreg is already zero-filled
    int B, A, C, R, T;
    LossHyper h;
};

// Role 1 (the first n_cls CTAs): grid-stride over (image, 256-row tile) items, one thread per row: the classification
// gradient for the real upstream gradient.  Role 2 (the other CTAs, kPosCtas per image): the gradient rows of the positives.
// Both compare the real upstream gradients with the ones the forward assumed and leave at once when they agree - the usual
// training step - so the whole backward is one launch that retires in one wave.
constexpr int kPosCtas = 8;
template <int VARIANT>
__global__ void __launch_bounds__(256, 4) focal_bwd_kernel(const __grid_constant__ ClsGradArgs p,
                                                           const __grid_constant__ PosArgs q, int n_cls, int roles) {
    if ((int)blockIdx.x < n_cls) {
        if (!(roles & 1)) return;
        const float go0 = __ldg(p.grad_out + 0) * (p.grad_scale ? __ldg(p.grad_scale + 0) : 1.0f);
        if (p.have_dcls && go0 == p.e0) return;
        const int lane = threadIdx.x & 31;
        const int64_t items = (int64_t)p.T * p.B;
        for (int64_t item = blockIdx.x; item < items; item += n_cls) {
            const int b = (int)(item / p.T);
            const int a = (int)(item - (int64_t)b * p.T) * 256 + threadIdx.x;
            const bool valid = a < p.A;
            const int64_t row = (int64_t)b * p.A + a;
            if (!p.have_dcls) {
                const int nrows = min(32, p.A - (a - lane));
                if (nrows > 0) {
                    if (p.R == 12) zero_rows<12>(p.dreg + (row - lane) * 12, nrows, lane);
                    else           zero_rows<4>(p.dreg + (row - lane) * 4, nrows, lane);
                }
            }
            if (!valid) continue;
            const int code = code_of_row(p.src, b, a);
            const float npos = (float)__ldg(p.npos + b);
            const float s_cls = go0 / ((float)p.B * fmaxf(npos, 1.0f));
            stream_row_general<true>(p.cls + row * p.C, p.dcls + row * p.C, p.C, code, s_cls, p.h);
        }
        return;
    }
    if (!(roles & 2)) return;
    const float go1 = __ldg(q.grad_out + 1) * (q.grad_scale ? __ldg(q.grad_scale + 1) : 1.0f);
    const float go2 = (VARIANT == G3D_VARIANT_3D) ? __ldg(q.grad_out + 2) * (q.grad_scale ? __ldg(q.grad_scale + 2) : 1.0f) : 0.0f;
    if (q.have_rows && go1 == q.e1 && (VARIANT != G3D_VARIANT_3D || go2 == q.e2)) return;
    const int idx = (int)blockIdx.x - n_cls;
    const int b = idx / kPosCtas, x = idx - b * kPosCtas;
    const int n = min(__ldg(q.npos + b), q.A);
    const float npos = (float)n;
    const float per_pos = (VARIANT == G3D_VARIANT_3D) ? 20.0f : 4.0f;
    const float s_reg = go1 / ((float)q.B * per_pos * npos);
    float s_vp = 0.0f;
    if (VARIANT == G3D_VARIANT_3D) s_vp = go2 / ((float)count_nonempty(q.gt_count, q.B) * npos * 3.0f);
    positives_rows<VARIANT, false>(q, b, n, s_reg, s_vp, true, x * 256 + threadIdx.x, kPosCtas * 256);
}

// =====================================================================================================================
// workspace
// =====================================================================================================================
struct FocalWorkspace {
    float4* gt_box;
    int32_t* gt_row;
    int32_t* gt_cls;
    int32_t* gt_count;
    float* gt_tab;       // [B][Gmax][kTabW]
    double* partials;    // [B][T]
    int32_t* pos_anchor; // [B][A]
    int32_t* pos_gt;     // [B][A]
    uint32_t* keys;      // [B][Ap]  GT-centric assignment: best (IoU, GT index) key per anchor
    uint32_t* mask;      // [B][MW]  GT-centric assignment: chunk mask (directly behind the keys: one fill covers both)
    uint8_t* code8;      // [B][Ap]  anchor-centric assignment: byte codes
    int32_t* counters;   // zeroed per call: [B] image tickets, [1] batch ticket, [B] npos, [B] nonfinite flags, then
                         // (8-byte aligned) [B][4] int64 fixed-point sums
    int64_t n_counters;  // number of int32 words to zero
    int32_t* prev_npos;  // [B] NOT zeroed per call: rows of dreg written by the last gradient step (persistent dreg)
    int64_t Ap;          // per-image pitch of keys / code8 (multiple of 32)
    int64_t MW;          // chunk-mask words per image
    int64_t key_fill_bytes;   // keys + mask: zeroed together
    int64_t bytes;
};

static FocalWorkspace carve(void* base, int64_t B, int64_t A, int64_t Gmax) {
    FocalWorkspace w;
    const int64_t T = ceil_div(A, kRowsPerCta);
    w.Ap = align_up(A, 32);
    int64_t off = 0;
    char* p = (char*)base;
    w.gt_box = (float4*)(p + off);   off += align_up(B * Gmax * 16, 256);
    w.gt_row = (int32_t*)(p + off);  off += align_up(B * Gmax * 4, 256);
    w.gt_cls = (int32_t*)(p + off);  off += align_up(B * Gmax * 4, 256);
    w.gt_count = (int32_t*)(p + off); off += align_up(B * 4, 256);
    w.gt_tab = (float*)(p + off);    off += align_up(B * Gmax * kTabW * 4, 256);
    w.partials = (double*)(p + off); off += align_up(B * T * 8, 256);
    w.pos_anchor = (int32_t*)(p + off); off += align_up(B * A * 4, 256);
    w.pos_gt = (int32_t*)(p + off);  off += align_up(B * A * 4, 256);
    w.MW = ceil_div(w.Ap, 1024);
    w.keys = (uint32_t*)(p + off);   off += align_up(B * w.Ap * 4, 256);
    w.mask = (uint32_t*)(p + off);   off += align_up(B * w.MW * 4, 1024);
    w.key_fill_bytes = (char*)w.mask - (char*)w.keys + align_up(B * w.MW * 4, 16);
    w.code8 = (uint8_t*)(p + off);   off += align_up(B * w.Ap, 256);
    w.counters = (int32_t*)(p + off);
    const int64_t head = align_up(3 * B + 1, 2);          // int32 words before the int64 sums
    w.n_counters = head + 8 * B;
    off += align_up(w.n_counters * 4, 256);
    w.prev_npos = (int32_t*)(p + off); off += align_up(B * 4, 256);
    w.bytes = off;
    return w;
}
static inline int32_t* ws_npos(const FocalWorkspace& w, int64_t B) { return w.counters + B + 1; }
static CodeSrc make_src(const FocalWorkspace& w, bool gt_centric, int64_t Gmax, int64_t C, const LossHyper& h) {
    CodeSrc s;
    s.code8 = gt_centric ? nullptr : w.code8;
    s.keys = w.keys; s.mask = w.mask; s.gt_cls = w.gt_cls;
    s.Ap = (int)w.Ap; s.MW = (int)w.MW; s.Gmax = (int)Gmax; s.C = (int)C;
    s.pos_thr = h.pos_thr; s.neg_thr_bits = h.neg_thr_bits;
    return s;
}
static inline int32_t* ws_nonfinite(const FocalWorkspace& w, int64_t B) { return w.counters + 2 * B + 1; }
static inline long long* ws_acc(const FocalWorkspace& w, int64_t B) { return (long long*)(w.counters + align_up(3 * B + 1, 2)); }

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

// The same reduction WITHOUT a collective library call: every rank stores its 5 statistics straight into the exchange
// buffer of every peer over NVLink (peer-mapped symmetric memory), publishes them with a system-scope fence + an epoch
// flag, and waits for the epoch flags of the others in its own buffer - one launch of one warp per rank instead of an
// NCCL all-gather plus the combine kernel, and nothing between the loss kernels and it but stream order.
// Exchange buffer of a rank (doubles): [2 parities][world][8] = {5 statistics, epoch flag (as bits), 2 unused}, then one
// local epoch counter.  Two parities: a peer can be at most one call ahead of this rank (its next call needs this rank's
// statistics of that call), so the slot it overwrites then is never the one still being read.  Zero on first use.
__global__ void exchange_shard_stats_kernel(const double* __restrict__ stats, const unsigned long long* __restrict__ peers,
                                            int world, int rank, float* __restrict__ losses, float* __restrict__ scale) {
    const int lane = threadIdx.x;
    double* mine = reinterpret_cast<double*>(peers[rank]);
    unsigned long long* epoch_ctr = reinterpret_cast<unsigned long long*>(mine + 2 * world * 8);
    const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(epoch_ctr) + 1ull;
    const int parity = (int)(e & 1ull);
    double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    bool ok = true;
    if (lane < world) {
        volatile double* dst = reinterpret_cast<volatile double*>(peers[lane]) + ((int64_t)parity * world + rank) * 8;
#pragma unroll
        for (int k = 0; k < 5; ++k) dst[k] = stats[k];
        __threadfence_system();
        reinterpret_cast<volatile unsigned long long*>(dst)[5] = e;
        const volatile double* src = mine + ((int64_t)parity * world + lane) * 8;
        const long long t0 = clock64();
        while (reinterpret_cast<const volatile unsigned long long*>(src)[5] != e) {
            if (clock64() - t0 > 6000000000LL) { ok = false; break; }       // ~3 s: a peer never arrived - do not hang the GPU
            __nanosleep(100);
        }
        __threadfence_system();
#pragma unroll
        for (int k = 0; k < 5; ++k) v[k] = src[k];
    }
    ok = __all_sync(0xffffffffu, ok);
    double t[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, bl = 0.0, nl = 0.0;
    for (int r = 0; r < world; ++r) {                       // fixed (rank) summation order: the same bits on every rank
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const double x = __shfl_sync(0xffffffffu, v[k], r);
            t[k] += x;
            if (r == rank && k == 3) bl = x;
            if (r == rank && k == 4) nl = x;
        }
    }
    if (lane == 0) {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        losses[0] = ok ? (float)(t[0] / t[3]) : (float)nan;
        losses[1] = ok ? (float)(t[1] / t[3]) : (float)nan;
        losses[2] = ok ? (float)(t[2] / t[4]) : (float)nan;
        scale[0] = scale[1] = (float)(bl / t[3]);
        scale[2] = nl > 0.0 ? (float)(nl / t[4]) : 0.0f;
        *reinterpret_cast<volatile unsigned long long*>(epoch_ctr) = e;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------------------------------
// hyper_host (nullable): G3D_HYPER_COUNT floats {alpha, gamma, pos_thr, neg_thr, beta, top_weighting, clamp_min,
// clamp_max}; null = the reference's constants (losses.py:28-30,56,121,124,346-348)
static int make_hyper(const float* hyper_host, LossHyper& h) {
    float alpha = 0.25f, gamma = 2.0f, pos_thr = 0.5f, neg_thr = 0.4f, top_w = 0.5f;
    double beta = 1.0 / 9.0;
    float pmin = (float)1e-4, pmax = (float)(1.0 - 1e-4);
    if (hyper_host) {
        alpha = hyper_host[0]; gamma = hyper_host[1]; pos_thr = hyper_host[2]; neg_thr = hyper_host[3];
        if (hyper_host[4] != (float)(1.0 / 9.0)) beta = (double)hyper_host[4];
        top_w = hyper_host[5]; pmin = hyper_host[6]; pmax = hyper_host[7];
    }
    G3D_REQUIRE(alpha >= 0.0f && alpha <= 1.0f && gamma >= 0.0f && beta > 0.0, "alpha in [0,1], gamma >= 0, beta > 0 expected");
    G3D_REQUIRE(neg_thr > 0.0f && neg_thr <= pos_thr && pos_thr <= 1.0f, "0 < neg_thr <= pos_thr <= 1 expected");
    G3D_REQUIRE(pmin > 0.0f && pmin < pmax && pmax < 1.0f, "0 < clamp_min < clamp_max < 1 expected");
    h.alpha = alpha;
    h.one_minus_alpha = 1.0f - alpha;
    h.gamma = gamma;
    h.gamma_is_two = gamma == 2.0f ? 1 : 0;
    h.pmin = pmin; h.pmax = pmax;
    h.pos_thr = pos_thr; h.neg_thr = neg_thr;
    h.cull_mul = (float)(1.04 / (double)neg_thr);          // 2.6 for 0.4: IoU <= 0.3847
    h.group_cull = 0.95f * neg_thr;                        // 0.38
    const double t = 0.9625 * (double)neg_thr;             // 0.385
    h.win_q = (float)(t / (1.0 + t));
    memcpy(&h.neg_thr_bits, &neg_thr, 4);
    h.sl1_beta = (float)beta;
    h.sl1_quad = (float)(0.5 / beta);                      // 4.5
    h.sl1_off = (float)(0.5 * beta);                       // 0.5 / 9
    h.sl1_slope = (float)(1.0 / beta);                     // 9
    h.top_w = top_w;
    return G3D_OK;
}

static int check_focal_shapes(int64_t B, int64_t A, int64_t C, int64_t R, int64_t Gmax, int64_t W, int variant) {
    G3D_REQUIRE(variant == G3D_VARIANT_2D || variant == G3D_VARIANT_3D, "unknown variant");
    G3D_REQUIRE(B >= 1 && A >= 1 && C >= 1 && Gmax >= 0, "sizes must be positive");
    G3D_REQUIRE(B <= 65535 && A < ((int64_t)1 << 31) - kRowsPerCta && Gmax < (1 << 30) && C < (1 << 20), "size out of range");
    G3D_REQUIRE(B * align_up(A, 32) < ((int64_t)1 << 40), "B x A out of range");
    if (variant == G3D_VARIANT_3D) {
        G3D_REQUIRE(R == 12, "3D variant needs 12 regression outputs per anchor");
        G3D_REQUIRE(W >= 21, "3D variant needs >= 21 annotation columns");
    } else {
        G3D_REQUIRE(R == 4, "2D variant needs 4 regression outputs per anchor");
        G3D_REQUIRE(W >= 5, "2D variant needs >= 5 annotation columns");
    }
    return G3D_OK;
}


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

static bool use_gt_centric(const double* pyramid_host, int64_t A, int64_t C, int64_t Gmax, const LossHyper& h, Pyramid& pyr) {
    return Gmax >= 1 && Gmax <= 256 && h.neg_thr >= 0.3f && C <= 250 && !g_force_anchor_centric.load() &&
           load_pyramid(pyramid_host, A, pyr);
}

}  // namespace g3d

using namespace g3d;

extern "C" int g3d_set_tuning(int key, int64_t value) {
    switch (key) {
        case G3D_TUNE_PDL: g_pdl.store((int)value); return G3D_OK;
        case G3D_TUNE_FORCE_ANCHOR_CENTRIC: g_force_anchor_centric.store((int)value); return G3D_OK;
        default: set_error("g3d_set_tuning: unknown key %d", key); return G3D_ERR_INVALID;
    }
}

extern "C" int g3d_combine_shard_stats(const double* gathered, int64_t world, int64_t rank, float* losses, float* scale,
                                       int device, void* stream) {
    G3D_REQUIRE(gathered && losses && scale, "null pointer");
    G3D_REQUIRE(world >= 1 && rank >= 0 && rank < world && world < (1 << 20), "bad world / rank");
    G3D_GUARD(device);
    combine_shard_stats_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(gathered, (int)world, (int)rank, losses, scale);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int64_t g3d_exchange_buffer_doubles(int64_t world) { return world < 1 ? G3D_ERR_INVALID : 2 * world * 8 + 8; }

extern "C" int g3d_exchange_shard_stats(const double* shard_stats, const void* peer_ptrs_dev, int64_t world, int64_t rank,
                                        float* losses, float* scale, int device, void* stream) {
    G3D_REQUIRE(shard_stats && peer_ptrs_dev && losses && scale, "null pointer");
    G3D_REQUIRE(world >= 1 && world <= 32 && rank >= 0 && rank < world, "bad world / rank (one warp exchanges: world <= 32)");
    G3D_GUARD(device);
    exchange_shard_stats_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(shard_stats, (const unsigned long long*)peer_ptrs_dev,
                                                                    (int)world, (int)rank, losses, scale);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int64_t g3d_focal_workspace_bytes(int64_t B, int64_t A, int64_t Gmax) {
    if (B < 0 || A < 0 || Gmax < 0) return G3D_ERR_INVALID;
    return carve(nullptr, B, A, Gmax).bytes;
}

extern "C" int g3d_focal_loss_fwd_bwd(const float* cls, const float* reg, const float* anchors, const float* ann,
                                      int64_t B, int64_t A, int64_t C, int64_t R, int64_t Gmax, int64_t W, int variant,
                                      const float* hyper_host, const float* grad_expected_host, float* losses,
                                      float* per_image, int32_t* assign, int32_t* gt_count_out, double* shard_stats,
                                      float* dcls, float* dreg, void* workspace, int64_t workspace_bytes,
                                      const double* pyramid_host, void* const* trace_events, int n_trace_events,
                                      int dreg_state, int device, void* stream) {
    int rc = check_focal_shapes(B, A, C, R, Gmax, W, variant);
    if (rc != G3D_OK) return rc;
    G3D_REQUIRE(cls && reg && anchors && losses && per_image && workspace, "null pointer");
    G3D_REQUIRE(dreg_state == G3D_DREG_UNDEFINED || dreg_state == G3D_DREG_CLEAN, "unknown dreg_state");
    G3D_REQUIRE(Gmax == 0 || ann, "null annotations");
    G3D_REQUIRE((dcls == nullptr) == (dreg == nullptr), "dcls and dreg must both be given or both be null");
    G3D_REQUIRE((dcls == nullptr) || grad_expected_host, "gradient buffers need grad_expected_host[3]");
    G3D_REQUIRE(n_trace_events >= 0 && n_trace_events <= 6 && (n_trace_events == 0 || trace_events), "bad trace events");
    LossHyper h;
    rc = make_hyper(hyper_host, h);
    if (rc != G3D_OK) return rc;
    FocalWorkspace w = carve(workspace, B, A, Gmax);
    G3D_REQUIRE(workspace_bytes >= w.bytes, "workspace too small (see g3d_focal_workspace_bytes)");
    G3D_REQUIRE(((uintptr_t)cls % 32) == 0 && ((uintptr_t)reg % 16) == 0 && ((uintptr_t)anchors % 16) == 0 &&
                    ((uintptr_t)workspace % 256) == 0 && ((uintptr_t)per_image % 16) == 0 && ((uintptr_t)dcls % 32) == 0 &&
                    ((uintptr_t)dreg % 16) == 0 && ((uintptr_t)assign % 4) == 0,
                "cls/dcls must be 32-byte, reg/dreg/anchors/per_image 16-byte and the workspace 256-byte aligned");
    G3D_GUARD(device);
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = sm_count(device);
    const bool grad = dcls != nullptr;
    int32_t* npos = ws_npos(w, B);
    auto trace = [&](int i) -> cudaError_t {
        return (i < n_trace_events) ? cudaEventRecord((cudaEvent_t)trace_events[i], st) : cudaSuccess;
    };

    Pyramid pyr;
    const bool gt_centric = use_gt_centric(pyramid_host, A, C, Gmax, h, pyr);

    G3D_CUDA(trace(0));
    // ---- K0
    PrologueArgs pr;
    pr.ann = ann; pr.gt_box = w.gt_box; pr.gt_row = w.gt_row; pr.gt_cls = w.gt_cls; pr.gt_tab = w.gt_tab;
    pr.gt_count = w.gt_count; pr.zero_ptr = w.counters; pr.zero_n = (int)w.n_counters;
    pr.B = (int)B; pr.Gmax = (int)Gmax; pr.W = (int)W; pr.variant = variant;
    pr.fill = FillSlice{nullptr, 0};
    pr.nfill = 0;
    // persistent dreg: zero the previous step's positive rows instead of filling all of it (GT-centric path; the
    // anchor-centric kernel writes every row anyway)
    const bool dreg_clean = grad && gt_centric && dreg_state == G3D_DREG_CLEAN;
    pr.dreg = dreg_clean ? dreg : nullptr; pr.prev_npos = w.prev_npos; pr.pos_anchor = w.pos_anchor;
    pr.A = (int)A; pr.R = (int)R;
    if (gt_centric) {
        pr.nfill = sms * 8;      // fill CTAs zero the keys and the chunk mask
        pr.fill = FillSlice{(char*)w.keys, (long long)w.key_fill_bytes};
    }
    loss_prologue_kernel<<<(unsigned)(B + pr.nfill), kTile, 0, st>>>(pr);
    G3D_LAUNCH_CHECK();
    G3D_CUDA(trace(1));

    // ---- K1 (+ K2)
    if (gt_centric) {
        PairArgs pa;
        pa.anchors = (const float4*)anchors; pa.gt_box = w.gt_box; pa.gt_count = w.gt_count; pa.keys = w.keys;
        pa.mask = w.mask;
        pa.B = (int)B; pa.A = (int)A; pa.Ap = (int)w.Ap; pa.MW = (int)w.MW; pa.Gmax = (int)Gmax;
        pa.win_q = h.win_q; pa.cull_mul = h.cull_mul; pa.neg_thr = h.neg_thr; pa.neg_thr_bits = h.neg_thr_bits;
        assign_pairs_kernel<<<(unsigned)ceil_div(B * Gmax, 8), 256, 0, st>>>(pa, pyr);
        G3D_LAUNCH_CHECK();
        G3D_CUDA(trace(2));
        ResolveArgs ra;
        ra.keys = w.keys; ra.mask = w.mask; ra.gt_row = w.gt_row; ra.assign = assign;
        ra.pos_anchor = w.pos_anchor; ra.pos_gt = w.pos_gt; ra.npos = npos;
        ra.B = (int)B; ra.A = (int)A; ra.Ap = (int)w.Ap; ra.MW = (int)w.MW; ra.Gmax = (int)Gmax;
        ra.tiles = (int)ceil_div(w.Ap, kResolveTile);
        ra.pos_thr = h.pos_thr; ra.neg_thr_bits = h.neg_thr_bits;
        assign_resolve_kernel<<<(unsigned)((int64_t)ra.tiles * B), 256, 0, st>>>(ra);
        G3D_LAUNCH_CHECK();
        G3D_CUDA(trace(3));
    } else {
        AssignCodesArgs q;
        q.anchors = (const float4*)anchors; q.gt_box = w.gt_box; q.gt_row = w.gt_row; q.gt_cls = w.gt_cls;
        q.gt_count = w.gt_count; q.code8 = w.code8; q.assign = assign; q.npos = npos; q.pos_anchor = w.pos_anchor;
        q.pos_gt = w.pos_gt; q.dreg = dreg; q.B = (int)B; q.A = (int)A; q.Ap = (int)w.Ap; q.Gmax = (int)Gmax;
        q.R = (int)R; q.C = (int)C;
        q.pos_thr = h.pos_thr; q.neg_thr = h.neg_thr; q.cull_mul = h.cull_mul; q.group_cull = h.group_cull;
        const dim3 agrid((unsigned)ceil_div(A, kTile), (unsigned)ceil_div(B, kImgPerCta));
        assign_codes_kernel<<<agrid, kTile, 0, st>>>(q);
        G3D_LAUNCH_CHECK();
        G3D_CUDA(trace(2));
        G3D_CUDA(trace(3));
    }

    // ---- K3: stream CTAs + positives CTAs + reductions
    StreamArgs p;
    p.cls = cls; p.npos = npos; p.partials = w.partials; p.dcls = dcls;
    p.dreg = (gt_centric && grad && !dreg_clean) ? dreg : nullptr;   // (the anchor-centric kernel wrote the zeros already)
    p.src = make_src(w, gt_centric, Gmax, C, h);
    p.g0 = grad ? grad_expected_host[0] : 0.0f;
    p.B = (int)B; p.A = (int)A; p.C = (int)C; p.R = (int)R; p.T = (int)ceil_div(A, kRowsPerCta);
    p.h = h;
    PosArgs pp;
    pp.reg = reg; pp.anchors = (const float4*)anchors; pp.gt_tab = w.gt_tab; pp.pos_anchor = w.pos_anchor;
    pp.pos_gt = w.pos_gt; pp.npos = npos; pp.gt_count = w.gt_count; pp.acc = ws_acc(w, B); pp.nonfinite = ws_nonfinite(w, B);
    pp.dreg = dreg; pp.g1 = grad ? grad_expected_host[1] : 0.0f; pp.g2 = grad ? grad_expected_host[2] : 0.0f;
    pp.prev_npos = grad ? w.prev_npos : nullptr;
    pp.grad_out = nullptr; pp.grad_scale = nullptr; pp.e1 = pp.e2 = 0.0f; pp.have_rows = 0;
    pp.B = (int)B; pp.A = (int)A; pp.R = (int)R; pp.Gmax = (int)Gmax; pp.h = h;
    pp.partials = w.partials; pp.counters = w.counters; pp.losses = losses; pp.per_image = per_image;
    pp.gt_count_out = gt_count_out; pp.shard_stats = shard_stats; pp.T = p.T;
    const dim3 sgrid((unsigned)p.T, (unsigned)B);
    if (C == 8 && h.gamma_is_two) {
        if (gt_centric) {
            if (grad) focal_stream8_kernel<true, true><<<sgrid, kTile, 0, st>>>(p);
            else      focal_stream8_kernel<false, true><<<sgrid, kTile, 0, st>>>(p);
        } else {
            if (grad) focal_stream8_kernel<true, false><<<sgrid, kTile, 0, st>>>(p);
            else      focal_stream8_kernel<false, false><<<sgrid, kTile, 0, st>>>(p);
        }
    } else {
        if (grad) focal_stream_generic_kernel<true><<<sgrid, kTile, 0, st>>>(p);
        else      focal_stream_generic_kernel<false><<<sgrid, kTile, 0, st>>>(p);
    }
    G3D_LAUNCH_CHECK();
    G3D_CUDA(trace(4));

    // ---- K4: positives + reductions, as a programmatic dependent launch behind K3
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(64, (unsigned)B);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_pdl.load() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (variant == G3D_VARIANT_3D) G3D_CUDA(cudaLaunchKernelEx(&cfg, positives_kernel<G3D_VARIANT_3D>, pp));
    else                           G3D_CUDA(cudaLaunchKernelEx(&cfg, positives_kernel<G3D_VARIANT_2D>, pp));
    G3D_CUDA(trace(5));
    return G3D_OK;
}

extern "C" int g3d_focal_loss_fwd(const float* cls, const float* reg, const float* anchors, const float* ann,
                                  int64_t B, int64_t A, int64_t C, int64_t R, int64_t Gmax, int64_t W, int variant,
                                  const float* hyper_host, float* losses, float* per_image, int32_t* assign,
                                  int32_t* gt_count_out, void* workspace, int64_t workspace_bytes,
                                  const double* pyramid_host, int device, void* stream) {
    return g3d_focal_loss_fwd_bwd(cls, reg, anchors, ann, B, A, C, R, Gmax, W, variant, hyper_host, nullptr, losses,
                                  per_image, assign, gt_count_out, nullptr, nullptr, nullptr, workspace, workspace_bytes,
                                  pyramid_host, nullptr, 0, G3D_DREG_UNDEFINED, device, stream);
}

extern "C" int g3d_focal_loss_bwd(const float* cls, const float* reg, const float* anchors, const float* ann,
                                  int64_t B, int64_t A, int64_t C, int64_t R, int64_t Gmax, int64_t W, int variant,
                                  const float* hyper_host, const float* grad_out, const float* grad_scale,
                                  int have_grads, const float* grad_expected_host, const void* workspace,
                                  int64_t workspace_bytes, const double* pyramid_host, float* dcls, float* dreg,
                                  int device, void* stream) {
    int rc = check_focal_shapes(B, A, C, R, Gmax, W, variant);
    if (rc != G3D_OK) return rc;
    G3D_REQUIRE(cls && reg && anchors && grad_out && workspace && dcls && dreg, "null pointer");
    G3D_REQUIRE(Gmax == 0 || ann, "null annotations");
    G3D_REQUIRE(!have_grads || grad_expected_host, "have_grads needs grad_expected_host[3]");
    G3D_REQUIRE(((uintptr_t)cls % 32) == 0 && ((uintptr_t)dcls % 32) == 0 && ((uintptr_t)dreg % 16) == 0 &&
                    ((uintptr_t)reg % 16) == 0 && ((uintptr_t)anchors % 16) == 0 && ((uintptr_t)workspace % 256) == 0,
                "cls/dcls must be 32-byte, reg/dreg/anchors 16-byte and the workspace 256-byte aligned");
    LossHyper h;
    rc = make_hyper(hyper_host, h);
    if (rc != G3D_OK) return rc;
    FocalWorkspace w = carve(const_cast<void*>(workspace), B, A, Gmax);
    G3D_REQUIRE(workspace_bytes >= w.bytes, "workspace too small (pass the forward's workspace, untouched)");
    G3D_GUARD(device);
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = sm_count(device);
    Pyramid pyr;
    const bool gt_centric = use_gt_centric(pyramid_host, A, C, Gmax, h, pyr);    // the decision the forward took
    ClsGradArgs p;
    p.cls = cls; p.grad_out = grad_out; p.grad_scale = grad_scale; p.npos = ws_npos(w, B);
    p.src = make_src(w, gt_centric, Gmax, C, h);
    p.dcls = dcls; p.dreg = dreg; p.e0 = have_grads ? grad_expected_host[0] : 0.0f; p.have_dcls = have_grads ? 1 : 0;
    p.B = (int)B; p.A = (int)A; p.C = (int)C; p.R = (int)R;
    p.T = (int)ceil_div(A, 256);
    p.h = h;
    const int64_t items = (int64_t)p.T * B;
    const int n_cls = (int)(items < (int64_t)sms * 4 ? items : (int64_t)sms * 4);
    PosArgs pp;
    pp.reg = reg; pp.anchors = (const float4*)anchors; pp.gt_tab = w.gt_tab; pp.pos_anchor = w.pos_anchor;
    pp.pos_gt = w.pos_gt; pp.npos = ws_npos(w, B); pp.gt_count = w.gt_count; pp.acc = nullptr; pp.nonfinite = nullptr;
    pp.dreg = dreg; pp.prev_npos = nullptr; pp.g1 = pp.g2 = 0.0f; pp.grad_out = grad_out; pp.grad_scale = grad_scale;
    pp.e1 = have_grads ? grad_expected_host[1] : 0.0f; pp.e2 = have_grads ? grad_expected_host[2] : 0.0f;
    pp.have_rows = have_grads ? 1 : 0;
    pp.B = (int)B; pp.A = (int)A; pp.R = (int)R; pp.Gmax = (int)Gmax; pp.h = h;
    pp.partials = nullptr; pp.counters = nullptr; pp.losses = nullptr; pp.per_image = nullptr; pp.gt_count_out = nullptr;
    pp.shard_stats = nullptr; pp.T = 0;
    const unsigned n_pos = (unsigned)(kPosCtas * B);
    auto launch = [&](int roles) {
        const unsigned grid = (roles == 1) ? (unsigned)n_cls : (unsigned)n_cls + n_pos;
        if (variant == G3D_VARIANT_3D) focal_bwd_kernel<G3D_VARIANT_3D><<<grid, 256, 0, st>>>(p, pp, n_cls, roles);
        else                           focal_bwd_kernel<G3D_VARIANT_2D><<<grid, 256, 0, st>>>(p, pp, n_cls, roles);
    };
    if (have_grads) {
        launch(3);                 // both roles verify; usually one wave of early exits
    } else {
        launch(1);                 // zero-fill of dreg + dcls first ...
        G3D_LAUNCH_CHECK();
        launch(2);                 // ... then the rows of the positives
    }
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}
