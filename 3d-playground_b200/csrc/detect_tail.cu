// detect_tail.cu — the whole detection tail up to the one host decision, as ONE library call.
//
// ResNet.forward's post-processing (3D model.py:346-397, 2D retinanet/model.py:270-311) is, on this side: score filter ->
// candidates with their NMS boxes decoded on the fly -> per-(image, class) sort + NMS -> offsets of the kept rows.
//   g3d_detect_tail_short  the usual case, segments of at most 1024 candidates: after the score filter (filter.cu) ONE
//                          launch does everything else per segment in shared memory (tail_short_kernel, below), one more
//                          the offsets and the summary - 4 launches per batch with the assembly;
//   g3d_detect_tail        any segment length (up to 16 384 candidates, the 3D model's threshold ladder keeps 10 000
//                          per class): the entry points of filter.cu / nms.cu issued back to back from C++ on the
//                          caller's stream (~10 launches; from Python each would cost ~10 us of host time).
// Both carve their temporaries out of one caller-owned workspace and leave the integers the host needs (number of
// detections, largest candidate count, segments the short path left over) in `summary`.
#include "box_decode.cuh"
#include "nms_common.cuh"

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
        summary[2] = 0;             // the general chain takes every segment
        __threadfence_system();     // `summary` may be mapped host memory that the host polls: entry 3 is written last
        summary[3] = 0;
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

// ---------------------------------------------------------------------------------------------------------------------
// Short segments - what a detector normally produces: some hundreds of candidates per (image, class).  ONE CTA per
// segment runs the whole chain in shared memory, one launch instead of gather + sort + greedy NMS (3 launches, the
// candidate boxes through global memory in between, ~200 us at BASELINE configs[2]):
//   1. the segment's candidates (arrival order of the score filter) get the key (score descending, anchor ascending) -
//      the order torchvision's stable sort gives the reference's boolean-mask gather - and are sorted by it;
//   2. candidate tables (score, anchor) are written in that order, the NMS boxes are decoded into shared memory;
//   3. the suppression pairs are found through a 16 x 16 grid over the box centres instead of all n^2/2 pair tests:
//      IoU(a,b) > t implies (in exact arithmetic) overlap_x > t max(w_a,w_b), hence |cx_a - cx_b| < w (1-t)/t for the
//      width w of EITHER box, and the same in y.  Box j therefore finds every higher-ranked box that suppresses it
//      among the boxes whose centre lies in a window of that size around its own centre; those few are put to the
//      exact test of nms.cu (same arithmetic, same result).  The window uses 0.99 t and a 1 % + 1e-5 |coordinate| margin,
//      orders of magnitude more than the rounding of the float IoU (16 ulp) and of the centres;
//   4. greedy NMS on the (sparse) suppression lists as a fixed point: a box is kept once every higher-ranked box that
//      suppresses it is removed, removed once one of them is kept; the lowest undecided rank is always decidable, so
//      the iteration ends with exactly the greedy result (a handful of rounds for clustered detections);
//   5. the kept ranks, in rank order, are the keep list.
// Segments this path does not take - more than kShortCap candidates, more than kEdgePool suppression pairs, or boxes
// outside the range where the window argument was checked (sides below 1e-10 or coordinates above 1e15) - are marked
// keep_count = -1 and counted in summary[2]; the caller then runs the general chain (g3d_detect_tail).
constexpr int kShortCap = 1024;
constexpr int kShortThreads = 1024;   // one thread per candidate
constexpr int kShortCtasPerSm = 2;
constexpr int kCells = 16;
constexpr int kEdgePool = 6144;
constexpr int kSlotChunk = 32;       // edge slots a warp reserves at a time
constexpr int kPairChunk = 256;      // candidate pairs a warp takes at a time
constexpr int kNarrowCells = 16;    // largest window (in grid cells) of a box handled cell by cell
constexpr uint16_t kNone = 0xffffu;
enum : uint8_t { kUndecided = 0, kKept = 1, kRemoved = 2 };

struct ShortSmem {
    union {
        uint64_t keys[2][kShortCap];                                // phase 1: the sort's two exchange buffers
        struct { uint16_t src[kEdgePool], next[kEdgePool]; } edge;  // phases 3-4: suppressor rank, next edge of the same box
    };
    float4 box[kShortCap];
    int head[kShortCap];             // first edge of the box's suppressor list, kNone = empty (phase 2b: its grid cell)
    uint16_t cell_item[kShortCap];   // ranks ordered by grid cell
    uint16_t wide[kShortCap];        // boxes whose pair search is done by a whole warp
    uint8_t status[kShortCap];
    int cell_start[kCells * kCells + 1];
    int cell_fill[kCells * kCells];
    unsigned cell_win[kCells * kCells];   // union of the windows of the cell's narrow boxes: column mask | row mask << 16
    float red[5][kShortThreads / 32];
    int wcount[kShortThreads / 32];
    int flag, edge_count, wide_count, next_cell, total_chunks;
};

__device__ __forceinline__ int cell_coord(float v, float lo, float scale) {
    return min(kCells - 1, max(0, (int)((v - lo) * scale)));   // monotone in v (float -> int conversion saturates)
}

#ifdef G3D_TAIL_CLOCKS
#define G3D_CLK(k) do { if (tid == 0) keep[base + cap - 16 + (k)] = clock64(); } while (0)
#else
#define G3D_CLK(k) do { } while (0)
#endif

// 1024 keys, one per thread, ascending: the bitonic network with the 40 steps inside a warp done by shuffles and the 15
// wider ones through shared memory (two buffers alternate, one barrier per step).
__device__ __forceinline__ uint64_t sort1024(uint64_t key, uint64_t (*buf)[kShortCap], int tid) {
    int flip = 0;
#pragma unroll
    for (int k = 2; k <= kShortCap; k <<= 1) {
        const bool up = (tid & k) == 0;
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            uint64_t other;
            if (j >= 32) {
                buf[flip][tid] = key;
                __syncthreads();
                other = buf[flip][tid ^ j];
                flip ^= 1;
            } else {
                const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)key, j);
                const uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(key >> 32), j);
                other = ((uint64_t)hi << 32) | lo;
            }
            const bool lower = (tid & j) == 0;
            key = ((key < other) == (lower == up)) ? key : other;
        }
    }
    return key;
}

__global__ void __launch_bounds__(kShortThreads, kShortCtasPerSm) tail_short_kernel(
    const float* __restrict__ scores, int inner, int64_t N, int64_t outer_pitch, const int32_t* __restrict__ idx,
    const int32_t* __restrict__ count, int cap, const BoxDecode dec, float thr, float win,
    int32_t* __restrict__ seg_offsets, float* __restrict__ cand_scores, int32_t* __restrict__ cand_src,
    int64_t* __restrict__ keep, int32_t* __restrict__ keep_count, int seg_base, int seg_total) {
    extern __shared__ __align__(16) unsigned char short_raw[];
    ShortSmem& sm = *reinterpret_cast<ShortSmem*>(short_raw);
    const int s = seg_base + blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t base = (int64_t)s * cap;
    const int n = min(__ldg(count + s), cap);
    if (tid == 0) {
        seg_offsets[s] = (int32_t)base;                             // fixed stride: no scan of the counts needed
        if (s == seg_total - 1) seg_offsets[s + 1] = (int32_t)(base + cap);
    }
    if (n <= 0 || n > kShortCap) {
        if (tid == 0) keep_count[s] = (n <= 0) ? 0 : -1;
        return;
    }
    const int64_t o = s / inner, c = s - o * inner;
    const float* __restrict__ sc = scores + o * outer_pitch + c;
    const bool have = tid < n;                                      // thread r owns rank r from phase 2 on

    G3D_CLK(0);
    // ---- 1. one candidate per thread: key (score descending, anchor ascending), sorted
    uint64_t key = ~0ull;
    if (have) {
        const int e = __ldg(idx + base + tid);
        key = make_key(__ldg(sc + (int64_t)e * inner), (uint32_t)e);
    }
    key = sort1024(key, sm.keys, tid);

    G3D_CLK(1);
    // ---- 2. candidate tables and boxes in sorted order; extent of the centres of the well-formed boxes
    float lox = INFINITY, loy = INFINITY, hix = -INFINITY, hiy = -INFINITY, amax = 0.0f;
    bool odd = false, shaped = false;
    float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
    if (have) {
        const int e = (int)(uint32_t)key;
        cand_src[base + tid] = e;
        cand_scores[base + tid] = __ldg(sc + (int64_t)e * inner);
        mine = decoded_nms_box(dec, o, N, e);
        sm.box[tid] = mine;
        const float w = mine.z - mine.x, h = mine.w - mine.y;
        shaped = w > 0.0f && h > 0.0f;  // the others (empty, inverted, NaN) never pass the test for thr >= 0: not in the grid
        if (shaped) {
            odd = !(w >= 1e-10f && h >= 1e-10f && w <= 1e15f && h <= 1e15f && fabsf(mine.x) <= 1e15f && fabsf(mine.y) <= 1e15f);
            lox = hix = mine.x + 0.5f * w;
            loy = hiy = mine.y + 0.5f * h;
            amax = fmaxf(fmaxf(fabsf(mine.x), fabsf(mine.z)), fmaxf(fabsf(mine.y), fabsf(mine.w)));
        }
    }
    const float mcx = lox, mcy = loy;                               // this thread's centre (if shaped)
    lox = warp_min(lox); loy = warp_min(loy); hix = warp_max(hix); hiy = warp_max(hiy); amax = warp_max(amax);
    odd = __any_sync(0xffffffffu, odd);
    if (tid == 0) { sm.flag = 0; sm.edge_count = 0; sm.wide_count = 0; sm.next_cell = 0; }
    if (tid < kCells * kCells) { sm.cell_fill[tid] = 0; sm.cell_win[tid] = 0u; }
    if (lane == 0) { sm.red[0][warp] = lox; sm.red[1][warp] = loy; sm.red[2][warp] = hix; sm.red[3][warp] = hiy; sm.red[4][warp] = amax; }
    __syncthreads();                // also: sm.keys is dead from here on (sm.edge takes its place)
    if (odd) sm.flag = 1;
    lox = warp_min(sm.red[0][lane]); loy = warp_min(sm.red[1][lane]);
    hix = warp_max(sm.red[2][lane]); hiy = warp_max(sm.red[3][lane]); amax = warp_max(sm.red[4][lane]);
    const float gx = (hix > lox) ? (float)kCells / (hix - lox) : 0.0f;
    const float gy = (hiy > loy) ? (float)kCells / (hiy - loy) : 0.0f;

    G3D_CLK(2);
    // ---- 2b. boxes by grid cell (counting sort)
    int cell = kNone;
    if (shaped) {
        cell = cell_coord(mcy, loy, gy) * kCells + cell_coord(mcx, lox, gx);
        atomicAdd(&sm.cell_fill[cell], 1);
    }
    if (have) sm.head[tid] = kNone;
    __syncthreads();
    if (sm.flag) {                  // uniform: a box outside the checked range - the general chain takes the segment
        if (tid == 0) keep_count[s] = -1;
        return;
    }
    if (warp == 0) {
        constexpr int kPer = kCells * kCells / 32;
        int v[kPer], sum = 0;
#pragma unroll
        for (int k = 0; k < kPer; ++k) { v[k] = sm.cell_fill[lane * kPer + k]; sum += v[k]; }
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        int run = incl - sum;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            sm.cell_start[lane * kPer + k] = run;
            sm.cell_fill[lane * kPer + k] = run;          // becomes the fill cursor
            run += v[k];
        }
        if (lane == 31) sm.cell_start[kCells * kCells] = run;
    }
    __syncthreads();
    if (shaped) sm.cell_item[atomicAdd(&sm.cell_fill[cell], 1)] = (uint16_t)tid;

    G3D_CLK(3);
    // ---- 3. suppressor lists: every pair of boxes that may suppress one another is put to the exact test once.
    // A box is NARROW if its window covers at most kNarrowCells grid cells, else WIDE (always, below t = 0.25: no window).
    //  - narrow x narrow: by grid cell.  The union of the windows of a cell's narrow boxes is a range of columns x rows;
    //    a warp takes the cell and tests its boxes against the boxes of that range that come later in cell order (the
    //    pair is seen from the earlier cell: the window argument holds from either side).  The m x len candidate pairs of
    //    a cell and a grid row are spread over the lanes as one flat index range - every lane busy;
    //  - wide boxes: a warp each, against every box in the window (other wide boxes: from the higher-ranked one only).
    const float eps = 1e-5f * amax;
    const bool dense = !(win < INFINITY);
    uint8_t* wide_flag = sm.status;               // phase 3 only; phase 4 initialises the status bytes afresh
    // one pair (called by all lanes; `valid` = this lane has one): straight-line up to the decision, so that the lanes
    // of a warp stay together.  The suppressing pairs of the warp's 32 then go into edge slots the warp reserves
    // kSlotChunk at a time - one shared-memory atomic per ~32 edges instead of one per edge.
    int slot_base = 0, slot_used = kSlotChunk;
    auto test = [&](bool valid, int j, int q) {
        const float4 a = sm.box[j], b = sm.box[q];
        const float w = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
        const float h = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
        const float inter = __fmul_rn(w, h);
        const float uni = __fsub_rn(__fadd_rn(box_area_rn(a.x, a.y, a.z, a.w), box_area_rn(b.x, b.y, b.z, b.w)), inter);
        const float p = __fmul_rn(uni, thr);
        bool hit = valid && w > 0.0f && h > 0.0f;
        // thr >= 0.25 (a window exists): the IEEE division is decided without dividing unless the quotient is within 2e-6
        // of the threshold - inter > fl(fl(u thr) 1.000002) implies inter / u > thr (1 + 1.8e-6), more than 15 ulp above
        // thr, so the rounded quotient is above thr; symmetrically below.  In between the division itself decides.
        const bool above = inter > __fmul_rn(p, 1.000002f), below = inter < __fmul_rn(p, 0.999998f);
        if (dense || !(above || below)) {
            if (hit) hit = __fdiv_rn(inter, uni) > thr;
        } else {
            hit = hit && above;
        }
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        if (hm == 0u) return;
        const int k = __popc(hm);
        if (slot_used + k > kSlotChunk) {
            if (lane == 0) slot_base = atomicAdd(&sm.edge_count, kSlotChunk);
            slot_base = __shfl_sync(0xffffffffu, slot_base, 0);
            slot_used = 0;
        }
        const int slot = slot_base + slot_used + __popc(hm & ((1u << lane) - 1u));
        slot_used += k;
        if (hit && slot < kEdgePool) {
            sm.edge.src[slot] = (uint16_t)min(j, q);
            sm.edge.next[slot] = (uint16_t)atomicExch(&sm.head[max(j, q)], slot);   // push on the lower-ranked box's list
        }
    };
    auto window = [&](const float4& b, float w, float h, int& x0, int& x1, int& y0, int& y1) {
        x0 = 0; x1 = kCells - 1; y0 = 0; y1 = kCells - 1;
        if (!dense) {
            const float cx = b.x + 0.5f * w, cy = b.y + 0.5f * h;
            const float rx = win * w + eps, ry = win * h + eps;
            x0 = cell_coord(cx - rx, lox, gx); x1 = cell_coord(cx + rx, lox, gx);
            y0 = cell_coord(cy - ry, loy, gy); y1 = cell_coord(cy + ry, loy, gy);
        }
    };
    if (shaped) {
        int x0, x1, y0, y1;
        window(mine, mine.z - mine.x, mine.w - mine.y, x0, x1, y0, y1);
        const bool wide = dense || (x1 - x0 + 1) * (y1 - y0 + 1) > kNarrowCells;
        wide_flag[tid] = wide;
        if (wide)
            sm.wide[atomicAdd(&sm.wide_count, 1)] = (uint16_t)tid;
        else
            atomicOr(&sm.cell_win[cell], (((2u << x1) - 1u) & ~((1u << x0) - 1u)) | ((((2u << y1) - 1u) & ~((1u << y0) - 1u)) << 16));
    }
    __syncthreads();
    // the candidate pairs of cell cc and grid row y are the flat range [0, m * len); the work is handed out in chunks of
    // kPairChunk pairs, numbered over (cell, row): cell_fill[cc] = first chunk of cell cc (scan below)
    auto cell_rows = [&](int cc, int& cs, int& m, int& xlo, int& xhi, int& yhi) {
        const unsigned wm = sm.cell_win[cc];
        cs = sm.cell_start[cc]; m = sm.cell_start[cc + 1] - cs;
        xlo = __ffs(wm & 0xffffu) - 1; xhi = 31 - __clz(wm & 0xffffu); yhi = 31 - __clz(wm >> 16);
        return wm != 0u;                                         // false: no narrow box in this cell
    };
    auto row_range = [&](int cc, int cs, int xlo, int xhi, int y, int& s0) {
        s0 = (y == cc / kCells) ? cs : sm.cell_start[y * kCells + xlo];   // own row: the cells to the left see cc themselves
        return sm.cell_start[y * kCells + xhi + 1] - s0;
    };
    if (tid < kCells * kCells) {
        int cs, m, xlo, xhi, yhi, chunks = 0;
        if (cell_rows(tid, cs, m, xlo, xhi, yhi))
            for (int y = tid / kCells; y <= yhi; ++y) {
                int s0;
                const int len = row_range(tid, cs, xlo, xhi, y, s0);
                chunks += (m * max(len, 0) + kPairChunk - 1) / kPairChunk;
            }
        sm.cell_fill[tid] = chunks;
    }
    __syncthreads();
    if (warp == 0) {
        constexpr int kPer = kCells * kCells / 32;
        int v[kPer], sum = 0;
#pragma unroll
        for (int k = 0; k < kPer; ++k) { v[k] = sm.cell_fill[lane * kPer + k]; sum += v[k]; }
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        int run = incl - sum;
#pragma unroll
        for (int k = 0; k < kPer; ++k) { sm.cell_fill[lane * kPer + k] = run; run += v[k]; }
        if (lane == 31) sm.total_chunks = run;
    }
    __syncthreads();
    G3D_CLK(4);
    const int total_chunks = sm.total_chunks;
    for (;;) {                                                   // narrow x narrow
        int g = 0;
        if (lane == 0) g = atomicAdd(&sm.next_cell, 1);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= total_chunks) break;
        // the cell that owns chunk g: the last one whose first chunk is <= g (two 32-way / 8-way votes)
        constexpr int kPer = kCells * kCells / 32;
        const int top = 31 - __clz(__ballot_sync(0xffffffffu, sm.cell_fill[lane * kPer] <= g));
        const unsigned low = __ballot_sync(0xffffffffu, lane < kPer && sm.cell_fill[top * kPer + (lane & (kPer - 1))] <= g);
        const int cc = top * kPer + 31 - __clz(low);
        int cs, m, xlo, xhi, yhi, s0 = 0, len = 0;
        cell_rows(cc, cs, m, xlo, xhi, yhi);
        int r = g - sm.cell_fill[cc];
        for (int y = cc / kCells; y <= yhi; ++y) {
            len = max(row_range(cc, cs, xlo, xhi, y, s0), 0);
            const int nch = (m * len + kPairChunk - 1) / kPairChunk;
            if (r < nch) break;
            r -= nch;
        }
        const int total = m * len, pend = min(total, (r + 1) * kPairChunk);
        const unsigned inv = 0xffffffffu / (unsigned)max(len, 1) + 1u;       // p / len == umulhi(p, inv) for p < 2^32 / len
        for (int p0 = r * kPairChunk; p0 < pend; p0 += 32) {
            const int p = min(p0 + lane, pend - 1);
            const int a = (len == 1) ? p : (int)__umulhi((unsigned)p, inv);
            const int ia = cs + a, ib = s0 + (p - a * len);
            const int j = sm.cell_item[ia], q = sm.cell_item[ib];
            test(p0 + lane < pend && ib > ia && !(wide_flag[j] | wide_flag[q]), j, q);
        }
    }
    G3D_CLK(5);
    for (int i = warp; i < sm.wide_count; i += kShortThreads / 32) {
        const int j = sm.wide[i];
        const float4 b = sm.box[j];
        int x0, x1, y0, y1;
        window(b, b.z - b.x, b.w - b.y, x0, x1, y0, y1);
        for (int y = y0; y <= y1; ++y) {
            const int kend = sm.cell_start[y * kCells + x1 + 1];              // the cells of a grid row are contiguous in the list
            for (int k0 = sm.cell_start[y * kCells + x0]; k0 < kend; k0 += 32) {
                const int q = sm.cell_item[min(k0 + lane, kend - 1)];
                test(k0 + lane < kend && q != j && !(wide_flag[q] && q < j), j, q);
            }
        }
    }
    __syncthreads();
    if (sm.edge_count > kEdgePool) {
        if (tid == 0) keep_count[s] = -1;
        return;
    }

    G3D_CLK(6);
    // ---- 4. greedy NMS as a fixed point over the suppressor lists (thread r decides rank r)
    volatile uint8_t* status = sm.status;
    const int first = have ? sm.head[tid] : kNone;              // (every wide_flag read is done: the bytes become the status)
    uint8_t st_mine = (first == kNone) ? kKept : kUndecided;
    if (have) status[tid] = st_mine;
    __syncthreads();
    for (;;) {
        int pending = 0;
        if (have && st_mine == kUndecided) {
            bool any_kept = false, all_removed = true;
            for (int e = first; e != kNone; e = sm.edge.next[e]) {
                const uint8_t st = status[sm.edge.src[e]];
                if (st == kKept) { any_kept = true; break; }
                all_removed &= (st == kRemoved);
            }
            if (any_kept) st_mine = kRemoved;
            else if (all_removed) st_mine = kKept;
            else pending = 1;
            if (!pending) status[tid] = st_mine;
        }
        if (!__syncthreads_or(pending)) break;
    }

    G3D_CLK(7);
    // ---- 5. keep list: the kept ranks in order (absolute positions in the candidate tables)
    const bool kept = have && st_mine == kKept;
    const unsigned bal = __ballot_sync(0xffffffffu, kept);
    if (lane == 0) sm.wcount[warp] = __popc(bal);
    __syncthreads();
    int incl = sm.wcount[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    const int before = __shfl_sync(0xffffffffu, incl, max(warp - 1, 0));
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (kept) keep[base + (warp ? before : 0) + __popc(bal & ((1u << lane) - 1u))] = base + tid;
    G3D_CLK(8);
    if (tid == 0) keep_count[s] = total;
}

// out_offsets = exclusive scan of max(keep_count, 0); summary = (detections, largest candidate count, segments left to
// the general chain (keep_count < 0), 0).  One CTA; S <= 2^24.
__global__ void __launch_bounds__(1024) tail_offsets_kernel(const int32_t* __restrict__ keep_count,
                                                            const int32_t* __restrict__ count, int S,
                                                            int32_t* __restrict__ out_offsets,
                                                            int32_t* __restrict__ summary) {
    __shared__ int wsum[32], wmax[32], wskip[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (S + 1023) / 1024;
    const int lo = min(S, tid * per), hi = min(S, lo + per);
    int sum = 0, mx = 0, skip = 0;
    for (int i = lo; i < hi; ++i) {
        const int kc = keep_count[i];
        sum += max(kc, 0);
        skip += kc < 0;
        mx = max(mx, count[i]);
    }
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    skip = __reduce_add_sync(0xffffffffu, skip);
    if (lane == 31) wsum[warp] = incl;
    if (lane == 0) { wmax[warp] = mx; wskip[warp] = skip; }
    __syncthreads();
    if (warp == 0) {
        const int v = wsum[lane];
        int inc2 = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, inc2, d);
            if (lane >= d) inc2 += up;
        }
        wsum[lane] = inc2 - v;
        const int m = __reduce_max_sync(0xffffffffu, wmax[lane]);
        const int k = __reduce_add_sync(0xffffffffu, wskip[lane]);
        if (lane == 31) {
            out_offsets[S] = inc2;
            summary[0] = inc2; summary[1] = m; summary[2] = k;
            __threadfence_system();     // `summary` may be mapped host memory that the host polls: entry 3 is written last
            summary[3] = 0;
        }
    }
    __syncthreads();
    int run = wsum[warp] + incl - sum;
    for (int i = lo; i < hi; ++i) {
        out_offsets[i] = run;
        run += max(keep_count[i], 0);
    }
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

extern "C" int g3d_detect_tail_short(const float* scores, int64_t outer, int64_t inner, int64_t N, int64_t outer_pitch,
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
    G3D_REQUIRE(iou_threshold >= 0.0, "the short path needs iou_threshold >= 0 (use g3d_detect_tail)");
    const int64_t S = outer * inner;
    G3D_REQUIRE(S < ((int64_t)1 << 24) && (S + 1) * cap < ((int64_t)1 << 31), "too many candidate slots");
    TailWorkspace w = carve_tail(workspace, S, cap);
    G3D_REQUIRE(workspace_bytes >= w.bytes, "workspace too small (see g3d_detect_tail_workspace_bytes)");
    BoxDecode d;
    int rc = make_box_decode(d, anchors, Ba, outer, reg, variant, mean_host, std_host, clip, clip_w, clip_h);
    if (rc) return rc;
    float thr_f = (float)iou_threshold;                    // same float threshold as g3d_nms_segmented
    if ((double)thr_f > iou_threshold) thr_f = nextafterf(thr_f, -INFINITY);
    // centre window of the pair search, in units of the box's own side: (1 - t') / t' with t' = 0.99 t, plus 1 %;
    // below t = 0.25 every pair of well-formed boxes is tested
    const float t2 = 0.99f * thr_f;
    const float win = thr_f >= 0.25f ? 1.01f * (1.0f - t2) / t2 : INFINITY;
    G3D_GUARD(device);
    cudaStream_t st = (cudaStream_t)stream;
    static bool attr_set[64] = {};
    if (device >= 0 && device < 64 && !attr_set[device]) {
        G3D_CUDA(cudaFuncSetAttribute(tail_short_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ShortSmem)));
        attr_set[device] = true;
    }
    auto segments = [&](cudaStream_t on, int64_t o0, int64_t no) {
        tail_short_kernel<<<(unsigned)(no * inner), kShortThreads, sizeof(ShortSmem), on>>>(
            scores, (int)inner, N, outer_pitch, w.idx, count, (int)cap, d, thr_f, win, seg_offsets, cand_scores, cand_src,
            keep, keep_count, (int)(o0 * inner), (int)S);
    };
    rc = g3d_filter_compact(scores, outer, inner, N, outer_pitch, thr, cap, w.idx, count, device, stream);
    if (rc) return rc;
    segments(st, 0, outer);
    G3D_LAUNCH_CHECK();
    tail_offsets_kernel<<<1, 1024, 0, st>>>(keep_count, count, (int)S, out_offsets, summary);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

