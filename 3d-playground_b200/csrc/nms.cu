// nms.cu — a11 greedy NMS with torchvision.ops.nms semantics, segmented (many independent box sets per launch).
//
// Replaces torchvision.ops.nms as called from retinanet/model.py:297, 3D model.py:383 and :336 (through batched_nms
// :19-57), perform_3D_detection_on_video_sequences.py:78, MC3D_crop_tracker.py:507,614,634, minimal_3D_track.py:515,535.
//
// Pipeline per segment, all on the device (no suppression bitmask in HBM, no device->host copy):
//   1. keys   : 64-bit key = (~orderable(score)) << 32 | index   -> ascending key order == stable descending score
//   2. sort   : bitonic sort of the keys (shared memory up to 16384 entries; global passes above that, S == 1 only)
//   3. greedy : one CTA walks the sorted boxes in blocks of 64: a 64x64 triangular IoU bit-matrix resolves the block
//               (warp-shuffle scan), then every thread sweeps the still-alive later boxes against the block's kept
//               boxes.  Work is O(kept x N) instead of O(N^2 / 2) and nothing but the boxes is ever stored.
#include "nms_common.cuh"

namespace g3d {

constexpr int kNmsThreads = 1024;
constexpr int kSortCap = 16384;        // longest segment the single-CTA shared-memory sort handles
constexpr int kSmemBoxCap = 4096;      // longest segment whose sorted boxes are cached in shared memory
constexpr int kLocalSort = 4096;       // tile of the multi-CTA (global) bitonic sort
constexpr int kShortSeg = 2048;        // two-size scheme (S >= 64): segments up to this length run on 256-thread CTAs

__device__ __forceinline__ float4 load_box(const float* __restrict__ boxes, int64_t stride, int64_t col, int64_t row) {
    const float* p = boxes + row * stride + col;
    return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}

// one CTA per segment, n <= kSortCap.  Segments whose length is outside (n_lo, n_hi] are left to the other launch of a
// two-size scheme: many short segments run 256-thread CTAs with little shared memory (several per SM), the rare long
// ones 1024-thread CTAs with the full buffer.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) seg_sort_kernel(const float* __restrict__ scores,
                                                           const int32_t* __restrict__ seg_offsets,
                                                           const float* __restrict__ boxes, int64_t stride, int64_t col,
                                                           int32_t* __restrict__ sidx, float4* __restrict__ sbox, int n_lo,
                                                           int n_hi) {
    extern __shared__ uint64_t sk[];
    const int seg = blockIdx.x;
    const int off = seg_offsets[seg];
    const int n = seg_offsets[seg + 1] - off;
    if (n <= n_lo || n > n_hi) return;
    int P = 2;
    while (P < n) P <<= 1;
    for (int i = threadIdx.x; i < P; i += blockDim.x)
        sk[i] = (i < n) ? make_key(__ldg(scores + off + i), (uint32_t)i) : ~0ull;
    __syncthreads();
    bitonic_smem(sk, P, 2, P, 0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int src = (int)(sk[i] & 0xffffffffu);
        sidx[off + i] = src;
        sbox[off + i] = load_box(boxes, stride, col, (int64_t)off + src);
    }
}

// ---- one segment of up to kRankCap boxes, sorted by RANK: the keys (score descending, index ascending on ties) are
// distinct, so the sorted position of element i is the number of elements before it.  Each thread counts them over all
// n scores held in shared memory as monotone 32-bit words (broadcast 16-byte reads, no barrier inside the loops):
// score_j >= score_i for the j below its warp's 32 indices, score_j > score_i above, the exact tie rule inside.  n^2
// 32-bit compares spread over n/128 CTAs where the one-CTA bitonic network needs 66 barrier-separated steps.
constexpr int kRankCap = 4096;
__device__ __forceinline__ uint32_t score_word(float score) { return (uint32_t)(make_key(score, 0u) >> 32); }   // ascending = better
__global__ void __launch_bounds__(128) rank_sort_kernel(const float* __restrict__ scores,
                                                        const int32_t* __restrict__ seg,
                                                        const float* __restrict__ boxes, int64_t stride, int64_t col,
                                                        int32_t* __restrict__ sidx, float4* __restrict__ sbox) {
    __shared__ __align__(16) uint32_t keys[kRankCap];
    const int off = seg[0], n = seg[1] - off;
    if ((int)(blockIdx.x * blockDim.x) >= n) return;
    const int n4 = (n + 3) & ~3;
    for (int j = threadIdx.x; j < n4; j += blockDim.x) keys[j] = (j < n) ? score_word(__ldg(scores + off + j)) : 0xffffffffu;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float4 box = (i < n) ? load_box(boxes, stride, col, (int64_t)off + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (i >= n) return;
    const uint32_t mine = keys[i];
    const int wbase = i & ~31;
    const uint4* k4 = reinterpret_cast<const uint4*>(keys);
    int rank = 0;
#pragma unroll 4
    for (int j = 0; j < (wbase >> 2); ++j) {                    // earlier indices win ties
        const uint4 k = k4[j];
        rank += (k.x <= mine) + (k.y <= mine) + (k.z <= mine) + (k.w <= mine);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {                              // the warp's own 32 indices
        const uint32_t k = keys[min(wbase + j, n4 - 1)];
        rank += (wbase + j < n) && (k < mine || (k == mine && wbase + j < i));
    }
#pragma unroll 4
    for (int j = (wbase + 32) >> 2; j < (n4 >> 2); ++j) {       // later indices lose ties (the 0xffffffff padding never counts)
        const uint4 k = k4[j];
        rank += (k.x < mine) + (k.y < mine) + (k.z < mine) + (k.w < mine);
    }
    sidx[off + rank] = i;
    sbox[off + rank] = box;
}

// ---- multi-CTA sort for one long segment (S == 1): keys in global memory, padded to a power of two
__global__ void __launch_bounds__(256) keys_init_kernel(const float* __restrict__ scores, int n, int P,
                                                        uint64_t* __restrict__ keys) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x)
        keys[i] = (i < n) ? make_key(__ldg(scores + i), (uint32_t)i) : ~0ull;
}
// full sort of each kLocalSort tile (first = 1), or the tail j < kLocalSort of merge stage k (first = 0)
__global__ void __launch_bounds__(kNmsThreads) bitonic_local_kernel(uint64_t* __restrict__ keys, int k, int first) {
    __shared__ uint64_t sk[kLocalSort];
    uint64_t* g = keys + (int64_t)blockIdx.x * kLocalSort;
    for (int i = threadIdx.x; i < kLocalSort; i += blockDim.x) sk[i] = g[i];
    __syncthreads();
    if (first) {
        // direction of a tile's final stage depends on its global position: emulate with the global index bit
        for (int kk = 2; kk <= kLocalSort; kk <<= 1) {
            for (int j = kk >> 1; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < (kLocalSort >> 1); t += blockDim.x) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int l = i | j;
                    const int64_t gi = (int64_t)blockIdx.x * kLocalSort + i;
                    const bool up = (gi & kk) == 0;
                    const uint64_t a = sk[i], b = sk[l];
                    if ((a > b) == up) { sk[i] = b; sk[l] = a; }
                }
                __syncthreads();
            }
        }
    } else {
        for (int j = kLocalSort >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (kLocalSort >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const int64_t gi = (int64_t)blockIdx.x * kLocalSort + i;
                const bool up = (gi & k) == 0;
                const uint64_t a = sk[i], b = sk[l];
                if ((a > b) == up) { sk[i] = b; sk[l] = a; }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < kLocalSort; i += blockDim.x) g[i] = sk[i];
}
__global__ void __launch_bounds__(256) bitonic_global_kernel(uint64_t* __restrict__ keys, int P, int k, int j) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < (P >> 1); t += gridDim.x * blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const uint64_t a = keys[i], b = keys[l];
        const bool up = (i & k) == 0;
        if ((a > b) == up) { keys[i] = b; keys[l] = a; }
    }
}
__global__ void __launch_bounds__(256) keys_gather_kernel(const uint64_t* __restrict__ keys, int n,
                                                          const float* __restrict__ boxes, int64_t stride, int64_t col,
                                                          int32_t* __restrict__ sidx, float4* __restrict__ sbox) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int src = (int)(keys[i] & 0xffffffffu);
        sidx[i] = src;
        sbox[i] = load_box(boxes, stride, col, src);
    }
}

// ---- greedy pass: one CTA per segment
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
    const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) nms_greedy_kernel(const float4* __restrict__ sbox,
                                                             const int32_t* __restrict__ sidx,
                                                             const int32_t* __restrict__ seg_offsets, float thr,
                                                             int relative, int removed_words, int box_cap,
                                                             int64_t* __restrict__ keep_out,
                                                             int32_t* __restrict__ keep_count, int n_lo, int n_hi) {
    extern __shared__ __align__(16) unsigned char dyn[];
    float4* cbox = reinterpret_cast<float4*>(dyn);  // [box_cap]: sorted boxes of segments with n <= box_cap
    uint32_t* removed = reinterpret_cast<uint32_t*>(dyn + sizeof(float4) * box_cap);
    __shared__ float4 dbox[64];
    __shared__ float darea[64];
    __shared__ uint64_t diag[64];
    __shared__ float4 kbox[64];
    __shared__ float karea[64];
    __shared__ uint64_t s_keep;
    __shared__ int dsrc[64];

    const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool thr_nonneg = thr >= 0.0f;
    const int off = seg_offsets[seg];
    const int n = seg_offsets[seg + 1] - off;
    if (n <= 0) {
        if (tid == 0 && n_lo < 0) keep_count[seg] = 0;
        return;
    }
    if (n <= n_lo || n > n_hi) return;   // the other launch of the two-size scheme handles this segment
    const float4* gbox = sbox + off;
    const int my_words = 2 * ((n + 63) >> 6);
    for (int i = tid; i < my_words; i += THREADS) removed[i] = 0u;
    // short segments keep their boxes in shared memory; long ones stream them from L2 (generic pointer)
    const float4* bsrc = gbox;
    if (n <= box_cap) {
        for (int i = tid; i < n; i += THREADS) cbox[i] = gbox[i];
        bsrc = cbox;
    }
    __syncthreads();

    int total_kept = 0;
    const int nblk = (n + 63) >> 6;
    for (int blk = 0; blk < nblk; ++blk) {
        const int base = blk << 6;
        const int m = min(64, n - base);
        const uint64_t mmask = (m == 64) ? ~0ull : ((1ull << m) - 1ull);
        const uint64_t remword = ((uint64_t)removed[2 * blk + 1] << 32) | removed[2 * blk];
        if ((remword & mmask) == mmask) continue;  // whole block already suppressed (uniform across the CTA)
        if (tid < m) {
            const float4 b = bsrc[base + tid];
            dbox[tid] = b;
            darea[tid] = box_area_rn(b.x, b.y, b.z, b.w);
            dsrc[tid] = sidx[off + base + tid];     // original indices: fetched here, off the serial section below
        }
        __syncthreads();
        // 64x64 strictly-upper-triangular suppression bits: 16 lanes per row, lane l16 covers columns l16 + 16q
#pragma unroll 1
        for (int i = tid >> 4; i < 64; i += THREADS >> 4) {
            const int l16 = tid & 15;
            const bool row_live = (i < m) && !((remword >> i) & 1ull);
            uint64_t word = 0ull;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = l16 + 16 * q;
                bool pred = false;
                if (row_live && j > i && j < m)
                    pred = suppresses(dbox[i], darea[i], dbox[j], darea[j], thr, thr_nonneg);
                const uint32_t bal = __ballot_sync(0xffffffffu, pred);
                const uint32_t half = (lane < 16) ? (bal & 0xffffu) : (bal >> 16);
                word |= (uint64_t)half << (16 * q);
            }
            if (l16 == 0) diag[i] = word;
        }
        __syncthreads();
        if (warp == 0) {
            // the serial part: 64 independent broadcast reads, then test bit i / OR word i (3 dependent instructions
            // per row; a shuffle per row costs several times that)
            uint64_t rem = remword;
#pragma unroll
            for (int g = 0; g < 4; ++g) {          // 16 rows at a time (the 1024-thread variant has 64 registers)
                uint64_t d[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) d[i] = diag[16 * g + i];
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (!((rem >> (16 * g + i)) & 1ull)) rem |= d[i];
            }
            const uint64_t keep = ~rem & mmask;
            if (lane == 0) s_keep = keep;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int i = lane + 32 * half;
                if ((keep >> i) & 1ull) {
                    const int pos = __popcll(keep & ((1ull << i) - 1ull));
                    kbox[pos] = dbox[i];
                    karea[pos] = darea[i];
                    const int64_t src = dsrc[i];
                    keep_out[off + total_kept + pos] = relative ? src : src + off;
                }
            }
        }
        __syncthreads();
        const int kc = __popcll(s_keep);
        total_kept += kc;
        if (kc > 0) {
            for (int j = base + 64 + tid; j < n; j += THREADS) {
                if ((removed[j >> 5] >> (j & 31)) & 1u) continue;
                const float4 bj = bsrc[j];
                const float aj = box_area_rn(bj.x, bj.y, bj.z, bj.w);
                for (int k = 0; k < kc; ++k) {
                    if (suppresses(kbox[k], karea[k], bj, aj, thr, thr_nonneg)) {
                        atomicOr(&removed[j >> 5], 1u << (j & 31));
                        break;
                    }
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) keep_count[seg] = total_kept;
}

// ---- single long segment (the trackers' NMS calls: one set of 10^2..10^4 boxes): the serial chain of the one-CTA
// greedy pass (n / 64 blocks x ~5 us) dominates, so the pair tests are spread over the whole GPU instead -
// a 64x64-block suppression bit-matrix in L2 (upper triangle only) - and ONE warp then walks the blocks: a 64-step
// shuffle scan on the diagonal words, then the kept rows' words are OR-ed into the running "removed" set (coalesced,
// independent loads).  Same pair test, same order, same keep list as the greedy kernel.
// The segment [seg[0], seg[1]) is read on the device (n <= the capacity the grid and the row stride nb were sized for), so
// a captured launch sequence serves any length up to its capacity.
__global__ void __launch_bounds__(64) nms_mask_kernel(const float4* __restrict__ sbox_all,
                                                      const int32_t* __restrict__ seg, int nb, float thr,
                                                      uint64_t* __restrict__ mask) {
    const int rb = blockIdx.y, cb = blockIdx.x, tid = threadIdx.x;
    if (cb < rb) return;
    const int off = seg[0], n = seg[1] - off;
    if (cb * 64 >= n || rb * 64 >= n) return;
    const float4* __restrict__ sbox = sbox_all + off;
    __shared__ float4 cbox[64];
    __shared__ float carea[64];
    const bool thr_nonneg = thr >= 0.0f;
    const int col = cb * 64 + tid;
    if (col < n) {
        const float4 b = sbox[col];
        cbox[tid] = b;
        carea[tid] = box_area_rn(b.x, b.y, b.z, b.w);
    }
    __syncthreads();
    const int row = rb * 64 + tid;
    if (row >= n) return;
    const float4 me = sbox[row];
    const float area = box_area_rn(me.x, me.y, me.z, me.w);
    const int m = min(64, n - cb * 64);
    uint64_t word = 0ull;
    for (int j = 0; j < m; ++j)
        if (cb * 64 + j > row && suppresses(me, area, cbox[j], carea[j], thr, thr_nonneg)) word |= 1ull << j;
    mask[(int64_t)row * nb + cb] = word;
}

// WPL = 64-bit words of the "removed" set per lane (nb <= 32 * WPL)
template <int WPL>
__global__ void __launch_bounds__(32) nms_mask_scan_kernel(const uint64_t* __restrict__ mask,
                                                           const int32_t* __restrict__ sidx_all,
                                                           const int32_t* __restrict__ seg, int nbs, int relative,
                                                           int64_t* __restrict__ keep_all,
                                                           int32_t* __restrict__ keep_count) {
    const int lane = threadIdx.x;
    const int off = seg[0], n = seg[1] - off;
    const int nb = (n + 63) >> 6;               // blocks in use; nbs = row stride of the matrix (capacity)
    const int32_t* __restrict__ sidx = sidx_all + off;
    int64_t* __restrict__ keep_out = keep_all + off;
    const int add = relative ? 0 : off;
    uint64_t remv[WPL];
#pragma unroll
    for (int k = 0; k < WPL; ++k) remv[k] = 0ull;
    int total = 0;
    // diagonal words of block 0 (each later block's are requested one block ahead: they depend on nothing)
    uint64_t dA = (lane < n) ? mask[(int64_t)lane * nbs] : 0ull;
    uint64_t dB = (lane + 32 < n) ? mask[(int64_t)(lane + 32) * nbs] : 0ull;
    int sA = (lane < n) ? sidx[lane] : 0, sB = (lane + 32 < n) ? sidx[lane + 32] : 0;   // original indices, same prefetch
    for (int blk = 0; blk < nb; ++blk) {
        const int base = blk << 6;
        const int m = min(64, n - base);
        const uint64_t mmask = (m == 64) ? ~0ull : ((1ull << m) - 1ull);
        const uint64_t cA = dA, cB = dB;
        const int tA = sA, tB = sB;
        if (blk + 1 < nb) {
            const int ra = base + 64 + lane, rb2 = ra + 32;
            dA = (ra < n) ? mask[(int64_t)ra * nbs + blk + 1] : 0ull;
            dB = (rb2 < n) ? mask[(int64_t)rb2 * nbs + blk + 1] : 0ull;
            sA = (ra < n) ? sidx[ra] : 0;
            sB = (rb2 < n) ? sidx[rb2] : 0;
        }
        uint64_t mine = 0ull;
#pragma unroll
        for (int k = 0; k < WPL; ++k)
            if (k == (blk >> 5)) mine = remv[k];
        uint64_t rem = shfl64(mine, blk & 31);
        if ((rem & mmask) == mmask) continue;          // whole block already suppressed
#pragma unroll
        for (int i = 0; i < 64; ++i) {
            const uint64_t d = shfl64(i < 32 ? cA : cB, i & 31);
            if (!((rem >> i) & 1ull)) rem |= d;
        }
        const uint64_t keep = ~rem & mmask;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int i = lane + 32 * half;
            if ((keep >> i) & 1ull) keep_out[total + __popcll(keep & ((1ull << i) - 1ull))] = (half ? tB : tA) + add;
        }
        total += __popcll(keep);
        // the kept rows suppress later blocks: their words are independent loads - keep kBatch rows in flight
        constexpr int kBatch = (WPL == 1) ? 16 : (WPL == 2 ? 8 : 4);
        uint64_t bits = keep;
        while (bits) {
            uint64_t v[kBatch][WPL];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const bool have = bits != 0ull;
                const int i = have ? __ffsll((long long)bits) - 1 : 0;
                bits &= bits - 1;      // 0 stays 0
                const uint64_t* rowp = mask + (int64_t)(base + i) * nbs;
#pragma unroll
                for (int k = 0; k < WPL; ++k) {
                    const int w = lane + 32 * k;
                    v[u][k] = (have && w > blk && w < nb) ? __ldg(rowp + w) : 0ull;
                }
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u)
#pragma unroll
                for (int k = 0; k < WPL; ++k) remv[k] |= v[u][k];
        }
    }
    if (lane == 0) keep_count[0] = total;
}

// Up to 64 blocks (n <= 4096): the scan above spends its time waiting - a 64-step shuffle chain per block, then
// dependent L2 loads of the kept rows' words.  Here warps 1-3 copy the whole 64-row strip of the NEXT block into shared
// memory while warp 0 - the only serial actor, so every instruction it does not execute is time saved -
// resolves the current block from shared memory: the 64 diagonal words are broadcast reads at immediate offsets (the
// strip's row stride SW is a compile-time constant), the chain tests one bit and ORs one word per row, and the kept
// rows (compacted into a small index list) OR their words into the running "removed" set with one conflict-free read
// per lane.  Same pair tests, same order, same keep list.
template <int WPL>   // 64-bit words of the removed set per lane: blocks <= 32 * WPL
__global__ void __launch_bounds__(128) nms_mask_scan_strip_kernel(const uint64_t* __restrict__ mask,
                                                                  const int32_t* __restrict__ sidx_all,
                                                                  const int32_t* __restrict__ seg, int nbs, int relative,
                                                                  int64_t* __restrict__ keep_all,
                                                                  int32_t* __restrict__ keep_count) {
    constexpr int SW = 32 * WPL;                                // strip row stride in words
    extern __shared__ __align__(16) uint64_t strip[];          // [2][64][SW]
    __shared__ int kept_rows[64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int off = seg[0], n = seg[1] - off;
    const int nb = (n + 63) >> 6;
    const int32_t* __restrict__ sidx = sidx_all + off;
    int64_t* __restrict__ keep_out = keep_all + off;
    const int add = relative ? 0 : off;
    // words [blk & ~1, nb) of the block's valid rows, by warps 1-3: every 16-byte piece is an independent load held in
    // registers, then stored (cp.async issues at ~200 cycles per instruction here and would be the critical path)
    auto stage = [&](int blk) {
        constexpr int CH = SW / 2, U = (64 * CH + 95) / 96;     // pieces per row, pieces per helper thread
        const int base = blk << 6, w0 = blk & ~1;
        const int rows = min(64, n - base), chunks = (nb - w0 + 1) >> 1;
        uint64_t* dst = strip + (size_t)(blk & 1) * 64 * SW + w0;
        const uint64_t* src = mask + (int64_t)base * nbs + w0;
        const int t = tid - 32;
        uint4 v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int u = t + 96 * k, r = u / CH, c = u % CH;
            if (u < 64 * CH && r < rows && c < chunks) v[k] = __ldcg(reinterpret_cast<const uint4*>(src + (int64_t)r * nbs + 2 * c));
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int u = t + 96 * k, r = u / CH, c = u % CH;
            if (u < 64 * CH && r < rows && c < chunks) *reinterpret_cast<uint4*>(dst + r * SW + 2 * c) = v[k];
        }
    };
    if (nb > 0 && warp != 0) stage(0);
    uint64_t r0 = 0ull, r1 = 0ull;                              // removed-set words lane and lane + 32 (warp 0)
    int total = 0;
    int sA = (lane < n) ? sidx[lane] : 0, sB = (lane + 32 < n) ? sidx[lane + 32] : 0;
    __syncthreads();
    for (int blk = 0; blk < nb; ++blk) {
        if (warp != 0) {
            if (blk + 1 < nb) stage(blk + 1);
        } else {
            const int base = blk << 6;
            const int m = min(64, n - base);
            const uint64_t mmask = (m == 64) ? ~0ull : ((1ull << m) - 1ull);
            const uint64_t* __restrict__ S = strip + (size_t)(blk & 1) * 64 * SW;
            const int tA = sA, tB = sB;
            if (blk + 1 < nb) {                                 // original indices of the next block, one block ahead
                const int ra = base + 64 + lane, rb = ra + 32;
                sA = (ra < n) ? sidx[ra] : 0;
                sB = (rb < n) ? sidx[rb] : 0;
            }
            uint64_t rem = shfl64((WPL == 1 || blk < 32) ? r0 : r1, blk & 31);
            if ((rem & mmask) != mmask) {
                // 64 independent broadcast reads first, then the chain: test bit i, OR the row's diagonal word
                const uint64_t* __restrict__ dg = S + blk;
                uint64_t d[64];
                if (m == 64) {
#pragma unroll
                    for (int i = 0; i < 64; ++i) d[i] = dg[i * SW];
                } else {
#pragma unroll
                    for (int i = 0; i < 64; ++i) d[i] = (i < m) ? dg[i * SW] : 0ull;
                }
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    if (!((rem >> i) & 1ull)) rem |= d[i];
            }
            const uint64_t keep = ~rem & mmask;
            const int kept = __popcll(keep);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int i = lane + 32 * half;
                if ((keep >> i) & 1ull) {
                    const int pos = __popcll(keep & ((1ull << i) - 1ull));
                    keep_out[total + pos] = (half ? tB : tA) + add;
                    kept_rows[pos] = i * SW;
                }
            }
            total += kept;
            __syncwarp();
            // the kept rows suppress later blocks: OR their words into the removed set, 8 independent reads at a time
            if (blk + 1 < nb) {
                const uint64_t* __restrict__ mine = S + lane;
                const bool u0 = lane > blk && lane < nb, u1 = WPL > 1 && lane + 32 > blk && lane + 32 < nb;
                for (int k0 = 0; k0 < kept; k0 += 8) {
                    uint64_t v0[8], v1[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const bool have = k0 + k < kept;
                        const int ro = have ? kept_rows[k0 + k] : 0;
                        v0[k] = (have && u0) ? mine[ro] : 0ull;
                        v1[k] = (WPL > 1 && have && u1) ? mine[ro + 32] : 0ull;
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        r0 |= v0[k];
                        if (WPL > 1) r1 |= v1[k];
                    }
                }
            }
            __syncwarp();
        }
        __syncthreads();
    }
    if (tid == 0) keep_count[0] = total;
}

struct NmsWorkspace {
    int32_t* sidx;
    float4* sbox;
    uint64_t* keys;
    uint64_t* mask;      // [n][ceil(n/64)] suppression bit-matrix (single-segment path only)
    int32_t* single;
    int64_t bytes;
};
// one segment of a few hundred to kSortCap boxes: bit-matrix + one-warp scan instead of the one-CTA greedy pass
static bool use_mask_path(int64_t S, int64_t max_seg_len, int64_t N) {
    return S == 1 && max_seg_len == N && N >= 256 && N <= kSortCap;
}
__global__ void single_segment_kernel(int32_t* seg, int n) { seg[0] = 0; seg[1] = n; }
static NmsWorkspace carve_nms(void* base, int64_t N, int64_t S, int64_t max_seg_len) {
    NmsWorkspace w;
    char* p = (char*)base;
    int64_t off = 0;
    w.sbox = (float4*)(p + off); off += align_up(N * 16, 256);
    w.sidx = (int32_t*)(p + off); off += align_up(N * 4, 256);
    w.keys = (uint64_t*)(p + off);
    if (max_seg_len > kSortCap) {
        int64_t P = kLocalSort;
        while (P < max_seg_len) P <<= 1;
        off += align_up(P * 8, 256);
    }
    w.mask = nullptr;
    if (use_mask_path(S, max_seg_len, N)) {
        w.mask = (uint64_t*)(p + off);
        off += align_up(max_seg_len * align_up(ceil_div(max_seg_len, 64), 2) * 8, 256);   // row stride: even word count
    }
    w.single = (int32_t*)(p + off);   // {0, N} for callers that pass seg_offsets == NULL with S == 1
    off += 256;
    w.bytes = off;
    return w;
}

}  // namespace g3d

using namespace g3d;

extern "C" int64_t g3d_nms_workspace_bytes(int64_t N, int64_t S, int64_t max_seg_len) {
    if (N < 0 || S < 0 || max_seg_len < 0) return G3D_ERR_INVALID;
    return carve_nms(nullptr, N, S, max_seg_len).bytes;
}

extern "C" int g3d_nms_segmented(const float* boxes, int64_t box_stride, int64_t box_col, const float* scores, int64_t N,
                                 const int32_t* seg_offsets, int64_t S, int64_t max_seg_len, double iou_threshold,
                                 int relative, int64_t* keep_out, int32_t* keep_count, void* workspace,
                                 int64_t workspace_bytes, int device, void* stream) {
    G3D_REQUIRE(N >= 0 && S >= 0 && max_seg_len >= 0 && box_stride >= 4 && box_col >= 0 && box_col + 4 <= box_stride,
                "bad size");
    G3D_REQUIRE(N < ((int64_t)1 << 31) && S < ((int64_t)1 << 31), "size out of range");
    if (S == 0) return G3D_OK;
    G3D_REQUIRE(keep_count && (seg_offsets || S == 1), "null pointer (seg_offsets may be NULL only for S == 1: one segment [0, N))");
    G3D_GUARD(device);
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0 || max_seg_len == 0) {
        G3D_CUDA(cudaMemsetAsync(keep_count, 0, sizeof(int32_t) * S, st));
        return G3D_OK;
    }
    G3D_REQUIRE(boxes && scores && keep_out && workspace, "null pointer");
    G3D_REQUIRE(max_seg_len <= N, "max_seg_len exceeds N");
    G3D_REQUIRE(((uintptr_t)workspace % 256) == 0, "workspace must be 256-byte aligned");
    NmsWorkspace w = carve_nms(workspace, N, S, max_seg_len);
    G3D_REQUIRE(workspace_bytes >= w.bytes, "workspace too small (see g3d_nms_workspace_bytes)");
    if (!seg_offsets) {
        single_segment_kernel<<<1, 1, 0, st>>>(w.single, (int)N);
        G3D_LAUNCH_CHECK();
        seg_offsets = w.single;
    }
    if (max_seg_len > kSortCap && S != 1) {
        set_error("g3d_nms_segmented: segments longer than %d boxes are only supported for S == 1", kSortCap);
        return G3D_ERR_UNSUPPORTED;
    }
    // (double)iou > iou_threshold  <=>  iou > thr_f  with thr_f the largest float <= iou_threshold (torchvision's CPU
    // kernel compares the float IoU against the double threshold)
    float thr_f = (float)iou_threshold;
    if ((double)thr_f > iou_threshold) thr_f = nextafterf(thr_f, -INFINITY);

    // many segments: short ones (<= kShortSeg boxes) on small CTAs, several per SM; the rest on 1024-thread CTAs
    const bool two_size = (S >= 64) && (max_seg_len > kShortSeg);
    const bool all_short = (S >= 64) && (max_seg_len <= kShortSeg);
    if (use_mask_path(S, max_seg_len, N) && N <= kRankCap) {
        rank_sort_kernel<<<(unsigned)ceil_div(N, 128), 128, 0, st>>>(scores, seg_offsets, boxes, box_stride, box_col, w.sidx,
                                                                   w.sbox);
        G3D_LAUNCH_CHECK();
    } else if (max_seg_len <= kSortCap) {
        int P = 2;
        while (P < max_seg_len) P <<= 1;
        const size_t smem = (size_t)P * 8;
        if (two_size || all_short) {
            int Ps = 2;
            while (Ps < (all_short ? max_seg_len : kShortSeg)) Ps <<= 1;
            seg_sort_kernel<256><<<(unsigned)S, 256, (size_t)Ps * 8, st>>>(scores, seg_offsets, boxes, box_stride, box_col,
                                                                         w.sidx, w.sbox, 0, kShortSeg);
            G3D_LAUNCH_CHECK();
        }
        if (!all_short) {
            G3D_CUDA(cudaFuncSetAttribute(seg_sort_kernel<kNmsThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            seg_sort_kernel<kNmsThreads><<<(unsigned)S, kNmsThreads, smem, st>>>(scores, seg_offsets, boxes, box_stride,
                                                                             box_col, w.sidx, w.sbox,
                                                                             two_size ? kShortSeg : 0, 0x7fffffff);
            G3D_LAUNCH_CHECK();
        }
    } else {
        int64_t P = kLocalSort;
        while (P < max_seg_len) P <<= 1;
        G3D_REQUIRE(P < ((int64_t)1 << 31), "segment too long");
        const int n = (int)N;  // S == 1: the single segment is [0, N)
        const int g256 = (int)(ceil_div(P, 256) < sm_count(device) * 16 ? ceil_div(P, 256) : sm_count(device) * 16);
        keys_init_kernel<<<g256, 256, 0, st>>>(scores, n, (int)P, w.keys);
        G3D_LAUNCH_CHECK();
        bitonic_local_kernel<<<(unsigned)(P / kLocalSort), kNmsThreads, 0, st>>>(w.keys, 0, 1);
        G3D_LAUNCH_CHECK();
        for (int64_t k = 2 * kLocalSort; k <= P; k <<= 1) {
            for (int64_t j = k >> 1; j >= kLocalSort; j >>= 1) {
                bitonic_global_kernel<<<g256, 256, 0, st>>>(w.keys, (int)P, (int)k, (int)j);
                G3D_LAUNCH_CHECK();
            }
            bitonic_local_kernel<<<(unsigned)(P / kLocalSort), kNmsThreads, 0, st>>>(w.keys, (int)k, 0);
            G3D_LAUNCH_CHECK();
        }
        keys_gather_kernel<<<g256, 256, 0, st>>>(w.keys, n, boxes, box_stride, box_col, w.sidx, w.sbox);
        G3D_LAUNCH_CHECK();
    }
    if (use_mask_path(S, max_seg_len, N)) {
        const int nb = (int)ceil_div(N, 64);      // capacity: the segment's own length is read from seg_offsets on the device
        const int nbs = (int)align_up(nb, 2);     // row stride of the bit-matrix (rows start 16-byte aligned)
        nms_mask_kernel<<<dim3((unsigned)nb, (unsigned)nb), 64, 0, st>>>(w.sbox, seg_offsets, nbs, thr_f, w.mask);
        G3D_LAUNCH_CHECK();
        if (nb <= 64) {
            const size_t smem = (size_t)2 * 64 * (nb <= 32 ? 32 : 64) * 8;   // two strips of 64 rows x SW words
            if (nb <= 32) {
                nms_mask_scan_strip_kernel<1><<<1, 128, smem, st>>>(w.mask, w.sidx, seg_offsets, nbs, relative, keep_out, keep_count);
            } else {
                G3D_CUDA(cudaFuncSetAttribute(nms_mask_scan_strip_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                nms_mask_scan_strip_kernel<2><<<1, 128, smem, st>>>(w.mask, w.sidx, seg_offsets, nbs, relative, keep_out, keep_count);
            }
        }
        else if (nb <= 128) nms_mask_scan_kernel<4><<<1, 32, 0, st>>>(w.mask, w.sidx, seg_offsets, nbs, relative, keep_out, keep_count);
        else                nms_mask_scan_kernel<8><<<1, 32, 0, st>>>(w.mask, w.sidx, seg_offsets, nbs, relative, keep_out, keep_count);
        G3D_LAUNCH_CHECK();
        return G3D_OK;
    }
    const int removed_words = (int)(2 * ceil_div(max_seg_len, 64));
    if (two_size || all_short) {
        const int cap_s = (int)(all_short ? max_seg_len : kShortSeg);
        const int words_s = (int)(2 * ceil_div(cap_s, 64));
        const size_t smem_s = sizeof(float4) * cap_s + (size_t)words_s * 4;
        nms_greedy_kernel<256><<<(unsigned)S, 256, smem_s, st>>>(w.sbox, w.sidx, seg_offsets, thr_f, relative, words_s, cap_s,
                                                               keep_out, keep_count, -1, kShortSeg);
        G3D_LAUNCH_CHECK();
    }
    if (!all_short) {
        const int box_cap = (int)(max_seg_len < kSmemBoxCap ? max_seg_len : kSmemBoxCap);
        const size_t smem = sizeof(float4) * box_cap + (size_t)removed_words * 4;
        G3D_REQUIRE(smem <= 220 * 1024, "segment too long for the shared-memory suppression bitset (max ~1.2M boxes)");
        G3D_CUDA(cudaFuncSetAttribute(nms_greedy_kernel<kNmsThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        nms_greedy_kernel<kNmsThreads><<<(unsigned)S, kNmsThreads, smem, st>>>(w.sbox, w.sidx, seg_offsets, thr_f, relative,
                                                                           removed_words, box_cap, keep_out, keep_count,
                                                                           two_size ? kShortSeg : -1, 0x7fffffff);
        G3D_LAUNCH_CHECK();
    }
    return G3D_OK;
}
