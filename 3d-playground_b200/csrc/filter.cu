// filter.cu — a10 score filtering: class max, the adaptive threshold ladder, and candidate compaction.
//
// Score vectors are described by (outer, inner, n): vector s = o * inner + c lives at
//     scores[o * outer_pitch + n * inner + c],  n in [0, N)
// which covers classification[B, A, C] taken per (image, class) (outer = B, inner = C, N = A, outer_pitch = A*C;
// retinanet/model.py:287-289, 3D model.py:365-374) and flat vectors (inner = 1; the MULTI_FRAME max-score vector of
// 3D model.py:320-328).
#include "box_decode.cuh"

namespace g3d {

constexpr int kMaxRungs = 512;

// ------------------------------------------------------------------------------------------------- class max (a10)
// scores, classes = torch.max(classification, dim=1)   (3D model.py:320): first maximal index on ties.
__global__ void __launch_bounds__(256) rowmax_kernel(const float* __restrict__ cls, int64_t rows, int C,
                                                     float* __restrict__ smax, int64_t* __restrict__ amax) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x) {
        const float* r = cls + i * C;
        float best = r[0];
        int bi = 0;
        for (int c = 1; c < C; ++c) {
            const float v = r[c];
            if (v > best) { best = v; bi = c; }
        }
        smax[i] = best;
        amax[i] = bi;
    }
}

// --------------------------------------------------------------------------------------------------- ladder (a10)
// rung index of a score: r(s) = #{k : t_k < s} for ascending rungs t (f32).  count(score > t_k) = #{r(s) >= k+1}.
__device__ __forceinline__ int rung_of(float s, const float* __restrict__ t, int L, float log2_t0, float inv_log2_step) {
    if (!(s > t[0])) return 0;
    int r = (int)((__log2f(s) - log2_t0) * inv_log2_step);
    r = max(1, min(r, L));
    while (r < L && t[r] < s) ++r;
    while (r > 0 && !(t[r - 1] < s)) --r;
    return r;
}

template <int INNER>
__global__ void __launch_bounds__(256) ladder_hist_kernel(const float* __restrict__ scores, int64_t N, int inner_rt,
                                                          int64_t outer_pitch, const float* __restrict__ rungs, int L,
                                                          float log2_t0, float inv_log2_step,
                                                          unsigned int* __restrict__ hist /*[S][L+1]*/) {
    extern __shared__ unsigned int s_mem[];
    const int inner = INNER > 0 ? INNER : inner_rt;
    float* s_t = reinterpret_cast<float*>(s_mem);           // [L]
    unsigned int* s_h = s_mem + L;                           // [inner][L+1]
    for (int i = threadIdx.x; i < L; i += blockDim.x) s_t[i] = rungs[i];
    for (int i = threadIdx.x; i < inner * (L + 1); i += blockDim.x) s_h[i] = 0u;
    __syncthreads();
    const int o = blockIdx.y;
    const float* base = scores + (int64_t)o * outer_pitch;
    const int64_t n_round = ((N + 31) / 32) * 32;  // whole warps stay in the loop: the match below is full-warp
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < n_round; n += (int64_t)gridDim.x * blockDim.x) {
        const bool in = n < N;
        float v[8];
        if (INNER == 8) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (in) {
                a = ld_stream(reinterpret_cast<const float4*>(base + n * 8));
                b = ld_stream(reinterpret_cast<const float4*>(base + n * 8) + 1);
            }
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
#pragma unroll
        for (int c = 0; c < inner; ++c) {
            int r = -1;
            if (in) {
                const float s = (INNER == 8) ? v[c & 7] : __ldg(base + n * inner + c);
                r = rung_of(s, s_t, L, log2_t0, inv_log2_step);
            }
            __syncwarp();
            // warp-aggregated histogram update: one shared atomic per distinct rung in the warp
            const unsigned peers = __match_any_sync(0xffffffffu, r);
            if (r >= 0 && (threadIdx.x & 31) == (__ffs(peers) - 1))
                atomicAdd(&s_h[c * (L + 1) + r], (unsigned)__popc(peers));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < inner * (L + 1); i += blockDim.x) {
        const unsigned int h = s_h[i];
        if (h) atomicAdd(&hist[(int64_t)o * inner * (L + 1) + i], h);
    }
}

// first rung whose surviving count is <= keep_max (the `while keep_count > keep` loop of 3D model.py:368-374)
__global__ void __launch_bounds__(256) ladder_pick_kernel(const unsigned int* __restrict__ hist, int S, int L,
                                                          const float* __restrict__ rungs, long long keep_max,
                                                          int32_t* __restrict__ rung_out, int32_t* __restrict__ count_out,
                                                          float* __restrict__ thr_out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const unsigned int* h = hist + (int64_t)s * (L + 1);
    long long total = 0;
    for (int j = 1; j <= L; ++j) total += h[j];
    // count_k = sum_{j >= k+1} h[j]; walk k upward
    long long cnt = total;  // k = 0
    int k = 0;
    while (k < L - 1 && cnt > keep_max) {
        ++k;
        cnt -= h[k];
    }
    if (cnt > keep_max) { rung_out[s] = -1; }
    else rung_out[s] = k;
    count_out[s] = (int32_t)(cnt > 0x7fffffffLL ? 0x7fffffffLL : cnt);
    thr_out[s] = rungs[k];
}

// ------------------------------------------------------------------------------------------------ compaction (a10)
// Append the element index n of every score > thr[s] to idx_out[s][*] (arrival order; the NMS front-end sorts by
// (score, index) so the final order is that of the reference's boolean-mask gather followed by a stable sort).
template <int INNER>
__global__ void __launch_bounds__(256) compact_kernel(const float* __restrict__ scores, int64_t N, int inner_rt,
                                                      int64_t outer_pitch, const float* __restrict__ thr, int cap,
                                                      int32_t* __restrict__ idx_out, int32_t* __restrict__ count) {
    const int inner = INNER > 0 ? INNER : inner_rt;
    const int o = blockIdx.y;
    const float* base = scores + (int64_t)o * outer_pitch;
    const int lane = threadIdx.x & 31;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = ((N + 31) / 32) * 32;  // keep whole warps in the loop so the ballots are full-warp
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < n_round; n += nthreads) {
        const bool in = n < N;
        float v[8];
        if (INNER == 8) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (in) {
                a = ld_stream(reinterpret_cast<const float4*>(base + n * 8));
                b = ld_stream(reinterpret_cast<const float4*>(base + n * 8) + 1);
            }
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
#pragma unroll
        for (int c = 0; c < inner; ++c) {
            const int s = o * inner + c;
            bool pass = false;
            if (in) {
                const float sc = (INNER == 8) ? v[c & 7] : __ldg(base + n * inner + c);
                pass = sc > __ldg(thr + s);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, pass);
            if (bal) {
                int pos = 0;
                if (lane == __ffs(bal) - 1) pos = atomicAdd(count + s, __popc(bal));
                pos = __shfl_sync(0xffffffffu, pos, __ffs(bal) - 1) + __popc(bal & ((1u << lane) - 1u));
                if (pass && pos < cap) idx_out[(int64_t)s * cap + pos] = (int32_t)n;
            }
        }
    }
}

// 8 scores per row (classification[B,A,8]): one warp per 32 consecutive rows, lane l holding float4 l and l + 32 of the
// chunk's 64 - every global access of the warp is one contiguous 512-byte run - i.e. classes (l & 1) * 4 .. + 3 of rows
// l >> 1 and 16 + (l >> 1).  A chunk without a single passing score (the usual case: ~99 % of the rows are background)
// costs two loads, eight compares and one vote.
__global__ void __launch_bounds__(256) compact8_kernel(const float* __restrict__ scores, int64_t N, int64_t outer_pitch,
                                                       const float* __restrict__ thr, int cap,
                                                       int32_t* __restrict__ idx_out, int32_t* __restrict__ count) {
    const int o = blockIdx.y, lane = threadIdx.x & 31;
    const float4* base = reinterpret_cast<const float4*>(scores + (int64_t)o * outer_pitch);
    const int c0 = (lane & 1) * 4;
    const float4 t = make_float4(__ldg(thr + o * 8 + c0), __ldg(thr + o * 8 + c0 + 1), __ldg(thr + o * 8 + c0 + 2),
                                 __ldg(thr + o * 8 + c0 + 3));
    const int64_t nchunks = (N + 31) >> 5;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const unsigned lt = (1u << lane) - 1u;
    for (int64_t ch = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); ch < nchunks; ch += nwarps) {
        const int64_t n0 = ch << 5;
        const int nrows = (int)min((int64_t)32, N - n0);
        const int r0 = lane >> 1, r1 = 16 + (lane >> 1);
        float4 v0 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), v1 = v0;
        if (r0 < nrows) v0 = ld_stream(base + n0 * 2 + lane);
        if (r1 < nrows) v1 = ld_stream(base + n0 * 2 + 32 + lane);
        const bool p[2][4] = {{v0.x > t.x, v0.y > t.y, v0.z > t.z, v0.w > t.w}, {v1.x > t.x, v1.y > t.y, v1.z > t.z, v1.w > t.w}};
        const bool any = p[0][0] | p[0][1] | p[0][2] | p[0][3] | p[1][0] | p[1][1] | p[1][2] | p[1][3];
        if (!__any_sync(0xffffffffu, any)) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned bal = __ballot_sync(0xffffffffu, p[h][k]);
                if (bal == 0) continue;
                // even lanes hold class k, odd lanes class 4 + k
                const unsigned mine = bal & ((lane & 1) ? 0xaaaaaaaau : 0x55555555u);
                const int s = o * 8 + c0 + k;
                int pos = 0;
                const int leader = __ffs(mine) - 1;
                if (mine && lane == leader) pos = atomicAdd(count + s, __popc(mine));
                // both parities shuffle in the same instruction: each lane reads its own group's leader
                pos = __shfl_sync(0xffffffffu, pos, mine ? leader : lane) + __popc(mine & lt);
                if (p[h][k] && pos < cap) idx_out[(int64_t)s * cap + pos] = (int32_t)(n0 + (h ? r1 : r0));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------- candidate packing (a10)
// exclusive scan of min(count, cap) over the S segments -> seg_offsets[S+1]   (single CTA; S is small)
__global__ void __launch_bounds__(1024) seg_scan_kernel(const int32_t* __restrict__ count, int S, int cap,
                                                        int32_t* __restrict__ seg_offsets) {
    __shared__ int s_part[1024];
    const int per = (S + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(S, lo + per);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += min(count[i], cap);
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < 1024; ++i) { const int v = s_part[i]; s_part[i] = run; run += v; }
        seg_offsets[S] = run;
    }
    __syncthreads();
    int run = s_part[threadIdx.x];
    for (int i = lo; i < hi; ++i) { seg_offsets[i] = run; run += min(count[i], cap); }
}

// One CTA per segment: sort the (arrival-ordered) candidate indices ascending - the order of the reference's
// boolean-mask gather (3D model.py:380-382) - and pack score / box / source index contiguously.
__global__ void __launch_bounds__(1024) gather_candidates_kernel(const float* __restrict__ scores, int inner, int64_t N,
                                                                 int64_t outer_pitch, const float* __restrict__ boxes,
                                                                 int64_t box_stride, int64_t box_col,
                                                                 const int32_t* __restrict__ idx,
                                                                 const int32_t* __restrict__ count, int cap,
                                                                 const int32_t* __restrict__ seg_offsets,
                                                                 float* __restrict__ cand_scores,
                                                                 float4* __restrict__ cand_boxes,
                                                                 int32_t* __restrict__ cand_src, const BoxDecode dec) {
    extern __shared__ unsigned int s_idx[];
    const int s = blockIdx.x;
    const int n = min(count[s], cap);
    if (n <= 0) return;
    int P = 2;
    while (P < n) P <<= 1;
    for (int i = threadIdx.x; i < P; i += blockDim.x) s_idx[i] = (i < n) ? (unsigned)idx[(int64_t)s * cap + i] : 0xffffffffu;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const unsigned a = s_idx[i], b = s_idx[l];
                const bool up = (i & k) == 0;
                if ((a > b) == up) { s_idx[i] = b; s_idx[l] = a; }
            }
            __syncthreads();
        }
    }
    const int o = s / inner, c = s - o * inner;
    const int off = seg_offsets[s];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int64_t e = s_idx[i];
        cand_scores[off + i] = __ldg(scores + (int64_t)o * outer_pitch + e * inner + c);
        cand_src[off + i] = (int32_t)e;
        if (boxes) {
            const float* bp = boxes + ((int64_t)o * N + e) * box_stride + box_col;
            cand_boxes[off + i] = make_float4(__ldg(bp), __ldg(bp + 1), __ldg(bp + 2), __ldg(bp + 3));
        } else if (dec.reg) {
            cand_boxes[off + i] = decoded_nms_box(dec, o, N, e);
        }
    }
}

}  // namespace g3d

using namespace g3d;

extern "C" int g3d_rowmax(const float* cls, int64_t rows, int64_t C, float* smax, int64_t* amax, int device,
                          void* stream) {
    G3D_REQUIRE(rows >= 0 && C >= 1 && C < (1 << 20), "bad size");
    if (rows == 0) return G3D_OK;
    G3D_REQUIRE(cls && smax && amax, "null pointer");
    G3D_GUARD(device);
    const int64_t blocks = ceil_div(rows, 256);
    const int grid = (int)(blocks < (int64_t)sm_count(device) * 32 ? blocks : (int64_t)sm_count(device) * 32);
    rowmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cls, rows, (int)C, smax, amax);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int64_t g3d_ladder_workspace_bytes(int64_t S, int64_t L) {
    if (S < 0 || L < 0) return G3D_ERR_INVALID;
    return align_up(S * (L + 1) * 4, 256) + align_up(L * 4, 256);
}

static int grid_x_for(int64_t N, int64_t outer, int device) {
    int64_t gx = ceil_div(N, 256);
    const int64_t want = ceil_div((int64_t)sm_count(device) * 8, outer > 0 ? outer : 1);
    if (gx > want) gx = want;
    return (int)(gx < 1 ? 1 : gx);
}

extern "C" int g3d_threshold_ladder(const float* scores, int64_t outer, int64_t inner, int64_t N, int64_t outer_pitch,
                                    const float* thresholds_host, int64_t L, int64_t keep_max, int32_t* rung_out,
                                    int32_t* count_out, float* thr_out, void* workspace, int64_t workspace_bytes,
                                    int device, void* stream) {
    G3D_REQUIRE(outer >= 1 && inner >= 1 && N >= 0 && L >= 1 && L <= kMaxRungs, "bad size (1 <= L <= 512)");
    G3D_REQUIRE(outer <= 65535 && inner <= 64, "outer <= 65535, inner <= 64");
    G3D_REQUIRE(scores || N == 0, "null scores");
    G3D_REQUIRE(thresholds_host && rung_out && count_out && thr_out && workspace, "null pointer");
    const int64_t S = outer * inner;
    G3D_REQUIRE(workspace_bytes >= g3d_ladder_workspace_bytes(S, L), "workspace too small");
    G3D_REQUIRE(((uintptr_t)workspace % 256) == 0, "workspace must be 256-byte aligned");
    for (int64_t i = 1; i < L; ++i)
        G3D_REQUIRE(thresholds_host[i] >= thresholds_host[i - 1], "rungs must be ascending");
    G3D_REQUIRE(thresholds_host[0] > 0.0f, "first rung must be positive");
    G3D_REQUIRE(inner != 8 || ((uintptr_t)scores % 16 == 0 && outer_pitch % 4 == 0), "scores must be 16-byte aligned");
    G3D_GUARD(device);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int* hist = (unsigned int*)workspace;
    float* rungs = (float*)((char*)workspace + align_up(S * (L + 1) * 4, 256));
    G3D_CUDA(cudaMemsetAsync(hist, 0, S * (L + 1) * 4, st));
    G3D_CUDA(cudaMemcpyAsync(rungs, thresholds_host, L * 4, cudaMemcpyHostToDevice, st));
    const float log2_t0 = log2f(thresholds_host[0]);
    // geometric step estimated from the finite rungs (only a search hint: rung_of() fixes the guess up exactly)
    int64_t last = L - 1;
    while (last > 0 && !isfinite(thresholds_host[last])) --last;
    float step = last > 0 ? (log2f(thresholds_host[last]) - log2_t0) / (float)last : 1.0f;
    if (!(step > 1e-6f)) step = 1.0f;
    if (N > 0) {
        dim3 grid((unsigned)grid_x_for(N, outer, device), (unsigned)outer);
        const size_t smem = (size_t)(L + inner * (L + 1)) * 4;
        if (inner == 8) {
            G3D_CUDA(cudaFuncSetAttribute(ladder_hist_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ladder_hist_kernel<8><<<grid, 256, smem, st>>>(scores, N, 8, outer_pitch, rungs, (int)L, log2_t0, 1.0f / step, hist);
        } else {
            G3D_CUDA(cudaFuncSetAttribute(ladder_hist_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ladder_hist_kernel<0><<<grid, 256, smem, st>>>(scores, N, (int)inner, outer_pitch, rungs, (int)L, log2_t0, 1.0f / step, hist);
        }
        G3D_LAUNCH_CHECK();
    }
    ladder_pick_kernel<<<(unsigned)ceil_div(S, 256), 256, 0, st>>>(hist, (int)S, (int)L, rungs, (long long)keep_max,
                                                                   rung_out, count_out, thr_out);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int g3d_filter_compact(const float* scores, int64_t outer, int64_t inner, int64_t N, int64_t outer_pitch,
                                  const float* thr, int64_t cap, int32_t* idx_out, int32_t* count_out, int device,
                                  void* stream) {
    G3D_REQUIRE(outer >= 1 && inner >= 1 && N >= 0 && cap >= 0, "bad size");
    G3D_REQUIRE(outer <= 65535 && inner <= 64 && N < ((int64_t)1 << 31) && cap < ((int64_t)1 << 31), "size out of range");
    G3D_REQUIRE(thr && count_out && (idx_out || cap == 0), "null pointer");
    G3D_REQUIRE(scores || N == 0, "null scores");
    G3D_REQUIRE(inner != 8 || ((uintptr_t)scores % 16 == 0 && outer_pitch % 4 == 0), "scores must be 16-byte aligned");
    G3D_GUARD(device);
    cudaStream_t st = (cudaStream_t)stream;
    G3D_CUDA(cudaMemsetAsync(count_out, 0, outer * inner * 4, st));
    if (N == 0) return G3D_OK;
    dim3 grid((unsigned)grid_x_for(N, outer, device), (unsigned)outer);
    if (inner == 8)
        compact8_kernel<<<grid, 256, 0, st>>>(scores, N, outer_pitch, thr, (int)cap, idx_out, count_out);
    else
        compact_kernel<0><<<grid, 256, 0, st>>>(scores, N, (int)inner, outer_pitch, thr, (int)cap, idx_out, count_out);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

static int gather_candidates_launch(const float* scores, int64_t outer, int64_t inner, int64_t N, int64_t outer_pitch,
                                    const float* boxes, int64_t box_stride, int64_t box_col, const BoxDecode& dec,
                                    const int32_t* idx, const int32_t* count, int64_t cap, int32_t* seg_offsets,
                                    float* cand_scores, float* cand_boxes, int32_t* cand_src, int device, void* stream) {
    G3D_REQUIRE(outer >= 1 && inner >= 1 && N >= 0 && cap >= 1 && cap <= 16384, "bad size (1 <= cap <= 16384)");
    G3D_REQUIRE(outer * inner < ((int64_t)1 << 24), "too many segments");
    G3D_REQUIRE(scores && idx && count && seg_offsets && cand_scores && cand_src, "null pointer");
    G3D_REQUIRE(!boxes || (cand_boxes && box_stride >= 4 && box_col >= 0 && box_col + 4 <= box_stride), "bad box layout");
    G3D_REQUIRE(!(boxes || dec.reg) || ((uintptr_t)cand_boxes % 16) == 0, "cand_boxes must be 16-byte aligned");
    G3D_GUARD(device);
    cudaStream_t st = (cudaStream_t)stream;
    const int S = (int)(outer * inner);
    seg_scan_kernel<<<1, 1024, 0, st>>>(count, S, (int)cap, seg_offsets);
    G3D_LAUNCH_CHECK();
    int P = 2;
    while (P < cap) P <<= 1;
    const size_t smem = (size_t)P * 4;
    G3D_CUDA(cudaFuncSetAttribute(gather_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gather_candidates_kernel<<<S, 1024, smem, st>>>(scores, (int)inner, N, outer_pitch, boxes, box_stride, box_col, idx,
                                                    count, (int)cap, seg_offsets, cand_scores, (float4*)cand_boxes,
                                                    cand_src, dec);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int g3d_gather_candidates(const float* scores, int64_t outer, int64_t inner, int64_t N, int64_t outer_pitch,
                                     const float* boxes, int64_t box_stride, int64_t box_col, const int32_t* idx,
                                     const int32_t* count, int64_t cap, int32_t* seg_offsets, float* cand_scores,
                                     float* cand_boxes, int32_t* cand_src, int device, void* stream) {
    BoxDecode none;
    none.anchors = nullptr; none.reg = nullptr; none.variant = -1; none.per_image_anchors = 0; none.clip = 0;
    none.cw = none.ch = 0.f; none.mean = none.stdv = make_float4(0.f, 0.f, 0.f, 0.f);
    return gather_candidates_launch(scores, outer, inner, N, outer_pitch, boxes, box_stride, box_col, none, idx, count, cap,
                                    seg_offsets, cand_scores, cand_boxes, cand_src, device, stream);
}

extern "C" int g3d_gather_candidates_decoded(const float* scores, int64_t outer, int64_t inner, int64_t N,
                                             int64_t outer_pitch, const float* anchors, int64_t Ba, const float* reg,
                                             int variant, const float* mean_host, const float* std_host, int clip,
                                             float clip_w, float clip_h, const int32_t* idx, const int32_t* count,
                                             int64_t cap, int32_t* seg_offsets, float* cand_scores, float* cand_boxes,
                                             int32_t* cand_src, int device, void* stream) {
    BoxDecode d;
    int rc = make_box_decode(d, anchors, Ba, outer, reg, variant, mean_host, std_host, clip, clip_w, clip_h);
    if (rc != G3D_OK) return rc;
    G3D_REQUIRE(cand_boxes, "null pointer");
    return gather_candidates_launch(scores, outer, inner, N, outer_pitch, nullptr, 4, 0, d, idx, count, cap, seg_offsets,
                                    cand_scores, cand_boxes, cand_src, device, stream);
}

// ------------------------------------------------------------------------------------------- detection assembly
namespace g3d {

// One CTA per segment s = (image o, class c): its kept candidates (NMS order = descending score) become rows
// out_offsets[s] ... of the final tensors: score, class, image index and the box row decoded from the regression
// output (3D: 20 columns; 2D: 4 columns) - the reference's scores[keep] / boxes[keep] gathers and torch.cat
// (3D model.py:384-395, retinanet/model.py:298-309) in one launch.
__global__ void __launch_bounds__(128) assemble_detections_kernel(const int64_t* __restrict__ keep,
                                                                  const int32_t* __restrict__ keep_count,
                                                                  const int32_t* __restrict__ seg_offsets,
                                                                  const int32_t* __restrict__ out_offsets,
                                                                  const float* __restrict__ cand_scores,
                                                                  const int32_t* __restrict__ cand_src, int inner, int64_t N,
                                                                  const BoxDecode dec, float* __restrict__ out_scores,
                                                                  int64_t* __restrict__ out_classes,
                                                                  float* __restrict__ out_boxes,
                                                                  int64_t* __restrict__ out_image, int64_t capacity) {
    const int s = blockIdx.x;
    const int n = keep_count[s];
    const int64_t in0 = seg_offsets[s], out0 = out_offsets[s];
    const int64_t o = s / inner, c = s - o * inner;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const int64_t pos = keep[in0 + k];
        const int64_t e = cand_src[pos];
        const int64_t row = out0 + k;
        if (row >= capacity) break;            // speculative output buffers: the caller re-runs with the exact size
        out_scores[row] = cand_scores[pos];
        out_classes[row] = c;
        out_image[row] = o;
        const float4 an = __ldg(dec.anchors + (dec.per_image_anchors ? o * N + e : e));
        if (dec.variant == G3D_VARIANT_3D) {
            const float4* rp = reinterpret_cast<const float4*>(dec.reg + (o * N + e) * 12);
            const float4 q0 = __ldg(rp), q1 = __ldg(rp + 1), q2 = __ldg(rp + 2);
            const float r[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
            float p[20];
            decode3d_row(r, anchor_geom(an), p);
            float4* op = reinterpret_cast<float4*>(out_boxes + row * 20);
#pragma unroll
            for (int j = 0; j < 5; ++j) op[j] = make_float4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
        } else {
            const float4 dl = __ldg(reinterpret_cast<const float4*>(dec.reg + (o * N + e) * 4));
            reinterpret_cast<float4*>(out_boxes)[row] = decode2d_row(an, dl, dec.mean, dec.stdv, dec.clip, dec.cw, dec.ch);
        }
    }
}

}  // namespace g3d

extern "C" int g3d_exclusive_scan_i32(const int32_t* count, int64_t S, int32_t* offsets, int device, void* stream) {
    G3D_REQUIRE(S >= 0 && S < ((int64_t)1 << 24), "bad size");
    G3D_REQUIRE(offsets && (count || S == 0), "null pointer");
    G3D_GUARD(device);
    seg_scan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(count, (int)S, 0x7fffffff, offsets);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int g3d_assemble_detections(const int64_t* keep, const int32_t* keep_count, const int32_t* seg_offsets,
                                       const int32_t* out_offsets, const float* cand_scores, const int32_t* cand_src,
                                       int64_t outer, int64_t inner, int64_t N, const float* anchors, int64_t Ba,
                                       const float* reg, int variant, const float* mean_host, const float* std_host,
                                       int clip, float clip_w, float clip_h, float* out_scores, int64_t* out_classes,
                                       float* out_boxes, int64_t* out_image, int64_t capacity, int device, void* stream) {
    G3D_REQUIRE(outer >= 1 && inner >= 1 && N >= 0 && outer * inner < ((int64_t)1 << 24) && capacity >= 0, "bad size");
    G3D_REQUIRE(keep_count && seg_offsets && out_offsets, "null pointer");
    BoxDecode d;
    int rc = make_box_decode(d, anchors, Ba, outer, reg, variant, mean_host, std_host, clip, clip_w, clip_h);
    if (rc != G3D_OK) return rc;
    G3D_REQUIRE(((uintptr_t)out_boxes % 16) == 0, "out_boxes must be 16-byte aligned");
    G3D_GUARD(device);
    assemble_detections_kernel<<<(unsigned)(outer * inner), 128, 0, (cudaStream_t)stream>>>(
        keep, keep_count, seg_offsets, out_offsets, cand_scores, cand_src, (int)inner, N, d, out_scores, out_classes,
        out_boxes, out_image, capacity);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}
