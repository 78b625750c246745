// homography.cu — a12..a19: state <-> road-plane space <-> image projection of homography.py.
//
// All projective arithmetic is FP64, as in the reference (space_to_im / im_to_space promote with .double(),
// homography.py:401-402,452-453): a plain FP32 projection misses the 1e-5 parity bound where the projective
// denominator is small.  state_to_space is FP32 with the reference's separately rounded ops
// (homography.py:305-320), space_to_state rounds to FP32 on store (homography.py:281).
//
// The per-camera matrices (P[ncam][2][3][4], H[ncam][2][3][3]; index 1 = the second correspondence of
// Homography_Wrapper, homography.py:793-862) are staged into shared memory once per CTA: the camera index can differ
// per object (the list-of-names API, homography.py:406,457), which would serialise __constant__ accesses, and a
// caller-owned device array keeps the library free of global state.
#include "common.cuh"

namespace g3d {

struct StateF {
    float x, y, l, w, h, dir;
};
__device__ __forceinline__ StateF load_state(const float* __restrict__ states, int64_t S, int64_t i) {
    const float* r = states + i * S;
    StateF s;
    s.x = __ldg(r); s.y = __ldg(r + 1); s.l = __ldg(r + 2); s.w = __ldg(r + 3); s.h = __ldg(r + 4); s.dir = __ldg(r + 5);
    return s;
}
// i24_state_to_space (homography.py:305-320), corner k of 8
__device__ __forceinline__ float space_x(const StateF& s, int k) {
    return (k & 2) ? s.x : __fadd_rn(s.x, __fmul_rn(s.dir, s.l));            // corners 0,1,4,5 carry the +dir*l
}
__device__ __forceinline__ float space_y(const StateF& s, int k) {
    const float half = __fdiv_rn(__fmul_rn(s.dir, s.w), 2.0f);
    return (k & 1) ? __fadd_rn(s.y, half) : __fsub_rn(s.y, half);
}
__device__ __forceinline__ float space_z(const StateF& s, int k) { return (k & 4) ? -s.h : 0.0f; }

__device__ __forceinline__ void stage_mats(double* s_m, const double* __restrict__ g, int count) {
    for (int i = threadIdx.x; i < count; i += blockDim.x) s_m[i] = __ldg(g + i);
    __syncthreads();
}

__device__ __forceinline__ void project(const double* __restrict__ M /*3x4*/, double X, double Y, double Z, double& u,
                                        double& v) {
    const double a = M[0] * X + M[1] * Y + M[2] * Z + M[3];
    const double b = M[4] * X + M[5] * Y + M[6] * Z + M[7];
    const double c = M[8] * X + M[9] * Y + M[10] * Z + M[11];
    const double inv = 1.0 / c;
    u = a * inv;
    v = b * inv;
}
__device__ __forceinline__ void plane_map(const double* __restrict__ H /*3x3*/, double u, double v, double& x, double& y) {
    const double a = H[0] * u + H[1] * v + H[2];
    const double b = H[3] * u + H[4] * v + H[5];
    const double c = H[6] * u + H[7] * v + H[8];
    const double inv = 1.0 / c;
    x = a * inv;
    y = b * inv;
}

// ------------------------------------------------------------------------------------------- a12 state_to_space
__global__ void __launch_bounds__(256) state_to_space_kernel(const float* __restrict__ states, int64_t d, int64_t S,
                                                             float* __restrict__ out) {
    const int64_t total = d * 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i & 7);
        const StateF s = load_state(states, S, i >> 3);
        float* o = out + i * 3;
        o[0] = space_x(s, k);
        o[1] = space_y(s, k);
        o[2] = space_z(s, k);
    }
}

// ------------------------------------------------------------------------------------------- a13 space_to_im
template <typename T>
__global__ void __launch_bounds__(256) space_to_im_kernel(const T* __restrict__ pts, int64_t d, int m,
                                                          const double* __restrict__ P, int ncam,
                                                          const uint8_t* __restrict__ cam, int cam_const, int wrapper,
                                                          double2* __restrict__ out) {
    extern __shared__ double s_m[];
    stage_mats(s_m, P, ncam * 24);
    const int64_t total = d * m;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t obj = i / m;
        const int c = cam ? (int)__ldg(cam + obj) : cam_const;
        int sel = 0;
        if (wrapper) sel = (pts[obj * m * 3 + 1] > (T)60) ? 1 : 0;  // points[:,0,1] > 60 (homography.py:854)
        const T* p = pts + i * 3;
        double u, v;
        project(s_m + (c * 2 + sel) * 12, (double)p[0], (double)p[1], (double)p[2], u, v);
        out[i] = make_double2(u, v);
    }
}

// ------------------------------------------------------------------------------------------- a14 state_to_im
// One thread per (state[, camera]).  P.[X,Y,Z,1] is affine in the corner signs, so the three rows are assembled from
// shared partial products (2 x-terms, 2 y-terms, 1 z-term) instead of 8 full 3x4 products; each corner then costs one
// FP64 reciprocal and two multiplies.  The 8 (u,v) pairs of a thread are staged through shared memory with an XOR
// swizzle (conflict-free 16-byte stores at a 128-byte lane stride) so that every warp store covers whole 128-byte
// output rows: the kernel is bound by the 128 B/state it has to write.
constexpr int kS2IThreads = 256;

template <typename OUT2, bool ALL_CAMS>
__global__ void __launch_bounds__(kS2IThreads, 4) state_to_im_kernel(const float* __restrict__ states, int64_t d, int64_t S,
                                                                  const double* __restrict__ P, int ncam,
                                                                  const uint8_t* __restrict__ cam, int cam_const,
                                                                  int wrapper, OUT2* __restrict__ out) {
    extern __shared__ __align__(16) double s_m[];
    const int mat_doubles = (ncam * 24 + 1) & ~1;                 // keep the staging area 16-byte aligned
    OUT2* stage = reinterpret_cast<OUT2*>(s_m + mat_doubles) + (threadIdx.x & ~31) * 8;   // this warp's 256 slots
    stage_mats(s_m, P, ncam * 24);
    const int lane = threadIdx.x & 31;
    const int64_t items = ALL_CAMS ? d * ncam : d;
    const int64_t stride = (int64_t)gridDim.x * kS2IThreads;
    for (int64_t base = (int64_t)blockIdx.x * kS2IThreads + (threadIdx.x & ~31); base < items; base += stride) {
        const int64_t item = base + lane;
        if (item < items) {
            int64_t obj = item;
            int c;
            if (ALL_CAMS) {
                if (items < (int64_t)0x7fffffff) { obj = (uint32_t)item / (uint32_t)ncam; c = (int)((uint32_t)item - (uint32_t)obj * ncam); }
                else { obj = item / ncam; c = (int)(item - obj * ncam); }
            } else {
                c = cam ? (int)__ldg(cam + obj) : cam_const;
            }
            const StateF s = load_state(states, S, obj);
            const float ylo = space_y(s, 0), yhi = space_y(s, 1);
            const double xf = (double)space_x(s, 0), xb = (double)s.x, y0 = (double)ylo, y1 = (double)yhi, z = (double)(-s.h);
            const int sel = (wrapper && ylo > 60.0f) ? 1 : 0;      // points[:,0,1] > 60 (homography.py:854)
            const double* M = s_m + (c * 2 + sel) * 12;
            double ax[2][3], by[2][3], cz[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                ax[0][r] = M[4 * r] * xf;
                ax[1][r] = M[4 * r] * xb;
                by[0][r] = fma(M[4 * r + 1], y0, M[4 * r + 3]);
                by[1][r] = fma(M[4 * r + 1], y1, M[4 * r + 3]);
                cz[r] = M[4 * r + 2] * z;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int ix = (k >> 1) & 1, iy = k & 1;
                double v0 = ax[ix][0] + by[iy][0], v1 = ax[ix][1] + by[iy][1], v2 = ax[ix][2] + by[iy][2];
                if (k & 4) { v0 += cz[0]; v1 += cz[1]; v2 += cz[2]; }
                const double inv = 1.0 / v2;
                OUT2 o;
                o.x = v0 * inv;
                o.y = v1 * inv;
                stage[lane * 8 + (k ^ (lane & 7))] = o;
            }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int slot = lane + 32 * r;
            const int o_local = slot >> 3;
            const int k = (slot & 7) ^ (o_local & 7);
            const int64_t it = base + o_local;
            if (it < items) out[it * 8 + k] = stage[slot];
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------- a15 im_to_space
template <typename T>
__global__ void __launch_bounds__(256) im_to_space_kernel(const T* __restrict__ pts, const T* __restrict__ heights,
                                                          int64_t d, const double* __restrict__ H, int ncam,
                                                          const uint8_t* __restrict__ cam, int cam_const, int wrapper,
                                                          double* __restrict__ out) {
    extern __shared__ double s_m[];
    stage_mats(s_m, H, ncam * 18);
    const int64_t total = d * 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t obj = i >> 3;
        const int k = (int)(i & 7);
        const int c = cam ? (int)__ldg(cam + obj) : cam_const;
        int sel = 0;
        if (wrapper) {  // boxes[:,0,1] > 60 of the FIRST correspondence's result (homography.py:845)
            double x0, y0;
            plane_map(s_m + (c * 2) * 9, (double)pts[obj * 16], (double)pts[obj * 16 + 1], x0, y0);
            sel = (y0 > 60.0) ? 1 : 0;
        }
        double x, y;
        plane_map(s_m + (c * 2 + sel) * 9, (double)pts[i * 2], (double)pts[i * 2 + 1], x, y);
        double* o = out + i * 3;
        o[0] = x;
        o[1] = y;
        o[2] = (k & 4) ? (double)heights[obj] : 0.0;  // homography.py:427-429
    }
}

// ------------------------------------------------------------------------------------------- a16 space_to_state
// state of one object from its 4 bottom corners (x,y) and |z_bottom - z_top| (homography.py:274-303), in FP64
__device__ __forceinline__ void state_from_bottom(const double* bx, const double* by, double h, float* o) {
    const double fx = bx[0] + bx[1], rx = bx[2] + bx[3];
    const double signed_l = (fx - rx) / 2.0;
    o[0] = (float)(rx / 2.0);
    o[1] = (float)((((by[0] + by[1]) + by[2]) + by[3]) / 4.0);
    o[2] = (float)fabs(signed_l);
    o[3] = (float)fabs(((by[0] + by[2]) - (by[1] + by[3])) / 2.0);
    o[4] = (float)h;
    o[5] = (signed_l > 0.0) ? 1.0f : ((signed_l < 0.0) ? -1.0f : (float)signed_l);  // torch.sign (0 -> 0, NaN -> NaN)
}

template <typename T>
__global__ void __launch_bounds__(256) space_to_state_kernel(const T* __restrict__ pts, int64_t d, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d; i += (int64_t)gridDim.x * blockDim.x) {
        const T* p = pts + i * 24;
        // computed in the input dtype as torch does, stored to float32
        const T fx = p[0] + p[3], rx = p[6] + p[9];
        const T signed_l = (fx - rx) / (T)2;
        T hs = (T)0;
#pragma unroll
        for (int k = 0; k < 4; ++k) hs += (T)fabs((double)(p[3 * k + 2] - p[3 * (k + 4) + 2]));
        float* o = out + i * 6;
        o[0] = (float)(rx / (T)2);
        o[1] = (float)((((p[1] + p[4]) + p[7]) + p[10]) / (T)4);
        o[2] = (float)fabs((double)signed_l);
        o[3] = (float)fabs((double)(((p[1] + p[7]) - (p[4] + p[10])) / (T)2));
        o[4] = (float)(hs / (T)4);
        o[5] = (signed_l > (T)0) ? 1.0f : ((signed_l < (T)0) ? -1.0f : (float)signed_l);
    }
}

// ------------------------------------------------------------------------------------------- a17 im_to_state (fused)
// Only the 4 bottom corners enter the state (x, y, l, w, dir) and h = |height| (z is 0 / +height): 4 plane maps per
// object instead of 8, no [d,8,3] intermediate.
template <typename T>
__device__ __forceinline__ void im_to_state_one(const T* __restrict__ p /*object's 16 image coords*/, double height,
                                                const double* __restrict__ Hc /*[2][9] of the camera*/, int wrapper,
                                                float* o) {
    double bx[4], by[4];
    const double* H = Hc;
#pragma unroll
    for (int k = 0; k < 4; ++k) plane_map(H, (double)p[2 * k], (double)p[2 * k + 1], bx[k], by[k]);
    if (wrapper && by[0] > 60.0) {  // re-map with the second correspondence (homography.py:840-847)
        H = Hc + 9;
#pragma unroll
        for (int k = 0; k < 4; ++k) plane_map(H, (double)p[2 * k], (double)p[2 * k + 1], bx[k], by[k]);
    }
    state_from_bottom(bx, by, fabs(height), o);
}

__device__ __forceinline__ double2 ld_half_line(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// Four lanes per object, lane q owning bottom corner q: the quad's loads cover one contiguous 32/64-byte run per
// object (coalesced), the four plane maps run in parallel and the state is assembled with quad shuffles.
template <typename T>
__global__ void __launch_bounds__(256) im_to_state_kernel(const T* __restrict__ pts, const T* __restrict__ heights,
                                                          int64_t d, const double* __restrict__ H, int ncam,
                                                          const uint8_t* __restrict__ cam, int cam_const, int wrapper,
                                                          float* __restrict__ out) {
    extern __shared__ double s_m[];
    stage_mats(s_m, H, ncam * 18);
    const int lane = threadIdx.x & 31, q = lane & 3, qbase = lane & ~3;
    const int64_t nquads = (int64_t)gridDim.x * (blockDim.x >> 2);
    const int64_t d_round = ((d + 7) >> 3) << 3;      // whole warps stay in the loop (full-mask shuffles)
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 2) + (threadIdx.x >> 2); i < d_round; i += nquads) {
        const bool in = i < d;
        double u = 0.0, v = 0.0, hgt = 0.0;
        int c = 0;
        if (in) {
            c = cam ? (int)__ldg(cam + i) : cam_const;
            if (sizeof(T) == 8) {
                // only the bottom face (first 64 of the object's 128 bytes) is needed: ask L2 for 64-byte fills so the
                // unused top-face half of every line is not fetched from HBM
                const double2 uv = ld_half_line(reinterpret_cast<const double2*>(pts + i * 16) + q);
                u = uv.x; v = uv.y;
            } else {
                const float2 uv = __ldg(reinterpret_cast<const float2*>(pts + i * 16) + q);
                u = (double)uv.x; v = (double)uv.y;
            }
            hgt = (double)heights[i];
        }
        double x, y;
        plane_map(s_m + c * 18, u, v, x, y);
        if (wrapper) {  // boxes[:,0,1] > 60 of the first correspondence decides for the whole object (:845)
            const double y_first = shfl_d(y, qbase);
            if (y_first > 60.0) plane_map(s_m + c * 18 + 9, u, v, x, y);
        }
        // quad reductions (space_to_state, homography.py:274-303)
        const double xpair = x + shfl_xor_d(x, 1);                 // lanes 0,1: x0+x1 (front); lanes 2,3: x2+x3 (rear)
        const double other = shfl_xor_d(xpair, 2);
        const double front = (q < 2) ? xpair : other, rear = (q < 2) ? other : xpair;
        const double ypair = y + shfl_xor_d(y, 1);
        const double ysum = ypair + shfl_xor_d(ypair, 2);
        const double ycol = y + shfl_xor_d(y, 2);                  // lanes 0,2: y0+y2 ; lanes 1,3: y1+y3
        const double ycol_o = shfl_xor_d(ycol, 1);
        const double even = (q & 1) ? ycol_o : ycol, odd = (q & 1) ? ycol : ycol_o;
        const double signed_l = (front - rear) / 2.0;
        float2 o;
        if (q == 0) {
            o = make_float2((float)(rear / 2.0), (float)(ysum / 4.0));
        } else if (q == 1) {
            o = make_float2((float)fabs(signed_l), (float)fabs((even - odd) / 2.0));
        } else {
            const float dir = (signed_l > 0.0) ? 1.0f : ((signed_l < 0.0) ? -1.0f : (float)signed_l);
            o = make_float2((float)fabs(hgt), dir);
        }
        if (in && q < 3) reinterpret_cast<float2*>(out + i * 6)[q] = o;
    }
}

// ------------------------------------------------------------------------------------------- a18 height_from_template
// |top - bottom|_1 of the corner means (homography.py:541-548), in the dtype torch would use for that operand
template <typename T>
__device__ __forceinline__ T im_height(const T* __restrict__ b) {
    T tx = (T)0, ty = (T)0, bx = (T)0, by = (T)0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        bx += b[2 * k]; by += b[2 * k + 1];
        tx += b[2 * (k + 4)]; ty += b[2 * (k + 4) + 1];
    }
    const T dx = tx / (T)4 - bx / (T)4, dy = ty / (T)4 - by / (T)4;
    return (T)sqrt((double)(dx * dx)) + (T)sqrt((double)(dy * dy));
}

template <typename TB, typename TH, typename BX, typename OUT>
__global__ void __launch_bounds__(256) height_from_template_kernel(const TB* __restrict__ tb, const TH* __restrict__ th,
                                                                   const BX* __restrict__ bx, int64_t d,
                                                                   OUT* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d; i += (int64_t)gridDim.x * blockDim.x) {
        const TB t_im = im_height<TB>(tb + i * 16);
        const BX b_im = im_height<BX>(bx + i * 16);
        // template_ratio in promote(TB, TH); result in promote(all) == OUT
        if (sizeof(TB) == 8 || sizeof(TH) == 8) {
            const double ratio = (double)t_im / (double)th[i];
            out[i] = (OUT)((double)b_im / ratio);
        } else {
            const float ratio = (float)t_im / (float)th[i];
            out[i] = (OUT)((OUT)b_im / (OUT)ratio);
        }
    }
}

// ------------------------------------------------------------------------------------------- two-pass refinement
// MC3D_crop_tracker.py:364-370: s0 = im_to_state(p, h0); repro = state_to_im(s0); h1 = height_from_template(repro, h0, p);
// s1 = im_to_state(p, h1).  One thread per object, everything in registers.
template <typename T>
__global__ void __launch_bounds__(256) im_to_state_refined_kernel(const T* __restrict__ pts, const T* __restrict__ heights,
                                                                  int64_t d, const double* __restrict__ H,
                                                                  const double* __restrict__ P, int ncam,
                                                                  const uint8_t* __restrict__ cam, int cam_const,
                                                                  int wrapper, float* __restrict__ out,
                                                                  double* __restrict__ heights_out) {
    extern __shared__ double s_m[];  // [ncam*18 H | ncam*24 P]
    for (int i = threadIdx.x; i < ncam * 18; i += blockDim.x) s_m[i] = __ldg(H + i);
    for (int i = threadIdx.x; i < ncam * 24; i += blockDim.x) s_m[ncam * 18 + i] = __ldg(P + i);
    __syncthreads();
    const double* sH = s_m;
    const double* sP = s_m + ncam * 18;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = cam ? (int)__ldg(cam + i) : cam_const;
        const T* p = pts + i * 16;
        const T h0 = heights[i];
        float o[6];
        im_to_state_one<T>(p, (double)h0, sH + c * 18, wrapper, o);
        // reproject the pass-1 state (float32, as the reference's intermediate) and measure its image height
        StateF s;
        s.x = o[0]; s.y = o[1]; s.l = o[2]; s.w = o[3]; s.h = o[4]; s.dir = o[5];
        int sel = 0;
        if (wrapper) sel = (space_y(s, 0) > 60.0f) ? 1 : 0;
        const double* M = sP + (c * 2 + sel) * 12;
        double tx = 0, ty = 0, bx = 0, by = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            double u, v;
            project(M, (double)space_x(s, k), (double)space_y(s, k), (double)space_z(s, k), u, v);
            if (k < 4) { bx += u; by += v; } else { tx += u; ty += v; }
        }
        const double dxt = tx / 4.0 - bx / 4.0, dyt = ty / 4.0 - by / 4.0;
        const double t_im = sqrt(dxt * dxt) + sqrt(dyt * dyt);
        const double ratio = t_im / (double)h0;
        const T b_im = im_height<T>(p);
        const double h1 = (double)b_im / ratio;
        if (heights_out) heights_out[i] = h1;
        im_to_state_one<T>(p, h1, sH + c * 18, wrapper, o);
        float2* dst = reinterpret_cast<float2*>(out + i * 6);
        dst[0] = make_float2(o[0], o[1]);
        dst[1] = make_float2(o[2], o[3]);
        dst[2] = make_float2(o[4], o[5]);
    }
}

// ------------------------------------------------------------------------------------------- a19 footprints / boxes
__global__ void __launch_bounds__(256) state_footprint_kernel(const float* __restrict__ states, int64_t d, int64_t S,
                                                              float4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d; i += (int64_t)gridDim.x * blockDim.x) {
        const StateF s = load_state(states, S, i);
        const float xf = space_x(s, 0), xb = space_x(s, 2), y0 = space_y(s, 0), y1 = space_y(s, 1);
        out[i] = make_float4(fminf(xf, xb), fminf(y0, y1), fmaxf(xf, xb), fmaxf(y0, y1));
    }
}

template <typename T>
__global__ void __launch_bounds__(256) corners_to_box_kernel(const T* __restrict__ pts, int64_t d, T* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d; i += (int64_t)gridDim.x * blockDim.x) {
        const T* p = pts + i * 16;
        T x0 = p[0], y0 = p[1], x1 = p[0], y1 = p[1];
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            x0 = p[2 * k] < x0 ? p[2 * k] : x0;          x1 = p[2 * k] > x1 ? p[2 * k] : x1;
            y0 = p[2 * k + 1] < y0 ? p[2 * k + 1] : y0;  y1 = p[2 * k + 1] > y1 ? p[2 * k + 1] : y1;
        }
        T* o = out + i * 4;
        o[0] = x0; o[1] = y0; o[2] = x1; o[3] = y1;
    }
}

static inline int grid_for(int64_t n, int device) {
    const int64_t b = ceil_div(n, 256);
    const int64_t cap = (int64_t)sm_count(device) * 16;
    return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}
static int check_cams(int64_t ncam, const void* cam, int cam_const) {
    G3D_REQUIRE(ncam >= 1 && ncam <= 256, "1 <= ncam <= 256");
    G3D_REQUIRE(cam || (cam_const >= 0 && cam_const < ncam), "camera index out of range");
    return G3D_OK;
}

}  // namespace g3d

using namespace g3d;

extern "C" int g3d_state_to_space(const float* states, int64_t d, int64_t S, float* out, int device, void* stream) {
    G3D_REQUIRE(d >= 0 && S >= 6, "states need >= 6 columns");
    if (d == 0) return G3D_OK;
    G3D_REQUIRE(states && out, "null pointer");
    G3D_GUARD(device);
    state_to_space_kernel<<<grid_for(d * 8, device), 256, 0, (cudaStream_t)stream>>>(states, d, S, out);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int g3d_space_to_im(const void* pts, int pts_is_f64, int64_t d, int64_t m, const double* P, int64_t ncam,
                               const uint8_t* cam, int cam_const, int wrapper, double* out, int device, void* stream) {
    G3D_REQUIRE(d >= 0 && m >= 1 && m < (1 << 20), "bad size");
    int rc = check_cams(ncam, cam, cam_const);
    if (rc) return rc;
    if (d == 0) return G3D_OK;
    G3D_REQUIRE(pts && P && out, "null pointer");
    G3D_REQUIRE(((uintptr_t)out % 16) == 0, "out must be 16-byte aligned");
    G3D_GUARD(device);
    const size_t smem = (size_t)ncam * 24 * 8;
    if (pts_is_f64) {
        G3D_CUDA(cudaFuncSetAttribute(space_to_im_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        space_to_im_kernel<double><<<grid_for(d * m, device), 256, smem, (cudaStream_t)stream>>>(
            (const double*)pts, d, (int)m, P, (int)ncam, cam, cam_const, wrapper, (double2*)out);
    } else {
        G3D_CUDA(cudaFuncSetAttribute(space_to_im_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        space_to_im_kernel<float><<<grid_for(d * m, device), 256, smem, (cudaStream_t)stream>>>(
            (const float*)pts, d, (int)m, P, (int)ncam, cam, cam_const, wrapper, (double2*)out);
    }
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int g3d_state_to_im(const float* states, int64_t d, int64_t S, const double* P, int64_t ncam,
                               const uint8_t* cam, int cam_const, int wrapper, int all_cams, void* out, int out_f32,
                               int device, void* stream) {
    G3D_REQUIRE(d >= 0 && S >= 6, "states need >= 6 columns");
    int rc = check_cams(ncam, cam, all_cams ? 0 : cam_const);
    if (rc) return rc;
    if (d == 0) return G3D_OK;
    G3D_REQUIRE(states && P && out, "null pointer");
    G3D_REQUIRE(((uintptr_t)out % 16) == 0, "out must be 16-byte aligned");
    G3D_GUARD(device);
    const int64_t items = d * (all_cams ? ncam : 1);
    const size_t mat_bytes = (size_t)((ncam * 24 + 1) & ~1) * 8;
    cudaStream_t st = (cudaStream_t)stream;
#define S2I_LAUNCH(OUT2, ALLC)                                                                                         \
    do {                                                                                                               \
        const size_t smem = mat_bytes + sizeof(OUT2) * 8 * kS2IThreads;                                                \
        G3D_CUDA(cudaFuncSetAttribute(state_to_im_kernel<OUT2, ALLC>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                      (int)smem));                                                                     \
        state_to_im_kernel<OUT2, ALLC><<<grid_for(items, device), kS2IThreads, smem, st>>>(states, d, S, P, (int)ncam, cam,    \
                                                                                   cam_const, wrapper, (OUT2*)out);    \
    } while (0)
    if (out_f32) {
        if (all_cams) S2I_LAUNCH(float2, true); else S2I_LAUNCH(float2, false);
    } else {
        if (all_cams) S2I_LAUNCH(double2, true); else S2I_LAUNCH(double2, false);
    }
#undef S2I_LAUNCH
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int g3d_im_to_space(const void* pts, const void* heights, int in_is_f64, int64_t d, const double* H,
                               int64_t ncam, const uint8_t* cam, int cam_const, int wrapper, double* out, int device,
                               void* stream) {
    G3D_REQUIRE(d >= 0, "bad size");
    int rc = check_cams(ncam, cam, cam_const);
    if (rc) return rc;
    if (d == 0) return G3D_OK;
    G3D_REQUIRE(pts && heights && H && out, "null pointer");
    G3D_GUARD(device);
    const size_t smem = (size_t)ncam * 18 * 8;
    if (in_is_f64)
        im_to_space_kernel<double><<<grid_for(d * 8, device), 256, smem, (cudaStream_t)stream>>>(
            (const double*)pts, (const double*)heights, d, H, (int)ncam, cam, cam_const, wrapper, out);
    else
        im_to_space_kernel<float><<<grid_for(d * 8, device), 256, smem, (cudaStream_t)stream>>>(
            (const float*)pts, (const float*)heights, d, H, (int)ncam, cam, cam_const, wrapper, out);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int g3d_space_to_state(const void* pts, int pts_is_f64, int64_t d, float* out, int device, void* stream) {
    G3D_REQUIRE(d >= 0, "bad size");
    if (d == 0) return G3D_OK;
    G3D_REQUIRE(pts && out, "null pointer");
    G3D_GUARD(device);
    if (pts_is_f64)
        space_to_state_kernel<double><<<grid_for(d, device), 256, 0, (cudaStream_t)stream>>>((const double*)pts, d, out);
    else
        space_to_state_kernel<float><<<grid_for(d, device), 256, 0, (cudaStream_t)stream>>>((const float*)pts, d, out);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int g3d_im_to_state(const void* pts, const void* heights, int in_is_f64, int64_t d, const double* H,
                               int64_t ncam, const uint8_t* cam, int cam_const, int wrapper, float* out, int device,
                               void* stream) {
    G3D_REQUIRE(d >= 0, "bad size");
    int rc = check_cams(ncam, cam, cam_const);
    if (rc) return rc;
    if (d == 0) return G3D_OK;
    G3D_REQUIRE(pts && heights && H && out, "null pointer");
    G3D_REQUIRE(((uintptr_t)out % 8) == 0, "out must be 8-byte aligned");
    G3D_GUARD(device);
    const size_t smem = (size_t)ncam * 18 * 8;
    G3D_REQUIRE(((uintptr_t)pts % 16) == 0, "pts must be 16-byte aligned");
    if (in_is_f64)
        im_to_state_kernel<double><<<grid_for(d * 4, device), 256, smem, (cudaStream_t)stream>>>(
            (const double*)pts, (const double*)heights, d, H, (int)ncam, cam, cam_const, wrapper, out);
    else
        im_to_state_kernel<float><<<grid_for(d * 4, device), 256, smem, (cudaStream_t)stream>>>(
            (const float*)pts, (const float*)heights, d, H, (int)ncam, cam, cam_const, wrapper, out);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int g3d_height_from_template(const void* tb, int tb_is_f64, const void* th, int th_is_f64, const void* bx,
                                        int bx_is_f64, int64_t d, void* out, int device, void* stream) {
    G3D_REQUIRE(d >= 0, "bad size");
    if (d == 0) return G3D_OK;
    G3D_REQUIRE(tb && th && bx && out, "null pointer");
    G3D_GUARD(device);
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid_for(d, device);
    const int key = (tb_is_f64 ? 4 : 0) | (th_is_f64 ? 2 : 0) | (bx_is_f64 ? 1 : 0);
#define HFT(TB, TH, BX, OUT) \
    height_from_template_kernel<TB, TH, BX, OUT><<<g, 256, 0, st>>>((const TB*)tb, (const TH*)th, (const BX*)bx, d, (OUT*)out)
    switch (key) {
        case 0: HFT(float, float, float, float); break;
        case 1: HFT(float, float, double, double); break;
        case 2: HFT(float, double, float, double); break;
        case 3: HFT(float, double, double, double); break;
        case 4: HFT(double, float, float, double); break;
        case 5: HFT(double, float, double, double); break;
        case 6: HFT(double, double, float, double); break;
        default: HFT(double, double, double, double); break;
    }
#undef HFT
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int g3d_im_to_state_refined(const void* pts, const void* heights, int in_is_f64, int64_t d, const double* H,
                                       const double* P, int64_t ncam, const uint8_t* cam, int cam_const, int wrapper,
                                       float* out, double* heights_out, int device, void* stream) {
    G3D_REQUIRE(d >= 0, "bad size");
    int rc = check_cams(ncam, cam, cam_const);
    if (rc) return rc;
    if (d == 0) return G3D_OK;
    G3D_REQUIRE(pts && heights && H && P && out, "null pointer");
    G3D_REQUIRE(((uintptr_t)out % 8) == 0, "out must be 8-byte aligned");
    G3D_GUARD(device);
    const size_t smem = (size_t)ncam * 42 * 8;
    if (in_is_f64) {
        G3D_CUDA(cudaFuncSetAttribute(im_to_state_refined_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        im_to_state_refined_kernel<double><<<grid_for(d, device), 256, smem, (cudaStream_t)stream>>>(
            (const double*)pts, (const double*)heights, d, H, P, (int)ncam, cam, cam_const, wrapper, out, heights_out);
    } else {
        G3D_CUDA(cudaFuncSetAttribute(im_to_state_refined_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        im_to_state_refined_kernel<float><<<grid_for(d, device), 256, smem, (cudaStream_t)stream>>>(
            (const float*)pts, (const float*)heights, d, H, P, (int)ncam, cam, cam_const, wrapper, out, heights_out);
    }
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int g3d_state_footprint(const float* states, int64_t d, int64_t S, float* out, int device, void* stream) {
    G3D_REQUIRE(d >= 0 && S >= 6, "states need >= 6 columns");
    if (d == 0) return G3D_OK;
    G3D_REQUIRE(states && out, "null pointer");
    G3D_REQUIRE(((uintptr_t)out % 16) == 0, "out must be 16-byte aligned");
    G3D_GUARD(device);
    state_footprint_kernel<<<grid_for(d, device), 256, 0, (cudaStream_t)stream>>>(states, d, S, (float4*)out);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int g3d_corners_to_box(const void* pts, int is_f64, int64_t d, void* out, int device, void* stream) {
    G3D_REQUIRE(d >= 0, "bad size");
    if (d == 0) return G3D_OK;
    G3D_REQUIRE(pts && out, "null pointer");
    G3D_GUARD(device);
    if (is_f64)
        corners_to_box_kernel<double><<<grid_for(d, device), 256, 0, (cudaStream_t)stream>>>((const double*)pts, d, (double*)out);
    else
        corners_to_box_kernel<float><<<grid_for(d, device), 256, 0, (cudaStream_t)stream>>>((const float*)pts, d, (float*)out);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}
