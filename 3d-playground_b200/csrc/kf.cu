// kf.cu — SURVEY §8(f)-4: Torch_KF.predict / Torch_KF.update of util_track/kf.py (:292-336, :339-403), the batched
// Kalman filter the trackers run between the geometry kernels.  The reference builds per-object copies of F / H / Q / R
// with .repeat() and pushes them through bmm and a batched inverse on the CPU; here one thread owns one object and its
// S x S covariance lives in registers (S = 6 states, M = 5 measurements in the trackers; up to 8 x 8 supported).
// Arithmetic is FP32 like the reference's (FP64 where torch promotes: the innovation z + mu_R - H x, and the Q scaling
// when dt is a per-object float64 tensor); sums run in index order with separately rounded products.
#include "common.cuh"

namespace g3d {

constexpr int kKfMax = 8;

struct KfModel {
    float F[kKfMax * kKfMax];
    float Q[kKfMax * kKfMax];
    float H[kKfMax * kKfMax];
    float R[kKfMax * kKfMax];
    float mu_R[kKfMax];
};

// predict (kf.py:292-336): F_rep = F with F_rep[0,5] = D * dt;  X = F_rep X;  P = F_rep P F_rep^T + Q * dt / dt_default
template <int S>
__global__ void __launch_bounds__(128) kf_predict_kernel(float* __restrict__ X, float* __restrict__ P,
                                                         const float* __restrict__ D, const double* __restrict__ dt_arr,
                                                         double dt_scalar, double dt_default, double* __restrict__ T,
                                                         int64_t n, int s_rt, const KfModel m) {
    const int SS = (S > 0) ? S : s_rt;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t ii = i;
        const bool live = true;
        const double dt = dt_arr ? dt_arr[i] : dt_scalar;
        const double dt_over_default = dt / dt_default;
        float F[kKfMax][kKfMax], Pm[kKfMax][kKfMax], A[kKfMax][kKfMax], x[kKfMax], xn[kKfMax], Pn[kKfMax][kKfMax];
#pragma unroll
        for (int r = 0; r < kKfMax; ++r)
#pragma unroll
            for (int c = 0; c < kKfMax; ++c)
                if (r < SS && c < SS) F[r][c] = m.F[r * SS + c];
        if (S == 6) {   // 144-byte rows: nine 16-byte loads per object instead of 36 scalar ones
            const float4* prow = reinterpret_cast<const float4*>(P + i * 36);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const float4 v = prow[q];
                const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) Pm[(4 * q + k) / 6][(4 * q + k) % 6] = e[k];
            }
        } else {
#pragma unroll
            for (int r = 0; r < kKfMax; ++r)
#pragma unroll
                for (int c = 0; c < kKfMax; ++c)
                    if (r < SS && c < SS) Pm[r][c] = P[(i * SS + r) * SS + c];
        }
        if (SS > 5) F[0][5] = (float)((double)D[ii] * dt);           // kf.py:315 (hard-wired position / speed slots)
#pragma unroll
        for (int r = 0; r < kKfMax; ++r) if (r < SS) x[r] = X[ii * SS + r];
#pragma unroll
        for (int r = 0; r < kKfMax; ++r) {
            if (r >= SS) continue;
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < kKfMax; ++k) if (k < SS) acc = __fadd_rn(acc, __fmul_rn(F[r][k], x[k]));
            xn[r] = acc;
        }
        // A = F P
#pragma unroll
        for (int r = 0; r < kKfMax; ++r)
#pragma unroll
            for (int c = 0; c < kKfMax; ++c) {
                if (r >= SS || c >= SS) continue;
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k < kKfMax; ++k) if (k < SS) acc = __fadd_rn(acc, __fmul_rn(F[r][k], Pm[k][c]));
                A[r][c] = acc;
            }
        // P = A F^T + Q * dt / dt_default
#pragma unroll
        for (int r = 0; r < kKfMax; ++r)
#pragma unroll
            for (int c = 0; c < kKfMax; ++c) {
                if (r >= SS || c >= SS) continue;
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k < kKfMax; ++k) if (k < SS) acc = __fadd_rn(acc, __fmul_rn(A[r][k], F[c][k]));
                float out;
                if (dt_arr) {   // float * double tensor / python float: the scaling and the sum in double, then .float()
                    const double q = (double)m.Q[r * SS + c] * dt_over_default;   // = Q * dt / dt_default to 1 ulp of FP64
                    out = (float)((double)acc + q);
                } else {        // float tensor * python scalar: FP32 with the scalar rounded to FP32
                    const float q = __fdiv_rn(__fmul_rn(m.Q[r * SS + c], (float)dt), (float)dt_default);
                    out = __fadd_rn(acc, q);
                }
                Pn[r][c] = out;
            }
        if (S == 6) {
            float4* prow = reinterpret_cast<float4*>(P + i * 36);
#pragma unroll
            for (int q = 0; q < 9; ++q)
                prow[q] = make_float4(Pn[(4 * q) / 6][(4 * q) % 6], Pn[(4 * q + 1) / 6][(4 * q + 1) % 6],
                                      Pn[(4 * q + 2) / 6][(4 * q + 2) % 6], Pn[(4 * q + 3) / 6][(4 * q + 3) % 6]);
        } else {
#pragma unroll
            for (int r = 0; r < kKfMax; ++r)
#pragma unroll
                for (int c = 0; c < kKfMax; ++c)
                    if (r < SS && c < SS) P[(i * SS + r) * SS + c] = Pn[r][c];
        }
        if (live) {
#pragma unroll
            for (int r = 0; r < kKfMax; ++r) if (r < SS) X[i * SS + r] = xn[r];
            if (T) T[i] += dt;
        }
    }
}

// predict for S == 6 and F = identity apart from F[0][5] (every filter of the reference: kf.py:58,315), written out: the
// products of the general kernel with the exact 0 / 1 entries are exact and adding +0 changes nothing, so
//   x0 += f x5;   row 0 of P += f * row 5;   then column 0 of P += f * column 5;   P += Q * dt / dt_default
// gives the same floats for finite inputs - in place, with 36 + a few live registers instead of three 6 x 6 matrices,
// hence twice the CTAs per SM for this latency-bound, 364-bytes-per-object kernel.
__global__ void __launch_bounds__(128, 6) kf_predict_fid6_kernel(float* __restrict__ X, float* __restrict__ P,
                                                                 const float* __restrict__ D, const double* __restrict__ dt_arr,
                                                                 double dt_scalar, double dt_default, double* __restrict__ T,
                                                                 int64_t n, const KfModel m) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double dt = dt_arr ? dt_arr[i] : dt_scalar;
        const double dt_over_default = dt / dt_default;
        const float f = (float)((double)D[i] * dt);                  // kf.py:315
        float p[36];
        float4* prow = reinterpret_cast<float4*>(P + i * 36);
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const float4 v = prow[q];
            p[4 * q] = v.x; p[4 * q + 1] = v.y; p[4 * q + 2] = v.z; p[4 * q + 3] = v.w;
        }
        float2* xrow = reinterpret_cast<float2*>(X + i * 6);
        const float2 x01 = xrow[0], x45 = xrow[2];
        xrow[0] = make_float2(__fadd_rn(x01.x, __fmul_rn(f, x45.y)), x01.y);
#pragma unroll
        for (int c = 0; c < 6; ++c) p[c] = __fadd_rn(p[c], __fmul_rn(f, p[30 + c]));          // A = F P: row 0
#pragma unroll
        for (int r = 0; r < 6; ++r) p[6 * r] = __fadd_rn(p[6 * r], __fmul_rn(p[6 * r + 5], f)); // A F^T: column 0
#pragma unroll
        for (int e = 0; e < 36; ++e) {
            if (dt_arr) {   // float * double tensor / python float: the scaling and the sum in double, then .float()
                p[e] = (float)((double)p[e] + (double)m.Q[e] * dt_over_default);
            } else {        // float tensor * python scalar: FP32 with the scalar rounded to FP32
                p[e] = __fadd_rn(p[e], __fdiv_rn(__fmul_rn(m.Q[e], (float)dt), (float)dt_default));
            }
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) prow[q] = make_float4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
        if (T) T[i] += dt;
    }
}

// update (kf.py:339-403) of the objects rows[j]:  y = z + mu_R - H x;  S = H P H^T + R;  K = P H^T S^-1;
// x += K y;  P = (I - K H) P
// HSEL: H = [I_M | 0] (the measurement is the first M states - every tracker configuration of the reference): H P, P H^T and
// K H are then sub-blocks of P and K; multiplying by the exact 0 / 1 entries gives the same floats, so the generic path and
// this one agree bit for bit on finite inputs, with 60 fewer live registers.
// m x m inverse by Gauss-Jordan on [A | I] with partial pivoting; arrays in local memory (the rare path of kf_update_kernel)
__device__ __noinline__ void invert_with_exchanges(float (*A)[kKfMax], float (*Ai)[kKfMax], int m) {
    for (int a = 0; a < m; ++a)
        for (int b = 0; b < m; ++b) Ai[a][b] = (a == b) ? 1.0f : 0.0f;
    for (int col = 0; col < m; ++col) {
        int piv = col;
        float best = fabsf(A[col][col]);
        for (int r = col + 1; r < m; ++r)
            if (fabsf(A[r][col]) > best) { best = fabsf(A[r][col]); piv = r; }
        if (piv != col)
            for (int c = 0; c < m; ++c) {
                const float t = A[col][c]; A[col][c] = A[piv][c]; A[piv][c] = t;
                const float u = Ai[col][c]; Ai[col][c] = Ai[piv][c]; Ai[piv][c] = u;
            }
        const float inv = 1.0f / A[col][col];
        for (int c = 0; c < m; ++c) { A[col][c] *= inv; Ai[col][c] *= inv; }
        for (int r = 0; r < m; ++r) {
            if (r == col) continue;
            const float f = A[r][col];
            for (int c = 0; c < m; ++c) {
                A[r][c] = fmaf(-f, A[col][c], A[r][c]);
                Ai[r][c] = fmaf(-f, Ai[col][c], Ai[r][c]);
            }
        }
    }
}

template <int S, int M, bool HSEL>
__global__ void __launch_bounds__(128, 5) kf_update_kernel(float* __restrict__ X, float* __restrict__ P,
                                                        const int64_t* __restrict__ rows, const double* __restrict__ z,
                                                        int64_t mcount, int s_rt, int m_rt, const KfModel mdl) {
    const int SS = (S > 0) ? S : s_rt, MM = (M > 0) ? M : m_rt;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < mcount; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = rows[j];
        float Pm[kKfMax][kKfMax], x[kKfMax], HP[kKfMax][kKfMax], Sm[kKfMax][kKfMax], y[kKfMax];
        // 144-byte rows: nine 16-byte loads per object instead of 36 scalar ones.  `volatile` + memory clobber: the matrix is
        // read TWICE (see below) and the second read must be a real load (L1 hit), not the first one's registers kept alive
        auto load_P6 = [&]() {
            const float4* prow = reinterpret_cast<const float4*>(P + i * 36);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                float4 v;
                asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(prow + q) : "memory");
                const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) Pm[(4 * q + k) / 6][(4 * q + k) % 6] = e[k];
            }
        };
        if (S == 6) {
            load_P6();
            const float2* xrow = reinterpret_cast<const float2*>(X + i * 6);
#pragma unroll
            for (int q = 0; q < 3; ++q) { const float2 v = xrow[q]; x[2 * q] = v.x; x[2 * q + 1] = v.y; }
        } else {
#pragma unroll
            for (int r = 0; r < kKfMax; ++r) {
                if (r < SS) x[r] = X[i * SS + r];
#pragma unroll
                for (int c = 0; c < kKfMax; ++c)
                    if (r < SS && c < SS) Pm[r][c] = P[(i * SS + r) * SS + c];
            }
        }
        // innovation: double(z) + mu_R - float(x H^T), then .float()
#pragma unroll
        for (int a = 0; a < kKfMax; ++a) {
            if (a >= MM) continue;
            float hx = 0.0f;
            if (HSEL) hx = x[a];
            else {
#pragma unroll
                for (int k = 0; k < kKfMax; ++k) if (k < SS) hx = __fadd_rn(hx, __fmul_rn(x[k], mdl.H[a * SS + k]));
            }
            y[a] = (float)(z[j * MM + a] + (double)mdl.mu_R[a] - (double)hx);
        }
        // HP = H P  [M,S];  S = HP H^T + R  [M,M]
#pragma unroll
        for (int a = 0; a < kKfMax; ++a)
#pragma unroll
            for (int c = 0; c < kKfMax; ++c) {
                if (a >= MM || c >= SS) continue;
                float acc = 0.0f;
                if (HSEL) acc = Pm[a][c];
                else {
#pragma unroll
                    for (int k = 0; k < kKfMax; ++k) if (k < SS) acc = __fadd_rn(acc, __fmul_rn(mdl.H[a * SS + k], Pm[k][c]));
                }
                HP[a][c] = acc;
            }
#pragma unroll
        for (int a = 0; a < kKfMax; ++a)
#pragma unroll
            for (int b = 0; b < kKfMax; ++b) {
                if (a >= MM || b >= MM) continue;
                float acc = 0.0f;
                if (HSEL) acc = HP[a][b];
                else {
#pragma unroll
                    for (int k = 0; k < kKfMax; ++k) if (k < SS) acc = __fadd_rn(acc, __fmul_rn(HP[a][k], mdl.H[b * SS + k]));
                }
                Sm[a][b] = __fadd_rn(acc, mdl.R[a * MM + b]);
            }
        // S^-1 by Gauss-Jordan (torch: LU with partial pivoting; both backward stable).  In place: once column `col` is
        // eliminated it is a unit vector, and its storage takes column `col` of the inverse - 25 live values instead of the
        // 50 of an augmented [S | I] (the same operations on the same numbers: a structural 1 times inv, a structural 0 in
        // an FMA).  An innovation covariance is diagonally dominant and never asks for a row exchange; if one of its columns
        // does (a larger entry below the pivot), the flag sends this object to the out-of-line pivoting version below.
        bool exchange = false;
#pragma unroll
        for (int col = 0; col < kKfMax; ++col) {
            if (col >= MM) continue;
            const float best = fabsf(Sm[col][col]);
#pragma unroll
            for (int r = 0; r < kKfMax; ++r)
                if (r > col && r < MM && fabsf(Sm[r][col]) > best) exchange = true;
            const float inv = 1.0f / Sm[col][col];
            Sm[col][col] = 1.0f;
#pragma unroll
            for (int c = 0; c < kKfMax; ++c) Sm[col][c] *= inv;
#pragma unroll
            for (int r = 0; r < kKfMax; ++r) {
                if (r == col || r >= MM) continue;
                const float f = Sm[r][col];
                Sm[r][col] = 0.0f;
#pragma unroll
                for (int c = 0; c < kKfMax; ++c) Sm[r][c] = fmaf(-f, Sm[col][c], Sm[r][c]);
            }
        }
        // The covariance was only needed for S so far; holding its 36 floats through the elimination (50 more live values)
        // is what pins the kernel at 128 registers.  It is read again here instead - the 144 bytes are still in L1.
        if (S == 6) load_P6();
        if (exchange) {                 // rare: S again from P, inverted with row exchanges, in local memory
            float Sp[kKfMax][kKfMax], Sq[kKfMax][kKfMax];
#pragma unroll
            for (int a = 0; a < kKfMax; ++a)
#pragma unroll
                for (int b = 0; b < kKfMax; ++b) {
                    if (a >= MM || b >= MM) { Sp[a][b] = 0.0f; continue; }
                    float acc = 0.0f;
                    if (HSEL) acc = Pm[a][b];
                    else {
                        for (int k2 = 0; k2 < SS; ++k2) {
                            float hp = 0.0f;
                            for (int k = 0; k < SS; ++k) hp = __fadd_rn(hp, __fmul_rn(mdl.H[a * SS + k], Pm[k][k2]));
                            acc = __fadd_rn(acc, __fmul_rn(hp, mdl.H[b * SS + k2]));
                        }
                    }
                    Sp[a][b] = __fadd_rn(acc, mdl.R[a * MM + b]);
                }
            invert_with_exchanges(Sp, Sq, MM);
#pragma unroll
            for (int a = 0; a < kKfMax; ++a)
#pragma unroll
                for (int b = 0; b < kKfMax; ++b) Sm[a][b] = Sq[a][b];
        }
        // Row by row (a full K and a full copy of the new P would cost 60 + 36 more live registers - the difference between
        // four and six CTAs per SM for this latency-bound kernel):  PHt_r = P_r H^T;  K_r = PHt_r S^-1;  x_r += K_r y;
        // P_r <- ((I - K H) P)_r.  Every new row is formed from the ORIGINAL matrix (Pm), as the reference's bmm does; the
        // dot products accumulate with FMA in index order (what the reference's CPU GEMM does too: parity is 1e-5, not bits).
#pragma unroll
        for (int r = 0; r < kKfMax; ++r) {
            if (r >= SS) continue;
            float PHt[kKfMax], K[kKfMax], IKH[kKfMax], Pn[kKfMax];
#pragma unroll
            for (int a = 0; a < kKfMax; ++a) {
                if (a >= MM) continue;
                float acc = 0.0f;
                if (HSEL) acc = Pm[r][a];
                else {
#pragma unroll
                    for (int k = 0; k < kKfMax; ++k) if (k < SS) acc = __fadd_rn(acc, __fmul_rn(Pm[r][k], mdl.H[a * SS + k]));
                }
                PHt[a] = acc;
            }
#pragma unroll
            for (int a = 0; a < kKfMax; ++a) {
                if (a >= MM) continue;
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k < kKfMax; ++k) if (k < MM) acc = fmaf(PHt[k], Sm[k][a], acc);
                K[a] = acc;
            }
            {   // x += K y
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k < kKfMax; ++k) if (k < MM) acc = fmaf(K[k], y[k], acc);
                X[i * SS + r] = __fadd_rn(x[r], acc);
            }
#pragma unroll
            for (int c = 0; c < kKfMax; ++c) {
                if (c >= SS) continue;
                float acc = 0.0f;
                if (HSEL) acc = (c < MM) ? K[c < kKfMax ? c : 0] : 0.0f;
                else {
#pragma unroll
                    for (int k = 0; k < kKfMax; ++k) if (k < MM) acc = __fadd_rn(acc, __fmul_rn(K[k], mdl.H[k * SS + c]));
                }
                IKH[c] = __fsub_rn((r == c) ? 1.0f : 0.0f, acc);
            }
#pragma unroll
            for (int c = 0; c < kKfMax; ++c) {
                if (c >= SS) continue;
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k < kKfMax; ++k) if (k < SS) acc = fmaf(IKH[k], Pm[k][c], acc);
                Pn[c] = acc;
            }
            if (S == 6) {      // 24-byte rows of a 16-byte aligned matrix: three 8-byte stores
                float2* prow = reinterpret_cast<float2*>(P + i * 36 + r * 6);
                prow[0] = make_float2(Pn[0], Pn[1]);
                prow[1] = make_float2(Pn[2], Pn[3]);
                prow[2] = make_float2(Pn[4], Pn[5]);
            } else {
#pragma unroll
                for (int c = 0; c < kKfMax; ++c)
                    if (c < SS) P[(i * SS + r) * SS + c] = Pn[c];
            }
        }
    }
}

static int fill_model(KfModel& m, const float* F, const float* Q, const float* H, const float* R, const float* mu_R, int S,
                      int M) {
    for (int i = 0; i < kKfMax * kKfMax; ++i) m.F[i] = m.Q[i] = m.H[i] = m.R[i] = 0.0f;
    for (int i = 0; i < kKfMax; ++i) m.mu_R[i] = 0.0f;
    if (F) for (int i = 0; i < S * S; ++i) m.F[i] = F[i];
    if (Q) for (int i = 0; i < S * S; ++i) m.Q[i] = Q[i];
    if (H) for (int i = 0; i < M * S; ++i) m.H[i] = H[i];
    if (R) for (int i = 0; i < M * M; ++i) m.R[i] = R[i];
    if (mu_R) for (int i = 0; i < M; ++i) m.mu_R[i] = mu_R[i];
    return 0;
}

}  // namespace g3d

using namespace g3d;

extern "C" int g3d_kf_predict(float* X, float* P, const float* D, const double* dt_per_object, double dt_scalar,
                              double dt_default, double* T, int64_t n, int64_t S, const float* F_host,
                              const float* Q_host, int device, void* stream) {
    G3D_REQUIRE(n >= 0 && S >= 1 && S <= kKfMax, "state size must be 1..8");
    if (n == 0) return G3D_OK;
    G3D_REQUIRE(X && P && F_host && Q_host, "null pointer");
    G3D_REQUIRE(S <= 5 || D, "direction vector needed (F[0,5] = D * dt)");
    G3D_REQUIRE(((uintptr_t)P % 16) == 0, "P must be 16-byte aligned");
    G3D_GUARD(device);
    KfModel m;
    fill_model(m, F_host, Q_host, nullptr, nullptr, nullptr, (int)S, 0);
    const int sms = sm_count(device);
    const int grid = (int)(ceil_div(n, 128) < (int64_t)sms * 16 ? ceil_div(n, 128) : (int64_t)sms * 16);
    bool fid = S == 6;                       // F == identity apart from [0][5] (which the kernel overwrites with D * dt)?
    for (int64_t r = 0; r < S && fid; ++r)
        for (int64_t c = 0; c < S; ++c)
            if (!(r == 0 && c == 5) && F_host[r * S + c] != ((r == c) ? 1.0f : 0.0f)) { fid = false; break; }
    if (fid && ((uintptr_t)X % 8) == 0)
        kf_predict_fid6_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(X, P, D, dt_per_object, dt_scalar, dt_default, T, n, m);
    else if (S == 6)
        kf_predict_kernel<6><<<grid, 128, 0, (cudaStream_t)stream>>>(X, P, D, dt_per_object, dt_scalar, dt_default, T, n, 6, m);
    else
        kf_predict_kernel<0><<<grid, 128, 0, (cudaStream_t)stream>>>(X, P, D, dt_per_object, dt_scalar, dt_default, T, n,
                                                                   (int)S, m);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}

extern "C" int g3d_kf_update(float* X, float* P, const int64_t* rows, const double* z, int64_t m_count, int64_t S, int64_t M,
                             const float* H_host, const float* R_host, const float* mu_R_host, int device, void* stream) {
    G3D_REQUIRE(m_count >= 0 && S >= 1 && S <= kKfMax && M >= 1 && M <= kKfMax, "state / measurement size must be 1..8");
    if (m_count == 0) return G3D_OK;
    G3D_REQUIRE(X && P && rows && z && H_host && R_host, "null pointer");
    G3D_REQUIRE(((uintptr_t)P % 16) == 0 && ((uintptr_t)X % 8) == 0, "P must be 16-byte and X 8-byte aligned");
    G3D_GUARD(device);
    KfModel m;
    fill_model(m, nullptr, nullptr, H_host, R_host, mu_R_host, (int)S, (int)M);
    const int sms = sm_count(device);
    // 5 CTAs of 128 threads are resident per SM (launch bounds): exactly one wave, every thread the same number of objects
    const int grid = (int)(ceil_div(m_count, 128) < (int64_t)sms * 5 ? ceil_div(m_count, 128) : (int64_t)sms * 5);
    bool hsel = M <= S;                      // H == [I_M | 0] exactly?
    for (int64_t a = 0; a < M && hsel; ++a)
        for (int64_t k = 0; k < S; ++k)
            if (H_host[a * S + k] != ((a == k) ? 1.0f : 0.0f)) { hsel = false; break; }
    cudaStream_t st = (cudaStream_t)stream;
    if (S == 6 && M == 5 && hsel) kf_update_kernel<6, 5, true><<<grid, 128, 0, st>>>(X, P, rows, z, m_count, 6, 5, m);
    else if (S == 6 && M == 5)    kf_update_kernel<6, 5, false><<<grid, 128, 0, st>>>(X, P, rows, z, m_count, 6, 5, m);
    else if (hsel)                kf_update_kernel<0, 0, true><<<grid, 128, 0, st>>>(X, P, rows, z, m_count, (int)S, (int)M, m);
    else                          kf_update_kernel<0, 0, false><<<grid, 128, 0, st>>>(X, P, rows, z, m_count, (int)S, (int)M, m);
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}
