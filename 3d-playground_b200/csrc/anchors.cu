// anchors.cu — device-side anchor table generation (reference retinanet/anchors.py:21-40,109-129).
//
// The reference rebuilds the [A,4] table in numpy on every forward and copies it host->device (6.2 MB at 1080p).  Here
// one launch writes it straight into HBM: anchor i -> (level, cell row, cell column, shape) by integer arithmetic, value =
// float32( base_shape[level][shape][c] + (cell + 0.5) * stride ) with the sum formed in FP64 exactly like numpy's
// float64 `anchors + shifts` followed by `.astype(np.float32)`.  The per-level base shapes (9 x 4 doubles at the default
// ratios/scales: they need sqrt and 2**(1/3), formed on the host by the same numpy expressions as the reference) travel
// in the kernel parameter block, so the launch reads nothing from memory and writes 16 B per anchor.
#include "common.cuh"

namespace g3d {

constexpr int kMaxLevels = 8;
constexpr int kMaxLevelShapes = 96;   // sum over levels of shapes per level (3 KB of kernel parameters)

struct AnchorTable {
    double shape[kMaxLevelShapes][4];  // [level * S + s] = (x1, y1, x2, y2) around the origin
    double stride[kMaxLevels];
    long long first[kMaxLevels + 1];   // first anchor index of each level; first[L] = A
    int cols[kMaxLevels];
    int L, S;
};

__global__ void __launch_bounds__(256) generate_anchors_kernel(const __grid_constant__ AnchorTable tab,
                                                               float4* __restrict__ out) {
    const long long A = tab.first[tab.L];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < A; i += (long long)gridDim.x * blockDim.x) {
        int lvl = 0;
#pragma unroll
        for (int l = 1; l < kMaxLevels; ++l)
            if (l < tab.L && i >= tab.first[l]) lvl = l;
        const long long local = i - tab.first[lvl];
        const long long cell = local / tab.S;
        const int s = (int)(local - cell * tab.S);
        const long long row = cell / tab.cols[lvl];
        const int col = (int)(cell - row * tab.cols[lvl]);
        // (np.arange(n) + 0.5) * stride, anchors.py:110-111 (FP64; exact for every sane stride)
        const double sx = __dmul_rn((double)col + 0.5, tab.stride[lvl]);
        const double sy = __dmul_rn((double)row + 0.5, tab.stride[lvl]);
        const double* b = tab.shape[lvl * tab.S + s];
        float4 v;
        v.x = (float)__dadd_rn(b[0], sx);
        v.y = (float)__dadd_rn(b[1], sy);
        v.z = (float)__dadd_rn(b[2], sx);
        v.w = (float)__dadd_rn(b[3], sy);
        out[i] = v;
    }
}

}  // namespace g3d

extern "C" int g3d_generate_anchors(const double* shapes_host, const double* strides_host, const int64_t* rows_host,
                                    const int64_t* cols_host, int64_t L, int64_t S, float* anchors, int64_t A,
                                    int device, void* stream) {
    using namespace g3d;
    G3D_REQUIRE(shapes_host && strides_host && rows_host && cols_host, "null host table");
    G3D_REQUIRE(L >= 1 && L <= kMaxLevels, "1 <= levels <= 8");
    G3D_REQUIRE(S >= 1 && L * S <= kMaxLevelShapes, "levels * shapes per level must be <= 96");
    AnchorTable tab;
    tab.L = (int)L;
    tab.S = (int)S;
    long long first = 0;
    for (int l = 0; l < kMaxLevels; ++l) {
        tab.first[l] = first;
        tab.stride[l] = 0.0;
        tab.cols[l] = 1;
        if (l >= L) continue;
        G3D_REQUIRE(rows_host[l] >= 0 && cols_host[l] >= 0 && cols_host[l] < (1ll << 31), "bad level grid");
        tab.stride[l] = strides_host[l];
        tab.cols[l] = cols_host[l] > 0 ? (int)cols_host[l] : 1;
        first += rows_host[l] * cols_host[l] * S;
    }
    tab.first[kMaxLevels] = first;
    tab.first[L] = first;
    for (int64_t k = 0; k < L * S; ++k)
        for (int c = 0; c < 4; ++c) tab.shape[k][c] = shapes_host[k * 4 + c];
    G3D_REQUIRE(first == A, "A must equal sum over levels of rows * cols * S");
    if (A == 0) return G3D_OK;
    G3D_REQUIRE(anchors, "null output");
    G3D_GUARD(device);
    const int64_t blocks = ceil_div(A, 256);
    generate_anchors_kernel<<<(unsigned)(blocks < sm_count(device) * 8 ? blocks : sm_count(device) * 8), 256, 0, (cudaStream_t)stream>>>(
        tab, reinterpret_cast<float4*>(anchors));
    G3D_LAUNCH_CHECK();
    return G3D_OK;
}
