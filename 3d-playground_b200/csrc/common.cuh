// common.cuh — shared helpers for the libgeom3d kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/geom3d.h"

namespace g3d {

void set_error(const char* fmt, ...);
int sm_count(int device);     // SMs of `device`, queried once per device (148 on B200); grids are sized in multiples of it

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != device && cudaSetDevice(device) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

#define G3D_REQUIRE(cond, msg)                                             \
    do {                                                                   \
        if (!(cond)) {                                                     \
            g3d::set_error("%s: %s", __func__, msg);                       \
            return G3D_ERR_INVALID;                                        \
        }                                                                  \
    } while (0)

#define G3D_CUDA(call)                                                     \
    do {                                                                   \
        cudaError_t e__ = (call);                                          \
        if (e__ != cudaSuccess) {                                          \
            g3d::set_error("%s: %s -> %s", __func__, #call, cudaGetErrorString(e__)); \
            return G3D_ERR_CUDA;                                           \
        }                                                                  \
    } while (0)

#define G3D_GUARD(device)                                                  \
    g3d::DeviceGuard guard__(device);                                      \
    if (!guard__.ok) {                                                     \
        g3d::set_error("%s: cannot select CUDA device %d", __func__, device); \
        return G3D_ERR_CUDA;                                               \
    }

#define G3D_LAUNCH_CHECK()                                                 \
    do {                                                                   \
        cudaError_t e__ = cudaGetLastError();                              \
        if (e__ != cudaSuccess) {                                          \
            g3d::set_error("%s: kernel launch -> %s", __func__, cudaGetErrorString(e__)); \
            return G3D_ERR_CUDA;                                           \
        }                                                                  \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t align_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---------------------------------------------------------------------------------------------------------------
// Exactly-rounded FP32 building blocks.  PyTorch evaluates every arithmetic op of the reference as its own rounded
// FP32 operation; nvcc would contract a*b+c into an FMA and change the last bit, so wherever a result feeds an index
// decision (argmax, >= 0.5, NMS suppression) the op order is spelled out with the _rn intrinsics (never contracted).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float box_area_rn(float x1, float y1, float x2, float y2) {
    return __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
}

// calc_iou of the reference (losses.py:5-22): intersection / clamp(area_a + area_b - intersection, 1e-8)
__device__ __forceinline__ float iou_retinanet(const float4& a, float area_a, const float4& b, float area_b) {
    float iw = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
    float ih = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
    iw = fmaxf(iw, 0.0f);
    ih = fmaxf(ih, 0.0f);
    const float inter = __fmul_rn(iw, ih);
    float ua = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    ua = fmaxf(ua, 1e-8f);
    return __fdiv_rn(inter, ua);
}

// torchvision nms IoU: inter / (area_i + area_j - inter), no clamp of the union
__device__ __forceinline__ float iou_torchvision(const float4& a, float area_a, const float4& b, float area_b) {
    const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
    const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
    const float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}

// ---------------------------------------------------------------------------------------------------------------
// warp / block reductions
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming (read-once) 128-bit load that does not pollute L1
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

}  // namespace g3d
