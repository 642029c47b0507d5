// Shared helpers for the sm_100a kernels of libdeepfm_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/deepfm_b200.h"

namespace dfm {

void set_error(const char* fmt, ...);

#define DFM_CHECK_CUDA(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            dfm::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                  \
                           cudaGetErrorString(_e));                                       \
            return DFM_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

#define DFM_CHECK_LAUNCH() DFM_CHECK_CUDA(cudaGetLastError())

#define DFM_REQUIRE(cond, code, ...)                                                      \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            dfm::set_error(__VA_ARGS__);                                                  \
            return (code);                                                                \
        }                                                                                 \
    } while (0)

int sm_count();

static inline int next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- V-wide (1 or 4 floats) register vectors: 128-bit loads/stores when V == 4 ----------
template <int V>
struct VecF {
    float v[V];
};

template <int V>
__device__ __forceinline__ VecF<V> vzero() {
    VecF<V> r;
#pragma unroll
    for (int i = 0; i < V; ++i) r.v[i] = 0.f;
    return r;
}

template <int V>
__device__ __forceinline__ VecF<V> vload(const float* p);  // read-only path, keeps L1/L2
template <>
__device__ __forceinline__ VecF<4> vload<4>(const float* p) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    VecF<4> r;
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
}
template <>
__device__ __forceinline__ VecF<1> vload<1>(const float* p) {
    VecF<1> r;
    r.v[0] = __ldg(p);
    return r;
}

template <>
__device__ __forceinline__ VecF<2> vload<2>(const float* p) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    VecF<2> r;
    r.v[0] = t.x; r.v[1] = t.y;
    return r;
}

template <int V>
__device__ __forceinline__ VecF<V> vload_stream(const float* p);  // touched once: evict-first
template <>
__device__ __forceinline__ VecF<4> vload_stream<4>(const float* p) {
    float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    VecF<4> r;
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
}
template <>
__device__ __forceinline__ VecF<1> vload_stream<1>(const float* p) {
    VecF<1> r;
    r.v[0] = __ldcs(p);
    return r;
}

template <int V>
__device__ __forceinline__ void vstore(float* p, const VecF<V>& x);
template <>
__device__ __forceinline__ void vstore<4>(float* p, const VecF<4>& x) {
    *reinterpret_cast<float4*>(p) = make_float4(x.v[0], x.v[1], x.v[2], x.v[3]);
}
template <>
__device__ __forceinline__ void vstore<1>(float* p, const VecF<1>& x) {
    *p = x.v[0];
}
template <>
__device__ __forceinline__ void vstore<2>(float* p, const VecF<2>& x) {
    *reinterpret_cast<float2*>(p) = make_float2(x.v[0], x.v[1]);
}

template <int V>
__device__ __forceinline__ void vstore_stream(float* p, const VecF<V>& x);  // st.global.cs
template <>
__device__ __forceinline__ void vstore_stream<4>(float* p, const VecF<4>& x) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(x.v[0], x.v[1], x.v[2], x.v[3]));
}
template <>
__device__ __forceinline__ void vstore_stream<1>(float* p, const VecF<1>& x) {
    __stcs(p, x.v[0]);
}

// Sum over the G (power of two, <= 32) consecutive lanes of a lane group.
__device__ __forceinline__ float group_sum(float x, int G, unsigned mask) {
    for (int o = G >> 1; o > 0; o >>= 1) x += __shfl_xor_sync(mask, x, o);
    return x;
}

__device__ __forceinline__ unsigned group_mask(int G) {
    if (G >= 32) return 0xffffffffu;
    unsigned lane = threadIdx.x & 31u;
    return ((1u << G) - 1u) << (lane & ~(unsigned)(G - 1));
}

}  // namespace dfm
