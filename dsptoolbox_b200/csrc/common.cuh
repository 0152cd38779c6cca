// Shared helpers for the b2dsp kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b2dsp.h"

// ------------------------------------------------------------------ error plumbing
void b2_set_error(const char *fmt, ...);

#define B2_REQUIRE(cond, code, ...)      \
    do {                                 \
        if (!(cond)) {                   \
            b2_set_error(__VA_ARGS__);   \
            return (code);               \
        }                                \
    } while (0)

#define B2_CUDA_OK(expr)                                                              \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            b2_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                         __FILE__, __LINE__);                                         \
            return B2_ECUDA;                                                          \
        }                                                                             \
    } while (0)

#define B2_LAUNCH_OK(name)                                                            \
    do {                                                                              \
        cudaError_t _e = cudaGetLastError();                                          \
        if (_e != cudaSuccess) {                                                      \
            b2_set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));    \
            return B2_ECUDA;                                                          \
        }                                                                             \
    } while (0)

// ------------------------------------------------------------------ complex helpers
template <typename T> struct cx_of;
template <> struct cx_of<float> { using type = float2; };
template <> struct cx_of<double> { using type = double2; };
template <typename T> using cx = typename cx_of<T>::type;

template <typename T> __host__ __device__ __forceinline__ cx<T> mk(T re, T im) {
    cx<T> r; r.x = re; r.y = im; return r;
}
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
    C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
// a * (-i)
template <typename C> __device__ __forceinline__ C cmul_mi(C a) { C r; r.x = a.y; r.y = -a.x; return r; }
template <typename C> __device__ __forceinline__ C cconj(C a) { C r; r.x = a.x; r.y = -a.y; return r; }

static inline int ilog2_exact(int64_t v) {
    int l = 0;
    while ((int64_t(1) << l) < v) ++l;
    return ((int64_t(1) << l) == v) ? l : -1;
}
