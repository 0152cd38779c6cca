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

// ------------------------------------------------------------------ per-device one-time setup
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count are properties of a DEVICE, so a
// process that drives several GPUs (or switches devices between calls) needs them once per device and
// kernel instantiation.  A function-local `static DeviceOnce once;` + b2_device_once() does that; the
// races between host threads are benign (the setup is idempotent).
constexpr int B2_MAX_DEVICES = 64;
struct DeviceOnce {
    bool done[B2_MAX_DEVICES];
    int sms[B2_MAX_DEVICES];
};
template <class Kern> int b2_device_once(DeviceOnce &once, Kern kern, int dynamic_smem, int *num_sms) {
    int dev = 0;
    B2_CUDA_OK(cudaGetDevice(&dev));
    const bool cached = dev >= 0 && dev < B2_MAX_DEVICES;
    if (!cached || !once.done[dev]) {
        if (dynamic_smem > 0)
            B2_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dynamic_smem));
        int sms = 0;
        B2_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (cached) { once.sms[dev] = sms; once.done[dev] = true; }
        if (num_sms) *num_sms = sms;
        return B2_OK;
    }
    if (num_sms) *num_sms = once.sms[dev];
    return B2_OK;
}

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

// ------------------------------------------------------------------ spectra layout (K1 writes, K2 / auto-power read)
// [frame chunk of 16][slot group][channel][8 frame pairs][4 slots][2 frames], one complex value each.
// A "slot" is a permuted bin index: the real-FFT untangling produces bins p and NC - p together (NC = wl / 2,
// p = 0 .. NC/2 - 1), so bin p takes slot p and bin NC - p slot NC/2 + p (the upper half of the spectrum in
// DESCENDING order, starting with the Nyquist bin); the self-paired bin NC/2 follows at slot NC.  Four
// neighbouring slots x one frame pair are one 128-byte line (fp64): the four lanes that untangle four neighbouring p
// of a frame pair fill whole lines with their 32-byte stores, for both halves of the spectrum.
__host__ __device__ __forceinline__ int b2_bin_slot(int f, int NC) {
    if (2 * f < NC) return f;
    if (2 * f == NC) return NC;
    return NC / 2 + (NC - f);
}
__host__ __device__ __forceinline__ int b2_slot_groups(int NC) { return (NC + 1 + 3) / 4; }
constexpr int B2_GROUP_ELEMS = 64;                                   // 8 frame pairs x 4 slots x 2 frames
// element offset of (slot s, frame k of the chunk) inside a (chunk, group, channel) block
__host__ __device__ __forceinline__ int b2_slot_frame_offset(int s, int k) {
    return (k >> 1) * 8 + (s & 3) * 2 + (k & 1);
}

static inline int ilog2_exact(int64_t v) {
    int l = 0;
    while ((int64_t(1) << l) < v) ++l;
    return ((int64_t(1) << l) == v) ? l : -1;
}
