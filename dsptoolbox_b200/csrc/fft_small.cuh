// Small in-register DFTs (natural order in, natural order out) shared by the shared-memory Stockham FFT of the
// Welch stage (welch_fft.cu) and the global-memory large FFT (big_fft.cu).
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------- small DFTs, natural order
template <typename C> __device__ __forceinline__ void dft2(C &a, C &b) {
    C t = a; a = cadd(t, b); b = csub(t, b);
}
template <typename C> __device__ __forceinline__ void dft4(C &v0, C &v1, C &v2, C &v3) {
    C t0 = cadd(v0, v2), t1 = csub(v0, v2), t2 = cadd(v1, v3), t3 = cmul_mi(csub(v1, v3));
    v0 = cadd(t0, t2); v2 = csub(t0, t2); v1 = cadd(t1, t3); v3 = csub(t1, t3);
}
template <typename T> __device__ __forceinline__ void dft8(cx<T> *v) {
    using C = cx<T>;
    C e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    C o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    const T h = T(0.70710678118654752440);
    // o1 *= (1 - i)/sqrt2 ; o2 *= -i ; o3 *= (-1 - i)/sqrt2
    C w1 = mk<T>((o1.x + o1.y) * h, (o1.y - o1.x) * h);
    C w2 = cmul_mi(o2);
    C w3 = mk<T>((o3.y - o3.x) * h, -(o3.x + o3.y) * h);
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, w1); v[5] = csub(e1, w1);
    v[2] = cadd(e2, w2); v[6] = csub(e2, w2);
    v[3] = cadd(e3, w3); v[7] = csub(e3, w3);
}
// 16-point DFT as 4 x 4: column DFT4s, twiddles W16^(n2*k1), row DFT4s; natural-order output.
template <typename T> __device__ __forceinline__ void dft16(cx<T> *v) {
    using C = cx<T>;
    const T c1 = T(0.92387953251128675613), s1 = T(0.38268343236508977173);   // cos, sin(pi/8)
    const T h = T(0.70710678118654752440);
    C y[4][4];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
        y[n2][0] = v[n2]; y[n2][1] = v[4 + n2]; y[n2][2] = v[8 + n2]; y[n2][3] = v[12 + n2];
        dft4(y[n2][0], y[n2][1], y[n2][2], y[n2][3]);
    }
    // y[n2][k1] *= W16^(n2*k1), W16 = exp(-i pi/8)
    auto mulw = [&](C a, T wr, T wi) { return mk<T>(a.x * wr - a.y * wi, a.x * wi + a.y * wr); };
    y[1][1] = mulw(y[1][1], c1, -s1);  y[1][2] = mulw(y[1][2], h, -h);    y[1][3] = mulw(y[1][3], s1, -c1);
    y[2][1] = mulw(y[2][1], h, -h);    y[2][2] = cmul_mi(y[2][2]);        y[2][3] = mulw(y[2][3], -h, -h);
    y[3][1] = mulw(y[3][1], s1, -c1);  y[3][2] = mulw(y[3][2], -h, -h);   y[3][3] = mulw(y[3][3], -c1, s1);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        C a = y[0][k1], b = y[1][k1], c = y[2][k1], d = y[3][k1];
        dft4(a, b, c, d);
        v[k1] = a; v[k1 + 4] = b; v[k1 + 8] = c; v[k1 + 12] = d;
    }
}
template <typename T, int R> __device__ __forceinline__ void dftR(cx<T> *v) {
    if constexpr (R == 16) dft16<T>(v);
    else if constexpr (R == 8) dft8<T>(v);
    else if constexpr (R == 4) dft4(v[0], v[1], v[2], v[3]);
    else dft2(v[0], v[1]);
}

