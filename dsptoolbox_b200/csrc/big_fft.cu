// Large and arbitrary-length real FFT along the time axis, batched over channels (float64).
//
// Replaces the whole-recording `np.fft.rfft(time_data, axis=0, n=fft_length)` of the reference's single-FFT
// spectrum (classes/signal.py:903-911: fft_length = next_fast_len(L) or L, i.e. NOT a power of two) that feeds
// `_csm_fft` (standard/_spectral_methods.py:374-443), and FFT lengths beyond the shared-memory kernel of the Welch
// stage.  Two layers:
//   * power-of-two lengths P: out-of-place Stockham autosort passes (radix 16/8/4/2) through global memory, one
//     butterfly per thread, coalesced reads and writes, twiddles from a two-level table (4096 fine x P/4096 coarse
//     entries, one complex product) so the table stays L2-resident for any P;
//   * any other length n: Bluestein's chirp-z transform on top of it,
//       X[k] = w[k] * sum_t (x[t] w[t]) conj(w[k - t]),   w[t] = exp(-i pi t^2 / n),
//     a circular convolution of length P >= 2n - 1 done with two forward FFTs (the inverse one as conj o FFT o conj);
//     the chirp phase t^2 mod 2n is reduced in 64-bit integers, so it is exact for any n < 2^31.
// HBM-bound: every pass reads and writes the batch once (16 B per complex value each way).
#include "common.cuh"
#include "fft_small.cuh"

namespace {

constexpr int LO_BITS = 12, LO = 1 << LO_BITS;

struct PlanLayout {
    int64_t n, P;
    bool bluestein;
    int64_t off_lo, off_hi, off_w, off_fb, total;      // in complex128 elements
};

inline int64_t next_pow2(int64_t v) { int64_t p = 1; while (p < v) p <<= 1; return p; }

inline PlanLayout plan_layout(int64_t n) {
    PlanLayout L;
    L.n = n;
    L.bluestein = (n & (n - 1)) != 0;
    L.P = L.bluestein ? next_pow2(2 * n - 1) : n;
    L.off_lo = 0;
    L.off_hi = LO;
    const int64_t nhi = L.P > LO ? L.P / LO : 1;
    L.off_w = L.off_hi + nhi;
    L.off_fb = L.off_w + (L.bluestein ? n : 0);
    L.total = L.off_fb + (L.bluestein ? L.P : 0);
    return L;
}

// exp(-2 pi i t / P) from the two-level table
__device__ __forceinline__ double2 twiddle(const double2 *__restrict__ lo, const double2 *__restrict__ hi, int64_t t) {
    const double2 a = __ldg(lo + (t & (LO - 1)));
    const double2 b = __ldg(hi + (t >> LO_BITS));
    return cmul(a, b);
}

__global__ void twiddle_tables_kernel(double2 *lo, double2 *hi, int64_t P) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t nhi = P > LO ? P / LO : 1;
    if (i < LO) {
        double s, c;
        sincospi(-2.0 * double(i % P) / double(P), &s, &c);       // exact argument reduction inside sincospi
        lo[i] = make_double2(c, s);
    }
    if (i < nhi) {
        double s, c;
        sincospi(-2.0 * double((i * LO) % P) / double(P), &s, &c);
        hi[i] = make_double2(c, s);
    }
}

// One Stockham pass over a batch of rows of P complex values: radix R, NS = product of the earlier radices.
//   y[(j - k) R + k + r NS] = DFT_R_r( x[j + r P/R] w^(r k) ),  k = j mod NS, w = exp(-2 pi i / (NS R))
template <int R>
__global__ void __launch_bounds__(256)
stockham_global_kernel(const double2 *__restrict__ x, double2 *__restrict__ y, int64_t P, int64_t NS,
                       const double2 *__restrict__ lo, const double2 *__restrict__ hi) {
    const int64_t J = P / R;
    const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= J) return;
    const double2 *xb = x + int64_t(blockIdx.y) * P;
    double2 *yb = y + int64_t(blockIdx.y) * P;
    double2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = xb[j + r * J];
    const int64_t k = j & (NS - 1);
    if (NS > 1) {
        // twiddles w^r: one table product, powers by products (depth <= 4: a few ulp)
        const double2 w1 = twiddle(lo, hi, k * (P / (NS * R)));
        if (R == 2) {
            v[1] = cmul(v[1], w1);
        } else {
            const double2 w2 = cmul(w1, w1), w3 = cmul(w1, w2);
            v[1] = cmul(v[1], w1); v[2] = cmul(v[2], w2); v[3] = cmul(v[3], w3);
            if (R >= 8) {
                const double2 w4 = cmul(w2, w2);
                v[4] = cmul(v[4], w4); v[5] = cmul(v[5], cmul(w4, w1));
                v[6] = cmul(v[6], cmul(w4, w2)); v[7] = cmul(v[7], cmul(w4, w3));
                if (R == 16) {
                    const double2 w8 = cmul(w4, w4), w12 = cmul(w8, w4);
                    v[8] = cmul(v[8], w8); v[9] = cmul(v[9], cmul(w8, w1));
                    v[10] = cmul(v[10], cmul(w8, w2)); v[11] = cmul(v[11], cmul(w8, w3));
                    v[12] = cmul(v[12], w12); v[13] = cmul(v[13], cmul(w12, w1));
                    v[14] = cmul(v[14], cmul(w12, w2)); v[15] = cmul(v[15], cmul(w12, w3));
                }
            }
        }
    }
    dftR<double, R>(v);
    const int64_t j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) yb[j0 + r * NS] = v[r];
}

// in-place-looking forward FFT of `batch` rows of length P (power of two): the result ends in *res (one of a, b)
int fft_pow2(double2 *a, double2 *b, int64_t P, int batch, const double2 *lo, const double2 *hi, cudaStream_t st,
             double2 **res) {
    double2 *src = a, *dst = b;
    int64_t NS = 1;
    while (NS < P) {
        const int64_t rem = P / NS;
        const int R = rem >= 16 ? 16 : int(rem);
        const int64_t J = P / R;
        dim3 grid((unsigned)((J + 255) / 256), (unsigned)batch);
        switch (R) {
        case 16: stockham_global_kernel<16><<<grid, 256, 0, st>>>(src, dst, P, NS, lo, hi); break;
        case 8: stockham_global_kernel<8><<<grid, 256, 0, st>>>(src, dst, P, NS, lo, hi); break;
        case 4: stockham_global_kernel<4><<<grid, 256, 0, st>>>(src, dst, P, NS, lo, hi); break;
        default: stockham_global_kernel<2><<<grid, 256, 0, st>>>(src, dst, P, NS, lo, hi); break;
        }
        B2_LAUNCH_OK("b2_rfft_any(pass)");
        NS *= R;
        double2 *t = src; src = dst; dst = t;
    }
    *res = src;
    return B2_OK;
}

// chirp w[t] = exp(-i pi t^2 / n), t < n, with t^2 reduced mod 2n in integers
__global__ void chirp_kernel(double2 *w, int64_t n) {
    const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const unsigned long long q = (unsigned long long)(t) * (unsigned long long)(t) % (unsigned long long)(2 * n);
    double s, c;
    sincospi(double(q) / double(n), &s, &c);
    w[t] = make_double2(c, -s);
}
// filter b[t] = conj(w[|t|]) laid out circularly over P, zero elsewhere
__global__ void chirp_filter_kernel(const double2 *__restrict__ w, double2 *b, int64_t n, int64_t P) {
    const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= P) return;
    double2 v = make_double2(0.0, 0.0);
    if (t < n) v = cconj(w[t]);
    else if (P - t < n) v = cconj(w[P - t]);
    b[t] = v;
}

// a[b][t] = x[b][t] (* w[t]) for t < n_use, zero up to P.  Threads run over (t, b) with b fastest, which is the
// contiguous direction of a (time, channel) recording.
// Row b of the batch starts at (b / rpg) * stride_group + (b % rpg) * stride_b: the whole recording (one group,
// rows = channels) as well as Welch/STFT frames (group = channel, rows = frames, stride_b = hop) fit this form.
// Optional per-sample window and per-row mean (of the windowed row, removed from the first n_in samples: the
// reference windows first and de-means second, _spectral_methods.py:257-261).
template <typename IN>
__global__ void load_rows_kernel(const IN *__restrict__ x, int64_t stride_t, int64_t stride_b, int64_t stride_group,
                                 int rpg, int64_t n_in, int64_t n_use, int batch, const double *__restrict__ window,
                                 const double *__restrict__ means, const double2 *__restrict__ w,
                                 double2 *__restrict__ a, int64_t P) {
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= P * batch) return;
    const int64_t t = idx / batch;
    const int b = int(idx - t * batch);
    double2 v = make_double2(0.0, 0.0);
    if (t < n_use) {
        double s = double(x[int64_t(b / rpg) * stride_group + int64_t(b % rpg) * stride_b + t * stride_t]);
        if (window) s *= window[t];
        if (means && t < n_in) s -= means[b];
        v = w ? make_double2(s * w[t].x, s * w[t].y) : make_double2(s, 0.0);
    }
    a[int64_t(b) * P + t] = v;
}

// mean over the n_in samples of every (windowed) row: one warp per row
template <typename IN>
__global__ void row_means_kernel(const IN *__restrict__ x, int64_t stride_t, int64_t stride_b, int64_t stride_group,
                                 int rpg, int64_t n_in, int batch, const double *__restrict__ window,
                                 double *__restrict__ means) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= batch) return;
    const IN *row = x + int64_t(b / rpg) * stride_group + int64_t(b % rpg) * stride_b;
    double acc = 0.0;
    for (int64_t t = lane; t < n_in; t += 32) acc += double(row[t * stride_t]) * (window ? window[t] : 1.0);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) means[b] = acc / double(n_in);
}

// c = conj(fa * fb): the inverse transform is then conj(FFT(c)) / P
__global__ void pointwise_conj_kernel(const double2 *__restrict__ fa, const double2 *__restrict__ fb, double2 *c,
                                      int64_t P, int batch) {
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= P * batch) return;
    const int64_t t = idx % P;
    c[idx] = cconj(cmul(fa[idx], fb[t]));
}

// out[k][b] = scale * (bluestein ? w[k] conj(y[b][k]) / P : y[b][k]) for k < nbins; threads over (k, b), b fastest
// The DC bin and, for an even length, the Nyquist bin of a real signal are real: their imaginary part is written as an
// exact +0 like a real-input FFT does (the complex transform leaves a rounding residue there, and the element-wise
// square root of `_csm_fft` sits on its branch cut for exactly these bins).
__global__ void store_bins_kernel(const double2 *__restrict__ y, const double2 *__restrict__ w, int64_t P, int64_t n, int64_t nbins,
                                  int batch, double scale, double2 *__restrict__ out, int64_t stride_k, int64_t stride_b,
                                  int64_t stride_group, int rpg) {
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= nbins * batch) return;
    const int64_t k = idx / batch;
    const int b = int(idx - k * batch);
    double2 v = y[int64_t(b) * P + k];
    if (w) v = cmul(w[k], make_double2(v.x / double(P), -v.y / double(P)));
    if (k == 0 || 2 * k == n) v.y = 0.0;
    out[k * stride_k + int64_t(b / rpg) * stride_group + int64_t(b % rpg) * stride_b] = make_double2(v.x * scale, v.y * scale);
}

}  // namespace

extern "C" int64_t b2_fft_plan_bytes(int64_t n) {
    if (n < 2 || n >= (int64_t(1) << 30)) { b2_set_error("b2_fft_plan_bytes: length %lld outside [2, 2^30)", (long long)n); return B2_EINVAL; }
    return plan_layout(n).total * 16;
}

extern "C" int64_t b2_fft_work_bytes(int64_t n, int batch) {
    if (n < 2 || n >= (int64_t(1) << 30) || batch < 1) { b2_set_error("b2_fft_work_bytes: bad n / batch"); return B2_EINVAL; }
    return 2 * plan_layout(n).P * int64_t(batch) * 16 + ((int64_t(batch) * 8 + 255) & ~int64_t(255));   // + per-row means
}

extern "C" int b2_fft_plan(int64_t n, void *plan_dev, void *work_dev, void *stream) {
    B2_REQUIRE(plan_dev && work_dev, B2_EINVAL, "b2_fft_plan: null pointer");
    B2_REQUIRE(n >= 2 && n < (int64_t(1) << 30), B2_EINVAL, "b2_fft_plan: length %lld outside [2, 2^30)", (long long)n);
    cudaStream_t st = (cudaStream_t)stream;
    const PlanLayout L = plan_layout(n);
    double2 *plan = (double2 *)plan_dev, *lo = plan + L.off_lo, *hi = plan + L.off_hi;
    const int64_t nhi = L.P > LO ? L.P / LO : 1;
    const int64_t nt = nhi > LO ? nhi : LO;
    twiddle_tables_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(lo, hi, L.P);
    B2_LAUNCH_OK("b2_fft_plan(twiddles)");
    if (L.bluestein) {
        double2 *w = plan + L.off_w, *fb = plan + L.off_fb, *wa = (double2 *)work_dev, *wb = wa + L.P;
        chirp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, n);
        chirp_filter_kernel<<<(unsigned)((L.P + 255) / 256), 256, 0, st>>>(w, wa, n, L.P);
        B2_LAUNCH_OK("b2_fft_plan(chirp)");
        double2 *res = nullptr;
        if (int rc = fft_pow2(wa, wb, L.P, 1, lo, hi, st, &res)) return rc;
        B2_CUDA_OK(cudaMemcpyAsync(fb, res, size_t(L.P) * 16, cudaMemcpyDeviceToDevice, st));
    }
    return B2_OK;
}

extern "C" int b2_rfft_any(const void *x_dev, int in_dtype, int64_t stride_t, int64_t stride_b, int64_t stride_group,
                           int rows_per_group, int64_t n_in, int64_t n, int batch, const void *window_dev, int detrend,
                           const void *plan_dev, void *work_dev, double scale, void *out_dev, int64_t out_stride_k,
                           int64_t out_stride_b, int64_t out_stride_group, void *stream) {
    B2_REQUIRE(x_dev && plan_dev && work_dev && out_dev, B2_EINVAL, "b2_rfft_any: null pointer");
    B2_REQUIRE(n >= 2 && n < (int64_t(1) << 30) && n_in >= 1 && batch >= 1 && rows_per_group >= 1, B2_EINVAL,
               "b2_rfft_any: bad sizes n=%lld n_in=%lld batch=%d", (long long)n, (long long)n_in, batch);
    B2_REQUIRE(in_dtype == B2_F64 || in_dtype == B2_F32, B2_EINVAL, "b2_rfft_any: bad input dtype %d", in_dtype);
    cudaStream_t st = (cudaStream_t)stream;
    const PlanLayout L = plan_layout(n);
    const double2 *plan = (const double2 *)plan_dev, *lo = plan + L.off_lo, *hi = plan + L.off_hi;
    const double2 *w = L.bluestein ? plan + L.off_w : nullptr, *fb = L.bluestein ? plan + L.off_fb : nullptr;
    double2 *wa = (double2 *)work_dev, *wb = wa + L.P * int64_t(batch);
    const int64_t n_use = n_in < n ? n_in : n;               // rfft(x, n): crop or zero-pad (numpy semantics)
    const int64_t tot = L.P * int64_t(batch);
    B2_REQUIRE((tot + 255) / 256 <= 0x7fffffff, B2_EINVAL, "b2_rfft_any: batch too large for one call");
    const unsigned blocks = (unsigned)((tot + 255) / 256);
    const double *win = (const double *)window_dev;
    double *means = nullptr;
    if (detrend) {
        // the per-row means live behind the two work buffers (b2_fft_work_bytes reserves them)
        means = reinterpret_cast<double *>(wb + L.P * int64_t(batch));
        const unsigned mb = (unsigned)((batch + 7) / 8);
        if (in_dtype == B2_F64) row_means_kernel<double><<<mb, 256, 0, st>>>((const double *)x_dev, stride_t, stride_b, stride_group, rows_per_group, n_in, batch, win, means);
        else row_means_kernel<float><<<mb, 256, 0, st>>>((const float *)x_dev, stride_t, stride_b, stride_group, rows_per_group, n_in, batch, win, means);
        B2_LAUNCH_OK("b2_rfft_any(means)");
    }
    if (in_dtype == B2_F64) load_rows_kernel<double><<<blocks, 256, 0, st>>>((const double *)x_dev, stride_t, stride_b, stride_group, rows_per_group, n_in, n_use, batch, win, means, w, wa, L.P);
    else load_rows_kernel<float><<<blocks, 256, 0, st>>>((const float *)x_dev, stride_t, stride_b, stride_group, rows_per_group, n_in, n_use, batch, win, means, w, wa, L.P);
    B2_LAUNCH_OK("b2_rfft_any(load)");
    double2 *res = nullptr;
    if (int rc = fft_pow2(wa, wb, L.P, batch, lo, hi, st, &res)) return rc;
    if (L.bluestein) {
        double2 *other = res == wa ? wb : wa;
        pointwise_conj_kernel<<<blocks, 256, 0, st>>>(res, fb, other, L.P, batch);
        B2_LAUNCH_OK("b2_rfft_any(pointwise)");
        double2 *res2 = nullptr;
        if (int rc = fft_pow2(other, res, L.P, batch, lo, hi, st, &res2)) return rc;
        res = res2;
    }
    const int64_t nbins = n / 2 + 1;
    const int64_t tb = nbins * int64_t(batch);
    store_bins_kernel<<<(unsigned)((tb + 255) / 256), 256, 0, st>>>(res, w, L.P, n, nbins, batch, scale, (double2 *)out_dev,
                                                                    out_stride_k, out_stride_b, out_stride_group, rows_per_group);
    B2_LAUNCH_OK("b2_rfft_any(store)");
    return B2_OK;
}
