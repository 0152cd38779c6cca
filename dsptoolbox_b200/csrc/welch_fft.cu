// K0 (channel gather / zero-pad / convert) and K1 (frame + window + detrend + real FFT).
//
// Replaces, for every channel at once, what the reference does per channel PAIR:
//   _get_framed_signal  (standard/_framed_signal_representation.py:50-65)
//   window, mean removal, np.fft.rfft in _welch (standard/_spectral_methods.py:126-148)
//
// K1 design (HBM-bound): one CTA owns FR consecutive frames of ONE channel.  The frames are
// read from the channel-major padded copy with vector loads, multiplied by the window on the
// way into shared memory, packed as wl/2 complex points (z[n] = y[2n] + i*y[2n+1]), and
// transformed by an in-place shared-memory Stockham FFT (radix 8/4/2, register-staged,
// autosorting).  The real-FFT untangling is fused into the store, which writes, for every
// bin, the FR frames as one contiguous run of the [frame chunk][bin][channel][16 frames] spectra
// layout that the rank-K kernel consumes without transposition.
#include "common.cuh"
#include "fft_small.cuh"

namespace {

// ---------------------------------------------------------------- compile-time geometry
template <typename T, int LOG2NC> struct Geo {
    static constexpr int NC = 1 << LOG2NC;              // complex points = wl / 2
    static constexpr int N = 2 * NC;                    // wl
    static constexpr int CXB = int(sizeof(cx<T>));
#ifndef B2_FFT_VPT64
#define B2_FFT_VPT64 16
#endif
#ifndef B2_FFT_VPT32
#define B2_FFT_VPT32 16
#endif
    static constexpr int VPT = sizeof(T) == 8 ? B2_FFT_VPT64 : B2_FFT_VPT32;  // complex values held per thread
    static constexpr int RMAX = VPT >= 16 ? 16 : 8;                          // largest butterfly held in registers
    static constexpr int R1 = NC >= RMAX ? RMAX : NC;                        // radix of the first pass
    // One pad slot per R1 points: the first pass writes with a stride of R1 points between lanes, which the pad turns
    // into R1 + 1 -- odd, so the 8 (fp64) or 16 (fp32) lanes of one shared-memory wavefront fall into distinct
    // 16-byte / 8-byte bank groups.  (One slot per 8 points under a radix-16 pass gave a stride of 18 and two-way
    // conflicts: 29 % of all shared wavefronts of the fp64 nfft-4096 kernel, profiles/r2_k1_ncu.txt.)
    static constexpr int PADSH = R1 >= 16 ? 4 : 3;
    static constexpr int NCP = NC + (NC >> PADSH);
    static constexpr int PITCH = NCP + 1;               // per-frame pitch (staggers frames)
    static constexpr int frames_fit(int fr) { return fr * PITCH * CXB <= 74 * 1024; }   // 3 CTAs per SM
    static constexpr int FR = frames_fit(16) ? 16 : frames_fit(8) ? 8 : frames_fit(4) ? 4 : frames_fit(2) ? 2 : 1;
    static constexpr int NT_RAW = FR * NC / VPT;
    static constexpr int NT = NT_RAW < 64 ? 64 : (NT_RAW > 1024 ? 1024 : NT_RAW);
    static constexpr int SMEM = FR * PITCH * CXB;
    // two resident CTAs (128 registers per thread) where the CTA is small enough: 6.20 instead of 7.85 ms for
    // 256 channels x 60 s x nfft 4096 in fp64 (profiles/r2_stage1_variants.txt); three (80 registers) spills and loses
#ifndef B2_FFT_MINB
#define B2_FFT_MINB 2
#endif
    static constexpr int MINB = NT * B2_FFT_MINB <= 512 ? B2_FFT_MINB : 1;
};

template <int SH> __device__ __forceinline__ int padidx(int i) { return i + (i >> SH); }

// Two neighbouring complex values with one store: 256-bit for fp64 (sm_100 has 32-byte global accesses), 128-bit for fp32.
__device__ __forceinline__ void store_pair(double2 *p, double2 a, double2 b) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}
__device__ __forceinline__ void store_pair(float2 *p, float2 a, float2 b) {
    *reinterpret_cast<float4 *>(p) = make_float4(a.x, a.y, b.x, b.y);
}

// One Stockham pass over all FR frames in the CTA: radix R, NS = product of earlier radices.
template <typename T, int LOG2NC, int R, int NS>
__device__ __forceinline__ void stockham_pass(cx<T> *s, const cx<T> *__restrict__ tw, int tid) {
    using G = Geo<T, LOG2NC>;
    constexpr int NC = G::NC, J = NC / R;                // butterflies per frame
    constexpr int TOTAL = G::FR * J;
    constexpr int ITEMS = (TOTAL + G::NT - 1) / G::NT;
    cx<T> v[ITEMS][R];
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        int item = tid + it * G::NT;
        if (item < TOTAL) {
            int fr = item / J, j = item - fr * J;
            int k = j & (NS - 1);
            cx<T> *base = s + fr * G::PITCH;
#pragma unroll
            for (int r = 0; r < R; ++r) v[it][r] = base[padidx<G::PADSH>(j + r * J)];
            if constexpr (NS > 1) {
                // twiddles w^r, w = exp(-2 pi i k / (NS R)): one table read, powers by products
                // (depth <= 3 multiplications, so the error stays at a few ulp)
                constexpr int TSTEP = G::N / (NS * R);
                const cx<T> w1 = __ldg(tw + k * TSTEP);
                if constexpr (R == 2) {
                    v[it][1] = cmul(v[it][1], w1);
                } else {
                    const cx<T> w2 = cmul(w1, w1), w3 = cmul(w1, w2);
                    v[it][1] = cmul(v[it][1], w1);
                    v[it][2] = cmul(v[it][2], w2);
                    v[it][3] = cmul(v[it][3], w3);
                    if constexpr (R >= 8) {
                        const cx<T> w4 = cmul(w2, w2);
                        v[it][4] = cmul(v[it][4], w4);
                        v[it][5] = cmul(v[it][5], cmul(w4, w1));
                        v[it][6] = cmul(v[it][6], cmul(w4, w2));
                        v[it][7] = cmul(v[it][7], cmul(w4, w3));
                        if constexpr (R == 16) {
                            const cx<T> w8 = cmul(w4, w4);
                            v[it][8] = cmul(v[it][8], w8);
                            v[it][9] = cmul(v[it][9], cmul(w8, w1));
                            v[it][10] = cmul(v[it][10], cmul(w8, w2));
                            v[it][11] = cmul(v[it][11], cmul(w8, w3));
                            const cx<T> w12 = cmul(w8, w4);
                            v[it][12] = cmul(v[it][12], w12);
                            v[it][13] = cmul(v[it][13], cmul(w12, w1));
                            v[it][14] = cmul(v[it][14], cmul(w12, w2));
                            v[it][15] = cmul(v[it][15], cmul(w12, w3));
                        }
                    }
                }
            }
            dftR<T, R>(v[it]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        int item = tid + it * G::NT;
        if (item < TOTAL) {
            int fr = item / J, j = item - fr * J;
            int k = j & (NS - 1);
            int j0 = (j - k) * R + k;
            cx<T> *base = s + fr * G::PITCH;
#pragma unroll
            for (int r = 0; r < R; ++r) base[padidx<G::PADSH>(j0 + r * NS)] = v[it][r];
        }
    }
    __syncthreads();
}

// First pass fused with the load: every butterfly reads its R inputs straight from global memory
// (frame samples x window, packed as complex pairs; coalesced because consecutive butterflies read
// consecutive pairs), so the time-domain frame never makes a round trip through shared memory.
template <typename T, int LOG2NC, int R>
__device__ __forceinline__ void first_pass_from_global(cx<T> *s, const T *__restrict__ xrow, int step, int64_t k0,
                                                       int64_t nframes, const T *__restrict__ window, int vec_ok,
                                                       int tid) {
    using G = Geo<T, LOG2NC>;
    using C = cx<T>;
    constexpr int NC = G::NC, J = NC / R;
    constexpr int TOTAL = G::FR * J;
    constexpr int ITEMS = (TOTAL + G::NT - 1) / G::NT;
    C v[ITEMS][R];
    const C *w2 = reinterpret_cast<const C *>(window);
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int item = tid + it * G::NT;
        if (item < TOTAL) {
            const int fr = item / J, j = item - fr * J;
            const int64_t k = k0 + fr;
            if (k < nframes) {
                const T *src = xrow + k * int64_t(step);
                if (vec_ok) {
                    const C *src2 = reinterpret_cast<const C *>(src);
#pragma unroll
                    for (int r = 0; r < R; ++r) v[it][r] = __ldg(src2 + j + r * J);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const C w = __ldg(w2 + j + r * J);
                        v[it][r] = mk<T>(v[it][r].x * w.x, v[it][r].y * w.y);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int n = j + r * J;
                        v[it][r] = mk<T>(src[2 * n] * window[2 * n], src[2 * n + 1] * window[2 * n + 1]);
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) v[it][r] = mk<T>(T(0), T(0));
            }
            dftR<T, R>(v[it]);
            C *base = s + fr * G::PITCH;
#pragma unroll
            for (int r = 0; r < R; ++r) base[padidx<G::PADSH>(j * R + r)] = v[it][r];
        }
    }
    __syncthreads();
}

// Run the remaining passes: radix 16 while >= 16 points remain, then one radix 8 / 4 / 2 pass.
template <typename T, int LOG2NC, int NS>
__device__ __forceinline__ void stockham_all(cx<T> *s, const cx<T> *__restrict__ tw, int tid) {
    constexpr int NC = 1 << LOG2NC;
    constexpr int REM = NC / NS;
    if constexpr (REM >= 16 && Geo<T, LOG2NC>::RMAX >= 16) {
        stockham_pass<T, LOG2NC, 16, NS>(s, tw, tid);
        stockham_all<T, LOG2NC, NS * 16>(s, tw, tid);
    } else if constexpr (REM >= 16) {
        stockham_pass<T, LOG2NC, 8, NS>(s, tw, tid);
        stockham_all<T, LOG2NC, NS * 8>(s, tw, tid);
    } else if constexpr (REM == 8) {
        stockham_pass<T, LOG2NC, 8, NS>(s, tw, tid);
    } else if constexpr (REM == 4) {
        stockham_pass<T, LOG2NC, 4, NS>(s, tw, tid);
    } else if constexpr (REM == 2) {
        stockham_pass<T, LOG2NC, 2, NS>(s, tw, tid);
    }
}

template <typename T, int LOG2NC>
__global__ void __launch_bounds__(Geo<T, LOG2NC>::NT, Geo<T, LOG2NC>::MINB)
welch_rfft_kernel(const T *__restrict__ xc, int64_t Lp, int step, int64_t nframes,
                  const T *__restrict__ window, const cx<T> *__restrict__ tw, int detrend,
                  T scale, cx<T> *__restrict__ X, int M, int64_t Kp, int vec_ok) {
    using G = Geo<T, LOG2NC>;
    using C = cx<T>;
    constexpr int NC = G::NC, FR = G::FR, NT = G::NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *s = reinterpret_cast<C *>(smem_raw);

    const int tid = threadIdx.x;
    // CTA order: frame group inside a 16-frame chunk fastest, then channel, then chunk -- the CTAs
    // that complete one 16-frame run of the spectra layout run together (L2 merges their stores)
    // and a whole wave stays inside one chunk window.
    constexpr int GPC = 16 / (FR < 16 ? FR : 16);
    const int64_t bid = blockIdx.x;
    const int gi = int(bid % GPC);
    const int64_t rest = bid / GPC;
    const int m = int(rest % M);
    const int64_t k0 = (rest / M) * 16 + int64_t(gi) * FR;
    const int64_t nb = NC + 1;
    const T *xrow = xc + int64_t(m) * Lp;

    // ---- wl/2-point complex FFT of every frame; the first pass reads global memory directly ----
    constexpr int R1 = NC >= G::RMAX ? G::RMAX : NC;
    first_pass_from_global<T, LOG2NC, R1>(s, xrow, step, k0, nframes, window, vec_ok, tid);
    stockham_all<T, LOG2NC, R1>(s, tw, tid);

    // ---- untangle to the wl/2+1 real-FFT bins and store; bins k and wl/2-k share one pair of reads ----
    // X[k] = 0.5 (a + conj b) - 0.5 i w_k (a - conj b),  a = Z[k], b = Z[NC-k], w_k = exp(-2 pi i k / wl)
    // and, because w_{NC-k} = -conj(w_k), X[NC-k] = conj(0.5 (a + conj b) + 0.5 i w_k (a - conj b)).
    // Spectra layout [frame chunk of 16][bin][channel][16 frames]: all CTAs of one frame chunk write
    // inside one (nb * M * 16)-element window and the rank-K kernel reads whole tiles contiguously.
    const T half = T(0.5) * scale;
    // Removing the frame mean (_spectral_methods.py:136-137) changes the DC bin only: the DFT of a constant
    // vanishes at every other bin.  The reference's own DC value is rounding noise.
    auto untangle = [&](const C *base, int kp, C &lo, C &hi) {
        const C a = base[padidx<G::PADSH>(kp)];
        const C b = cconj(base[padidx<G::PADSH>((NC - kp) & (NC - 1))]);
        const C w = __ldg(tw + kp);
        const C sum = cadd(a, b), dif = cmul(w, csub(a, b));      // dif = w (a - conj b)
        lo = mk<T>((sum.x + dif.y) * half, (sum.y - dif.x) * half);
        hi = mk<T>((sum.x - dif.y) * half, -(sum.y + dif.x) * half);   // bin NC - kp (kp = 0: the Nyquist bin)
        if (detrend && kp == 0) lo = mk<T>(T(0), T(0));
    };
    if constexpr (FR >= 2) {
        // two neighbouring frames of one bin per thread: one 32-byte (fp64) / 16-byte (fp32) store instead of two --
        // the scattered 16-byte stores were 40 % of the LSU wavefronts of this kernel
        constexpr int NP = FR / 2;
        for (int idx = tid; idx < (NC / 2 + 1) * NP; idx += NT) {
            const int pr = idx % NP, kp = idx / NP;
            C lo0, hi0, lo1, hi1;
            untangle(s + (2 * pr) * G::PITCH, kp, lo0, hi0);
            untangle(s + (2 * pr + 1) * G::PITCH, kp, lo1, hi1);
            const int64_t k = k0 + 2 * pr;
            store_pair(X + (((k >> 4) * nb + kp) * M + m) * 16 + (k & 15), lo0, lo1);
            if (2 * kp != NC) store_pair(X + (((k >> 4) * nb + (NC - kp)) * M + m) * 16 + (k & 15), hi0, hi1);
        }
    } else {
        for (int kp = tid; kp < NC / 2 + 1; kp += NT) {
            C lo, hi;
            untangle(s, kp, lo, hi);
            X[(((k0 >> 4) * nb + kp) * M + m) * 16 + (k0 & 15)] = lo;
            if (2 * kp != NC) X[(((k0 >> 4) * nb + (NC - kp)) * M + m) * 16 + (k0 & 15)] = hi;
        }
    }
}

template <typename T>
__global__ void twiddle_kernel(cx<T> *tw, int N) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < N) {
        double sn, cs;
        sincospi(2.0 * double(t) / double(N), &sn, &cs);
        tw[t] = mk<T>(T(cs), T(-sn));
    }
}

// ---------------------------------------------------------------- K0
// time-major source (stride_m == 1 typical): 32x32 tile transpose through shared memory.
template <typename IN, typename OUT>
__global__ void gather_transpose_kernel(const IN *__restrict__ x, int64_t stride_t, int64_t stride_m,
                                        int64_t t0, int64_t nt, int M, OUT *__restrict__ out, int64_t Lp) {
    __shared__ OUT tile[32][33];
    const int64_t tb = int64_t(blockIdx.x) * 32;
    const int mb = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
#pragma unroll
    for (int r = 0; r < 32; r += 8) {
        int64_t t = tb + ty + r;
        int m = mb + tx;
        OUT v = OUT(0);
        if (t < nt && m < M) v = OUT(x[(t0 + t) * stride_t + int64_t(m) * stride_m]);
        tile[ty + r][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 32; r += 8) {
        int m = mb + ty + r;
        int64_t t = tb + tx;
        if (m < M && t < Lp) out[int64_t(m) * Lp + t] = tile[tx][ty + r];
    }
}

// channel-major source (stride_t == 1): plain convert/copy with zero fill.
template <typename IN, typename OUT>
__global__ void gather_rows_kernel(const IN *__restrict__ x, int64_t stride_m, int64_t t0, int64_t nt,
                                   OUT *__restrict__ out, int64_t Lp) {
    const int m = blockIdx.y;
    const IN *src = x + int64_t(m) * stride_m + t0;
    OUT *dst = out + int64_t(m) * Lp;
    for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < Lp; t += int64_t(gridDim.x) * blockDim.x)
        dst[t] = t < nt ? OUT(src[t]) : OUT(0);
}

template <typename IN, typename OUT>
int launch_gather(const void *x, int64_t stride_t, int64_t stride_m, int64_t t0, int64_t nt, int M,
                  void *out, int64_t Lp, cudaStream_t st) {
    if (stride_t == 1 && stride_m != 1) {
        int64_t bx = (Lp + 255) / 256;
        if (bx > 4096) bx = 4096;
        gather_rows_kernel<IN, OUT><<<dim3((unsigned)bx, M), 256, 0, st>>>(
            (const IN *)x, stride_m, t0, nt, (OUT *)out, Lp);
    } else {
        dim3 grid((unsigned)((Lp + 31) / 32), (unsigned)((M + 31) / 32));
        gather_transpose_kernel<IN, OUT><<<grid, dim3(32, 8), 0, st>>>(
            (const IN *)x, stride_t, stride_m, t0, nt, M, (OUT *)out, Lp);
    }
    B2_LAUNCH_OK("b2_prepare_channels");
    return B2_OK;
}

template <typename T, int LOG2NC>
int launch_rfft(const void *xc, int64_t Lp, int M, int step, int64_t nframes, const void *window,
                const void *tw, int detrend, double scale, void *X, int64_t Kp, cudaStream_t st) {
    using G = Geo<T, LOG2NC>;
    auto kern = welch_rfft_kernel<T, LOG2NC>;
    static DeviceOnce once;          // per instantiation and device
    if (int rc = b2_device_once(once, kern, G::SMEM, nullptr)) return rc;
    B2_REQUIRE(Kp % G::FR == 0, B2_EINVAL, "b2_welch_rfft: Kp=%lld must be a multiple of %d", (long long)Kp, G::FR);
    int vec_ok = (step % 2 == 0) && (Lp % 2 == 0) && (reinterpret_cast<uintptr_t>(xc) % (2 * sizeof(T)) == 0) &&
                 (reinterpret_cast<uintptr_t>(window) % (2 * sizeof(T)) == 0);
    const int64_t nblk = (Kp / G::FR) * int64_t(M);
    B2_REQUIRE(Kp % 16 == 0 && nblk <= 0x7fffffff, B2_EINVAL, "b2_welch_rfft: Kp=%lld must be a multiple of 16 (grid too large otherwise)", (long long)Kp);
    dim3 grid((unsigned)nblk);
    kern<<<grid, G::NT, G::SMEM, st>>>((const T *)xc, Lp, step, nframes, (const T *)window,
                                      (const cx<T> *)tw, detrend, T(scale), (cx<T> *)X, M, Kp, vec_ok);
    B2_LAUNCH_OK("b2_welch_rfft");
    return B2_OK;
}

template <typename T>
int dispatch_rfft(int log2nc, const void *xc, int64_t Lp, int M, int step, int64_t nframes,
                  const void *window, const void *tw, int detrend, double scale, void *X, int64_t Kp,
                  cudaStream_t st) {
    switch (log2nc) {
#define B2_CASE(L) case L: return launch_rfft<T, L>(xc, Lp, M, step, nframes, window, tw, detrend, scale, X, Kp, st);
        B2_CASE(2) B2_CASE(3) B2_CASE(4) B2_CASE(5) B2_CASE(6) B2_CASE(7) B2_CASE(8)
        B2_CASE(9) B2_CASE(10) B2_CASE(11) B2_CASE(12) B2_CASE(13)
#undef B2_CASE
    default: break;
    }
    b2_set_error("b2_welch_rfft: window length 2^%d is outside the shared-memory FFT range [8, 16384]", log2nc + 1);
    return B2_EUNSUPPORTED;
}

template <typename T> int frame_tile(int log2nc) {
    switch (log2nc) {
#define B2_CASE(L) case L: return Geo<T, L>::FR;
        B2_CASE(2) B2_CASE(3) B2_CASE(4) B2_CASE(5) B2_CASE(6) B2_CASE(7) B2_CASE(8)
        B2_CASE(9) B2_CASE(10) B2_CASE(11) B2_CASE(12) B2_CASE(13)
#undef B2_CASE
    default: return -1;
    }
}

}  // namespace

extern "C" int b2_welch_frames(int64_t L, int wl, int step, int64_t *frames, int64_t *pad) {
    B2_REQUIRE(L > 0 && wl > 0 && step > 0, B2_EINVAL, "b2_welch_frames: L, wl and step must be positive");
    if (frames) *frames = (L + step - 1) / step;
    if (pad) *pad = wl - (L % step);
    return B2_OK;
}

extern "C" int b2_welch_frame_tile(int wl, int precision) {
    int l2 = ilog2_exact(wl);
    if (l2 < 3 || l2 > 14) { b2_set_error("b2_welch_frame_tile: unsupported window length %d", wl); return B2_EUNSUPPORTED; }
    return precision == B2_F64 ? frame_tile<double>(l2 - 1) : frame_tile<float>(l2 - 1);
}

extern "C" int b2_fft_twiddles(int wl, int precision, void *twiddle_dev, void *stream) {
    B2_REQUIRE(twiddle_dev != nullptr, B2_EINVAL, "b2_fft_twiddles: null output");
    B2_REQUIRE(ilog2_exact(wl) >= 3, B2_EINVAL, "b2_fft_twiddles: wl=%d must be a power of two >= 8", wl);
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == B2_F64) twiddle_kernel<double><<<(wl + 255) / 256, 256, 0, st>>>((double2 *)twiddle_dev, wl);
    else if (precision == B2_F32) twiddle_kernel<float><<<(wl + 255) / 256, 256, 0, st>>>((float2 *)twiddle_dev, wl);
    else { b2_set_error("b2_fft_twiddles: bad precision %d", precision); return B2_EINVAL; }
    B2_LAUNCH_OK("b2_fft_twiddles");
    return B2_OK;
}

extern "C" int b2_prepare_channels(const void *x_dev, int in_dtype, int64_t stride_t, int64_t stride_m,
                                   int64_t t0, int64_t nt, int M, void *out_dev, int64_t Lp,
                                   int precision, void *stream) {
    B2_REQUIRE(x_dev && out_dev, B2_EINVAL, "b2_prepare_channels: null pointer");
    B2_REQUIRE(M > 0 && nt >= 0 && Lp >= nt && t0 >= 0, B2_EINVAL,
               "b2_prepare_channels: bad sizes M=%d nt=%lld Lp=%lld t0=%lld", M, (long long)nt, (long long)Lp, (long long)t0);
    B2_REQUIRE(stride_t > 0 && stride_m > 0, B2_EINVAL, "b2_prepare_channels: strides must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    if (in_dtype == B2_F64 && precision == B2_F64) return launch_gather<double, double>(x_dev, stride_t, stride_m, t0, nt, M, out_dev, Lp, st);
    if (in_dtype == B2_F64 && precision == B2_F32) return launch_gather<double, float>(x_dev, stride_t, stride_m, t0, nt, M, out_dev, Lp, st);
    if (in_dtype == B2_F32 && precision == B2_F64) return launch_gather<float, double>(x_dev, stride_t, stride_m, t0, nt, M, out_dev, Lp, st);
    if (in_dtype == B2_F32 && precision == B2_F32) return launch_gather<float, float>(x_dev, stride_t, stride_m, t0, nt, M, out_dev, Lp, st);
    b2_set_error("b2_prepare_channels: bad dtype/precision %d/%d", in_dtype, precision);
    return B2_EINVAL;
}

extern "C" int b2_welch_rfft(const void *xc_dev, int64_t Lp, int M, int wl, int step, int64_t nframes,
                             const void *window_dev, const void *twiddle_dev, int detrend,
                             double fft_scale, void *X_out_dev, int64_t Kp, int precision, void *stream) {
    B2_REQUIRE(xc_dev && window_dev && twiddle_dev && X_out_dev, B2_EINVAL, "b2_welch_rfft: null pointer");
    int l2 = ilog2_exact(wl);
    B2_REQUIRE(l2 >= 3 && l2 <= 18, B2_EINVAL,
               "b2_welch_rfft: Window length should be a power of 2 between [8, 262_144] (got %d)", wl);
    B2_REQUIRE(step > 0 && step <= wl, B2_EINVAL, "b2_welch_rfft: step=%d must be in (0, wl]", step);
    B2_REQUIRE(M > 0 && M <= 65535, B2_EINVAL, "b2_welch_rfft: M=%d out of range", M);
    B2_REQUIRE(nframes >= 0 && nframes <= Kp, B2_EINVAL, "b2_welch_rfft: nframes=%lld exceeds Kp=%lld", (long long)nframes, (long long)Kp);
    B2_REQUIRE(nframes == 0 || Lp >= (nframes - 1) * int64_t(step) + wl, B2_EINVAL,
               "b2_welch_rfft: padded length Lp=%lld too short for %lld frames", (long long)Lp, (long long)nframes);
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == B2_F64) return dispatch_rfft<double>(l2 - 1, xc_dev, Lp, M, step, nframes, window_dev, twiddle_dev, detrend, fft_scale, X_out_dev, Kp, st);
    if (precision == B2_F32) return dispatch_rfft<float>(l2 - 1, xc_dev, Lp, M, step, nframes, window_dev, twiddle_dev, detrend, fft_scale, X_out_dev, Kp, st);
    b2_set_error("b2_welch_rfft: bad precision %d", precision);
    return B2_EINVAL;
}
