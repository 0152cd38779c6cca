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
// autosorting).  The real-FFT untangling is fused into the store, which writes frame pairs of four neighbouring
// bin slots as whole 128-byte lines of the [frame chunk][slot group][channel][8 pairs][4 slots][2 frames] spectra
// layout (common.cuh) that the rank-K kernel consumes without transposition.
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

// v[r] *= w^r, r = 1 .. R-1: one table value, the powers by products (depth <= 3 multiplications, so the error
// stays at a few ulp)
template <typename T, int R> __device__ __forceinline__ void twiddle_powers(cx<T> *v, const cx<T> w1) {
    if constexpr (R == 2) {
        v[1] = cmul(v[1], w1);
    } else {
        const cx<T> w2 = cmul(w1, w1), w3 = cmul(w1, w2);
        v[1] = cmul(v[1], w1);
        v[2] = cmul(v[2], w2);
        v[3] = cmul(v[3], w3);
        if constexpr (R >= 8) {
            const cx<T> w4 = cmul(w2, w2);
            v[4] = cmul(v[4], w4);
            v[5] = cmul(v[5], cmul(w4, w1));
            v[6] = cmul(v[6], cmul(w4, w2));
            v[7] = cmul(v[7], cmul(w4, w3));
            if constexpr (R == 16) {
                const cx<T> w8 = cmul(w4, w4);
                v[8] = cmul(v[8], w8);
                v[9] = cmul(v[9], cmul(w8, w1));
                v[10] = cmul(v[10], cmul(w8, w2));
                v[11] = cmul(v[11], cmul(w8, w3));
                const cx<T> w12 = cmul(w8, w4);
                v[12] = cmul(v[12], w12);
                v[13] = cmul(v[13], cmul(w12, w1));
                v[14] = cmul(v[14], cmul(w12, w2));
                v[15] = cmul(v[15], cmul(w12, w3));
            }
        }
    }
}

// One Stockham pass over all FR frames in the CTA: radix R, NS = product of earlier radices.
// The LAST pass (NS * R == NC) leaves its trailing barrier to the caller, which issues the twiddle loads of the
// untangling step in front of it.
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
                twiddle_powers<T, R>(v[it], __ldg(tw + k * TSTEP));
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
    if constexpr (NS * R < NC) __syncthreads();
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

// Radix of the pass that follows when NS points are already combined (0: none left).
template <typename T, int LOG2NC> __host__ __device__ constexpr int next_radix(int ns) {
    const int rem = (1 << LOG2NC) / ns;
    if (rem <= 1) return 0;
    if (rem >= 16) return Geo<T, LOG2NC>::RMAX >= 16 ? 16 : 8;
    return rem;
}
// Radix of the LAST pass after a first pass of radix r1 (0: the first pass is the only one).
template <typename T, int LOG2NC> __host__ __device__ constexpr int last_radix(int r1) {
    int ns = r1, r = 0;
    while (next_radix<T, LOG2NC>(ns) != 0) { r = next_radix<T, LOG2NC>(ns); ns *= r; }
    return r;
}

// Run the remaining passes: radix 16 while >= 16 points remain, then one radix 8 / 4 / 2 pass.  With SKIP_LAST the
// last pass is left to last_pass_untangle().
template <typename T, int LOG2NC, int NS, bool SKIP_LAST>
__device__ __forceinline__ void stockham_all(cx<T> *s, const cx<T> *__restrict__ tw, int tid) {
    constexpr int R = next_radix<T, LOG2NC>(NS);
    if constexpr (R != 0) {
        constexpr bool last = next_radix<T, LOG2NC>(NS * R) == 0;
        if constexpr (!(last && SKIP_LAST)) {
            stockham_pass<T, LOG2NC, R, NS>(s, tw, tid);
            stockham_all<T, LOG2NC, NS * R, SKIP_LAST>(s, tw, tid);
        }
    }
}

// X[p] = 0.5 (a + conj b) - 0.5 i w_p (a - conj b),  a = Z[p], b = Z[NC-p], w_p = exp(-2 pi i p / wl)
// and, because w_{NC-p} = -conj(w_p), X[NC-p] = conj(0.5 (a + conj b) + 0.5 i w_p (a - conj b)).
// `bc` is conj(b); p = 0 gives the DC bin (lo) and the Nyquist bin (hi).
template <typename T>
__device__ __forceinline__ void untangle_pair(cx<T> a, cx<T> bc, cx<T> w, T half, cx<T> &lo, cx<T> &hi) {
    const cx<T> sum = cadd(a, bc), dif = cmul(w, csub(a, bc));
    lo = mk<T>((sum.x + dif.y) * half, (sum.y - dif.x) * half);
    hi = mk<T>((sum.x - dif.y) * half, -(sum.y + dif.x) * half);
}

// Last Stockham pass (radix R, NS = NC / R) fused with the untangling.  After its butterfly a thread holds
// Z[j + J r], r = 0 .. R-1 (J = NC / R), of BOTH frames of a pair; the partner of p = j + J r is
// NC - p = (J - j) + J (R-1-r), held by thread J - j.  Every thread publishes its upper half (r >= R/2) in shared
// memory, reads the partner values of its lower half from there and writes bins p and NC - p for r < R/2: half the
// shared-memory traffic of a full pass plus an untangling pass, and one latency chain less.  Thread j = 0 pairs with
// itself (r <-> R - r) and owns the DC / Nyquist pair (r = 0) and the self-paired bin NC/2 (r = R/2).
template <typename T, int LOG2NC, int R>
__device__ __forceinline__ void last_pass_untangle(cx<T> *s, const cx<T> *__restrict__ tw, int tid, T half, int detrend,
                                                   cx<T> *Xc, int64_t gstride, int64_t k0) {
    using G = Geo<T, LOG2NC>;
    using C = cx<T>;
    constexpr int NC = G::NC, J = NC / R, H = R / 2, NP = G::FR / 2;
    constexpr int TOTAL = NP * J;
    constexpr int ITEMS = (TOTAL + G::NT - 1) / G::NT;
    C v[ITEMS][2][R];
    C wu[ITEMS][H];
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int item = tid + it * G::NT;
        if (item < TOTAL) {
            const int pr = item / J, j = item - pr * J;
            const C w1 = __ldg(tw + 2 * j);               // exp(-2 pi i j / NC)
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                const C *base = s + (2 * pr + f) * G::PITCH;
#pragma unroll
                for (int r = 0; r < R; ++r) v[it][f][r] = base[padidx<G::PADSH>(j + r * J)];
                twiddle_powers<T, R>(v[it][f], w1);
                dftR<T, R>(v[it][f]);
            }
#pragma unroll
            for (int r = 0; r < H; ++r) wu[it][r] = __ldg(tw + j + r * J);
        }
    }
    __syncthreads();
    // exchange buffer E[(frame, q)][j], q = r - R/2: lanes write and read consecutive elements, no padding needed
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int item = tid + it * G::NT;
        if (item < TOTAL) {
            const int pr = item / J, j = item - pr * J;
#pragma unroll
            for (int f = 0; f < 2; ++f)
#pragma unroll
                for (int q = 0; q < H; ++q) s[((2 * pr + f) * H + q) * J + j] = v[it][f][H + q];
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int item = tid + it * G::NT;
        if (item < TOTAL) {
            const int pr = item / J, j = item - pr * J;
            const bool self = (j == 0);
            const int jj = self ? 0 : J - j;
            const int64_t k = k0 + 2 * pr;
            C *dst = Xc + b2_slot_frame_offset(0, int(k & 15));
#pragma unroll
            for (int r = 0; r < H; ++r) {
                const int p = j + r * J;
                const int q = self ? H - r : H - 1 - r;                 // partner's r' - R/2 (unused for p = 0)
                C lo[2], hi[2];
#pragma unroll
                for (int f = 0; f < 2; ++f) {
                    const C a = v[it][f][r];
                    const C bc = (self && r == 0) ? cconj(a) : cconj(s[((2 * pr + f) * H + (q < H ? q : 0)) * J + jj]);
                    untangle_pair<T>(a, bc, wu[it][r], half, lo[f], hi[f]);
                }
                // Removing the frame mean (_spectral_methods.py:136-137) changes the DC bin only: the DFT of a
                // constant vanishes at every other bin.  The reference's own DC value is rounding noise.
                if (detrend && p == 0) lo[0] = lo[1] = mk<T>(T(0), T(0));
                store_pair(dst + (p >> 2) * gstride + (p & 3) * 2, lo[0], lo[1]);
                store_pair(dst + ((NC / 2 + p) >> 2) * gstride + ((NC / 2 + p) & 3) * 2, hi[0], hi[1]);
            }
            if (self) {                                                  // bin NC/2: a = Z[NC/2], b = a, w = -i
                C lo[2], hi[2];
#pragma unroll
                for (int f = 0; f < 2; ++f)
                    untangle_pair<T>(v[it][f][H], cconj(v[it][f][H]), mk<T>(T(0), T(-1)), half, lo[f], hi[f]);
                store_pair(dst + (NC >> 2) * gstride + (NC & 3) * 2, lo[0], lo[1]);
            }
        }
    }
}

template <typename T, int LOG2NC>
__global__ void __launch_bounds__(Geo<T, LOG2NC>::NT, Geo<T, LOG2NC>::MINB)
welch_rfft_kernel(const T *__restrict__ xc, int64_t Lp, int step, int64_t nframes,
                  const T *__restrict__ window, const cx<T> *__restrict__ tw, int detrend,
                  T scale, cx<T> *__restrict__ X, int M, int vec_ok) {
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
    const T *xrow = xc + int64_t(m) * Lp;

    // ---- wl/2-point complex FFT of every frame; the first pass reads global memory directly ----
    constexpr int R1 = NC >= G::RMAX ? G::RMAX : NC;
    constexpr int RL = last_radix<T, LOG2NC>(R1);
#ifndef B2_FFT_FUSED
#define B2_FFT_FUSED 1
#endif
    constexpr bool FUSED = B2_FFT_FUSED && FR >= 2 && RL >= 2;
    first_pass_from_global<T, LOG2NC, R1>(s, xrow, step, k0, nframes, window, vec_ok, tid);
    stockham_all<T, LOG2NC, R1, FUSED>(s, tw, tid);

    // ---- untangle to the wl/2+1 real-FFT bins and store; bins p and wl/2-p share one pair of reads ----
    // Spectra layout (common.cuh): [frame chunk][slot group][channel][8 frame pairs][4 slots][2 frames]; the lanes
    // of a warp take neighbouring p, so that four of them fill one 128-byte line with the frame pair of four
    // neighbouring slots -- for the lower half of the spectrum (slot p) and for the upper half (slot NC/2 + p, bins
    // descending) alike.  In the [bin][channel][16 frames] order of r1 every 32-byte store of a warp went to a line
    // of its own: 55 LSU wavefronts per store instruction, 43 % of the wavefronts of the kernel, whose LSU data
    // pipe was 87 % busy (profiles/r2_welch_rfft_ncu_full_v2.json).
    const T half = T(0.5) * scale;
    constexpr int NG = (NC + 1 + 3) / 4;
    C *Xc = X + ((k0 >> 4) * int64_t(NG) * M + m) * B2_GROUP_ELEMS;       // (chunk, group 0, channel m)
    const int64_t gstride = int64_t(M) * B2_GROUP_ELEMS;
    if constexpr (FUSED) {
        last_pass_untangle<T, LOG2NC, RL>(s, tw, tid, half, detrend, Xc, gstride, k0);
    } else {
        auto zat = [&](int fr, int p) { return s[fr * G::PITCH + padidx<G::PADSH>(p)]; };
        auto slot_ptr = [&](int slot, int64_t k) { return Xc + (slot >> 2) * gstride + b2_slot_frame_offset(slot, int(k & 15)); };
        constexpr int NP = FR >= 2 ? FR / 2 : 1;              // frame pairs (FR = 1: single frames)
        constexpr int ITEMS = (NC / 2) * NP;                  // (p = 0 .. NC/2 - 1) x pairs
        constexpr int ITER = (ITEMS + NT - 1) / NT;
        C w[ITER];
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int idx = tid + it * NT;
            w[it] = __ldg(tw + (idx < ITEMS ? idx / NP : 0));
        }
        __syncthreads();                                      // the barrier the last FFT pass left out
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int idx = tid + it * NT;
            if (idx < ITEMS) {
                const int pr = idx % NP, p = idx / NP, pn = (NC - p) & (NC - 1);
                if constexpr (FR >= 2) {
                    C lo0, hi0, lo1, hi1;
                    untangle_pair<T>(zat(2 * pr, p), cconj(zat(2 * pr, pn)), w[it], half, lo0, hi0);
                    untangle_pair<T>(zat(2 * pr + 1, p), cconj(zat(2 * pr + 1, pn)), w[it], half, lo1, hi1);
                    if (detrend && p == 0) lo0 = lo1 = mk<T>(T(0), T(0));
                    const int64_t k = k0 + 2 * pr;
                    store_pair(slot_ptr(p, k), lo0, lo1);
                    store_pair(slot_ptr(NC / 2 + p, k), hi0, hi1);
                } else {
                    C lo, hi;
                    untangle_pair<T>(zat(0, p), cconj(zat(0, pn)), w[it], half, lo, hi);
                    if (detrend && p == 0) lo = mk<T>(T(0), T(0));
                    *slot_ptr(p, k0) = lo;
                    *slot_ptr(NC / 2 + p, k0) = hi;
                }
            }
        }
        if (tid < NP) {                                       // the self-paired bin NC/2: w = -i
            const C mi = mk<T>(T(0), T(-1));
            C lo0, hi0, lo1, hi1;
            if constexpr (FR >= 2) {
                untangle_pair<T>(zat(2 * tid, NC / 2), cconj(zat(2 * tid, NC / 2)), mi, half, lo0, hi0);
                untangle_pair<T>(zat(2 * tid + 1, NC / 2), cconj(zat(2 * tid + 1, NC / 2)), mi, half, lo1, hi1);
                store_pair(slot_ptr(NC, k0 + 2 * tid), lo0, lo1);
            } else {
                untangle_pair<T>(zat(0, NC / 2), cconj(zat(0, NC / 2)), mi, half, lo0, hi0);
                *slot_ptr(NC, k0) = lo0;
            }
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
                                      (const cx<T> *)tw, detrend, T(scale), (cx<T> *)X, M, vec_ok);
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

extern "C" int64_t b2_welch_spectra_elems(int wl, int M, int64_t Kp) {
    if (ilog2_exact(wl) < 3 || M <= 0 || Kp <= 0 || Kp % 16 != 0) {
        b2_set_error("b2_welch_spectra_elems: wl=%d must be a power of two >= 8, M=%d positive, Kp=%lld a positive multiple of 16", wl, M, (long long)Kp);
        return -1;
    }
    return (Kp / 16) * int64_t(b2_slot_groups(wl / 2)) * M * B2_GROUP_ELEMS;
}

extern "C" int b2_welch_bin_slot(int wl, int bin) {
    if (ilog2_exact(wl) < 3 || bin < 0 || bin > wl / 2) {
        b2_set_error("b2_welch_bin_slot: bin %d outside [0, %d] or wl not a power of two >= 8", bin, wl / 2);
        return -1;
    }
    return b2_bin_slot(bin, wl / 2);
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
