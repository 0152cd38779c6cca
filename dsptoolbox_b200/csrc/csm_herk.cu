// K2 (per-bin Hermitian rank-K accumulation on the FP64 tensor pipe) and K2e (finalize).
//
// Replaces the M(M+1)/2 `_welch` calls of _csm_welch (standard/_spectral_methods.py:351-369):
// with X_f in C^{M x K} the spectra of bin f over all channels and frames,
//   S_f[j,i] = sum_k X_f[j,k] conj(X_f[i,k])   (j >= i),
// i.e. the reference's conj(rfft(x_i)) * rfft(x_j) summed over frames (:141-152).
//
// K2 design: one CTA = one 64x64 block of the lower block-triangle of one bin; 4 warps, each
// a 32x32 sub-block as 4x4 m8n8k4 FP64 MMA tiles with real and imaginary accumulators.
// The complex product is mapped onto real DMMAs:
//   Sr += Ar*Br^T + Ai*Bi^T          Si += Ai*Br^T + Ar*(-Bi)^T
// Frames are streamed in chunks of 16 through a cp.async double buffer.  Spectra layout (common.cuh):
// [frame chunk][slot group][channel][8 frame pairs][4 slots][2 frames] -- a 64-row tile of one bin is 64 x 8 runs of
// 32 bytes (fp64), one per (channel, frame pair); the CTAs of the four bins of a slot group run next to each other
// (blockIdx.y) and use up the 128-byte lines between them in L2.
// The sums are kept PACKED: S[f][j (j + 1) / 2 + i] for j >= i (row-major lower triangle, M (M + 1) / 2
// complex128 values per bin) -- half the bytes of the square, and exactly the payload of the
// multi-GPU all-reduce (dist.py), which therefore needs no pack / unpack pass.
#include "common.cuh"

namespace {

constexpr int BM = 64;        // block rows/cols
constexpr int KC = 16;        // frames per chunk
constexpr int HERK_THREADS = 128;

template <typename TIN> struct HerkSmem {
    // row pitch in bytes: KC elements + padding chosen so that the 8 rows x 4 k fragment
    // pattern of m8n8k4 is bank-conflict free (see fragment loads below)
    static constexpr int ELEM = int(sizeof(cx<TIN>));
    static constexpr int PITCH = KC * ELEM + (sizeof(TIN) == 8 ? 64 : 32);
    static constexpr int TILE = BM * PITCH;
    static constexpr int BYTES = 2 /*stages*/ * 2 /*A,B*/ * TILE;
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <typename TIN> __device__ __forceinline__ double2 load_frag(const unsigned char *p);
template <> __device__ __forceinline__ double2 load_frag<double>(const unsigned char *p) {
    return *reinterpret_cast<const double2 *>(p);
}
template <> __device__ __forceinline__ double2 load_frag<float>(const unsigned char *p) {
    float2 v = *reinterpret_cast<const float2 *>(p);
    return make_double2(double(v.x), double(v.y));
}

// Issue the cp.async copies of one 64-row x KC-frame tile (rows >= M are left untouched: they
// were zeroed once before the main loop).  Xc points at this bin's slot inside [chunk][slot group]: channel m
// starts B2_GROUP_ELEMS elements further, frame pair p 8 elements further, the odd frame of a pair 1 element.
template <typename TIN>
__device__ __forceinline__ void load_tile(unsigned char *dst, const cx<TIN> *__restrict__ Xc, int row0,
                                          int M, int tid) {
    using SM = HerkSmem<TIN>;
    constexpr int SEG_PER_ROW = KC * SM::ELEM / 16;        // 16-byte segments per row
    constexpr int TOTAL = BM * SEG_PER_ROW;
#pragma unroll
    for (int s = tid; s < TOTAL; s += HERK_THREADS) {
        int r = s / SEG_PER_ROW, seg = s - r * SEG_PER_ROW;
        int row = row0 + r;
        if (row < M) {
            const int k = seg * (16 / SM::ELEM);               // first frame of this 16-byte segment
            cp_async16(dst + r * SM::PITCH + seg * 16, Xc + int64_t(row) * B2_GROUP_ELEMS + b2_slot_frame_offset(0, k));
        }
    }
}

template <typename TIN>
__global__ void __launch_bounds__(HERK_THREADS)
herk_dmma_kernel(const cx<TIN> *__restrict__ X, int M, int64_t Kp, int nb, int bin0, double2 *__restrict__ S,
                 int accumulate) {
    using SM = HerkSmem<TIN>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wj = warp >> 1, wi = warp & 1;

    // decode the lower-triangular block pair
    const int p = blockIdx.x;
    int bj = int((sqrtf(8.0f * float(p) + 1.0f) - 1.0f) * 0.5f);
    while ((bj + 1) * (bj + 2) / 2 <= p) ++bj;
    while (bj * (bj + 1) / 2 > p) --bj;
    const int bi = p - bj * (bj + 1) / 2;
    const bool diag = (bi == bj);
    const int f = bin0 + blockIdx.y;
    const int rowA0 = bj * BM, rowB0 = bi * BM;
    const int slot = b2_bin_slot(f, nb - 1);
    const cx<TIN> *Xf = X + int64_t(slot >> 2) * M * B2_GROUP_ELEMS + (slot & 3) * 2;      // chunk 0 of this bin
    const int64_t chunk_stride = int64_t(b2_slot_groups(nb - 1)) * M * B2_GROUP_ELEMS;      // elements between frame chunks

    unsigned char *As[2] = {smem, smem + 2 * SM::TILE};
    unsigned char *Bs[2] = {diag ? As[0] : smem + SM::TILE, diag ? As[1] : smem + 3 * SM::TILE};

    // rows past M contribute zeros
    for (int i = tid; i < SM::BYTES / 16; i += HERK_THREADS) reinterpret_cast<int4 *>(smem)[i] = make_int4(0, 0, 0, 0);
    __syncthreads();

    double sr[4][4][2], si[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { sr[a][b][0] = sr[a][b][1] = 0.0; si[a][b][0] = si[a][b][1] = 0.0; }

    const int nchunks = int(Kp / KC);
    load_tile<TIN>(As[0], Xf, rowA0, M, tid);
    if (!diag) load_tile<TIN>(Bs[0], Xf, rowB0, M, tid);
    cp_async_commit();

    const int frow = lane >> 2, fk = lane & 3;
    // Diagonal blocks are computed in full: skipping their above-diagonal sub-tiles was measured
    // (r1) to be ~2 % SLOWER -- the CTA's duration is set by its fully-populated warp tile.
    for (int c = 0; c < nchunks; ++c) {
        const int st = c & 1;
        if (c + 1 < nchunks) {
            const cx<TIN> *Xn = Xf + int64_t(c + 1) * chunk_stride;
            load_tile<TIN>(As[st ^ 1], Xn, rowA0, M, tid);
            if (!diag) load_tile<TIN>(Bs[st ^ 1], Xn, rowB0, M, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const unsigned char *At = As[st] + (wj * 32 + frow) * SM::PITCH + fk * SM::ELEM;
        const unsigned char *Bt = Bs[st] + (wi * 32 + frow) * SM::PITCH + fk * SM::ELEM;
#pragma unroll
        for (int ks = 0; ks < KC / 4; ++ks) {
            double2 a[4], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a[q] = load_frag<TIN>(At + q * 8 * SM::PITCH + ks * 4 * SM::ELEM);
                b[q] = load_frag<TIN>(Bt + q * 8 * SM::PITCH + ks * 4 * SM::ELEM);
            }
#pragma unroll
            for (int jb = 0; jb < 4; ++jb)
#pragma unroll
                for (int ib = 0; ib < 4; ++ib) {
                    dmma(sr[jb][ib][0], sr[jb][ib][1], a[jb].x, b[ib].x);
                    dmma(sr[jb][ib][0], sr[jb][ib][1], a[jb].y, b[ib].y);
                    dmma(si[jb][ib][0], si[jb][ib][1], a[jb].y, b[ib].x);
                    dmma(si[jb][ib][0], si[jb][ib][1], a[jb].x, -b[ib].y);
                }
        }
        __syncthreads();
    }

    // ---- epilogue: C fragment (row = lane/4, cols 2*(lane%4)+{0,1}) ----
    double2 *Sf = S + int64_t(f) * (int64_t(M) * (M + 1) / 2);      // packed lower triangle of this bin
#pragma unroll
    for (int jb = 0; jb < 4; ++jb) {
        const int j = rowA0 + wj * 32 + jb * 8 + (lane >> 2);
        if (j >= M) continue;
        double2 *Srow = Sf + int64_t(j) * (j + 1) / 2;
#pragma unroll
        for (int ib = 0; ib < 4; ++ib) {
            const int i0 = rowB0 + wi * 32 + ib * 8 + 2 * (lane & 3);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = i0 + e;
                if (i < M && i <= j) {          // rows >= cols only: the rest is implied by Hermitian symmetry
                    double2 v = make_double2(sr[jb][ib][e], si[jb][ib][e]);
                    double2 *dst = Srow + i;
                    if (accumulate) { double2 o = *dst; v.x += o.x; v.y += o.y; }
                    *dst = v;
                }
            }
        }
    }
}

// Principal complex square root (numpy's np.sqrt on complex128, _spectral_methods.py:170-171).
__device__ __forceinline__ double2 csqrt_principal(double2 z) {
    const double a = z.x, b = z.y;
    if (a == 0.0 && b == 0.0) return make_double2(0.0, b);
    const double r = hypot(a, b);
    if (a >= 0.0) {
        const double t = sqrt(0.5 * (a + r));
        return make_double2(t, b / (2.0 * t));
    }
    const double t = sqrt(0.5 * (r - a));
    return make_double2(fabs(b) / (2.0 * t), copysign(t, b));
}

// K2e: one 32x32 tile of the lower triangle per CTA (read from the packed sums); mirrored tile written
// through smem.
__global__ void __launch_bounds__(256)
csm_finalize_kernel(const double2 *__restrict__ S, int M, int nb, double inv_frames, double factor,
                    int halve_edges, int amplitude_sqrt, double2 *__restrict__ out) {
    __shared__ double2 tile[32][33];
    const int p = blockIdx.x;
    int tj = int((sqrtf(8.0f * float(p) + 1.0f) - 1.0f) * 0.5f);
    while ((tj + 1) * (tj + 2) / 2 <= p) ++tj;
    while (tj * (tj + 1) / 2 > p) --tj;
    const int ti = p - tj * (tj + 1) / 2;
    const int f = blockIdx.y;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const bool edge = halve_edges && (f == 0 || f == nb - 1);
    const double2 *Sf = S + int64_t(f) * (int64_t(M) * (M + 1) / 2);
    double2 *of = out + int64_t(f) * M * M;
#pragma unroll
    for (int r = 0; r < 32; r += 8) {
        const int j = tj * 32 + ty + r, i = ti * 32 + tx;
        double2 v = make_double2(0.0, 0.0);
        if (j < M && i < M && i <= j) {
            v = Sf[int64_t(j) * (j + 1) / 2 + i];
            // order of the reference: mean, *factor, /2 on the edge bins, sqrt
            v.x *= inv_frames; v.y *= inv_frames;
            v.x *= factor; v.y *= factor;
            if (edge) { v.x *= 0.5; v.y *= 0.5; }
            if (i == j) v.y = 0.0;
            if (amplitude_sqrt) v = (i == j) ? make_double2(sqrt(v.x), 0.0) : csqrt_principal(v);
            of[int64_t(j) * M + i] = v;
        }
        tile[ty + r][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 32; r += 8) {
        // mirrored element: row i = ti*32 + (ty+r), col j = tj*32 + tx, strictly above the diagonal
        const int i = ti * 32 + ty + r, j = tj * 32 + tx;
        if (i < M && j < M && i < j) {
            double2 v = tile[tx][ty + r];
            of[int64_t(i) * M + j] = make_double2(v.x, -v.y);
        }
    }
}

// Auto-spectra only (the CSM diagonal): P[f][m] (+)= sum_k |X_f[m,k]|^2.  Replaces `_welch(x, None)`
// (standard/_spectral_methods.py:146-152) for Signal.get_spectrum's Welch branch, where the full
// rank-K update would be M times too much work.  One warp per pair of channels of one bin and 16 frames.
template <typename TIN>
__global__ void __launch_bounds__(256)
autopower_kernel(const cx<TIN> *__restrict__ X, int M, int64_t Kp, int nb, double *__restrict__ P, int accumulate) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pairs = (M + 1) / 2;
    const int64_t item = int64_t(blockIdx.x) * 8 + warp;            // (bin, channel pair)
    if (item >= int64_t(nb) * pairs) return;
    const int f = int(item / pairs), m = 2 * int(item % pairs) + (lane >> 4);
    const int nchunks = int(Kp / KC);
    double acc = 0.0;
    if (m < M) {
        const int slot = b2_bin_slot(f, nb - 1);
        const cx<TIN> *src = X + (int64_t(slot >> 2) * M + m) * B2_GROUP_ELEMS + b2_slot_frame_offset(slot, lane & 15);
        const int64_t chunk_stride = int64_t(b2_slot_groups(nb - 1)) * M * B2_GROUP_ELEMS;
        for (int c = 0; c < nchunks; ++c) {
            const cx<TIN> v = src[int64_t(c) * chunk_stride];
            acc += double(v.x) * double(v.x) + double(v.y) * double(v.y);
        }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((lane & 15) == 0 && m < M) {
        double *dst = P + int64_t(f) * M + m;
        *dst = accumulate ? *dst + acc : acc;
    }
}

// mean, factor, DC/Nyquist halving, sqrt for amplitude scalings (_spectral_methods.py:152,164-171)
__global__ void auto_finalize_kernel(const double *__restrict__ P, int M, int nb, double inv_frames, double factor,
                                     int halve_edges, int amplitude_sqrt, double *__restrict__ out) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= int64_t(nb) * M) return;
    const int f = int(i / M);
    double v = P[i] * inv_frames;
    v *= factor;
    if (halve_edges && (f == 0 || f == nb - 1)) v *= 0.5;
    out[i] = amplitude_sqrt ? sqrt(v) : v;
}

// CSM of ONE spectrum per channel (no averaging): csm[f][j][i] = X[f][j] conj(X[f][i]), the whole Hermitian matrix
// written directly -- `_csm_fft` (_spectral_methods.py:374-443).  Scalings other than FFTBackward: DC and last bin
// halved, times the conversion factor, element-wise principal sqrt for the amplitude scalings (:430-443).
__global__ void csm_rank1_kernel(const double2 *__restrict__ X, int M, int64_t nb, double factor, int halve_edges,
                                 int amplitude_sqrt, double2 *__restrict__ out) {
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t mm = int64_t(M) * M;
    if (idx >= nb * mm) return;
    const int64_t f = idx / mm;
    const int r = int(idx - f * mm);
    const int j = r / M, i = r - j * M;
    // The reference fills the lower triangle with conj(X_i) X_j and adds the conjugate transpose of that
    // (upper triangle still zero) -- which decides the SIGN OF A ZERO imaginary part (real DC / Nyquist bins), and
    // with it the branch the element-wise square root of the amplitude scalings takes for a negative real product:
    // lower: the product's own zero; upper: 0 + (-imag) = +0.  Evaluated the same way here.
    const int lo = j > i ? i : j, hi = j > i ? j : i;
    const double2 a = X[f * M + hi], b = X[f * M + lo];
    double2 v = make_double2(b.x * a.x + b.y * a.y, __dadd_rn(__dmul_rn(b.x, a.y), __dmul_rn(-b.y, a.x)));   // conj(b) a
    if (j < i) v.y = __dadd_rn(0.0, -v.y);
    if (i == j) v.y = 0.0;
    if (halve_edges) {
        // numpy divides / multiplies the complex matrix by a real number as by the complex number (x + 0i): for finite
        // values that only touches zero imaginary parts -- (vi - vr * 0) / 2 turns a -0 behind a negative real part
        // into +0 -- which is exactly what decides the branch of the square root below
        if (f == 0 || f == nb - 1) {
            v.y = __dmul_rn(__dsub_rn(v.y, __dmul_rn(v.x, 0.0)), 0.5);
            v.x *= 0.5;
        }
        v.y = __dadd_rn(__dmul_rn(v.x, 0.0), __dmul_rn(v.y, factor));
        v.x *= factor;
    }
    if (i == j) v.y = 0.0;
    if (amplitude_sqrt) v = (i == j) ? make_double2(sqrt(v.x), 0.0) : csqrt_principal(v);
    out[idx] = v;
}

template <typename TIN>
int launch_herk(const void *X, int M, int64_t Kp, int nb, int bin0, int nbins, void *S, int accumulate, cudaStream_t st) {
    using SM = HerkSmem<TIN>;
    auto kern = herk_dmma_kernel<TIN>;
    static DeviceOnce once;          // per instantiation and device
    if (int rc = b2_device_once(once, kern, SM::BYTES, nullptr)) return rc;
    const int nbk = (M + BM - 1) / BM;
    dim3 grid((unsigned)(nbk * (nbk + 1) / 2), (unsigned)nbins);
    kern<<<grid, HERK_THREADS, SM::BYTES, st>>>((const cx<TIN> *)X, M, Kp, nb, bin0, (double2 *)S, accumulate);
    B2_LAUNCH_OK("b2_csm_herk");
    return B2_OK;
}

}  // namespace

extern "C" int b2_csm_herk(const void *X_dev, int M, int64_t Kp, int nb, int bin0, int nbins, void *S_dev,
                           int accumulate, int precision, void *stream) {
    B2_REQUIRE(X_dev && S_dev, B2_EINVAL, "b2_csm_herk: null pointer");
    B2_REQUIRE(M > 0 && nb >= 5 && (nb - 1) % 4 == 0, B2_EINVAL, "b2_csm_herk: bad M=%d / nb=%d (nb = wl/2 + 1, wl a power of two >= 8)", M, nb);
    B2_REQUIRE(bin0 >= 0 && nbins >= 0 && bin0 + nbins <= nb && nbins <= 65535, B2_EINVAL,
               "b2_csm_herk: bin range [%d, %d) outside [0, %d) (at most 65535 bins per call)", bin0, bin0 + nbins, nb);
    B2_REQUIRE(Kp > 0 && Kp % KC == 0, B2_EINVAL, "b2_csm_herk: Kp=%lld must be a positive multiple of %d", (long long)Kp, KC);
    if (nbins == 0) return B2_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == B2_F64) return launch_herk<double>(X_dev, M, Kp, nb, bin0, nbins, S_dev, accumulate, st);
    if (precision == B2_F32) return launch_herk<float>(X_dev, M, Kp, nb, bin0, nbins, S_dev, accumulate, st);
    b2_set_error("b2_csm_herk: bad precision %d", precision);
    return B2_EINVAL;
}

extern "C" int b2_csm_finalize(const void *S_dev, int M, int nb, double inv_frames, double factor,
                               int halve_edges, int amplitude_sqrt, void *csm_out_dev, void *stream) {
    B2_REQUIRE(S_dev && csm_out_dev, B2_EINVAL, "b2_csm_finalize: null pointer");
    B2_REQUIRE(M > 0 && nb > 0 && nb <= 65535, B2_EINVAL, "b2_csm_finalize: bad M=%d / nb=%d", M, nb);
    const int nt = (M + 31) / 32;
    dim3 grid((unsigned)(nt * (nt + 1) / 2), (unsigned)nb);
    csm_finalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const double2 *)S_dev, M, nb, inv_frames, factor,
                                                                halve_edges, amplitude_sqrt, (double2 *)csm_out_dev);
    B2_LAUNCH_OK("b2_csm_finalize");
    return B2_OK;
}

extern "C" int b2_welch_autopower(const void *X_dev, int M, int64_t Kp, int nb, void *P_dev, int accumulate,
                                  int precision, void *stream) {
    B2_REQUIRE(X_dev && P_dev, B2_EINVAL, "b2_welch_autopower: null pointer");
    B2_REQUIRE(M > 0 && nb >= 5 && (nb - 1) % 4 == 0, B2_EINVAL, "b2_welch_autopower: bad M=%d / nb=%d (nb = wl/2 + 1, wl a power of two >= 8)", M, nb);
    B2_REQUIRE(Kp > 0 && Kp % KC == 0, B2_EINVAL, "b2_welch_autopower: Kp=%lld must be a positive multiple of %d", (long long)Kp, KC);
    const int64_t items = int64_t(nb) * ((M + 1) / 2);
    const unsigned blocks = (unsigned)((items + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == B2_F64) autopower_kernel<double><<<blocks, 256, 0, st>>>((const double2 *)X_dev, M, Kp, nb, (double *)P_dev, accumulate);
    else if (precision == B2_F32) autopower_kernel<float><<<blocks, 256, 0, st>>>((const float2 *)X_dev, M, Kp, nb, (double *)P_dev, accumulate);
    else { b2_set_error("b2_welch_autopower: bad precision %d", precision); return B2_EINVAL; }
    B2_LAUNCH_OK("b2_welch_autopower");
    return B2_OK;
}

extern "C" int b2_auto_finalize(const void *P_dev, int M, int nb, double inv_frames, double factor, int halve_edges,
                                int amplitude_sqrt, void *out_dev, void *stream) {
    B2_REQUIRE(P_dev && out_dev, B2_EINVAL, "b2_auto_finalize: null pointer");
    B2_REQUIRE(M > 0 && nb > 0, B2_EINVAL, "b2_auto_finalize: bad M=%d / nb=%d", M, nb);
    const int64_t n = int64_t(nb) * M;
    auto_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const double *)P_dev, M, nb, inv_frames, factor, halve_edges, amplitude_sqrt, (double *)out_dev);
    B2_LAUNCH_OK("b2_auto_finalize");
    return B2_OK;
}

extern "C" int b2_csm_rank1(const void *spectrum_dev, int M, int64_t nb, double factor, int halve_edges,
                            int amplitude_sqrt, void *csm_out_dev, void *stream) {
    B2_REQUIRE(spectrum_dev && csm_out_dev, B2_EINVAL, "b2_csm_rank1: null pointer");
    B2_REQUIRE(M > 0 && nb > 0, B2_EINVAL, "b2_csm_rank1: bad M=%d / nb=%lld", M, (long long)nb);
    const int64_t n = nb * int64_t(M) * M;
    B2_REQUIRE((n + 255) / 256 <= 0x7fffffff, B2_EINVAL, "b2_csm_rank1: matrix too large for one launch");
    csm_rank1_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const double2 *)spectrum_dev, M, nb, factor, halve_edges, amplitude_sqrt, (double2 *)csm_out_dev);
    B2_LAUNCH_OK("b2_csm_rank1");
    return B2_OK;
}
