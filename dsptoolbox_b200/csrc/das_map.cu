// K3a (operand preparation), K3 (delay-and-sum map, CUDA-core variants) and the dense
// steering-vector kernel.
//
// Replaces BeamformerDASFrequency.get_beamformer_map (beamforming/beamforming.py:841-878):
// the reference materialises h (F, M, G) complex128 (:858) and calls np.linalg.multi_dot per
// (grid point, bin) (:863-868).  Here h is synthesised on the fly and never leaves the SM.
//
// Algebra: with h_m = a_m exp(-i k r_m) (the common phase exp(+i k r_0) of :1515-1703 cancels
// in the quadratic form) and C Hermitian,
//   P = Re(h^H C h) = Re sum_{m>=n} conj(h_m) L[m,n] h_n,   L = 2*tril(C,-1) + diag(C),
// so only the lower triangle is touched.  The CUDA-core kernels below are the strict
// (fp64) path and the fall-back for shapes the tensor-core kernel does not take.
#include "common.cuh"
#include "das_common.cuh"

namespace {

// ---------------------------------------------------------------- K3a
template <typename T>
__global__ void das_prepare_simt_kernel(const double2 *__restrict__ csm, int M, int Mp, int bin0,
                                        double gain, int remove_diag, cx<T> *__restrict__ op) {
    const int f = blockIdx.z, m = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Mp) return;
    cx<T> v = mk<T>(T(0), T(0));
    if (m < M && n < M && n <= m) {
        double2 c = csm[(int64_t(bin0 + f) * M + m) * M + n];
        if (n == m) {
            if (!remove_diag) v = mk<T>(T(c.x * gain), T(0));     // diagonal of a Hermitian matrix is real
        } else {
            v = mk<T>(T(2.0 * gain * c.x), T(2.0 * gain * c.y));
        }
    }
    op[(int64_t(f) * Mp + m) * Mp + n] = v;
}

// ---------------------------------------------------------------- K3 (CUDA cores)
template <typename T> struct SimtCfg {
    static constexpr int TG = sizeof(T) == 8 ? 16 : 32;   // grid points per CTA
    static constexpr int GT = 8;                          // grid points per thread (register tile)
    static constexpr int THREADS = 256;
};

template <typename T>
__global__ void __launch_bounds__(SimtCfg<T>::THREADS)
das_map_simt_kernel(const cx<T> *__restrict__ op, int M, int Mp, int nbins,
                    const double *__restrict__ freqs, const double *__restrict__ weights, double inv_c,
                    const double *__restrict__ mic, double cx0, double cy0, double cz0,
                    const double *__restrict__ grid, int64_t G, int steering, int clamp_negative,
                    double *__restrict__ map_out, double *__restrict__ per_bin) {
    using CF = SimtCfg<T>;
    constexpr int TG = CF::TG, GT = CF::GT, NTHR = CF::THREADS, NW = NTHR / 32;
    extern __shared__ __align__(16) unsigned char smem[];
    cx<T> *H = reinterpret_cast<cx<T> *>(smem);                       // [TG][Mp]
    T *roc = reinterpret_cast<T *>(H + TG * Mp);                      // [TG][Mp]  r / c  (seconds)
    T *amp = roc + TG * Mp;                                           // [TG][Mp]
    double *part = reinterpret_cast<double *>(amp + TG * Mp);         // [TG][NG]
    const int NG = Mp / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t g0 = int64_t(blockIdx.x) * TG;

    // ---- geometry: one warp per grid point, lanes over microphones ----
    for (int gl = warp; gl < TG; gl += NW) {
        const int64_t g = g0 + gl;
        double gx = 0, gy = 0, gz = 0;
        const bool gvalid = g < G;
        if (gvalid) { gx = grid[3 * g]; gy = grid[3 * g + 1]; gz = grid[3 * g + 2]; }
        double inv_r2_sum = 0.0;
        for (int m = lane; m < Mp; m += 32) {
            double r = 0.0;
            if (m < M && gvalid) {
                r = das_distance(gx, gy, gz, mic[3 * m], mic[3 * m + 1], mic[3 * m + 2]);
                inv_r2_sum += 1.0 / (r * r);
            }
            roc[gl * Mp + m] = T(r * inv_c);
            amp[gl * Mp + m] = T(r);          // temporarily the distance itself
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) inv_r2_sum += __shfl_xor_sync(0xffffffffu, inv_r2_sum, o);
        const double r0 = gvalid ? das_distance(gx, gy, gz, cx0, cy0, cz0) : 1.0;
        for (int m = lane; m < Mp; m += 32) {
            double a = 0.0;
            if (m < M && gvalid) {
                const double r = das_distance(gx, gy, gz, mic[3 * m], mic[3 * m + 1], mic[3 * m + 2]);
                a = das_amplitude(steering, r, r0, inv_r2_sum, M);
            }
            amp[gl * Mp + m] = T(a);
        }
    }
    __syncthreads();

    double acc_map = 0.0;      // thread tid < TG owns grid point tid
    for (int fi = 0; fi < nbins; ++fi) {
        const double freq = freqs[fi];
        // ---- steering values of this bin for the whole tile ----
        for (int e = tid; e < TG * Mp; e += NTHR) {
            const double cyc = freq * double(roc[e]);
            const double fr = cyc - rint(cyc);
            T sn, cs;
            das_sincospi(T(2.0 * fr), &sn, &cs);
            const T a = amp[e];
            H[e] = mk<T>(a * cs, -a * sn);            // a * exp(-i k r)
        }
        __syncthreads();

        // ---- P[g] = Re sum_{m>=n} conj(H[g,m]) L[m,n] H[g,n] ----
        const cx<T> *Lf = op + int64_t(fi) * Mp * Mp;
        const int ncombo = NG * (TG / GT);
        for (int cb = warp; cb < ncombo; cb += NW) {
            const int ng = cb % NG, gg = cb / NG;
            const int n = ng * 32 + lane;
            cx<T> t[GT];
#pragma unroll
            for (int q = 0; q < GT; ++q) t[q] = mk<T>(T(0), T(0));
            const cx<T> *Hg = H + gg * GT * Mp;
            for (int m = ng * 32; m < M; ++m) {
                const cx<T> l = __ldg(Lf + int64_t(m) * Mp + n);
#pragma unroll
                for (int q = 0; q < GT; ++q) {
                    const cx<T> h = Hg[q * Mp + m];
                    // conj(h) * l
                    t[q].x += h.x * l.x + h.y * l.y;
                    t[q].y += h.x * l.y - h.y * l.x;
                }
            }
#pragma unroll
            for (int q = 0; q < GT; ++q) {
                const cx<T> h = Hg[q * Mp + n];
                double p = double(t[q].x * h.x - t[q].y * h.y);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
                if (lane == 0) part[(gg * GT + q) * NG + ng] = p;
            }
        }
        __syncthreads();
        if (tid < TG) {
            double p = 0.0;
            for (int ng = 0; ng < NG; ++ng) p += part[tid * NG + ng];
            if (clamp_negative && p < 0.0) p = 0.0;                 // beamforming.py:871-872
            acc_map += weights[fi] * p;                               // Simpson weights, :875-876
            if (per_bin && g0 + tid < G) per_bin[(g0 + tid) * nbins + fi] = p;
        }
        // H is rewritten next iteration; `part` is read above by tid < TG only after the sync
        __syncthreads();
    }
    if (tid < TG && g0 + tid < G) map_out[g0 + tid] = acc_map;
}

template <typename T>
int launch_simt(const void *op, int M, int nbins, const void *freqs, const void *weights, double c,
                const void *mic, const double *center, const void *grid, int64_t G, int steering,
                int clamp_negative, void *map_out, void *per_bin, cudaStream_t st) {
    using CF = SimtCfg<T>;
    const int Mp = das_padded_mics(M);
    const size_t smem = size_t(CF::TG) * Mp * (sizeof(cx<T>) + 2 * sizeof(T)) + size_t(CF::TG) * (Mp / 32) * sizeof(double);
    B2_REQUIRE(smem <= 227 * 1024, B2_EUNSUPPORTED, "b2_das_map: %d microphones exceed the CUDA-core kernel's shared memory", M);
    auto kern = das_map_simt_kernel<T>;
    B2_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t nblk = (G + CF::TG - 1) / CF::TG;
    B2_REQUIRE(nblk <= 0x7fffffff, B2_EINVAL, "b2_das_map: too many grid points");
    kern<<<(unsigned)nblk, CF::THREADS, smem, st>>>((const cx<T> *)op, M, Mp, nbins, (const double *)freqs,
                                                    (const double *)weights, 1.0 / c, (const double *)mic,
                                                    center[0], center[1], center[2], (const double *)grid, G,
                                                    steering, clamp_negative, (double *)map_out, (double *)per_bin);
    B2_LAUNCH_OK("b2_das_map");
    return B2_OK;
}

// ---------------------------------------------------------------- dense steering vectors
__global__ void steering_vector_kernel(const double *__restrict__ k, int nbins, const double *__restrict__ mic,
                                       int M, double cx0, double cy0, double cz0,
                                       const double *__restrict__ grid, int64_t G, int steering,
                                       double2 *__restrict__ h) {
    const int64_t g = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (g >= G) return;
    const double gx = grid[3 * g], gy = grid[3 * g + 1], gz = grid[3 * g + 2];
    double s = 0.0;
    for (int j = 0; j < M; ++j) {
        const double rj = das_distance(gx, gy, gz, mic[3 * j], mic[3 * j + 1], mic[3 * j + 2]);
        s += 1.0 / (rj * rj);
    }
    const double r = das_distance(gx, gy, gz, mic[3 * m], mic[3 * m + 1], mic[3 * m + 2]);
    const double r0 = das_distance(gx, gy, gz, cx0, cy0, cz0);
    const double a = das_amplitude(steering, r, r0, s, M);
    for (int f = 0; f < nbins; ++f) {
        double sn, cs;
        sincos(k[f] * (r - r0), &sn, &cs);
        h[(int64_t(f) * M + m) * G + g] = make_double2(a * cs, -a * sn);
    }
}

}  // namespace

// tensor-core variants live in das_map_tc.cu
int64_t b2_das_tc_operand_bytes(int M, int nbins, int mode);
int b2_das_tc_prepare(const void *csm_dev, int M, int nb, int bin0, int nbins, int remove_diag,
                      void *op_dev, int mode, cudaStream_t st);
int b2_das_tc_map(const void *op_dev, int M, int nbins, const void *freqs_dev, const void *weights_dev,
                  double c, const void *mic_xyz_dev, const double *center_xyz, const void *grid_xyz_dev,
                  int64_t G, int steering, int clamp_negative, void *map_out_dev, void *per_bin_out_dev,
                  int mode, cudaStream_t st);

extern "C" int64_t b2_das_operand_bytes(int M, int nbins, int mode) {
    if (M <= 0 || nbins <= 0) { b2_set_error("b2_das_operand_bytes: bad sizes"); return B2_EINVAL; }
    const int64_t Mp = das_padded_mics(M);
    if (mode == B2_MAP_SIMT_F64) return int64_t(nbins) * Mp * Mp * 16;
    if (mode == B2_MAP_SIMT_F32) return int64_t(nbins) * Mp * Mp * 8;
    if (mode >= B2_MAP_TC_TF32 && mode <= B2_MAP_TC_F16X3) return b2_das_tc_operand_bytes(M, nbins, mode);
    b2_set_error("b2_das_operand_bytes: bad mode %d", mode);
    return B2_EINVAL;
}

extern "C" int b2_das_prepare(const void *csm_dev, int M, int nb, int bin0, int nbins, int remove_diag,
                              void *op_dev, int mode, void *stream) {
    B2_REQUIRE(csm_dev && op_dev, B2_EINVAL, "b2_das_prepare: null pointer");
    B2_REQUIRE(M > 1, B2_EINVAL, "b2_das_prepare: at least two microphones are needed (got %d)", M);
    B2_REQUIRE(bin0 >= 0 && nbins > 0 && bin0 + nbins <= nb, B2_EINVAL,
               "b2_das_prepare: bins [%d, %d) outside [0, %d)", bin0, bin0 + nbins, nb);
    B2_REQUIRE(nbins <= 65535, B2_EINVAL, "b2_das_prepare: too many bins");
    cudaStream_t st = (cudaStream_t)stream;
    if (mode >= B2_MAP_TC_TF32 && mode <= B2_MAP_TC_F16X3)
        return b2_das_tc_prepare(csm_dev, M, nb, bin0, nbins, remove_diag, op_dev, mode, st);
    const int Mp = das_padded_mics(M);
    const double gain = remove_diag ? double(M) / double(M - 1) : 1.0;   // beamforming.py:843
    dim3 grid((Mp + 127) / 128, Mp, nbins);
    if (mode == B2_MAP_SIMT_F64)
        das_prepare_simt_kernel<double><<<grid, 128, 0, st>>>((const double2 *)csm_dev, M, Mp, bin0, gain, remove_diag, (double2 *)op_dev);
    else if (mode == B2_MAP_SIMT_F32)
        das_prepare_simt_kernel<float><<<grid, 128, 0, st>>>((const double2 *)csm_dev, M, Mp, bin0, gain, remove_diag, (float2 *)op_dev);
    else { b2_set_error("b2_das_prepare: bad mode %d", mode); return B2_EINVAL; }
    B2_LAUNCH_OK("b2_das_prepare");
    return B2_OK;
}

extern "C" int b2_das_map(const void *op_dev, int M, int nbins, const void *freqs_dev,
                          const void *weights_dev, double c, const void *mic_xyz_dev,
                          const double *center_xyz, const void *grid_xyz_dev, int64_t G, int steering,
                          int clamp_negative, void *map_out_dev, void *per_bin_out_dev, int mode,
                          void *stream) {
    B2_REQUIRE(op_dev && freqs_dev && weights_dev && mic_xyz_dev && center_xyz && grid_xyz_dev && map_out_dev,
               B2_EINVAL, "b2_das_map: null pointer");
    B2_REQUIRE(M > 1 && nbins > 0 && G > 0, B2_EINVAL, "b2_das_map: bad sizes M=%d nbins=%d G=%lld", M, nbins, (long long)G);
    B2_REQUIRE(c > 0, B2_EINVAL, "b2_das_map: Speed of sound should be bigger than 0");
    B2_REQUIRE(steering >= B2_STEER_CLASSIC && steering <= B2_STEER_TRUE_LOCATION, B2_EINVAL,
               "b2_das_map: Incorrect formulation %d", steering);
    cudaStream_t st = (cudaStream_t)stream;
    switch (mode) {
    case B2_MAP_SIMT_F64:
        return launch_simt<double>(op_dev, M, nbins, freqs_dev, weights_dev, c, mic_xyz_dev, center_xyz, grid_xyz_dev, G, steering, clamp_negative, map_out_dev, per_bin_out_dev, st);
    case B2_MAP_SIMT_F32:
        return launch_simt<float>(op_dev, M, nbins, freqs_dev, weights_dev, c, mic_xyz_dev, center_xyz, grid_xyz_dev, G, steering, clamp_negative, map_out_dev, per_bin_out_dev, st);
    case B2_MAP_TC_TF32:
    case B2_MAP_TC_TF32X3:
    case B2_MAP_TC_MIXED:
    case B2_MAP_TC_F16X3:
        return b2_das_tc_map(op_dev, M, nbins, freqs_dev, weights_dev, c, mic_xyz_dev, center_xyz, grid_xyz_dev, G, steering, clamp_negative, map_out_dev, per_bin_out_dev, mode, st);
    default: break;
    }
    b2_set_error("b2_das_map: bad mode %d", mode);
    return B2_EINVAL;
}

extern "C" int b2_steering_vector(const void *wave_numbers_dev, int nbins, const void *mic_xyz_dev, int M,
                                  const double *center_xyz, const void *grid_xyz_dev, int64_t G,
                                  int steering, void *h_out_dev, void *stream) {
    B2_REQUIRE(wave_numbers_dev && mic_xyz_dev && center_xyz && grid_xyz_dev && h_out_dev, B2_EINVAL,
               "b2_steering_vector: null pointer");
    B2_REQUIRE(M > 0 && M <= 65535 && nbins > 0 && G > 0, B2_EINVAL, "b2_steering_vector: bad sizes");
    B2_REQUIRE(steering >= B2_STEER_CLASSIC && steering <= B2_STEER_TRUE_LOCATION, B2_EINVAL,
               "b2_steering_vector: Incorrect formulation %d", steering);
    dim3 grid((unsigned)((G + 127) / 128), (unsigned)M);
    steering_vector_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(
        (const double *)wave_numbers_dev, nbins, (const double *)mic_xyz_dev, M, center_xyz[0], center_xyz[1],
        center_xyz[2], (const double *)grid_xyz_dev, G, steering, (double2 *)h_out_dev);
    B2_LAUNCH_OK("b2_steering_vector");
    return B2_OK;
}
