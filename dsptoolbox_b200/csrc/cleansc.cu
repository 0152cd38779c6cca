// CLEAN-SC iteration kernels (float64): the per-bin source component and its subtraction from the dirty map.
//
// Replaces the body of `_clean_sc_deconvolve` (beamforming/_beamforming.py:246-295), which the reference runs bin
// after bin; here all bins of the band iterate in lock-step and one iteration is two launches:
//   b2_cleansc_component   stopping criterion (matrix 1-norms of the degraded CSM, :259), steering vector to the peak
//                          point, D w / P_max, the 20 fixed-point steps for h (:274-278), G = h h^H P_max (:280-283)
//                          and the update of the degraded CSM (:291-294) -- one CTA per bin;
//   b2_das_rank1_update    map[g] -= Re(h_g^H G h_g) * safety_factor for every grid point (:285-289).  G has rank one,
//                          so Re(h_g^H G h_g) = P_max (|h_g^H h|^2 - [diagonal removed] sum_m |h_gm|^2 |h_m|^2): O(G F M)
//                          instead of the O(G F M^2) of the general quadratic form (the general fp64 kernel took 80 % of
//                          the GPU time of a CLEAN-SC map, profiles/r2_cleansc_profile.txt).
// The steering values of the map update are synthesised exactly as in das_map_simt_kernel<double> (das_map.cu).
#include "common.cuh"
#include "das_common.cuh"

namespace {

constexpr int CSC_THREADS = 256;

__device__ double csc_block_sum(double v, double *scratch) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < CSC_THREADS / 32; ++w) r += scratch[w];
    return r;
}
__device__ double csc_block_max(double v, double *scratch) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = scratch[0];
    for (int w = 1; w < CSC_THREADS / 32; ++w) r = fmax(r, scratch[w]);
    return r;
}

// largest column sum of |D| (np.linalg.norm(D, ord=1))
__device__ double csc_norm1(const double2 *__restrict__ D, int M, double *scratch) {
    double best = 0.0;
    for (int j = threadIdx.x; j < M; j += CSC_THREADS) {
        double s = 0.0;
        for (int i = 0; i < M; ++i) { const double2 v = D[int64_t(i) * M + j]; s += hypot(v.x, v.y); }
        best = fmax(best, s);
    }
    return csc_block_max(best, scratch);
}

__global__ void __launch_bounds__(CSC_THREADS)
cleansc_component_kernel(double2 *__restrict__ D0, double2 *__restrict__ D1, int M, const int64_t *__restrict__ peak,
                         const double *__restrict__ pmax, unsigned char *__restrict__ active,
                         const double *__restrict__ wave_numbers, const double *__restrict__ mic, double cx0, double cy0,
                         double cz0, const double *__restrict__ grid, int steering, double safety_factor, int remove_diag,
                         double2 *__restrict__ h_out, double *__restrict__ scale_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *w = reinterpret_cast<double2 *>(smem_raw);       // steering vector to the peak point
    double2 *d = w + M;                                       // D w / P_max
    double2 *h = d + M;
    double *scratch = reinterpret_cast<double *>(h + M);
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double2 *D0f = D0 + int64_t(f) * M * M, *D1f = D1 + int64_t(f) * M * M;
    double2 *hf = h_out + int64_t(f) * M;

    // ---- stopping criterion: the degraded CSM stopped shrinking (a finished bin stays finished)
    bool act = active[f] != 0;
    if (act) {
        const double n1 = csc_norm1(D1f, M, scratch), n0 = csc_norm1(D0f, M, scratch);
        act = !(n1 >= n0);
    }
    if (!act) {
        for (int m = tid; m < M; m += CSC_THREADS) hf[m] = make_double2(0.0, 0.0);
        if (tid == 0) { active[f] = 0; scale_out[f] = 0.0; }
        return;
    }
    const double pm = pmax[f];

    // ---- steering vector to the peak point (distances by the expansion formula of helpers/other.py:131-153, like
    //      the reference's h array; amplitudes of beamforming.py:1515-1703)
    const int64_t g = peak[f];
    const double gx = grid[3 * g], gy = grid[3 * g + 1], gz = grid[3 * g + 2];
    const double gg = gx * gx + gy * gy + gz * gz;
    auto dist = [&](double bx, double by, double bz) {
        return sqrt(gg + (bx * bx + by * by + bz * bz) - 2.0 * (gx * bx + gy * by + gz * bz));
    };
    const double rt0 = dist(cx0, cy0, cz0);
    double part = 0.0;
    for (int m = tid; m < M; m += CSC_THREADS) {
        const double r = dist(mic[3 * m], mic[3 * m + 1], mic[3 * m + 2]);
        part += 1.0 / (r * r);
    }
    const double rtj = csc_block_sum(part, scratch);
    const double k = wave_numbers[f];
    for (int m = tid; m < M; m += CSC_THREADS) {
        const double r = dist(mic[3 * m], mic[3 * m + 1], mic[3 * m + 2]);
        double a;
        switch (steering) {
        case B2_STEER_CLASSIC: a = 1.0 / double(M); break;
        case B2_STEER_INVERSE: a = r / double(M) / rt0; break;
        case B2_STEER_TRUE_POWER: a = 1.0 / rt0 / r / rtj; break;
        default: a = 1.0 / r / sqrt(double(M) * rtj); break;
        }
        double sn, cs;
        sincos(k * (r - rt0), &sn, &cs);
        w[m] = make_double2(a * cs, -a * sn);
    }
    __syncthreads();

    // ---- d = D1 w / P_max (one warp per row), h = w
    for (int m = warp; m < M; m += CSC_THREADS / 32) {
        double sx = 0.0, sy = 0.0;
        for (int n = lane; n < M; n += 32) {
            const double2 a = D1f[int64_t(m) * M + n], b = w[n];
            sx += a.x * b.x - a.y * b.y;
            sy += a.x * b.y + a.y * b.x;
        }
        for (int o = 16; o > 0; o >>= 1) { sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); }
        if (lane == 0) { d[m] = make_double2(sx / pm, sy / pm); h[m] = w[m]; }
    }
    __syncthreads();

    // ---- 20 fixed-point steps: H = |h|^2, h = (d + H w) / sqrt(1 + sum_m H_m |w_m|^2)
    for (int it = 0; it < 20; ++it) {
        double s = 0.0;
        for (int m = tid; m < M; m += CSC_THREADS) {
            const double H = h[m].x * h[m].x + h[m].y * h[m].y;
            s += H * (w[m].x * w[m].x + w[m].y * w[m].y);
        }
        const double den = sqrt(1.0 + csc_block_sum(s, scratch));
        for (int m = tid; m < M; m += CSC_THREADS) {
            const double H = h[m].x * h[m].x + h[m].y * h[m].y;
            h[m] = make_double2((d[m].x + H * w[m].x) / den, (d[m].y + H * w[m].y) / den);
        }
        __syncthreads();
    }

    // ---- G = h h^H P_max (diagonal removed on request); D0 <- D1, D1 <- D1 - safety_factor G
    for (int idx = tid; idx < M * M; idx += CSC_THREADS) {
        const int i = idx / M, j = idx - i * M;
        const double2 a = h[i], b = h[j];
        double2 gv = make_double2((a.x * b.x + a.y * b.y) * pm, (a.y * b.x - a.x * b.y) * pm);   // h_i conj(h_j) P_max
        if (remove_diag && i == j) gv = make_double2(0.0, 0.0);
        const double2 v = D1f[idx];
        D0f[idx] = v;
        D1f[idx] = make_double2(v.x - safety_factor * gv.x, v.y - safety_factor * gv.y);
    }
    for (int m = tid; m < M; m += CSC_THREADS) hf[m] = h[m];
    if (tid == 0) scale_out[f] = pm;
}

// out[g][f] += alpha * scale_f * (|sum_m conj(H_gm) v_fm|^2 - [remove_diag] sum_m |H_gm|^2 |v_fm|^2)
constexpr int R1_TG = 16;                 // grid points per CTA, as in das_map_simt_kernel<double>
__global__ void __launch_bounds__(256)
das_rank1_update_kernel(const double2 *__restrict__ vec, const double *__restrict__ scale, int M, int Mp, int nbins,
                        const double *__restrict__ freqs, double inv_c, const double *__restrict__ mic, double cx0,
                        double cy0, double cz0, const double *__restrict__ grid, int64_t G, int steering,
                        int remove_diag, double alpha, double *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2 *H = reinterpret_cast<double2 *>(smem);                    // [TG][Mp]
    double *roc = reinterpret_cast<double *>(H + R1_TG * Mp);          // [TG][Mp]  r / c  (seconds)
    double *amp = roc + R1_TG * Mp;                                    // [TG][Mp]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = 256 / 32;
    const int64_t g0 = int64_t(blockIdx.x) * R1_TG;

    // ---- geometry: one warp per grid point, lanes over microphones (same arithmetic as das_map_simt_kernel)
    for (int gl = warp; gl < R1_TG; gl += NW) {
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
            roc[gl * Mp + m] = r * inv_c;
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
            amp[gl * Mp + m] = a;
        }
    }
    __syncthreads();

    for (int fi = 0; fi < nbins; ++fi) {
        const double sc = scale[fi];
        if (sc == 0.0) continue;                                       // finished bin (uniform over the CTA)
        const double freq = freqs[fi];
        for (int e = tid; e < R1_TG * Mp; e += 256) {
            const double cyc = freq * roc[e];
            const double fr = cyc - rint(cyc);
            double sn, cs;
            sincospi(2.0 * fr, &sn, &cs);
            const double a = amp[e];
            H[e] = make_double2(a * cs, -a * sn);                      // a * exp(-i k r)
        }
        __syncthreads();
        const double2 *v = vec + int64_t(fi) * M;
        for (int gl = warp; gl < R1_TG; gl += NW) {
            double sx = 0.0, sy = 0.0, dg = 0.0;
            for (int m = lane; m < M; m += 32) {
                const double2 hv = H[gl * Mp + m], vv = v[m];
                sx += hv.x * vv.x + hv.y * vv.y;                       // conj(h) v
                sy += hv.x * vv.y - hv.y * vv.x;
                dg += (hv.x * hv.x + hv.y * hv.y) * (vv.x * vv.x + vv.y * vv.y);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sx += __shfl_xor_sync(0xffffffffu, sx, o);
                sy += __shfl_xor_sync(0xffffffffu, sy, o);
                dg += __shfl_xor_sync(0xffffffffu, dg, o);
            }
            if (lane == 0 && g0 + gl < G) {
                double p = sx * sx + sy * sy;
                if (remove_diag) p -= dg;
                out[(g0 + gl) * nbins + fi] += alpha * sc * p;
            }
        }
        __syncthreads();                                               // H is rewritten for the next bin
    }
}

}  // namespace

extern "C" int b2_cleansc_component(void *D0_dev, void *D1_dev, int M, int nbins, const void *peak_dev,
                                    const void *pmax_dev, void *active_dev, const void *wave_numbers_dev,
                                    const void *mic_xyz_dev, const double *center_xyz, const void *grid_xyz_dev,
                                    int steering, double safety_factor, int remove_diag, void *h_out_dev,
                                    void *scale_out_dev, void *stream) {
    B2_REQUIRE(D0_dev && D1_dev && peak_dev && pmax_dev && active_dev && wave_numbers_dev && mic_xyz_dev && center_xyz &&
               grid_xyz_dev && h_out_dev && scale_out_dev, B2_EINVAL, "b2_cleansc_component: null pointer");
    B2_REQUIRE(M > 1 && M <= 4096 && nbins > 0, B2_EINVAL, "b2_cleansc_component: bad sizes M=%d nbins=%d", M, nbins);
    B2_REQUIRE(steering >= B2_STEER_CLASSIC && steering <= B2_STEER_TRUE_LOCATION, B2_EINVAL,
               "b2_cleansc_component: Incorrect formulation %d", steering);
    const size_t smem = size_t(3) * M * sizeof(double2) + (CSC_THREADS / 32) * sizeof(double);
    static DeviceOnce once;
    if (int rc = b2_device_once(once, cleansc_component_kernel, int(3 * 4096 * sizeof(double2) + 64), nullptr)) return rc;
    cleansc_component_kernel<<<(unsigned)nbins, CSC_THREADS, smem, (cudaStream_t)stream>>>(
        (double2 *)D0_dev, (double2 *)D1_dev, M, (const int64_t *)peak_dev, (const double *)pmax_dev,
        (unsigned char *)active_dev, (const double *)wave_numbers_dev, (const double *)mic_xyz_dev, center_xyz[0],
        center_xyz[1], center_xyz[2], (const double *)grid_xyz_dev, steering, safety_factor, remove_diag,
        (double2 *)h_out_dev, (double *)scale_out_dev);
    B2_LAUNCH_OK("b2_cleansc_component");
    return B2_OK;
}

extern "C" int b2_das_rank1_update(const void *vectors_dev, const void *scale_dev, int M, int nbins,
                                   const void *freqs_dev, double c, const void *mic_xyz_dev, const double *center_xyz,
                                   const void *grid_xyz_dev, int64_t G, int steering, int remove_diag, double alpha,
                                   void *out_dev, void *stream) {
    B2_REQUIRE(vectors_dev && scale_dev && freqs_dev && mic_xyz_dev && center_xyz && grid_xyz_dev && out_dev, B2_EINVAL,
               "b2_das_rank1_update: null pointer");
    B2_REQUIRE(M > 1 && nbins > 0 && G > 0, B2_EINVAL, "b2_das_rank1_update: bad sizes M=%d nbins=%d G=%lld", M, nbins, (long long)G);
    B2_REQUIRE(c > 0, B2_EINVAL, "b2_das_rank1_update: Speed of sound should be bigger than 0");
    B2_REQUIRE(steering >= B2_STEER_CLASSIC && steering <= B2_STEER_TRUE_LOCATION, B2_EINVAL,
               "b2_das_rank1_update: Incorrect formulation %d", steering);
    const int Mp = das_padded_mics(M);
    const size_t smem = size_t(R1_TG) * Mp * (sizeof(double2) + 2 * sizeof(double));
    B2_REQUIRE(smem <= 227 * 1024, B2_EUNSUPPORTED, "b2_das_rank1_update: %d microphones exceed the kernel's shared memory", M);
    B2_CUDA_OK(cudaFuncSetAttribute(das_rank1_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t nblk = (G + R1_TG - 1) / R1_TG;
    B2_REQUIRE(nblk <= 0x7fffffff, B2_EINVAL, "b2_das_rank1_update: too many grid points");
    das_rank1_update_kernel<<<(unsigned)nblk, 256, smem, (cudaStream_t)stream>>>(
        (const double2 *)vectors_dev, (const double *)scale_dev, M, Mp, nbins, (const double *)freqs_dev, 1.0 / c,
        (const double *)mic_xyz_dev, center_xyz[0], center_xyz[1], center_xyz[2], (const double *)grid_xyz_dev, G,
        steering, remove_diag, alpha, (double *)out_dev);
    B2_LAUNCH_OK("b2_das_rank1_update");
    return B2_OK;
}
