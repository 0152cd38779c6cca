// Batched Hermitian eigen-decomposition of small matrices (one per frequency bin): parallel two-sided Jacobi in
// shared memory, one CTA per matrix.
//
// Replaces the per-bin LAPACK factorisations of the "other" gridded beamformers of the reference,
//   np.linalg.eigh(csm[f])   BeamformerOrthogonal  (beamforming/beamforming.py:1107)
//   np.linalg.svd(csm[f])    BeamformerFunctional  (beamforming/beamforming.py:1201; for a Hermitian matrix the
//                            singular triplets are (|lambda|, sign(lambda) v, v))
// which this package otherwise sends through torch.linalg (cuSOLVER: ~1 ms per 64 x 64 matrix, one after the other --
// 40 to 100 ms per map, far more than the map kernel itself).
//
// Algorithm: cyclic Jacobi with the round-robin ("tournament") ordering -- every step annihilates M/2 disjoint
// off-diagonal pairs (p, q) at once.  For a pair with a_pq = |a_pq| e^{i phi} the unitary rotation is
//   J = diag(1, e^{-i phi}) [[c, s], [-s, c]],   cot 2 theta = (a_qq - a_pp) / (2 |a_pq|),  t = tan theta (smaller root)
// (Numerical Recipes 11.1 after removing the phase).  A <- J^H A J is applied as a column phase (A J and V J; the
// columns of different pairs are disjoint) and a row phase (J^H A), separated by CTA barriers; the 2 x 2 pivot block
// is written exactly (a_pp - t |a_pq|, a_qq + t |a_pq|, 0).  Sweeps repeat until the largest off-diagonal magnitude is
// below M * 2^-52 of the largest diagonal magnitude (7 to 9 sweeps at M = 64; the convergence is quadratic).
// Rows are padded by one element so that the lanes of the column phase (consecutive rows) fall into distinct banks.
#include "common.cuh"

namespace {

constexpr int HEIG_THREADS = 256;
constexpr int HEIG_MAX_M = 80;            // A and V, (M + 1)-pitched complex128: 2 * 80 * 81 * 16 B = 207 KB
constexpr int HEIG_MAX_SWEEPS = 24;

struct Rot {                              // one pair of a step
    int p, q;                             // p < q; p < 0: nothing to do
    double cs, sn;                        // cos, sin
    double2 se, ce;                       // sin e^{-i phi}, cos e^{-i phi}
    double app, aqq;                      // the pivot block after the rotation
};

__device__ __forceinline__ double2 cmulr(double2 a, double r) { return make_double2(a.x * r, a.y * r); }

__device__ double block_max(double v, double *scratch) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();                                      // scratch may still be read from the previous call
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = scratch[0];
    for (int w = 1; w < HEIG_THREADS / 32; ++w) r = fmax(r, scratch[w]);
    return r;
}

__global__ void __launch_bounds__(HEIG_THREADS)
heig_jacobi_kernel(const double2 *__restrict__ Ain, int M, double *__restrict__ w_out, double2 *__restrict__ V_out,
                   int *__restrict__ sweeps_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int P = M + 1;                                  // row pitch in elements
    double2 *A = reinterpret_cast<double2 *>(smem_raw);
    double2 *V = A + M * P;
    Rot *rot = reinterpret_cast<Rot *>(V + M * P);
    double *scratch = reinterpret_cast<double *>(rot + (M + 1) / 2);
    const int tid = threadIdx.x;
    const int64_t f = blockIdx.x;
    const double2 *src = Ain + f * int64_t(M) * M;

    // the lower triangle defines the matrix (what the map kernel reads as well); the diagonal is real
    for (int idx = tid; idx < M * M; idx += HEIG_THREADS) {
        const int i = idx / M, j = idx - i * M;
        double2 v;
        if (i > j) v = src[idx];
        else if (i < j) v = cconj(src[j * M + i]);
        else v = make_double2(src[idx].x, 0.0);
        A[i * P + j] = v;
        V[i * P + j] = make_double2(i == j ? 1.0 : 0.0, 0.0);
    }
    __syncthreads();

    const int Me = M + (M & 1), half = Me / 2;            // an odd order plays with one dummy index (never paired)
    int sweep = 0;
    for (; sweep < HEIG_MAX_SWEEPS; ++sweep) {
        double off = 0.0, dmax = 0.0;
        for (int idx = tid; idx < M * M; idx += HEIG_THREADS) {
            const int i = idx / M, j = idx - i * M;
            const double2 v = A[i * P + j];
            const double m = fmax(fabs(v.x), fabs(v.y));
            if (i == j) dmax = fmax(dmax, m); else off = fmax(off, m);
        }
        off = block_max(off, scratch);
        dmax = block_max(dmax, scratch);
        if (off <= double(M) * 2.220446049250313e-16 * dmax) break;

        for (int s = 0; s < Me - 1; ++s) {
            // ---- rotation parameters of the Me / 2 pairs of this step
            if (tid < half) {
                int p, q;
                if (tid == 0) { p = Me - 1; q = s; }
                else { p = (s + tid) % (Me - 1); q = (s - tid + (Me - 1)) % (Me - 1); }
                if (p > q) { const int t = p; p = q; q = t; }
                Rot r;
                r.p = -1; r.q = q; r.cs = 1.0; r.sn = 0.0;
                r.se = make_double2(0.0, 0.0); r.ce = make_double2(1.0, 0.0); r.app = r.aqq = 0.0;
                if (q < M) {
                    const double a = A[p * P + p].x, b = A[q * P + q].x;
                    const double2 c = A[p * P + q];
                    const double ac = hypot(c.x, c.y);
                    if (ac > 1e-300 && ac > 1e-18 * (fabs(a) + fabs(b))) {
                        const double tau = (b - a) / (2.0 * ac);
                        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        const double cs = 1.0 / sqrt(1.0 + t * t), sn = t * cs;
                        const double2 e = make_double2(c.x / ac, -c.y / ac);      // e^{-i phi}
                        r.p = p; r.cs = cs; r.sn = sn;
                        r.se = cmulr(e, sn); r.ce = cmulr(e, cs);
                        r.app = a - t * ac; r.aqq = b + t * ac;
                    }
                }
                rot[tid] = r;
            }
            __syncthreads();
            // ---- columns: [x_p, x_q] <- [cs x_p - se x_q, sn x_p + ce x_q] for the rows of A and V
            for (int item = tid; item < half * M; item += HEIG_THREADS) {
                const int k = item / M, i = item - k * M;
                const Rot r = rot[k];
                if (r.p < 0) continue;
                double2 *row = A + i * P;
#pragma unroll
                for (int which = 0; which < 2; ++which) {
                    const double2 xp = row[r.p], xq = row[r.q];
                    const double2 sq = cmul(r.se, xq), cq = cmul(r.ce, xq);
                    row[r.p] = make_double2(r.cs * xp.x - sq.x, r.cs * xp.y - sq.y);
                    row[r.q] = make_double2(r.sn * xp.x + cq.x, r.sn * xp.y + cq.y);
                    row = V + i * P;
                }
            }
            __syncthreads();
            // ---- rows of A: [a_p, a_q] <- [cs a_p - conj(se) a_q, sn a_p + conj(ce) a_q]; exact pivot block
            for (int item = tid; item < half * M; item += HEIG_THREADS) {
                const int k = item / M, j = item - k * M;
                const Rot r = rot[k];
                if (r.p < 0) continue;
                double2 *rp = A + r.p * P + j, *rq = A + r.q * P + j;
                if (j == r.p) { *rp = make_double2(r.app, 0.0); *rq = make_double2(0.0, 0.0); continue; }
                if (j == r.q) { *rp = make_double2(0.0, 0.0); *rq = make_double2(r.aqq, 0.0); continue; }
                const double2 ap = *rp, aq = *rq;
                const double2 sq = cmul(cconj(r.se), aq), cq = cmul(cconj(r.ce), aq);
                *rp = make_double2(r.cs * ap.x - sq.x, r.cs * ap.y - sq.y);
                *rq = make_double2(r.sn * ap.x + cq.x, r.sn * ap.y + cq.y);
            }
            __syncthreads();
        }
    }

    for (int m = tid; m < M; m += HEIG_THREADS) w_out[f * M + m] = A[m * P + m].x;
    double2 *dst = V_out + f * int64_t(M) * M;
    for (int idx = tid; idx < M * M; idx += HEIG_THREADS) {
        const int i = idx / M, j = idx - i * M;
        dst[idx] = V[i * P + j];
    }
    if (sweeps_out && tid == 0) sweeps_out[f] = sweep;
}

int heig_smem_bytes(int M) {
    return 2 * M * (M + 1) * int(sizeof(double2)) + ((M + 1) / 2) * int(sizeof(Rot)) + (HEIG_THREADS / 32) * int(sizeof(double)) + 16;
}

}  // namespace

extern "C" int b2_heig_max_order(void) { return HEIG_MAX_M; }

extern "C" int b2_heig_jacobi(const void *A_dev, int M, int nb, void *w_out_dev, void *V_out_dev, void *sweeps_out_dev,
                              void *stream) {
    B2_REQUIRE(A_dev && w_out_dev && V_out_dev, B2_EINVAL, "b2_heig_jacobi: null pointer");
    B2_REQUIRE(M >= 1 && nb >= 0, B2_EINVAL, "b2_heig_jacobi: bad M=%d / nb=%d", M, nb);
    if (M > HEIG_MAX_M) {
        b2_set_error("b2_heig_jacobi: order %d exceeds the shared-memory solver's limit of %d", M, HEIG_MAX_M);
        return B2_EUNSUPPORTED;
    }
    if (nb == 0) return B2_OK;
    static DeviceOnce once;
    if (int rc = b2_device_once(once, heig_jacobi_kernel, heig_smem_bytes(HEIG_MAX_M), nullptr)) return rc;
    heig_jacobi_kernel<<<(unsigned)nb, HEIG_THREADS, heig_smem_bytes(M), (cudaStream_t)stream>>>(
        (const double2 *)A_dev, M, (double *)w_out_dev, (double2 *)V_out_dev, (int *)sweeps_out_dev);
    B2_LAUNCH_OK("b2_heig_jacobi");
    return B2_OK;
}
