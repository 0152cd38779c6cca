// K3 on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Replaces the G x F np.linalg.multi_dot loop of BeamformerDASFrequency.get_beamformer_map
// (beamforming/beamforming.py:863-868) with the dense steering array of :858 never stored.
//
// Per CTA: a tile of 128 grid points (UMMA M = 128, one TMEM lane per grid point).  For every
// bin f and every block of NB matrix columns n:
//
//     T[g, n] = sum_{m >= n} conj(H[g, m]) L_f[m, n]            (L = 2*tril(C,-1) + diag(C))
//
// is a real GEMM  D[128 x 2NB] = A[128 x 2M] * B^T  with
//     A[g, 2m+{0,1}]   = (Re, Im) of H[g, m] = a_gm exp(-i k_f r_gm)   synthesised by 4 warps
//                        straight into the 128B-swizzled K-major shared-memory operand tile,
//     B[2n,     2m+{0,1}] = ( Re L[m,n],  Im L[m,n])                    -> D[:, 2n]     = Re T
//     B[2n + 1, 2m+{0,1}] = ( Im L[m,n], -Re L[m,n])                    -> D[:, 2n + 1] = Im T
// (operand images pre-swizzled by b2_das_prepare and streamed with cp.async.bulk), accumulated
// in TMEM as fp32.  L is lower triangular: K chunks (16 microphones) entirely above the diagonal
// are skipped, and a chunk that straddles it only touches its first w = 16, 32, ... columns, so
// the MMA is issued with N = 2w and only the first 2w operand rows are fetched.  Chunks are
// walked from the last one down so that the first MMA of an accumulator (the one that
// overwrites instead of accumulating) is the widest.
// The epilogue warps read D back with tcgen05.ld, re-synthesise H[g, n] and reduce
//     P[g, f] = sum_n Re(T[g, n] H[g, n]),   clamp,   map[g] += w_f P[g, f]
// while the next column block accumulates into the second TMEM buffer.
// Split-TF32 (PASSES = 3): every operand is split x = hi + lo (both TF32-exact) and
// hi*hi + lo*hi + hi*lo is accumulated, which restores ~fp32 accuracy.
// Mixed mode (PASSES = 2): hi*hi stays a TF32 product, but the two correction products only need
// ~8 significant bits, so they run as ONE kind::f16 (BF16) product of twice the depth at twice
// the TF32 rate: the second operand image of a stage holds, per row, [bf16(x - hi) | bf16(x)] for
// A and [bf16(y) | bf16(y - hi)] for B, so that sum_k A2[k] B2[k] = lo_A*y + x*lo_B.  Tensor work
// drops from 3 to 2 TF32-equivalents per product at ~2^-20 relative accuracy.
//
// Warp roles (576 threads): 0-7 steering synthesis (two threads per grid point), 8-15 epilogue
// (two warps per TMEM lane quadrant), 16 MMA issue + TMEM allocation, 17 operand loader.
// Pipelines: smem ring full/empty mbarriers (loader + synthesis -> MMA -> tcgen05.commit), TMEM
// full/empty mbarriers (MMA -> epilogue).  The bin-independent part of the steering vector
// (distance r and amplitude a of every (grid point, microphone) pair of the tile) is computed
// once per tile in fp64 and parked in an L2-resident scratch table, so the per-bin work is one
// compensated phase product, one sincos and the TF32 split.
#include <cuda_bf16.h>

#include <type_traits>

#include "common.cuh"
#include "das_common.cuh"

namespace {

constexpr int TILE_G = 128;
constexpr int KCH = 16;                       // microphones per K chunk = 32 reals = one 128 B row
constexpr int A_TILE = TILE_G * 128;          // bytes of one A operand image
constexpr int TC_THREADS = 576;
constexpr int NGEN = 256;                     // synthesis threads
constexpr int NEPI = 256;                     // epilogue threads
constexpr int MAX_CTAS = 160;                 // scratch tables are sized for this many persistent CTAs
constexpr int MAX_MICS = 1024;

__host__ __device__ constexpr int tc_nb(int M) { return M <= 32 ? 32 : (M <= 64 ? 64 : 128); }

template <int NB, int PASSES> struct TcCfg {
    static constexpr int UMMA_N = 2 * NB;
    static constexpr int B_TILE = UMMA_N * 128;
    static constexpr int NSPLIT = PASSES == 1 ? 1 : 2;                 // operand images per stage (hi[, correction])
    static constexpr int STAGE = NSPLIT * (A_TILE + B_TILE);
    static constexpr int NSTAGE_RAW = (208 * 1024) / STAGE;
    static constexpr int NSTAGE = NSTAGE_RAW > 6 ? 6 : NSTAGE_RAW;
    static constexpr int TMEM_COLS = 2 * UMMA_N;                       // two accumulator buffers
    static constexpr int SMEM = 1024 /*align slack*/ + NSTAGE * STAGE + 2 * TILE_G * 4 + 256;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// Wait for the phase with the given parity.  try_wait suspends the thread in hardware until the
// phase completes or the time hint expires, so waiting warps do not burn issue slots spinning
// (the first version of this kernel spent ~20 % of its executed instructions in wait loops).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(0x989680u)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t *r = reinterpret_cast<uint32_t *>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t *r = reinterpret_cast<uint32_t *>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// K-major, 128B-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 | LBO(=1)<<16 | SBO(=1024B>>4)<<32 | version(=1)<<46 | layout SWIZZLE_128B(=2)<<61
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    return uint64_t((saddr & 0x3FFFF) >> 4) | (uint64_t(1) << 16) | (uint64_t(64) << 32) | (uint64_t(1) << 46) |
           (uint64_t(2) << 61);
}
// kind::tf32 instruction descriptor: fp32 accumulate, TF32 x TF32, both operands K-major
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
// kind::f16 instruction descriptor: fp32 accumulate, BF16 x BF16, both operands K-major
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo_elem, float hi_elem) {   // lo_elem first in memory
    __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);
    return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ void sts128u(uint32_t saddr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// byte offset of the 16-byte chunk j (0..7) of row r inside a 128B-swizzled K-major tile
__host__ __device__ __forceinline__ int sw128_offset(int r, int j) { return (r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4); }

// ---------------------------------------------------------------- per-tile geometry table
// tab[((m >> 3) * 128 + row) * 8 + (m & 7)] = r = |grid_row - mic_m| (fp32 of an fp64 result): the
// bin-independent part of the steering phase.  The amplitude follows from r and three per-row
// constants, |h| = c0 + c1 / r + c2 * r  (the four formulations of beamforming.py:1515-1703).
__device__ __forceinline__ int tab_index(int m, int row) { return (((m >> 3) * TILE_G) + row) * 8 + (m & 7); }

// The four formulations of beamforming.py:1515-1703 all have |h_m| = c_g * u(r_m) with ONE per-grid-point
// constant c_g and u = 1 (classic), r (inverse) or 1/r (true power / true location).  Because the map is
// quadratic in h, the kernels work with h' = u(r) exp(-i phi) and scale the finished power by c_g^2.
enum { AMP_ONE = 0, AMP_R = 1, AMP_RINV = 2 };
__host__ __device__ inline int amp_kind(int steering) {
    return steering == B2_STEER_CLASSIC ? AMP_ONE : (steering == B2_STEER_INVERSE ? AMP_R : AMP_RINV);
}

// Fill r for the microphones with ((m >> 3) & 1) == half of one grid point (if `write`); returns c_g^2
// (0 for rows past the end of the grid).  Padded microphones get r = 1 (their operand columns are zero).
__device__ __forceinline__ double row_table(float *__restrict__ tab, int row, int half, bool write, int64_t g,
                                            int64_t G, const double *__restrict__ grid,
                                            const double *__restrict__ mic, int M, int Mk, double cx0, double cy0,
                                            double cz0, int steering) {
    const bool valid = g < G;
    double x = 0, y = 0, z = 0, c = 0.0;
    if (valid) {
        x = grid[3 * g]; y = grid[3 * g + 1]; z = grid[3 * g + 2];
        double s = 0.0;
        for (int m = 0; m < M; ++m) {
            const double dx = x - mic[3 * m], dy = y - mic[3 * m + 1], dz = z - mic[3 * m + 2];
            s += 1.0 / (dx * dx + dy * dy + dz * dz);
        }
        const double r0 = das_distance(x, y, z, cx0, cy0, cz0);
        switch (steering) {
        case B2_STEER_CLASSIC: c = 1.0 / double(M); break;
        case B2_STEER_INVERSE: c = 1.0 / (double(M) * r0); break;
        case B2_STEER_TRUE_POWER: c = 1.0 / (r0 * s); break;
        default: c = 1.0 / sqrt(double(M) * s); break;
        }
    }
    if (write) {
        for (int m8 = half * 8; m8 < Mk; m8 += 16) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = m8 + i;
                float v = 1.0f;
                if (valid && m < M) v = float(das_distance(x, y, z, mic[3 * m], mic[3 * m + 1], mic[3 * m + 2]));
                tab[tab_index(m, row)] = v;
            }
        }
    }
    return c * c;
}

// h'[g, m] = u(r) exp(-i 2 pi q r), q = f / c in cycles per metre: returns (Re, Im).
// The phase is reduced to a fraction of a cycle before the MUFU sincos.  q r < 2^9 cycles, so the
// fp32 product carries an absolute phase error <= ~1e-4 rad per element -- the same size as the
// fp32 rounding of r itself; being uncorrelated over microphones it changes the map by < 1e-6 of
// its maximum (first-order phase perturbations cancel at a coherent peak).
template <int AMP>
__device__ __forceinline__ float2 steer(float r, float q) {
    const float p = q * r;
    const float rn = (p + 12582912.0f) - 12582912.0f;         // round to nearest integer (|p| < 2^22)
    const float fr = p - rn;
    float sn, cs;
    __sincosf(6.28318530717958647692f * fr, &sn, &cs);
    if (AMP == AMP_ONE) return make_float2(cs, -sn);
    float u = r;
    if (AMP == AMP_RINV) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(u) : "f"(r));
    return make_float2(u * cs, -u * sn);
}

// top 11 significant bits by truncation: the residual x - trunc(x) is exact and is carried by the
// correction operand, so no rounding step is needed in the split modes
__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__device__ __forceinline__ void sts128(uint32_t saddr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---------------------------------------------------------------- the kernel
// work unit -> (tile of 128 grid points, contiguous range of bins); units are dealt round-robin
#define B2_UNIT_DECODE                                                        \
    const int tile = unit / nsplit, upart = unit - tile * nsplit;            \
    const int f_begin = int(int64_t(nbins) * upart / nsplit);                \
    const int f_end = int(int64_t(nbins) * (upart + 1) / nsplit);

template <int NB, int PASSES>
__global__ void __launch_bounds__(TC_THREADS, 1)
das_map_tc_kernel(const unsigned char *__restrict__ op, float *__restrict__ scratch, int M, int nbins,
                  const double *__restrict__ freqs, const double *__restrict__ weights, double inv_c,
                  const double *__restrict__ mic, double cx0, double cy0, double cz0,
                  const double *__restrict__ grid, int64_t G, int steering, int clamp_negative,
                  double *__restrict__ map_out, double *__restrict__ per_bin, int ntiles, int nsplit) {
    using CF = TcCfg<NB, PASSES>;
    constexpr int UMMA_N = CF::UMMA_N, NSTAGE = CF::NSTAGE;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float *part = reinterpret_cast<float *>(smem + NSTAGE * CF::STAGE);                // [2][128] epilogue partials
    uint64_t *bars = reinterpret_cast<uint64_t *>(part + 2 * TILE_G);
    uint64_t *full = bars, *empty = bars + NSTAGE, *tfull = bars + 2 * NSTAGE, *tempty = bars + 2 * NSTAGE + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * NSTAGE + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NKC = (M + KCH - 1) / KCH;          // K chunks
    const int Mk = NKC * KCH;
    const int NBLK = (M + NB - 1) / NB;           // column blocks
    const int akind = amp_kind(steering);
    float *tab = scratch + size_t(blockIdx.x) * TILE_G * Mk;

    // stage s: [A_hi][A_lo?][B_hi][B_lo?]
    auto a_tile = [&](int s, int split) { return smem + s * CF::STAGE + split * A_TILE; };
    auto b_tile = [&](int s, int split) { return smem + s * CF::STAGE + CF::NSPLIT * A_TILE + split * CF::B_TILE; };

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], NGEN + 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], NEPI); }
        fence_barrier_init();
    }
    if (warp == 16) tmem_alloc(tmem_slot, CF::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // =========================== steering synthesis -> A operand tiles ===========================
        const int r = tid & 127, half = tid >> 7;            // row of the tile, which 8 of a chunk's 16 mics
        const uint32_t a_base = smem_u32(smem);
        uint32_t it = 0;
        for (int unit = blockIdx.x; unit < ntiles * nsplit; unit += gridDim.x) {
            B2_UNIT_DECODE
            // the epilogue warps are done with the previous tile's table before it is overwritten
            asm volatile("bar.sync 1, 512;" ::: "memory");
            row_table(tab, r, half, true, int64_t(tile) * TILE_G + r, G, grid, mic, M, Mk, cx0, cy0, cz0, steering);
            __threadfence_block();
            asm volatile("bar.sync 1, 512;" ::: "memory");
            // distances of this thread's 8 microphones of a chunk: 32 contiguous bytes, coalesced over rows
            auto load_r = [&](int kc, float4 &lo, float4 &hi) {
                const float4 *src = reinterpret_cast<const float4 *>(tab + tab_index(kc * KCH + half * 8, r));
                lo = __ldcg(src);
                hi = __ldcg(src + 1);
            };
            float4 r_lo, r_hi;
            load_r(NKC - 1, r_lo, r_hi);
            for (int fi = f_begin; fi < f_end; ++fi) {
                const float bf = float(freqs[fi] * inv_c);
                for (int nbk = 0; nbk < NBLK; ++nbk) {
                    const int kc0 = (nbk * NB) / KCH;
                    for (int kc = NKC - 1; kc >= kc0; --kc, ++it) {
                        const int s = it % NSTAGE;
                        const float rr[8] = {r_lo.x, r_lo.y, r_lo.z, r_lo.w, r_hi.x, r_hi.y, r_hi.z, r_hi.w};
                        // prefetch the next chunk's distances (the walk only depends on kc)
                        int kn = kc - 1;
                        if (kn < kc0) kn = NKC - 1;
                        load_r(kn, r_lo, r_hi);
                        float4 hi4[4], lo4[4];
                        uint4 mixlo[2], mixb[2];          // mixed mode: bf16 residuals / bf16 values of 8 mics
                        auto synth = [&](auto amp_tag) {
                            constexpr int AMP = decltype(amp_tag)::value;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 h0 = steer<AMP>(rr[2 * j], bf);
                                const float2 h1 = steer<AMP>(rr[2 * j + 1], bf);
                                if (PASSES == 1) hi4[j] = make_float4(to_tf32(h0.x), to_tf32(h0.y), to_tf32(h1.x), to_tf32(h1.y));
                                else hi4[j] = make_float4(trunc_tf32(h0.x), trunc_tf32(h0.y), trunc_tf32(h1.x), trunc_tf32(h1.y));
                                if (PASSES == 3)   // residuals are exact in fp32; the tensor core keeps their top 11 bits
                                    lo4[j] = make_float4(h0.x - hi4[j].x, h0.y - hi4[j].y, h1.x - hi4[j].z, h1.y - hi4[j].w);
                                if (PASSES == 2) {
                                    const uint32_t l0 = pack_bf16(h0.x - hi4[j].x, h0.y - hi4[j].y);
                                    const uint32_t l1 = pack_bf16(h1.x - hi4[j].z, h1.y - hi4[j].w);
                                    const uint32_t b0 = pack_bf16(h0.x, h0.y), b1 = pack_bf16(h1.x, h1.y);
                                    uint32_t *pl = reinterpret_cast<uint32_t *>(&mixlo[j >> 1]) + 2 * (j & 1);
                                    uint32_t *pb = reinterpret_cast<uint32_t *>(&mixb[j >> 1]) + 2 * (j & 1);
                                    pl[0] = l0; pl[1] = l1; pb[0] = b0; pb[1] = b1;
                                }
                            }
                        };
                        if (akind == AMP_RINV) synth(std::integral_constant<int, AMP_RINV>{});
                        else if (akind == AMP_R) synth(std::integral_constant<int, AMP_R>{});
                        else synth(std::integral_constant<int, AMP_ONE>{});
                        mbar_wait(&empty[s], ((it / NSTAGE) & 1) ^ 1);
                        const uint32_t ahi = a_base + s * CF::STAGE;
                        const uint32_t alo = ahi + (CF::NSPLIT - 1) * A_TILE;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int off = sw128_offset(r, half * 4 + j);
                            sts128(ahi + off, hi4[j]);
                            if (PASSES == 3) sts128(alo + off, lo4[j]);
                        }
                        if (PASSES == 2) {
                            // correction image row = [bf16(h - hi) of the chunk's 16 mics | bf16(h) of the 16 mics]:
                            // this thread's 8 mics are 32 bytes = 16-byte chunks 2*half, 2*half+1 of each 64-byte part
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                sts128u(alo + sw128_offset(r, 2 * half + c), mixlo[c]);
                                sts128u(alo + sw128_offset(r, 4 + 2 * half + c), mixb[c]);
                            }
                        }
                        fence_proxy_async();                 // generic-proxy writes -> visible to the MMA (async proxy)
                        mbar_arrive(&full[s]);
                    }
                }
            }
        }
    } else if (warp < 16) {
        // =========================== epilogue: TMEM -> P[g,f] -> map ===========================
        const int ew = warp - 8;
        const int quad = ew & 3, half = ew >> 2;             // TMEM lane quadrant, which half of the columns
        const int r = quad * 32 + lane;
        const uint32_t lane_base = uint32_t(quad * 32) << 16;
        constexpr int CG_PER_HALF = (NB / 32 + 1) / 2;       // 32-column groups per epilogue half
        uint32_t acc_it = 0;
        for (int unit = blockIdx.x; unit < ntiles * nsplit; unit += gridDim.x) {
            B2_UNIT_DECODE
            const int64_t g = int64_t(tile) * TILE_G + r;
            asm volatile("bar.sync 1, 512;" ::: "memory");
            const double c2 = row_table(tab, r, 0, false, g, G, grid, mic, M, Mk, cx0, cy0, cz0, steering);
            asm volatile("bar.sync 1, 512;" ::: "memory");   // table of this tile is complete
            double acc_map = 0.0;
            for (int fi = f_begin; fi < f_end; ++fi) {
                const float bf = float(freqs[fi] * inv_c);
                float p0 = 0.f, p1 = 0.f;
                for (int nbk = 0; nbk < NBLK; ++nbk, ++acc_it) {
                    const int buf = acc_it & 1;
                    mbar_wait(&tfull[buf], (acc_it >> 1) & 1);
                    tc_fence_after();
                    const uint32_t tcol = tmem_base + lane_base + buf * UMMA_N;
#pragma unroll 1
                    for (int c = 0; c < 2 * CG_PER_HALF; ++c) {          // 16 columns (microphones) at a time
                        const int col = half * CG_PER_HALF * 32 + c * 16;
                        const int n0 = nbk * NB + col;
                        if (col >= NB || n0 >= Mk) break;
                        float t[32];                                     // (Re T, Im T) of 16 columns n
                        tmem_ld32(tcol + 2 * col, t);
                        // distances of the 16 microphones n0 .. n0+15 of this row (two 32-byte runs)
                        float4 rq[4];
#pragma unroll
                        for (int k8 = 0; k8 < 2; ++k8) {
                            const float4 *src = reinterpret_cast<const float4 *>(tab + tab_index(n0 + 8 * k8, r));
                            rq[2 * k8] = __ldcg(src);
                            rq[2 * k8 + 1] = __ldcg(src + 1);
                        }
                        tmem_ld_wait();
                        auto reduce = [&](auto amp_tag) {
                            constexpr int AMP = decltype(amp_tag)::value;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                // padded microphones have all-zero operand rows, so T is 0 there: no bounds test
                                const float2 h0 = steer<AMP>(rq[j].x, bf), h1 = steer<AMP>(rq[j].y, bf);
                                const float2 h2 = steer<AMP>(rq[j].z, bf), h3 = steer<AMP>(rq[j].w, bf);
                                p0 = fmaf(t[8 * j], h0.x, p0);
                                p0 = fmaf(-t[8 * j + 1], h0.y, p0);
                                p1 = fmaf(t[8 * j + 2], h1.x, p1);
                                p1 = fmaf(-t[8 * j + 3], h1.y, p1);
                                p0 = fmaf(t[8 * j + 4], h2.x, p0);
                                p0 = fmaf(-t[8 * j + 5], h2.y, p0);
                                p1 = fmaf(t[8 * j + 6], h3.x, p1);
                                p1 = fmaf(-t[8 * j + 7], h3.y, p1);
                            }
                        };
                        if (akind == AMP_RINV) reduce(std::integral_constant<int, AMP_RINV>{});
                        else if (akind == AMP_R) reduce(std::integral_constant<int, AMP_R>{});
                        else reduce(std::integral_constant<int, AMP_ONE>{});
                    }
                    tc_fence_before();
                    mbar_arrive(&tempty[buf]);
                }
                // combine the two column halves of this row
                float *pp = part + (fi & 1) * TILE_G;
                if (half == 1) pp[r] = p0 + p1;
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (half == 0) {
                    double p = (double(p0 + p1) + double(pp[r])) * c2;        // |h|^2 scale of this grid point
                    if (clamp_negative && p < 0.0) p = 0.0;                         // beamforming.py:871-872
                    acc_map += weights[fi] * p;                                       // :875-876
                    if (per_bin != nullptr && g < G) per_bin[g * nbins + fi] = p;
                }
            }
            // the bins of a tile may be split over several work units: the map is zeroed by the
            // launcher and every unit adds its weighted partial sum
            if (half == 0 && g < G) atomicAdd(map_out + g, acc_map);
        }
    } else if (warp == 16) {
        // =========================== MMA issue (one thread) ===========================
        if (lane == 0) {
            uint32_t it = 0, acc_it = 0;
            for (int unit = blockIdx.x; unit < ntiles * nsplit; unit += gridDim.x) {
            B2_UNIT_DECODE
                for (int fi = f_begin; fi < f_end; ++fi) {
                    for (int nbk = 0; nbk < NBLK; ++nbk, ++acc_it) {
                        const int buf = acc_it & 1;
                        mbar_wait(&tempty[buf], ((acc_it >> 1) & 1) ^ 1);
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + buf * UMMA_N;
                        const int kc0 = (nbk * NB) / KCH;
                        for (int kc = NKC - 1; kc >= kc0; --kc, ++it) {
                            const int s = it % NSTAGE;
                            // columns of this block that chunk kc can reach (lower triangle): w, N = 2w
                            const int w = min(NB, KCH * (kc - kc0 + 1));
                            const uint32_t idesc = idesc_tf32(TILE_G, 2 * w);
                            mbar_wait(&full[s], (it / NSTAGE) & 1);
                            tc_fence_after();
                            const uint64_t a_hi = smem_desc_sw128(smem_u32(a_tile(s, 0)));
                            const uint64_t b_hi = smem_desc_sw128(smem_u32(b_tile(s, 0)));
                            const uint64_t a_lo = smem_desc_sw128(smem_u32(a_tile(s, CF::NSPLIT - 1)));
                            const uint64_t b_lo = smem_desc_sw128(smem_u32(b_tile(s, CF::NSPLIT - 1)));
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {              // UMMA K = 8 tf32 = 32 bytes
                                const uint64_t adv = uint64_t(ks * 2);    // (32 B) >> 4
                                umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, (kc < NKC - 1 || ks > 0) ? 1u : 0u);
                                if (PASSES == 3) {
                                    umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
                                    umma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                                }
                                // mixed mode: both corrections as one BF16 product of depth 64 (4 x K=16)
                                if (PASSES == 2) umma_bf16(d_tmem, a_lo + adv, b_lo + adv, idesc_bf16(TILE_G, 2 * w), 1u);
                            }
                            umma_commit(&empty[s]);                       // frees the stage when the MMAs retire
                        }
                        umma_commit(&tfull[buf]);                         // accumulator complete -> epilogue
                    }
                }
            }
        }
    } else {
        // =========================== B operand loader (one thread) ===========================
        if (lane == 0) {
            uint32_t it = 0;
            const size_t bin_bytes = size_t(NBLK) * NKC * CF::NSPLIT * CF::B_TILE;
            for (int unit = blockIdx.x; unit < ntiles * nsplit; unit += gridDim.x) {
            B2_UNIT_DECODE
                for (int fi = f_begin; fi < f_end; ++fi) {
                    for (int nbk = 0; nbk < NBLK; ++nbk) {
                        const int kc0 = (nbk * NB) / KCH;
                        for (int kc = NKC - 1; kc >= kc0; --kc, ++it) {
                            const int s = it % NSTAGE;
                            const int w = min(NB, KCH * (kc - kc0 + 1));
#ifdef B2_EXPERIMENT_BDIV
                            const uint32_t bytes = uint32_t(2 * w) * 128 / B2_EXPERIMENT_BDIV;   // timing experiment only (wrong results)
#else
                            const uint32_t bytes = uint32_t(2 * w) * 128;          // first 2w rows of each image
#endif
                            mbar_wait(&empty[s], ((it / NSTAGE) & 1) ^ 1);
                            const unsigned char *src = op + size_t(fi) * bin_bytes + (size_t(nbk) * NKC + kc) * (CF::NSPLIT * CF::B_TILE);
                            mbar_arrive_expect_tx(&full[s], CF::NSPLIT * bytes);
                            bulk_g2s(b_tile(s, 0), src, bytes, &full[s]);
                            if (PASSES != 1) bulk_g2s(b_tile(s, 1), src + CF::B_TILE, bytes, &full[s]);
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 16) tmem_dealloc(tmem_base, CF::TMEM_COLS);
}

// ---------------------------------------------------------------- K3a for the tensor-core path
// One thread per 16-byte chunk of an operand image: (bin, column block, K chunk, row, chunk j).
template <int NB, int PASSES>
__global__ void das_prepare_tc_kernel(const double2 *__restrict__ csm, int M, int bin0, double gain, int remove_diag,
                                      unsigned char *__restrict__ op, int NKC, int NBLK) {
    using CF = TcCfg<NB, PASSES>;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;       // 0 .. 2NB-1
    if (row >= CF::UMMA_N) return;
    const int kc = blockIdx.y % NKC, nbk = blockIdx.y / NKC;
    const int fi = blockIdx.z;
    const bool imag_row = row & 1;                                // rows alternate Re T / Im T of column n
    const int n = nbk * NB + (row >> 1);
    unsigned char *tile = op + ((size_t(fi) * NBLK + nbk) * NKC + kc) * size_t(CF::NSPLIT * CF::B_TILE);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float v[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int m = kc * KCH + 2 * j + e;
            double lr = 0.0, li = 0.0;
            if (m < M && n < M && n <= m) {
                const double2 c = csm[(int64_t(bin0 + fi) * M + m) * M + n];
                if (n == m) { lr = remove_diag ? 0.0 : gain * c.x; }
                else { lr = 2.0 * gain * c.x; li = 2.0 * gain * c.y; }
            }
            // real rows: (Re L, Im L) ; imaginary rows: (Im L, -Re L)
            v[2 * e] = float(imag_row ? li : lr);
            v[2 * e + 1] = float(imag_row ? -lr : li);
        }
        const int off = sw128_offset(row, j);
        const float4 hi = make_float4(to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]));
        *reinterpret_cast<float4 *>(tile + off) = hi;
        if (PASSES == 3) {
            const float4 lo = make_float4(to_tf32(v[0] - hi.x), to_tf32(v[1] - hi.y), to_tf32(v[2] - hi.z), to_tf32(v[3] - hi.w));
            *reinterpret_cast<float4 *>(tile + CF::B_TILE + off) = lo;
        }
        if (PASSES == 2) {
            // correction image row = [bf16(y) of the chunk's 16 mics | bf16(y - hi)]: the 2 mics of this j are
            // 8 bytes at byte offset 8*j of each 64-byte part, i.e. half of 16-byte chunk j/2 (resp. 4 + j/2)
            unsigned char *img = tile + CF::B_TILE;
            uint2 vb = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
            uint2 vl = make_uint2(pack_bf16(v[0] - hi.x, v[1] - hi.y), pack_bf16(v[2] - hi.z, v[3] - hi.w));
            *reinterpret_cast<uint2 *>(img + sw128_offset(row, j >> 1) + 8 * (j & 1)) = vb;
            *reinterpret_cast<uint2 *>(img + sw128_offset(row, 4 + (j >> 1)) + 8 * (j & 1)) = vl;
        }
    }
}

template <int PASSES> int64_t images_bytes(int M, int nbins) {
    const int NKC = (M + KCH - 1) / KCH;
    const int nb = tc_nb(M), NBLK = (M + nb - 1) / nb;
    const int64_t btile = 2 * nb * 128;
    return int64_t(nbins) * NBLK * NKC * (PASSES == 1 ? 1 : 2) * btile;
}

template <int NB, int PASSES>
int launch_prepare(const void *csm, int M, int bin0, int nbins, int remove_diag, void *op, cudaStream_t st) {
    using CF = TcCfg<NB, PASSES>;
    const int NKC = (M + KCH - 1) / KCH, NBLK = (M + NB - 1) / NB;
    const double gain = remove_diag ? double(M) / double(M - 1) : 1.0;
    dim3 grid((CF::UMMA_N + 127) / 128, NKC * NBLK, nbins);
    das_prepare_tc_kernel<NB, PASSES><<<grid, 128, 0, st>>>((const double2 *)csm, M, bin0, gain, remove_diag,
                                                            (unsigned char *)op, NKC, NBLK);
    B2_LAUNCH_OK("b2_das_prepare(tc)");
    return B2_OK;
}

template <int NB, int PASSES>
int launch_map(const void *op, int M, int nbins, const void *freqs, const void *weights, double c, const void *mic,
               const double *center, const void *grid, int64_t G, int steering, int clamp_negative, void *map_out,
               void *per_bin, cudaStream_t st) {
    using CF = TcCfg<NB, PASSES>;
    auto kern = das_map_tc_kernel<NB, PASSES>;
    static bool attr_done = false;
    static int num_sms = 0;
    if (!attr_done) {
        B2_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM));
        int dev = 0;
        B2_CUDA_OK(cudaGetDevice(&dev));
        B2_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        attr_done = true;
    }
    const int64_t ntiles = (G + TILE_G - 1) / TILE_G;
    B2_REQUIRE(ntiles <= 0x7fffffff, B2_EINVAL, "b2_das_map: too many grid points");
    // Persistent grid, one CTA per SM.  When the tiles do not fill the last round of CTAs evenly
    // (e.g. 977 tiles on 148 SMs: 60 % of the SMs idle in round 7), the bins of every tile are split
    // into up to 4 work units so that the tail shrinks; partial maps are combined with atomicAdd.
    int cap = num_sms < MAX_CTAS ? num_sms : MAX_CTAS;
    int nsplit = 1;
    double best = 0.0;
    for (int sp = 1; sp <= 4 && nbins / sp >= 32; ++sp) {
        const int64_t units = ntiles * sp;
        const int64_t rounds = (units + cap - 1) / cap;
        const double eff = double(units) / double(rounds * cap);
        if (eff > best + 0.01) { best = eff; nsplit = sp; }
    }
    const int64_t units = ntiles * nsplit;
    const int nblocks = int(units < cap ? units : cap);
    B2_CUDA_OK(cudaMemsetAsync(map_out, 0, size_t(G) * sizeof(double), st));
    float *scratch = reinterpret_cast<float *>((unsigned char *)op + images_bytes<PASSES>(M, nbins));
    kern<<<nblocks, TC_THREADS, CF::SMEM, st>>>((const unsigned char *)op, scratch, M, nbins, (const double *)freqs,
                                                (const double *)weights, 1.0 / c, (const double *)mic, center[0],
                                                center[1], center[2], (const double *)grid, G, steering,
                                                clamp_negative, (double *)map_out, (double *)per_bin, int(ntiles), nsplit);
    B2_LAUNCH_OK("b2_das_map(tc)");
    return B2_OK;
}

template <int PASSES> int64_t operand_bytes(int M, int nbins) {
    const int NKC = (M + KCH - 1) / KCH;
    // operand images + one distance table of 128 x Mk floats per persistent CTA
    return images_bytes<PASSES>(M, nbins) + int64_t(MAX_CTAS) * TILE_G * NKC * KCH * 4;
}

}  // namespace

int64_t b2_das_tc_operand_bytes(int M, int nbins, int mode) {
    if (M > MAX_MICS) { b2_set_error("tensor-core map kernel supports at most %d microphones", MAX_MICS); return B2_EUNSUPPORTED; }
    return mode == B2_MAP_TC_TF32 ? operand_bytes<1>(M, nbins) : operand_bytes<3>(M, nbins);
}

#define B2_TC_DISPATCH(FN, ...)                                                      \
    do {                                                                             \
        const int nb_ = tc_nb(M);                                                    \
        if (mode == B2_MAP_TC_TF32X3) {                                              \
            if (nb_ == 32) return FN<32, 3>(__VA_ARGS__);                            \
            if (nb_ == 64) return FN<64, 3>(__VA_ARGS__);                            \
            return FN<128, 3>(__VA_ARGS__);                                          \
        } else if (mode == B2_MAP_TC_MIXED) {                                        \
            if (nb_ == 32) return FN<32, 2>(__VA_ARGS__);                            \
            if (nb_ == 64) return FN<64, 2>(__VA_ARGS__);                            \
            return FN<128, 2>(__VA_ARGS__);                                          \
        } else {                                                                     \
            if (nb_ == 32) return FN<32, 1>(__VA_ARGS__);                            \
            if (nb_ == 64) return FN<64, 1>(__VA_ARGS__);                            \
            return FN<128, 1>(__VA_ARGS__);                                          \
        }                                                                            \
    } while (0)

int b2_das_tc_prepare(const void *csm_dev, int M, int nb, int bin0, int nbins, int remove_diag, void *op_dev,
                      int mode, cudaStream_t st) {
    (void)nb;
    B2_REQUIRE(M <= MAX_MICS, B2_EUNSUPPORTED, "tensor-core map kernel supports at most %d microphones", MAX_MICS);
    B2_TC_DISPATCH(launch_prepare, csm_dev, M, bin0, nbins, remove_diag, op_dev, st);
}

int b2_das_tc_map(const void *op_dev, int M, int nbins, const void *freqs_dev, const void *weights_dev, double c,
                  const void *mic_xyz_dev, const double *center_xyz, const void *grid_xyz_dev, int64_t G,
                  int steering, int clamp_negative, void *map_out_dev, void *per_bin_out_dev, int mode,
                  cudaStream_t st) {
    B2_REQUIRE(M <= MAX_MICS, B2_EUNSUPPORTED, "tensor-core map kernel supports at most %d microphones", MAX_MICS);
    B2_TC_DISPATCH(launch_map, op_dev, M, nbins, freqs_dev, weights_dev, c, mic_xyz_dev, center_xyz, grid_xyz_dev, G,
                   steering, clamp_negative, map_out_dev, per_bin_out_dev, st);
}
