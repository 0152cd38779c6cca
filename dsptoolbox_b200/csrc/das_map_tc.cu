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
// walked from the last one down, so that the first MMA of an accumulator (the one that overwrites
// instead of accumulating) is the widest; two neighbouring column blocks share their A tiles
// (walk_pair below).
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
// FP16x3 mode (PASSES = 4): both operands are split into two FP16 halves x = f1 + f2 (22 significant
// bits) and f1*g1 + f1*g2 + f2*g1 is accumulated as three kind::f16 products: 1.5 TF32-equivalents
// of tensor work and -- what matters more, because the kernel is bound by shared-memory bandwidth
// (every operand byte is written once and read once per product) -- operand rows of 128 B per 16
// microphones instead of 256 B.  FP16's narrow exponent is handled by exact power-of-two scalings:
// one per bin for L (max |L| -> [2^11, 2^12)) and one per grid point for the amplitude u(r), both
// divided out of the finished power in fp64.
//
// Warp roles: 4*TPR steering-synthesis warps (TPR threads per grid point), 8 epilogue warps (two per
// TMEM lane quadrant), 1 MMA issue + TMEM allocation warp, 1 operand loader warp.
// Pipelines: two shared-memory rings with full/empty mbarriers -- A (synthesis -> MMA, 3+ stages) and
// B (loader -> MMA, 2-3 stages), both released by tcgen05.commit -- and TMEM full/empty mbarriers
// (MMA -> epilogue).  The bin-independent part of the steering vector
// (distance r and amplitude a of every (grid point, microphone) pair of the tile) is computed
// once per tile in fp64 and parked in an L2-resident scratch table, so the per-bin work is one
// compensated phase product, one sincos and the TF32 split.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "das_common.cuh"

namespace {

constexpr int TILE_G = 128;
constexpr int KCH = 16;                       // microphones per K chunk = 32 reals = one 128 B row
constexpr int A_TILE = TILE_G * 128;          // bytes of one A operand image
#ifndef B2_TPR
#define B2_TPR 2
#endif
constexpr int TPR = B2_TPR;                   // synthesis threads per grid point (1, 2 or 4)
constexpr int MPT = KCH / TPR;                // microphones of a chunk per synthesis thread
constexpr int G4 = MPT / 4;                   // ... as groups of 4 (16 bytes of the distance table)
constexpr int NGEN = TILE_G * TPR;            // synthesis threads
constexpr int NEPI = 256;                     // epilogue threads
constexpr int W_EPI0 = NGEN / 32;             // first epilogue warp
constexpr int W_MMA = W_EPI0 + NEPI / 32;     // two MMA issue warps (one per accumulator stream); the loader warp follows
constexpr int TC_THREADS = NGEN + NEPI + 96;

constexpr int MAX_CTAS = 160;                 // scratch tables are sized for this many persistent CTAs
constexpr int MAX_MICS = 1024;

__host__ __device__ constexpr int tc_nb(int M) { return M <= 32 ? 32 : (M <= 64 ? 64 : 128); }

template <int NB, int PASSES> struct TcCfg {
    static constexpr int UMMA_N = 2 * NB;
    static constexpr int B_TILE = UMMA_N * 128;
    static constexpr int NSPLIT = (PASSES == 1 || PASSES == 4) ? 1 : 2; // operand images per stage (hi[, correction])
    // Two independent shared-memory rings: A (synthesised steering tiles, small stages) and B (streamed
    // operand images, large stages).  The A ring is made as deep as the 227 KB allow, because the
    // synthesis warps and the tensor pipe have different per-chunk costs (the MMA width varies with the
    // chunk) and a deep ring is what lets them overlap instead of waiting for each other.
    static constexpr int A_STAGE = NSPLIT * A_TILE;
    static constexpr int B_STAGE = NSPLIT * B_TILE;
#ifdef B2_NSB
    static constexpr int NSB = B2_NSB;
#else
    static constexpr int NSB = B_STAGE >= 64 * 1024 ? 2 : 3;
#endif
    static constexpr int NSA_RAW = (224 * 1024 - NSB * B_STAGE) / A_STAGE;
#ifdef B2_NSA
    static constexpr int NSA = B2_NSA;
#else
    static constexpr int NSA = NSA_RAW > 10 ? 10 : NSA_RAW;
#endif
    // With a deep A ring, two neighbouring column blocks share the synthesised A tiles (see walk_pair);
    // GS = how many chunks the second block's stream lags behind (their tiles are held in the ring).
    static constexpr bool SHARE = NSA >= 6;
#ifdef B2_LAG
    static constexpr int GS = B2_LAG;
#else
    static constexpr int GS = NSA - 2;
#endif
    static constexpr int TMEM_COLS = 2 * UMMA_N;                       // two accumulator buffers
    static constexpr int SMEM = 1024 /*align slack*/ + NSA * A_STAGE + NSB * B_STAGE + 2 * TILE_G * 4 + 256;
};

#ifdef B2_MAP_PROFILE
// Debug instrumentation: cycles each role spends blocked on its barriers (CTA 0 only).
// [0] synthesis wait empty  [1] synthesis total  [2] mma wait full  [3] mma wait tmem-empty  [4] mma total
// [5] epilogue wait tmem-full  [6] epilogue total  [7] mma wait B ring full
__device__ unsigned long long g_map_prof[8];
#define PROF_T0() const long long _t0 = clock64()
#define PROF_ADD(slot, cond) do { if (cond) atomicAdd(&g_map_prof[slot], (unsigned long long)(clock64() - _t0)); } while (0)
#else
#define PROF_T0()
#define PROF_ADD(slot, cond)
#endif

#ifdef B2_MAP_TRACE
// Debug: event trace of CTA 0 for the bins [B2_TRACE_BIN0, B2_TRACE_BIN0 + 2): {role, event, index, clock}
#ifndef B2_TRACE_BIN0
#define B2_TRACE_BIN0 20
#endif
__device__ long long g_trace[4 * 2 * 1024];     // [role][record]{(ev << 16) | idx, clock}
#define TRACE(role, ev, idx, bin)                                                              \
    do {                                                                                       \
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (bin) >= B2_TRACE_BIN0 && (bin) < B2_TRACE_BIN0 + 2 && trace_n < 1024) { \
            g_trace[((role) * 1024 + trace_n) * 2] = ((long long)(ev) << 16) | (long long)(idx);                \
            g_trace[((role) * 1024 + trace_n) * 2 + 1] = clock64();                            \
            ++trace_n;                                                                         \
        }                                                                                      \
    } while (0)
#else
#define TRACE(role, ev, idx, bin)
#endif

#ifdef B2_DEBUG_HANG
__device__ unsigned int g_abort, g_nhang, g_hang[256], g_snapped, g_snap[16];
__device__ unsigned int g_prog[16];      // progress counters of CTA 0: [0] synth chunks [1] epilogue accumulators
                                         // [2] issue warp 0 steps [3] issue warp 1 steps [4] loader tiles, [8+] = 100*role + stage
#define DBG_PROG(slot, v) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_prog[slot] = (v); } while (0)
#else
#define DBG_PROG(slot, v)
#endif

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
#ifdef B2_DEBUG_HANG
    uint32_t spins = 0;
    unsigned long long t_begin;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
#endif
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
#ifdef B2_DEBUG_HANG
        // debug build: a barrier that stays closed for ~1 s is recorded and every waiter gives up, so the
        // kernel ends (with garbage results) and the host can print who was stuck where
        if (!done && ((++spins & 63) == 0)) {
            unsigned long long t_now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
            const bool aborting = *(volatile unsigned int *)&g_abort != 0;
            if (aborting && t_now - t_begin < 100000000ull) break;
            if (aborting || t_now - t_begin > 500000000ull) {   // this one wait has lasted 0.5 s (0.1 s when aborting)
                if ((threadIdx.x & 31) == 0) {
                    const unsigned int k = atomicAdd(&g_nhang, 1u);
                    if (k < 64) { g_hang[4 * k] = blockIdx.x; g_hang[4 * k + 1] = threadIdx.x >> 5; g_hang[4 * k + 2] = addr; g_hang[4 * k + 3] = parity; }
                }
                if (!aborting && (threadIdx.x & 31) == 0 && atomicCAS(&g_snapped, 0u, 1u) == 0u)
                    for (int q = 0; q < 16; ++q) g_snap[q] = *(volatile unsigned int *)&g_prog[q];
                __threadfence();
                *(volatile unsigned int *)&g_abort = 1;
                break;
            }
        }
#endif
    } while (!done);
}
// non-blocking probe of a phase
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
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
//   start>>4 | LBO(=1)<<16 | SBO(=1024B>>4)<<32 | version(=1)<<46 | layout SWIZZLE_128B(=2)<<61
// kept as (constant high word, low word = start>>4 | LBO<<16), so that stepping through ring stages and K
// slices is 32-bit arithmetic on the low word.
constexpr uint32_t DESC_HI = 64u | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFF) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t desc64(uint32_t lo) { return (uint64_t(DESC_HI) << 32) | lo; }
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
// kind::tf32 instruction descriptor: fp32 accumulate, TF32 x TF32, both operands K-major
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
// kind::f16 instruction descriptor: fp32 accumulate, BF16 x BF16, both operands K-major
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
// same, FP16 x FP16
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
    return (1u << 4) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
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
// kind::f16 MMA whose A operand is kept in / taken from the collector buffer: the next MMA with the same
// A slice (and another B) does not read A from shared memory again
__device__ __forceinline__ void umma_f16_afill(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f16_alastuse(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo_elem, float hi_elem) {   // lo_elem first in memory
    __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);
    return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ uint32_t pack_f16(__half2 v) { return *reinterpret_cast<uint32_t *>(&v); }
__device__ __forceinline__ void sts128u(uint32_t saddr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// byte offset of the 16-byte chunk j (0..7) of row r inside a 128B-swizzled K-major tile
__host__ __device__ __forceinline__ int sw128_offset(int r, int j) { return (r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4); }

// ---------------------------------------------------------------- per-tile geometry table
// tab[((m >> 2) * 128 + row) * 4 + (m & 3)] = d = |grid_row - mic_m| - |grid_row - array centre| (fp32 of an
// fp64 result): the bin-independent part of the steering phase.  The amplitude follows from r = r0 + d.
__device__ __forceinline__ int tab_index(int m, int row) { return (((m >> 2) * TILE_G) + row) * 4 + (m & 3); }

// The four formulations of beamforming.py:1515-1703 all have |h_m| = c_g * u(r_m) with ONE per-grid-point
// constant c_g and u = 1 (classic), r (inverse) or 1/r (true power / true location).  Because the map is
// quadratic in h, the kernels work with h' = u(r) exp(-i phi) and scale the finished power by c_g^2.
enum { AMP_ONE = 0, AMP_R = 1, AMP_RINV = 2 };
__host__ __device__ inline int amp_kind(int steering) {
    return steering == B2_STEER_CLASSIC ? AMP_ONE : (steering == B2_STEER_INVERSE ? AMP_R : AMP_RINV);
}

// Fill r for the microphones with ((m >> 2) % TPR) == part of one grid point (if `write`); returns c_g^2
// (0 for rows past the end of the grid).  Padded microphones get r = 1 (their operand columns are zero).
// *amp_scale = the power of two s_g that brings u(r) of this grid point near 1 (FP16 operands).
__device__ __forceinline__ double row_table(float *__restrict__ tab, int row, int part, bool write, int64_t g,
                                            int64_t G, const double *__restrict__ grid,
                                            const double *__restrict__ mic, int M, int Mk, double cx0, double cy0,
                                            double cz0, int steering, float *amp_scale, float *r0_out) {
    const bool valid = g < G;
    double x = 0, y = 0, z = 0, c = 0.0;
    float r0f = 1.0f;                     // fp32 distance grid point -> array centre (the reference's rt0)
    *amp_scale = 1.0f;
    if (valid) {
        x = grid[3 * g]; y = grid[3 * g + 1]; z = grid[3 * g + 2];
        double s = 0.0;
        for (int m = 0; m < M; ++m) {
            const double dx = x - mic[3 * m], dy = y - mic[3 * m + 1], dz = z - mic[3 * m + 2];
            s += 1.0 / (dx * dx + dy * dy + dz * dz);
        }
        const double r0 = das_distance(x, y, z, cx0, cy0, cz0);
        r0f = float(r0);
        if (steering != B2_STEER_CLASSIC && r0 > 0.0 && r0 < 1e30) {
            int e = ilogb(r0);                                     // r0 = 2^e * [1, 2)
            e = e < -60 ? -60 : (e > 60 ? 60 : e);
            *amp_scale = float(scalbn(1.0, steering == B2_STEER_INVERSE ? -e : e));
        }
        switch (steering) {
        case B2_STEER_CLASSIC: c = 1.0 / double(M); break;
        case B2_STEER_INVERSE: c = 1.0 / (double(M) * r0); break;
        case B2_STEER_TRUE_POWER: c = 1.0 / (r0 * s); break;
        default: c = 1.0 / sqrt(double(M) * s); break;
        }
    }
    if (write) {
        for (int m4 = part * 4; m4 < Mk; m4 += 4 * TPR) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = m4 + i;
                // d = r_m - r0 (beamforming.py:1557 etc. use exactly this difference in the phase): bounded by the array
                // aperture however far the grid point is, so the fp32 phase product q*d stays accurate; r_m itself
                // (amplitude) is recovered as r0 + d.  Padding: d with r0 + d = 1.
                double d = 1.0 - double(r0f);
                if (valid && m < M) d = das_distance(x, y, z, mic[3 * m], mic[3 * m + 1], mic[3 * m + 2]) - double(r0f);
                tab[tab_index(m, row)] = float(d);
            }
        }
    }
    *r0_out = r0f;
    return c * c;
}

// h'[g, m] = u(r) exp(-i 2 pi q d), d = r - r0, q = f / c in cycles per metre: returns (Re, Im).  (The common
// phase exp(+i 2 pi q r0) of a grid point cancels in h^H C h.)  The phase is reduced to a fraction of a cycle
// before the MUFU sincos.  |q d| is at most a few tens of cycles (d is bounded by the array aperture), so the
// fp32 product carries an absolute phase error of ~1e-5 rad per element; being uncorrelated over microphones
// it changes the map by < 1e-6 of its maximum (first-order phase perturbations cancel at a coherent peak).
template <int AMP>
__device__ __forceinline__ float2 steer(float d, float q, float r0, float us = 1.0f) {
    const float p = q * d;
    const float r = r0 + d;
    const float rn = (p + 12582912.0f) - 12582912.0f;         // round to nearest integer (|p| < 2^22)
    const float fr = p - rn;
    float sn, cs;
    __sincosf(6.28318530717958647692f * fr, &sn, &cs);
    if (AMP == AMP_ONE) return make_float2(cs, -sn);
    float u = r;
    if (AMP == AMP_RINV) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(u) : "f"(r));
    u *= us;                                                   // us = 1 except for the FP16 operands
    return make_float2(u * cs, -u * sn);
}

// top 11 significant bits by truncation: the residual x - trunc(x) is exact and is carried by the
// correction operand, so no rounding step is needed in the split modes
__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__device__ __forceinline__ void sts128(uint32_t saddr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---------------------------------------------------------------- the kernel
// Order of the tensor-core work of one bin for a PAIR of neighbouring column blocks (b0, b1 = b0 + 1).
// Chunks are walked downwards from the last one (kc = NKC-1: the widest, its MMA overwrites the
// accumulator).  Chunks kc >= kcB feed both blocks, chunks kcA <= kc < kcB feed b0 only.  The two
// accumulators are fed as two skewed streams: step i multiplies chunk NKC-1-i into b1 (first use of its A
// tile) and chunk NKC-1-(i-LAG) into b0 (last use, frees the tile).  So every steering tile is
// synthesised ONCE per bin although both blocks need it, the tensor pipe sees a steady mix of both
// streams, and the accumulator of b0 is not touched during the first LAG steps -- which is what gives the
// epilogue time to drain it from the previous bin.  A lone block is the case kcB = NKC.
//   fn(kc, target /*1: b1, 0: b0*/, first_use_of_A_tile, last_use_of_A_tile)
template <class F>
__device__ __forceinline__ void walk_pair(int NKC, int kcA, int kcB, int LAG, F &&fn) {
    const int n1 = NKC - kcB;                 // steps of the b1 stream (0 for a lone block)
    const int n0 = NKC - kcA;                 // steps of the b0 stream
    const int lag = n1 < LAG ? n1 : LAG;
    for (int i = 0; i < n0 + lag; ++i) {
        if (i < n1) fn(NKC - 1 - i, 1, true, false);
        const int j = i - lag;
        if (j >= 0) fn(NKC - 1 - j, 0, j >= n1, true);
    }
}

// work unit -> (contiguous range of bins, tile of 128 grid points); units are dealt round-robin and ordered
// bin-range-major: all CTAs work on the same range of bins at the same time, so the operand images of that
// range (sized to fit the L2) are fetched from HBM once per launch instead of once per wave of tiles
#define B2_UNIT_DECODE                                                        \
    const int upart = unit / ntiles, tile = unit - upart * ntiles;           \
    const int f_begin = int(int64_t(nbins) * upart / nsplit);                \
    const int f_end = int(int64_t(nbins) * (upart + 1) / nsplit);

template <int NB, int PASSES>
__global__ void __launch_bounds__(TC_THREADS, 1)
das_map_tc_kernel(const unsigned char *__restrict__ op, float *__restrict__ scratch,
                  const float *__restrict__ bin_scale, int M, int nbins,
                  const double *__restrict__ freqs, const double *__restrict__ weights, double inv_c,
                  const double *__restrict__ mic, double cx0, double cy0, double cz0,
                  const double *__restrict__ grid, int64_t G, int steering, int clamp_negative,
                  double *__restrict__ map_out, double *__restrict__ per_bin, int ntiles, int nsplit) {
    using CF = TcCfg<NB, PASSES>;
    constexpr int UMMA_N = CF::UMMA_N, NSA = CF::NSA, NSB = CF::NSB;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float *part = reinterpret_cast<float *>(smem + NSA * CF::A_STAGE + NSB * CF::B_STAGE);   // [2][128] epilogue partials
    uint64_t *bars = reinterpret_cast<uint64_t *>(part + 2 * TILE_G);
    uint64_t *fullA = bars, *emptyA = fullA + NSA, *fullB = emptyA + NSA, *emptyB = fullB + NSB;
    uint64_t *tfull = emptyB + NSB, *tempty = tfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef B2_MAP_TRACE
    int trace_n = 0;
#endif
    const int NKC = (M + KCH - 1) / KCH;          // K chunks
    const int Mk = NKC * KCH;
    const int NBLK = (M + NB - 1) / NB;           // column blocks
    const int akind = amp_kind(steering);
    float *tab = scratch + size_t(blockIdx.x) * TILE_G * Mk;
    // column blocks are processed in pairs that share their A tiles (CF::SHARE), else one by one
    constexpr bool SHARE = CF::SHARE;
    constexpr int GS = CF::GS;
    const int NPAIR = SHARE ? (NBLK + 1) / 2 : NBLK;
    auto pair_b0 = [&](int p) { return SHARE ? 2 * p : p; };
    auto pair_b1 = [&](int p) { return (SHARE && 2 * p + 1 < NBLK) ? 2 * p + 1 : -1; };
    auto first_chunk = [&](int blk) { return (blk * NB) / KCH; };          // first K chunk that reaches block blk

    // A ring stage: [A_hi][A_corr?]; B ring stage: [B_hi][B_corr?]
    auto a_tile = [&](int s, int split) { return smem + s * CF::A_STAGE + split * A_TILE; };
    auto b_tile = [&](int s, int split) { return smem + NSA * CF::A_STAGE + s * CF::B_STAGE + split * CF::B_TILE; };

    if (tid == 0) {
        for (int s = 0; s < NSA; ++s) { mbar_init(&fullA[s], NGEN / 32); mbar_init(&emptyA[s], 2); }   // both issue warps release
        for (int s = 0; s < NSB; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 2); }         // both issue warps release
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], NEPI / 32); }
        fence_barrier_init();
    }
    if (warp == W_MMA) tmem_alloc(tmem_slot, CF::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
#ifdef B2_MAP_PROFILE
    const long long t_kernel0 = clock64();
#endif

    if (warp < W_EPI0) {
        // =========================== steering synthesis -> A operand tiles ===========================
        // TPR threads per grid point; each handles MPT = 16 / TPR microphones of every chunk as G4 groups of 4
        const int r = tid & 127, part = tid >> 7;
        const uint32_t a_base = smem_u32(smem);
        // this thread's distances of chunk kc: G4 runs of 16 bytes, tab_index(kc * 16 + (part * G4 + gi) * 4, r)
        const float *tabp = tab + (part * G4 * TILE_G + r) * 4;
        uint32_t sa = 0, pha = 0;                            // A ring stage and its phase bit
        for (int unit = blockIdx.x; unit < ntiles * nsplit; unit += gridDim.x) {
            B2_UNIT_DECODE
            // the epilogue warps are done with the previous tile's table before it is overwritten
            asm volatile("bar.sync 1, %0;" ::"n"(NGEN + NEPI) : "memory");
            float us;                                     // power-of-two amplitude scale of this grid point
            float r0f;                                    // distance of this grid point to the array centre
            row_table(tab, r, part, true, int64_t(tile) * TILE_G + r, G, grid, mic, M, Mk, cx0, cy0, cz0, steering, &us, &r0f);
            if (PASSES != 4) us = 1.0f;
            __threadfence_block();
            asm volatile("bar.sync 1, %0;" ::"n"(NGEN + NEPI) : "memory");
            // Distances are prefetched TWO chunks ahead and the load is issued right after the arrive: the
            // proxy fence in front of the arrive is a MEMBAR that waits for every outstanding load of the
            // thread, so a load issued one chunk ahead would expose its whole L2 latency there.
            // Two register buffers alternate (the chunk loop is flat and unrolled by two); rotating one
            // buffer into the other would make the compiler wait for the newest load at the top of the body.
            float4 r_a[G4], r_b[G4];
            // cursor of the next chunk to prefetch: every pair walks kc = NKC-1 ... first chunk of its block b0
            int pf_p = 0, pf_kc = NKC - 1;
            auto prefetch = [&](float4 *dst) {
#pragma unroll
                for (int gi = 0; gi < G4; ++gi)
                    dst[gi] = __ldcg(reinterpret_cast<const float4 *>(tabp + (pf_kc * 4 + gi) * (TILE_G * 4)));
                if (pf_kc == first_chunk(pair_b0(pf_p))) { pf_kc = NKC - 1; if (++pf_p == NPAIR) pf_p = 0; }
                else --pf_kc;
            };
            prefetch(r_a);
            prefetch(r_b);
            int chunks_per_bin = 0;
            for (int p = 0; p < NPAIR; ++p) chunks_per_bin += NKC - first_chunk(pair_b0(p));
            const int total = (f_end - f_begin) * chunks_per_bin;
            int fi = f_begin, cib = 0;                       // bin of the current chunk, its index inside the bin
            float bf = float(freqs[fi] * inv_c);
            // one chunk: synthesise MPT microphones of this row, wait for the ring slot, store, signal
            auto chunk_body = [&](float4 (&r_cur)[G4]) {
                // probe the ring slot now so that the barrier round trip overlaps the arithmetic
                const bool slot_free = mbar_test(&emptyA[sa], pha ^ 1);
                if (warp == 0) TRACE(0, 0, cib, fi - f_begin);
                float rr[MPT];
#pragma unroll
                for (int gi = 0; gi < G4; ++gi) {
                    rr[4 * gi] = r_cur[gi].x; rr[4 * gi + 1] = r_cur[gi].y;
                    rr[4 * gi + 2] = r_cur[gi].z; rr[4 * gi + 3] = r_cur[gi].w;
                }
                float4 hi4[MPT / 2], lo4[MPT / 2];   // TF32 images: 2 microphones per 16 bytes
                uint4 img1[G4], img2[G4];            // 16-bit images: 4 microphones per 16 bytes
                auto synth = [&](auto amp_tag) {
                    constexpr int AMP = decltype(amp_tag)::value;
#pragma unroll
                    for (int j = 0; j < MPT / 2; ++j) {
                        const float2 h0 = steer<AMP>(rr[2 * j], bf, r0f, us);
                        const float2 h1 = steer<AMP>(rr[2 * j + 1], bf, r0f, us);
                        uint32_t *p1 = reinterpret_cast<uint32_t *>(&img1[j >> 1]) + 2 * (j & 1);
                        uint32_t *p2 = reinterpret_cast<uint32_t *>(&img2[j >> 1]) + 2 * (j & 1);
                        if (PASSES == 4) {           // img1 = first FP16 halves, img2 = second halves
                            const __half2 a1 = __floats2half2_rn(h0.x, h0.y), b1 = __floats2half2_rn(h1.x, h1.y);
                            const float2 af = __half22float2(a1), bfl = __half22float2(b1);
                            p1[0] = pack_f16(a1);
                            p1[1] = pack_f16(b1);
                            p2[0] = pack_f16(__floats2half2_rn(h0.x - af.x, h0.y - af.y));
                            p2[1] = pack_f16(__floats2half2_rn(h1.x - bfl.x, h1.y - bfl.y));
                            continue;
                        }
                        if (PASSES == 1) hi4[j] = make_float4(to_tf32(h0.x), to_tf32(h0.y), to_tf32(h1.x), to_tf32(h1.y));
                        else hi4[j] = make_float4(trunc_tf32(h0.x), trunc_tf32(h0.y), trunc_tf32(h1.x), trunc_tf32(h1.y));
                        if (PASSES == 3)   // residuals are exact in fp32; the tensor core keeps their top 11 bits
                            lo4[j] = make_float4(h0.x - hi4[j].x, h0.y - hi4[j].y, h1.x - hi4[j].z, h1.y - hi4[j].w);
                        if (PASSES == 2) {           // img1 = bf16 residuals, img2 = bf16 values
                            p1[0] = pack_bf16(h0.x - hi4[j].x, h0.y - hi4[j].y);
                            p1[1] = pack_bf16(h1.x - hi4[j].z, h1.y - hi4[j].w);
                            p2[0] = pack_bf16(h0.x, h0.y);
                            p2[1] = pack_bf16(h1.x, h1.y);
                        }
                    }
                };
                if (akind == AMP_RINV) synth(std::integral_constant<int, AMP_RINV>{});
                else if (akind == AMP_R) synth(std::integral_constant<int, AMP_R>{});
                else synth(std::integral_constant<int, AMP_ONE>{});
                if (!slot_free) { PROF_T0(); mbar_wait(&emptyA[sa], pha ^ 1); PROF_ADD(0, blockIdx.x == 0 && tid == 0); }
                if (warp == 0) TRACE(0, slot_free ? 1 : 2, cib, fi - f_begin);
                const uint32_t ahi = a_base + sa * CF::A_STAGE;
                const uint32_t alo = ahi + (CF::NSPLIT - 1) * A_TILE;
                if (PASSES != 4) {
#pragma unroll
                    for (int j = 0; j < MPT / 2; ++j) {
                        const int off = sw128_offset(r, part * (MPT / 2) + j);
                        sts128(ahi + off, hi4[j]);
                        if (PASSES == 3) sts128(alo + off, lo4[j]);
                    }
                }
                if (PASSES == 2 || PASSES == 4) {
                    // 16-bit image row = [img1 of the chunk's 16 mics | img2 of the 16 mics]: a group of 4 mics
                    // is 16 bytes = 16-byte chunk (part * G4 + gi) of each 64-byte half.  In FP16x3 mode it is
                    // the only image of the stage (alo == ahi).
#pragma unroll
                    for (int gi = 0; gi < G4; ++gi) {
                        sts128u(alo + sw128_offset(r, part * G4 + gi), img1[gi]);
                        sts128u(alo + sw128_offset(r, 4 + part * G4 + gi), img2[gi]);
                    }
                }
                // generic-proxy writes -> visible to the MMA (async proxy); one arrival per warp
                // (32 same-address arrivals serialise in the shared-memory atomic unit)
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&fullA[sa]);
                if (warp == 0) TRACE(0, 3, cib, fi - f_begin);
                if (++sa == NSA) { sa = 0; pha ^= 1; }
#ifdef B2_DEBUG_HANG
                if (warp == 0) { DBG_PROG(0, g_prog[0] + 1); }
#endif
                prefetch(r_cur);                     // the chunk after next uses this buffer again
                if (++cib == chunks_per_bin) {
                    cib = 0;
                    if (++fi < f_end) bf = float(freqs[fi] * inv_c);
                }
            };
            int i = 0;
#pragma unroll 1
            for (; i + 1 < total; i += 2) { chunk_body(r_a); chunk_body(r_b); }
            if (i < total) chunk_body(r_a);
        }
    } else if (warp < W_MMA) {
        // =========================== epilogue: TMEM -> P[g,f] -> map ===========================
        const int ew = warp - W_EPI0;
        const int quad = ew & 3, half = ew >> 2;             // TMEM lane quadrant, which half of the columns
        const int r = quad * 32 + lane;
        const uint32_t lane_base = uint32_t(quad * 32) << 16;
        constexpr int CG_PER_HALF = (NB / 32 + 1) / 2;       // 32-column groups per epilogue half
        uint32_t acc_it = 0;
        for (int unit = blockIdx.x; unit < ntiles * nsplit; unit += gridDim.x) {
            B2_UNIT_DECODE
            const int64_t g = int64_t(tile) * TILE_G + r;
            asm volatile("bar.sync 1, %0;" ::"n"(NGEN + NEPI) : "memory");
            float us;
            float r0f;
            double c2 = row_table(tab, r, 0, false, g, G, grid, mic, M, Mk, cx0, cy0, cz0, steering, &us, &r0f);
            if (PASSES == 4) c2 /= double(us);            // T carries the amplitude scale once (the epilogue's h does not)
            asm volatile("bar.sync 1, %0;" ::"n"(NGEN + NEPI) : "memory");   // table of this tile is complete
            double acc_map = 0.0;
            for (int fi = f_begin; fi < f_end; ++fi) {
                const float bf = float(freqs[fi] * inv_c);
                float p0 = 0.f, p1 = 0.f;
                for (int t = 0; t < NBLK; ++t, ++acc_it) {
                    // accumulators complete in the order b1, b0 within a pair (see walk_pair)
                    const int nbk = !SHARE ? t : ((t & 1) ? t - 1 : (t + 1 < NBLK ? t + 1 : t));
                    const int buf = acc_it & 1;
                    // this warp's groups of 8 columns (microphones) of the block: col0, col0 + 8, ...
                    const int col0 = half * CG_PER_HALF * 32;
                    const int cols_here = min(min(NB, Mk - nbk * NB) - col0, CG_PER_HALF * 32);
                    const int ngrp = cols_here > 0 ? cols_here >> 3 : 0;
                    // distances of the 8 microphones of a group for this row (two 16-byte runs of the table),
                    // fetched one group ahead -- the first group before the accumulator is even complete
                    const float *tabq = tab + tab_index(nbk * NB + col0, r);
                    float4 ra = make_float4(1.f, 1.f, 1.f, 1.f), rb = ra;
                    if (ngrp > 0) {
                        ra = __ldcg(reinterpret_cast<const float4 *>(tabq));
                        rb = __ldcg(reinterpret_cast<const float4 *>(tabq + TILE_G * 4));
                    }
                    if (ew == 0) TRACE(1, 0, nbk, fi - f_begin);
                    { PROF_T0(); mbar_wait(&tfull[buf], (acc_it >> 1) & 1); PROF_ADD(5, blockIdx.x == 0 && ew == 0 && lane == 0); }
                    if (ew == 0) TRACE(1, 1, nbk, fi - f_begin);
                    tc_fence_after();
                    const uint32_t tcol = tmem_base + lane_base + buf * UMMA_N + 2 * col0;
#pragma unroll 1
                    for (int c = 0; c < ngrp; ++c) {
                        float t[16];                                     // (Re T, Im T) of 8 columns n
                        tmem_ld16(tcol + 16 * c, t);
                        const float rq[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
                        if (c + 1 < ngrp) {
                            ra = __ldcg(reinterpret_cast<const float4 *>(tabq + (2 * c + 2) * (TILE_G * 4)));
                            rb = __ldcg(reinterpret_cast<const float4 *>(tabq + (2 * c + 3) * (TILE_G * 4)));
                        }
                        tmem_ld_wait();
                        auto reduce = [&](auto amp_tag) {
                            constexpr int AMP = decltype(amp_tag)::value;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                // padded microphones have all-zero operand rows, so T is 0 there: no bounds test
                                const float2 h0 = steer<AMP>(rq[2 * j], bf, r0f), h1 = steer<AMP>(rq[2 * j + 1], bf, r0f);
                                p0 = fmaf(t[4 * j], h0.x, p0);
                                p0 = fmaf(-t[4 * j + 1], h0.y, p0);
                                p1 = fmaf(t[4 * j + 2], h1.x, p1);
                                p1 = fmaf(-t[4 * j + 3], h1.y, p1);
                            }
                        };
                        if (akind == AMP_RINV) reduce(std::integral_constant<int, AMP_RINV>{});
                        else if (akind == AMP_R) reduce(std::integral_constant<int, AMP_R>{});
                        else reduce(std::integral_constant<int, AMP_ONE>{});
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[buf]);
                    if (ew == 0) TRACE(1, 2, nbk, fi - f_begin);
                    if (ew == 0) DBG_PROG(1, acc_it + 1);
                }
                // combine the two column halves of this row
                float *pp = part + (fi & 1) * TILE_G;
                if (half == 1) pp[r] = p0 + p1;
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (half == 0) {
                    double p = (double(p0 + p1) + double(pp[r])) * c2;        // |h|^2 scale of this grid point
                    if (PASSES == 4) p /= double(bin_scale[fi]);                    // power-of-two scale of L_f
                    if (clamp_negative && p < 0.0) p = 0.0;                         // beamforming.py:871-872
                    acc_map += weights[fi] * p;                                       // :875-876
                    if (per_bin != nullptr && g < G) per_bin[g * nbins + fi] = p;
                }
            }
            // the bins of a tile may be split over several work units: the map is zeroed by the
            // launcher and every unit adds its weighted partial sum
            if (half == 0 && g < G) atomicAdd(map_out + g, acc_map);
        }
    } else if (warp == W_MMA || warp == W_MMA + 1) {
        // =========================== MMA issue ===========================
        // Two issue warps, one per accumulator stream of a pair (warp W_MMA: block b0, warp W_MMA+1: block
        // b1): a single issue thread needs ~800 cycles of dependent instructions per step (three barrier
        // waits, descriptors, elect, 6 MMAs, commits) and was the pacing resource of the kernel.  Both warps
        // walk the same step sequence; each acts on the steps of its own stream.  Every A and B tile is released
        // by BOTH warps (emptyA / emptyB count 2): the warp that does not multiply a tile arrives on it directly.
        const int my = warp - W_MMA;
        // The whole warp walks the loop, so every address, descriptor and flag is warp-uniform and lives in
        // uniform registers; one elected lane issues the tcgen05 instructions.  (Running the loop inside an
        // `if (lane == 0)` made every tcgen05.mma cost ~13 instructions of register shuffling, and the issue
        // thread -- not the tensor pipe -- set the pace of the kernel: time grew linearly with the number of
        // MMA instructions, 110 cycles each.)
        const uint32_t a0 = desc_lo(smem_u32(smem)), b0d = desc_lo(smem_u32(smem) + NSA * CF::A_STAGE);
        uint32_t seqA = 0;                                   // A tiles consumed so far (ring sequence number)
        uint32_t sb = 0, phb = 0, acc_it = 0;
        for (int unit = blockIdx.x; unit < ntiles * nsplit; unit += gridDim.x) {
            B2_UNIT_DECODE
            (void)tile;
            for (int fi = f_begin; fi < f_end; ++fi) {
                for (int p = 0; p < NPAIR; ++p) {
                    const int blk0 = pair_b0(p), blk1 = pair_b1(p);
                    const int kcA = first_chunk(blk0), kcB = blk1 >= 0 ? first_chunk(blk1) : NKC;
                    // accumulators are handed out round-robin in completion order: b1 (if any), then b0
                    const uint32_t acc1 = acc_it, acc0 = acc_it + (blk1 >= 0 ? 1u : 0u);
                    acc_it = acc0 + 1;
                    const uint32_t seq0 = seqA;                  // ring sequence number of chunk NKC-1
                    seqA += uint32_t(NKC - kcA);
                    walk_pair(NKC, kcA, kcB, GS, [&](int kc, int tgt, bool first_use, bool last_use) {
                        (void)last_use;
                        DBG_PROG(2 + my, g_prog[2 + my] + 1);
                        if (tgt != my) {
                            // the other warp's step: this warp only has to release a tile it never multiplies
                            // (a b0-only chunk seen by the b1 warp) -- after the tile's phase is complete, so
                            // that the arrival cannot land in an earlier phase
                            if (my == 1 && first_use) {
                                const uint32_t seq = seq0 + uint32_t(NKC - 1 - kc);
                                const uint32_t sa = seq % NSA, pha = (seq / NSA) & 1;
                                mbar_wait(&fullA[sa], pha);
                                if (elect_one_sync()) mbar_arrive(&emptyA[sa]);
                                __syncwarp();
                            }
                            // The B tile of the step is not used by this warp, but the warp still waits for it
                            // and releases it (emptyB counts 2): a stage is only recycled after BOTH warps have
                            // passed it, so neither can get a whole ring round away from the other -- mbarrier
                            // waits only know the parity of a phase, and a warp one round ahead (or two behind)
                            // would mistake another phase for its own (scripts/sim_protocol.py reproduces the
                            // deadlock that follows).
                            mbar_wait(&fullB[sb], phb);
                            if (elect_one_sync()) mbar_arrive(&emptyB[sb]);
                            __syncwarp();
                            if (++sb == NSB) { sb = 0; phb ^= 1; }
                            return;
                        }
                        const uint32_t acc = tgt ? acc1 : acc0;
                        const uint32_t buf = acc & 1;
                        TRACE(2 + my, 0, kc, fi - f_begin);
                        if (kc == NKC - 1) {                     // first MMA of this accumulator: buffer drained?
                            PROF_T0(); mbar_wait(&tempty[buf], ((acc >> 1) & 1) ^ 1); PROF_ADD(3, blockIdx.x == 0 && lane == 0 && my == 0);
                        }
                        const uint32_t d_tmem = tmem_base + buf * UMMA_N;
                        const int kbase = tgt ? kcB : kcA;
                        // columns of this block that chunk kc can reach (lower triangle): w, N = 2w
                        const int w = min(NB, KCH * (kc - kbase + 1));
                        const uint32_t seq = seq0 + uint32_t(NKC - 1 - kc);
                        const uint32_t sa = seq % NSA, pha = (seq / NSA) & 1;
                        { PROF_T0(); mbar_wait(&fullB[sb], phb); PROF_ADD(7, blockIdx.x == 0 && lane == 0 && my == 0); }
                        { PROF_T0(); mbar_wait(&fullA[sa], pha); PROF_ADD(2, blockIdx.x == 0 && lane == 0 && my == 0); }
                        tc_fence_after();
                        TRACE(2 + my, 1, kc, fi - f_begin);
                        DBG_PROG(8 + my, 1000 * kc + 10 * tgt + 1);
                        // low words of the operand descriptors: first image, second (correction) image
                        const uint32_t a1 = a0 + sa * (CF::A_STAGE >> 4), b1 = b0d + sb * (CF::B_STAGE >> 4);
                        const uint32_t a2 = a1 + (CF::NSPLIT - 1) * (A_TILE >> 4), b2 = b1 + (CF::NSPLIT - 1) * (CF::B_TILE >> 4);
                        const uint32_t first = kc != NKC - 1 ? 1u : 0u;    // 0: the first MMA overwrites the accumulator
                        if (elect_one_sync()) {
                            if (PASSES == 4) {
                                // row = [f1 | f2] (A) resp. [g1 | g2] (B), 64 B each = 2 K-steps of 16 halfs:
                                // f1*g1 + f1*g2 + f2*g1
                                const uint32_t idh = idesc_f16(TILE_G, 2 * w);
#pragma unroll
                                for (uint32_t ks = 0; ks < 2; ++ks) {
                                    const uint32_t adv = ks * 2;               // (32 B) >> 4
                                    // the f1 slice of A is read from shared memory once for both products
                                    umma_f16_afill(d_tmem, desc64(a1 + adv), desc64(b1 + adv), idh, ks > 0 ? 1u : first);
                                    umma_f16_alastuse(d_tmem, desc64(a1 + adv), desc64(b1 + 4 + adv), idh, 1u);
                                    umma_bf16(d_tmem, desc64(a1 + 4 + adv), desc64(b1 + adv), idh, 1u);
                                }
                            } else {
                                const uint32_t idesc = idesc_tf32(TILE_G, 2 * w);
#pragma unroll
                                for (uint32_t ks = 0; ks < 4; ++ks) {          // UMMA K = 8 tf32 = 32 bytes
                                    const uint32_t adv = ks * 2;
                                    umma_tf32(d_tmem, desc64(a1 + adv), desc64(b1 + adv), idesc, ks > 0 ? 1u : first);
                                    if (PASSES == 3) {
                                        umma_tf32(d_tmem, desc64(a2 + adv), desc64(b1 + adv), idesc, 1u);
                                        umma_tf32(d_tmem, desc64(a1 + adv), desc64(b2 + adv), idesc, 1u);
                                    }
                                    // mixed mode: both corrections as one BF16 product of depth 64 (4 x K=16)
                                    if (PASSES == 2)
                                        umma_bf16(d_tmem, desc64(a2 + adv), desc64(b2 + adv), idesc_bf16(TILE_G, 2 * w), 1u);
                                }
                            }
                            // tcgen05.commit: the barriers are signalled when the MMAs issued so far retire
                            umma_commit(&emptyA[sa]);                      // this warp is done with the A tile
                            umma_commit(&emptyB[sb]);
                            if (kc == kbase) umma_commit(&tfull[buf]);     // accumulator complete -> epilogue
                        }
                        __syncwarp();
                        TRACE(2 + my, 2, kc, fi - f_begin);
                        DBG_PROG(8 + my, 1000 * kc + 10 * tgt + 2);
                        if (++sb == NSB) { sb = 0; phb ^= 1; }
                    });
                }
            }
        }
    } else {
        // =========================== B operand loader (one thread) ===========================
        if (lane == 0) {
            uint32_t sb = 0, phb = 0;
            const size_t bin_bytes = size_t(NBLK) * NKC * CF::NSPLIT * CF::B_TILE;
            for (int unit = blockIdx.x; unit < ntiles * nsplit; unit += gridDim.x) {
                B2_UNIT_DECODE
                (void)tile;
                for (int fi = f_begin; fi < f_end; ++fi) {
                    for (int p = 0; p < NPAIR; ++p) {
                        const int blk0 = pair_b0(p), blk1 = pair_b1(p);
                        const int kcA = first_chunk(blk0), kcB = blk1 >= 0 ? first_chunk(blk1) : NKC;
                        // the tiles are consumed in exactly the order of the MMA warp's walk
                        walk_pair(NKC, kcA, kcB, GS, [&](int kc, int tgt, bool, bool) {
                            const int blk = tgt ? blk1 : blk0;
                            const int w = min(NB, KCH * (kc - (tgt ? kcB : kcA) + 1));
                            const uint32_t bytes = uint32_t(2 * w) * 128;          // first 2w rows of each image
                            mbar_wait(&emptyB[sb], phb ^ 1);
                            const unsigned char *src = op + size_t(fi) * bin_bytes + (size_t(blk) * NKC + kc) * (CF::NSPLIT * CF::B_TILE);
                            mbar_arrive_expect_tx(&fullB[sb], CF::NSPLIT * bytes);
                            bulk_g2s(b_tile(sb, 0), src, bytes, &fullB[sb]);
                            if (CF::NSPLIT == 2) bulk_g2s(b_tile(sb, 1), src + CF::B_TILE, bytes, &fullB[sb]);
                            if (++sb == NSB) { sb = 0; phb ^= 1; }
#ifdef B2_DEBUG_HANG
                            if (blockIdx.x == 0) g_prog[4] += 1;
#endif
                        });
                    }
                }
            }
        }
    }

#ifdef B2_MAP_PROFILE
    if (blockIdx.x == 0) {
        const unsigned long long dt = (unsigned long long)(clock64() - t_kernel0);
        if (tid == 0) atomicAdd(&g_map_prof[1], dt);
        if (warp == W_MMA && lane == 0) atomicAdd(&g_map_prof[4], dt);
        if (warp == W_EPI0 && lane == 0) atomicAdd(&g_map_prof[6], dt);
    }
#endif
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == W_MMA) tmem_dealloc(tmem_base, CF::TMEM_COLS);
}

// ---------------------------------------------------------------- K3a for the tensor-core path
// FP16x3 mode: exact power-of-two scale per bin that maps max |L_f| into [2^11, 2^12) (one CTA per bin).
__global__ void das_bin_scale_kernel(const double2 *__restrict__ csm, int M, int bin0, double gain, int remove_diag,
                                     float *__restrict__ scale) {
    const double2 *c = csm + int64_t(bin0 + blockIdx.x) * M * M;
    double mx = 0.0;
    for (int i = threadIdx.x; i < M * M; i += blockDim.x) {
        const int m = i / M, n = i - m * M;
        if (n > m) continue;
        const double f = n == m ? (remove_diag ? 0.0 : gain) : 2.0 * gain;
        const double v = f * fmax(fabs(c[i].x), fabs(c[i].y));
        if (v < 1e300) mx = fmax(mx, v);                     // ignores inf / nan
    }
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < int(blockDim.x >> 5); ++w) mx = fmax(mx, red[w]);
        int k = 0;
        if (mx > 0.0) {
            k = 11 - ilogb(mx);                              // mx * 2^k in [2^11, 2^12)
            k = k < -100 ? -100 : (k > 100 ? 100 : k);
        }
        scale[blockIdx.x] = float(scalbn(1.0, k));
    }
}

// One thread per 16-byte chunk of an operand image: (bin, column block, K chunk, row, chunk j).
template <int NB, int PASSES>
__global__ void das_prepare_tc_kernel(const double2 *__restrict__ csm, int M, int bin0, double gain, int remove_diag,
                                      unsigned char *__restrict__ op, int NKC, int NBLK,
                                      const float *__restrict__ bin_scale) {
    using CF = TcCfg<NB, PASSES>;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;       // 0 .. 2NB-1
    if (row >= CF::UMMA_N) return;
    const int kc = blockIdx.y % NKC, nbk = blockIdx.y / NKC;
    const int fi = blockIdx.z;
    const bool imag_row = row & 1;                                // rows alternate Re T / Im T of column n
    const int n = nbk * NB + (row >> 1);
    unsigned char *tile = op + ((size_t(fi) * NBLK + nbk) * NKC + kc) * size_t(CF::NSPLIT * CF::B_TILE);
    if (PASSES == 4) gain *= double(bin_scale[fi]);
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
        if (PASSES == 4) {
            // row = [g1 of the chunk's 16 mics | g2 of the 16 mics] (FP16): the 2 mics of this j are 8 bytes
            // at byte offset 8*j of each 64-byte part
            const __half2 a1 = __floats2half2_rn(v[0], v[1]), b1 = __floats2half2_rn(v[2], v[3]);
            const float2 af = __half22float2(a1), bfl = __half22float2(b1);
            const uint2 g1 = make_uint2(pack_f16(a1), pack_f16(b1));
            const uint2 g2 = make_uint2(pack_f16(__floats2half2_rn(v[0] - af.x, v[1] - af.y)),
                                        pack_f16(__floats2half2_rn(v[2] - bfl.x, v[3] - bfl.y)));
            *reinterpret_cast<uint2 *>(tile + sw128_offset(row, j >> 1) + 8 * (j & 1)) = g1;
            *reinterpret_cast<uint2 *>(tile + sw128_offset(row, 4 + (j >> 1)) + 8 * (j & 1)) = g2;
            continue;
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
    return int64_t(nbins) * NBLK * NKC * ((PASSES == 1 || PASSES == 4) ? 1 : 2) * btile;
}
// operand buffer = [images][one distance table of 128 x Mk floats per persistent CTA][per-bin scales]
inline int64_t tables_bytes(int M) { return int64_t(MAX_CTAS) * TILE_G * ((M + KCH - 1) / KCH) * KCH * 4; }
template <int PASSES> float *tables_ptr(const void *op, int M, int nbins) {
    return reinterpret_cast<float *>((unsigned char *)op + images_bytes<PASSES>(M, nbins));
}
template <int PASSES> float *scales_ptr(const void *op, int M, int nbins) {
    return reinterpret_cast<float *>((unsigned char *)op + images_bytes<PASSES>(M, nbins) + tables_bytes(M));
}

template <int NB, int PASSES>
int launch_prepare(const void *csm, int M, int bin0, int nbins, int remove_diag, void *op, cudaStream_t st) {
    using CF = TcCfg<NB, PASSES>;
    const int NKC = (M + KCH - 1) / KCH, NBLK = (M + NB - 1) / NB;
    const double gain = remove_diag ? double(M) / double(M - 1) : 1.0;
    dim3 grid((CF::UMMA_N + 127) / 128, NKC * NBLK, nbins);
    float *scales = scales_ptr<PASSES>(op, M, nbins);
    if (PASSES == 4) {
        das_bin_scale_kernel<<<nbins, 256, 0, st>>>((const double2 *)csm, M, bin0, gain, remove_diag, scales);
        B2_LAUNCH_OK("b2_das_prepare(scale)");
    }
    das_prepare_tc_kernel<NB, PASSES><<<grid, 128, 0, st>>>((const double2 *)csm, M, bin0, gain, remove_diag,
                                                            (unsigned char *)op, NKC, NBLK, scales);
    B2_LAUNCH_OK("b2_das_prepare(tc)");
    return B2_OK;
}

template <int NB, int PASSES>
int launch_map(const void *op, int M, int nbins, const void *freqs, const void *weights, double c, const void *mic,
               const double *center, const void *grid, int64_t G, int steering, int clamp_negative, void *map_out,
               void *per_bin, cudaStream_t st) {
    using CF = TcCfg<NB, PASSES>;
    auto kern = das_map_tc_kernel<NB, PASSES>;
    static DeviceOnce once;          // per instantiation and device
    int num_sms = 0;
    if (int rc = b2_device_once(once, kern, CF::SMEM, &num_sms)) return rc;
    const int64_t ntiles = (G + TILE_G - 1) / TILE_G;
    B2_REQUIRE(ntiles <= 0x7fffffff, B2_EINVAL, "b2_das_map: too many grid points");
    // Persistent grid, one CTA per SM.  The bins are split into nsplit ranges and a work unit is (bin range,
    // tile); partial maps are combined with atomicAdd.  Two reasons to split:
    //  * L2 reuse: a range whose operand images fit the L2 (~64 MB of its 126 MB) is streamed from HBM once
    //    and then served from L2 to every tile.  Unsplit, the persistent CTAs drift apart over the 2048 bins
    //    of their tiles and each re-reads the operand from HBM (ncu: 928 GB per launch at 10^6 points).
    //  * tail: when the tiles do not fill the last round of CTAs evenly (e.g. 977 tiles on 148 SMs: 60 % of
    //    the SMs idle in round 7), up to 4 ranges shrink the idle tail.
    int cap = num_sms < MAX_CTAS ? num_sms : MAX_CTAS;
    int nsplit = 1;
    double best = 0.0;
    for (int sp = 1; sp <= 4 && nbins / sp >= 32; ++sp) {
        const int64_t units = ntiles * sp;
        const int64_t rounds = (units + cap - 1) / cap;
        const double eff = double(units) / double(rounds * cap);
        if (eff > best + 0.01) { best = eff; nsplit = sp; }
    }
    {
        const char *e = getenv("B2_MAP_L2_CHUNK_MB");      // tuning knob, read per launch (0 disables)
        const double chunk_mb = e ? atof(e) : 64.0;
        if (chunk_mb > 0.0 && ntiles > cap) {              // more than one wave of tiles: reuse pays
            const double total_mb = double(images_bytes<PASSES>(M, nbins)) / 1048576.0;
            int by_l2 = int(total_mb / chunk_mb + 0.999);
            const int most = nbins / 32 > 0 ? nbins / 32 : 1;      // at least 32 bins per unit
            if (by_l2 > most) by_l2 = most;
            if (by_l2 > nsplit) nsplit = by_l2;
        }
    }
    if (const char *e = getenv("B2_MAP_NSPLIT")) {         // test knob: force the number of bin ranges
        const int forced = atoi(e);
        if (forced >= 1) nsplit = forced < nbins ? forced : nbins;
    }
    const int64_t units = ntiles * nsplit;
    B2_REQUIRE(units <= 0x7fffffff, B2_EINVAL, "b2_das_map: too many work units (%lld tiles x %d bin ranges)",
               (long long)ntiles, nsplit);
    const int nblocks = int(units < cap ? units : cap);
    B2_CUDA_OK(cudaMemsetAsync(map_out, 0, size_t(G) * sizeof(double), st));
    float *scratch = tables_ptr<PASSES>(op, M, nbins);
    kern<<<nblocks, TC_THREADS, CF::SMEM, st>>>((const unsigned char *)op, scratch, scales_ptr<PASSES>(op, M, nbins), M, nbins,
                                                (const double *)freqs,
                                                (const double *)weights, 1.0 / c, (const double *)mic, center[0],
                                                center[1], center[2], (const double *)grid, G, steering,
                                                clamp_negative, (double *)map_out, (double *)per_bin, int(ntiles), nsplit);
    B2_LAUNCH_OK("b2_das_map(tc)");
    return B2_OK;
}

template <int PASSES> int64_t operand_bytes(int M, int nbins) {
    return images_bytes<PASSES>(M, nbins) + tables_bytes(M) + ((int64_t(nbins) * 4 + 255) & ~int64_t(255));
}

}  // namespace

#ifdef B2_DEBUG_HANG
extern "C" int b2_debug_hang_report(unsigned int *out) {      // out[0] = count, then 4 words per record
    unsigned int n = 0, z = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&n, g_nhang, 4);
    out[0] = n;
    cudaMemcpyFromSymbol(out + 1, g_hang, sizeof(unsigned int) * 240);
    cudaMemcpyFromSymbol(out + 241, g_snap, sizeof(unsigned int) * 16);
    unsigned int zz[16] = {0};
    cudaMemcpyToSymbol(g_prog, zz, sizeof(zz));
    cudaMemcpyToSymbol(g_snap, zz, sizeof(zz));
    cudaMemcpyToSymbol(g_snapped, &z, 4);
    cudaMemcpyToSymbol(g_nhang, &z, 4);
    cudaMemcpyToSymbol(g_abort, &z, 4);
    return 0;
}
#endif

#ifdef B2_MAP_TRACE
extern "C" int b2_debug_trace(long long *out, int max_records) {      // copies [4 roles][1024]{tag, clock}
    (void)max_records;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * 4 * 2 * 1024);
    static long long zeros[4 * 2 * 1024];
    cudaMemcpyToSymbol(g_trace, zeros, sizeof(zeros));
    return 4096;
}
#endif

#ifdef B2_MAP_PROFILE
extern "C" int b2_debug_map_profile(unsigned long long *out) {
    cudaMemcpyFromSymbol(out, g_map_prof, sizeof(unsigned long long) * 8);
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(g_map_prof, z, sizeof(z));
    return 0;
}
#endif

int64_t b2_das_tc_operand_bytes(int M, int nbins, int mode) {
    if (M > MAX_MICS) { b2_set_error("tensor-core map kernel supports at most %d microphones", MAX_MICS); return B2_EUNSUPPORTED; }
    if (mode == B2_MAP_TC_TF32) return operand_bytes<1>(M, nbins);
    if (mode == B2_MAP_TC_F16X3) return operand_bytes<4>(M, nbins);
    return operand_bytes<3>(M, nbins);
}

#define B2_TC_DISPATCH(FN, ...)                                                      \
    do {                                                                             \
        const int nb_ = tc_nb(M);                                                    \
        if (mode == B2_MAP_TC_TF32X3) {                                              \
            if (nb_ == 32) return FN<32, 3>(__VA_ARGS__);                            \
            if (nb_ == 64) return FN<64, 3>(__VA_ARGS__);                            \
            return FN<128, 3>(__VA_ARGS__);                                          \
        } else if (mode == B2_MAP_TC_F16X3) {                                        \
            if (nb_ == 32) return FN<32, 4>(__VA_ARGS__);                            \
            if (nb_ == 64) return FN<64, 4>(__VA_ARGS__);                            \
            return FN<128, 4>(__VA_ARGS__);                                          \
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
