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
//     A[g, 2m+{0,1}]   = (Re, Im) of H[g, m] = a_gm exp(-i k_f r_gm)   synthesised straight into the
//                        128B-swizzled K-major shared-memory operand tile,
//     B[2n,     2m+{0,1}] = ( Re L[m,n],  Im L[m,n])                    -> D[:, 2n]     =  Re T
//     B[2n + 1, 2m+{0,1}] = (-Im L[m,n],  Re L[m,n])                    -> D[:, 2n + 1] = -Im T
// (operand images pre-swizzled by b2_das_prepare and streamed with cp.async.bulk), accumulated
// in TMEM as fp32.  [This interleaved form is what the TF32 modes use; the default FP16x3 mode uses the PLANAR
// form described at TcCfg: rows [Re | Im] per FP16 half, one B row per matrix column, Re T and Im T as separate
// column ranges of the accumulator, the minus sign of (-Hi) Lr^T in the instruction descriptor.]  L is lower triangular: K chunks (16 microphones) entirely above the diagonal
// are skipped, and a chunk that straddles it only touches its first w = 16, 32, ... columns, so
// the MMA is issued with N = 2w and only the first 2w operand rows are fetched.  Chunks are
// walked from the last one down, so that the first MMA of an accumulator (the one that overwrites
// instead of accumulating) is the widest; two neighbouring column blocks share their A tiles
// (one MMA stream per block, see the issue role).
//
// There is NO separate epilogue role.  Columns [16 kc, 16 kc + 16) of T are final as soon as chunk kc
// itself has been multiplied in (chunks are walked downwards, L is lower triangular), so the thread that
// synthesised H[g, m] for the 16 microphones of chunk kc keeps those values in registers and, when the
// tensor core has retired that chunk's own step (mbarrier `tdone`), reads (Re T, -Im T) of exactly those
// microphones from TMEM with tcgen05.ld and accumulates
//     P[g, f] = sum_n Re(T[g, n] H[g, n]),   then per bin: clamp,   map[g] += w_f P[g, f]
// -- every steering element is synthesised once per bin instead of once for the operand and once more for
// the reduction.
//
// Warp roles: NG groups of 4 synthesis+reduction warps (a group covers the 128 TMEM lanes; warp w owns lanes
// 32 (w % 4) ...), 2 MMA issue warps (one per accumulator stream; the first also owns the TMEM allocation),
// 2 operand loader threads (one B ring per stream).  An issue warp takes two chunks per step (all their barriers in
// one round trip, then 24 MMAs).  The groups take the chunks ROUND-ROBIN: group j synthesises the
// whole tile of chunks j, j + NG, ... (16 microphones per thread) and reduces exactly those chunks DLAG own chunks
// later.  Every per-chunk hand-shake (ring slot wait, proxy fence, barrier arrival, tdone wait, TMEM load) is a
// serial latency of the warp that executes it; with every warp taking part in every chunk (the first two
// generations of this kernel) that latency, ~1400 cycles, WAS the chunk period.
// Pipelines: shared-memory rings with full/empty mbarriers -- A (synthesis -> MMA) and one B ring per stream
// (loader -> MMA), released by tcgen05.commit --, a ring of `tdone` mbarriers indexed by the chunk sequence number
// (MMA -> reduction of that chunk's columns), one `tempty` mbarrier per TMEM accumulator (all its columns read ->
// the next bin may overwrite it: a group releases an accumulator right after its last own chunk of the block) and
// a non-blocking hand-off of the per-bin partial sums between the groups.  setmaxnreg gives the synthesis
// warpgroups the registers the issue / loader warpgroup does not need.  The bin-independent part of the steering
// vector (the distance difference r0 - r_m of every (grid point, microphone) pair of the tile) is computed once per
// work unit in fp64 and parked in an L2-resident scratch table, so the per-bin work is one phase product, one
// sincos, one rcp and the operand split, in packed fp32 arithmetic (FMUL2 / FADD2 / FFMA2) two microphones at a time.
// scripts/sim_protocol.py models the barrier protocol.
//
// Split-TF32 (PASSES = 3): every operand is split x = hi + lo (both TF32-exact) and
// hi*hi + lo*hi + hi*lo is accumulated, which restores ~fp32 accuracy.
// Mixed mode (PASSES = 2): hi*hi stays a TF32 product, but the two correction products only need
// ~8 significant bits, so they run as ONE kind::f16 (BF16) product of twice the depth at twice
// the TF32 rate: the second operand image of a stage holds, per row, [bf16(x - hi) | bf16(x)] for
// A and [bf16(y) | bf16(y - hi)] for B, so that sum_k A2[k] B2[k] = lo_A*y + x*lo_B.  Tensor work
// drops from 3 to 2 TF32-equivalents per product at ~2^-20 relative accuracy.
// FP16x3 mode (PASSES = 4, the default): both operands are split into two FP16 halves x = f1 + f2 (22
// significant bits) and f1*g1 + f1*g2 + f2*g1 is accumulated as three kind::f16 products: 1.5
// TF32-equivalents of tensor work and operand rows of 128 B per 16 microphones instead of 256 B.  FP16's
// narrow exponent is handled by exact power-of-two scalings: one per bin for L (max |L| -> [2^11, 2^12))
// and one per grid point for the amplitude u(r), both divided out of the finished power in fp64.
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
#ifndef B2_NG
#define B2_NG 3
#endif
static_assert(B2_NG >= 1 && B2_NG <= 4, "1..4 groups of 4 warps");

constexpr int MAX_CTAS = 160;                 // scratch tables are sized for this many persistent CTAs
constexpr int MAX_MICS = 1024;
constexpr int PROG_MAX = 320;                 // chunk positions per bin (M = 1024, unpaired blocks: 64 + 56 + ... + 8 = 288)

__host__ __device__ constexpr int tc_nb(int M) { return M <= 32 ? 32 : (M <= 64 ? 64 : 128); }

template <int NB, int PASSES> struct TcCfg {
    static constexpr int UMMA_N = 2 * NB;
    // PLANAR complex layout (FP16x3 mode): operand rows hold [Re of the chunk's 16 microphones | Im of them] per
    // FP16 half instead of interleaved (Re, Im) pairs, and the B image has ONE row per matrix column n (not a Re-T and
    // an Im-T row that carry the same numbers): Re T and Im T come from different K slices of the same rows,
    //   Re T += Hr Lr^T + Hi Li^T,   Im T += Hr Li^T + (-Hi) Lr^T   (A negated by the instruction descriptor),
    // as twice as many MMAs of half the width -- the same tensor work for HALF the B bytes.  The operand stream from
    // L2 was what bound the kernel: 148 CTAs x 557 KB per (tile, bin) = 5.2 TB/s at 25 K cycles per bin.
    static constexpr bool PLANAR = PASSES == 4;
    static constexpr int B_ROWS = PLANAR ? NB : UMMA_N;
    static constexpr int B_TILE = B_ROWS * 128;
    static constexpr int NSPLIT = (PASSES == 1 || PASSES == 4) ? 1 : 2; // operand images per stage (hi[, correction])
    // Two independent shared-memory rings: A (synthesised steering tiles, small stages) and B (streamed
    // operand images, large stages).  The A ring is made as deep as the 227 KB allow, because the
    // synthesis warps and the tensor pipe have different per-chunk costs (the MMA width varies with the
    // chunk) and a deep ring is what lets them overlap instead of waiting for each other.
    static constexpr int A_STAGE = NSPLIT * A_TILE;
    static constexpr int B_STAGE = NSPLIT * B_TILE;
    // A ring and two B rings.  Each of the two MMA issue warps (stream 0: block b0 of a pair, stream 1: block b1)
    // has its OWN B ring and loader thread: with one shared ring in a fixed step order the loader could not fetch
    // the first tile of the next bin for one warp before the other warp had consumed its last tiles of this bin,
    // which chained the 24 steps of a bin into one serial sequence (~1000 cycles each: role profile, r2).
    static constexpr bool SHARE = A_STAGE <= 16 * 1024 && B_STAGE <= 32 * 1024;   // pairs of column blocks share their A tiles
#ifdef B2_NSB0
    static constexpr int NSB0 = B2_NSB0, NSB1 = B2_NSB1;
#else
    static constexpr int NSB0 = !SHARE ? 2 : (B_STAGE <= 16 * 1024 ? 4 : 2);
    static constexpr int NSB1 = !SHARE ? 0 : (B_STAGE <= 16 * 1024 ? 3 : 2);
#endif
    static constexpr int NSB = NSB0 + NSB1;
    static constexpr int NSA_RAW = (224 * 1024 - NSB * B_STAGE) / A_STAGE;
#ifdef B2_NSA
    static constexpr int NSA = B2_NSA;
#else
    static constexpr int NSA = NSA_RAW > 10 ? 10 : NSA_RAW;
#endif
    // Synthesis/reduction warp groups (chunks are dealt round-robin: group j synthesises chunks j, j + NG, ...).  A
    // group reduces a chunk DLAG of its OWN chunks after it synthesised it (it keeps the h' of DLAG chunks in
    // registers): synthesis -> fullA -> MMA issue -> tensor pipe -> tdone -> TMEM load is a dependent chain of
    // ~2000 cycles.  With DLAG = 1 a group reduces chunk i - NG right before it synthesises chunk i, i.e. right
    // after it synthesised chunk i - NG: it sits out that whole chain every time (ncu, r2: 31 % of the synthesis
    // warps' samples on the tdone wait, tensor pipe 54 % busy).  With DLAG = 2 the chain has one own chunk
    // (~2500 cycles of synthesis) to complete.
    // Before it synthesises chunk i a group passes the chunks (i - (DLAG + 1) NG, i - DLAG NG] of the event stream,
    // waiting on `tdone` of its own chunk and of any block's lowest chunk in that range.  mbarrier waits only see
    // the PARITY of a phase, so a tdone barrier must not complete a further use while a group still waits on it:
    // the tdone barriers form their OWN ring of TD slots, indexed by the chunk sequence number.  The tensor pipe
    // cannot run further ahead of the slowest group than the end of the accumulator after the one that group
    // is reading (tempty), i.e. less than 2 NKC <= TD chunks (scripts/sim_protocol.py models it).  The A ring only
    // has to cover synthesis -> MMA retirement: NSA >= DLAG NG keeps the groups from starving the issue warps.
    // Measured (r2, final kernel): DLAG = 2 vs 1: 12.9 vs 13.1 ms (256 mics, 37 888 points x 512 bins), 12.1 vs 12.4 ms
    // (64 mics, 148 000 points x 640 bins); DLAG = 3 needs a deeper A ring than fits.
#ifndef B2_DLAG
#define B2_DLAG 2
#endif
    static constexpr int TD = 2 * (MAX_MICS / KCH);
    static_assert((TD & (TD - 1)) == 0, "the tdone ring is indexed with a mask");
    // NB = 32 (at most 32 microphones: one or two chunks per bin): ONE group.  With more groups than chunks per bin a
    // group has bins without any chunk of its own, has to hand over several (empty) bin sums at once, and the blocking
    // part of that hand-off can close a wait cycle with the accumulator release (scripts/sim_protocol.py found it with two
    // accumulator buffers; with four it has not shown up, but three groups have nothing to gain on one or two chunks).
    static constexpr int NG = NB == 32 ? 1 : ((B2_NG <= NSA && B2_DLAG * B2_NG <= NSA) ? B2_NG : (NSA >= 4 ? 2 : 1));
    static constexpr int DLAG = B2_DLAG * NG <= NSA ? B2_DLAG : 1;
    static_assert(NG <= NSA && DLAG * NG <= NSA, "the A ring is too shallow for the chunks the groups keep in flight");
    static constexpr int NGEN = TILE_G * NG;                           // warp w owns TMEM lanes 32 (w % 4) ...
    static constexpr int NS_WARPS = NGEN / 32;
    static constexpr int W_MMA = NS_WARPS;            // two MMA issue warps, then their two B loader warps
    static constexpr int THREADS = NGEN + 128;
    static constexpr int PARTIAL_ROWS = NG > 1 ? NG - 1 : 1;           // per-bin partial sums handed to the first group
    // accumulator buffers: as many as the 512 TMEM columns hold, at most 4 (two for 128-column blocks; four for the
    // small-array kernels, whose bins are only a few chunks long -- the issue warp then runs up to three bins ahead of
    // the reduction instead of waiting for a buffer at the start of every other bin)
    static constexpr int NBUF = 512 / UMMA_N >= 4 ? 4 : 2;
    static constexpr int TMEM_COLS = NBUF * UMMA_N;
    // Register file split (setmaxnreg): the issue / loader warpgroup keeps REGS_IO registers per thread, the synthesis
    // warpgroups take the rest of the 64 K registers -- the h' of DLAG chunks (32 registers per chunk) live there.
    static constexpr bool SETMAXNREG = NG >= 3;      // 512 threads: 128 registers each at launch (fewer groups get enough without it)
    static constexpr int REGS_IO = 48;
    static constexpr int REGS_S_RAW = ((65536 - 128 * REGS_IO) / NGEN) / 8 * 8;
    static constexpr int REGS_S = REGS_S_RAW > 232 ? 232 : REGS_S_RAW;
    static constexpr int NBARS = 2 * NSA + TD + 2 * NSB + NBUF + 8;     // ... + per-quadrant hand-off of the bin sums (full, free)
    static constexpr int SMEM = NSA * A_STAGE + NSB * B_STAGE + PARTIAL_ROWS * TILE_G * 4 + PROG_MAX * 2 + NBARS * 8 + 16 + 16;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

#ifdef B2_MAP_PROFILE
// Debug instrumentation (variant build only): clock64 cycles CTA 0 spends per role and activity.
//  issue warp w (0: b0 lower, 1: b1, 2: b0 upper), slots 8 w + { 0 total, 1 wait tempty / tinit, 2 wait fullB, 3 wait fullA,
//                                                   4 issue + commits, 5 extra A releases, 6 steps }
//  synthesis group 0 / warp 0: 24 + { 0 total, 1 wait tdone, 2 TMEM load + FMAs, 3 wait emptyA, 4 synthesis + stores, 5 fence + arrive,
//                                     6 own chunks, 7 per-unit table }
__device__ unsigned long long g_map_prof[32];
// B2_MAP_TRACE (with B2_MAP_PROFILE): the same probe points also log (slot, clock64) of CTA 0's two issue warps and of
// one warp per synthesis group, events [TRACE_SKIP, TRACE_SKIP + TRACE_LEN) of each role: a steady-state timeline.
// (The probes themselves cost a few hundred cycles per step: read the timeline for its ORDER of events.)
#ifdef B2_MAP_TRACE
constexpr int TRACE_SKIP = 6000, TRACE_LEN = 2048;
__device__ long long g_map_trace[5][TRACE_LEN][2];
#define PROF_DECL long long _pt = 0; int _tn = 0
#define TRACE_EV(slot) do { if (_tn >= TRACE_SKIP && _tn < TRACE_SKIP + TRACE_LEN) { g_map_trace[prof_role][_tn - TRACE_SKIP][0] = (slot); g_map_trace[prof_role][_tn - TRACE_SKIP][1] = clock64(); } ++_tn; } while (0)
#else
#define PROF_DECL long long _pt = 0
#define TRACE_EV(slot)
#endif
#define PROF_T0() _pt = clock64()
#define PROF_ADD(slot, cond) do { if (cond) { const long long _n = clock64(); atomicAdd(&g_map_prof[slot], (unsigned long long)(_n - _pt)); _pt = _n; TRACE_EV(slot); } } while (0)
#define PROF_CNT(slot, cond) do { if (cond) atomicAdd(&g_map_prof[slot], 1ull); } while (0)
#else
#define PROF_DECL
#define PROF_T0()
#define PROF_ADD(slot, cond)
#define PROF_CNT(slot, cond)
#endif

// ---------------------------------------------------------------- PTX wrappers
// barriers and tiles are addressed by their 32-bit shared-space addresses (computed once per thread: a
// generic -> shared conversion in front of every barrier operation was 9 % of the executed instructions)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
                 : "memory");
}
// Wait for the phase with the given parity.  try_wait suspends the thread in hardware until the
// phase completes or the time hint expires, so waiting warps do not burn issue slots spinning.
// Watchdog: a wait that lasts longer than ~4 s can only be a protocol bug; the kernel then traps (the launch
// fails with an error) instead of hanging the GPU until somebody kills the process.
// B2_WAIT_HINT: pass a suspend-time hint (ptxas then emits TRYWAIT + NANOSLEEP.SYNCS + re-check); without it the
// wait is a plain loop of hardware-bounded TRYWAITs, like CUTLASS' ClusterBarrier::wait.
#ifdef B2_WAIT_HINT
#define B2_TRY_WAIT "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
#else
#define B2_TRY_WAIT "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#endif
__device__ __forceinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
    // (inlined: a real call cannot be register-allocated in a kernel that uses setmaxnreg.)  The inner loop is
    // nothing but try_wait + branch -- the poll loops of the issue / loader warps were 15 % of all instructions the
    // kernel executed when every poll also kept a spin counter and sampled the timer every 16th time.
    unsigned long long t_begin = 0, t_now;
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            ".reg .u32 n;\n\t"
            "mov.u32 n, 4096;\n\t"
            "B2_POLL_%=:\n\t"
            B2_TRY_WAIT
            "@p bra B2_DONE_%=;\n\t"
            "sub.u32 n, n, 1;\n\t"
            "setp.ne.u32 p, n, 0;\n\t"
            "@p bra B2_POLL_%=;\n\t"
            "mov.u32 %0, 0;\n\t"
            "bra B2_OUT_%=;\n\t"
            "B2_DONE_%=:\n\t"
            "mov.u32 %0, 1;\n\t"
            "B2_OUT_%=:\n\t"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(0x989680u)
            : "memory");
        if (done) return;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
        if (t_begin == 0) t_begin = t_now;
        else if (t_now - t_begin > 4000000000ull) __trap();
    }
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {          // non-blocking: has the phase completed?
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        B2_TRY_WAIT
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
    if (!done) mbar_wait_slow(bar, parity);
}
// Several barriers at once: the try_waits are in flight together (a try_wait that succeeds at once still costs ~150
// cycles of round trip; the issue warp paid it once per barrier, one after the other).
__device__ __forceinline__ void mbar_wait2(uint32_t bar_a, uint32_t par_a, uint32_t bar_b, uint32_t par_b) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%3], %4;\n\t"
        "and.pred p, p, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar_a), "r"(par_a), "r"(bar_b), "r"(par_b)
        : "memory");
    if (!done) {
        mbar_wait(bar_a, par_a);
        mbar_wait(bar_b, par_b);
    }
}
__device__ __forceinline__ void mbar_wait4(uint32_t bar_a, uint32_t par_a, uint32_t bar_b, uint32_t par_b, uint32_t bar_c,
                                           uint32_t par_c, uint32_t bar_d, uint32_t par_d) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p, q, r, t;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%3], %4;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 r, [%5], %6;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 t, [%7], %8;\n\t"
        "and.pred p, p, q;\n\t"
        "and.pred r, r, t;\n\t"
        "and.pred p, p, r;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar_a), "r"(par_a), "r"(bar_b), "r"(par_b), "r"(bar_c), "r"(par_c), "r"(bar_d), "r"(par_d)
        : "memory");
    if (!done) {
        mbar_wait(bar_a, par_a);
        mbar_wait(bar_b, par_b);
        mbar_wait(bar_c, par_c);
        mbar_wait(bar_d, par_d);
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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
// same, FP16 x FP16; bit 13 negates the A operand
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
    return (1u << 4) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
constexpr uint32_t IDESC_NEGATE_A = 1u << 13;
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
__device__ __forceinline__ void umma_f16_ause(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::use [%0], %1, %2, %3, p;\n\t"
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
__device__ __forceinline__ void sts128(uint32_t saddr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// byte offset of the 16-byte chunk j (0..7) of row r inside a 128B-swizzled K-major tile
__host__ __device__ __forceinline__ int sw128_offset(int r, int j) { return (r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4); }

// ---------------------------------------------------------------- per-tile geometry table
// tab[((m >> 2) * 128 + row) * 4 + (m & 3)] = dn = |grid_row - array centre| - |grid_row - mic_m| = r0 - r_m (fp32 of
// an fp64 result): the bin-independent part of the steering phase, NEGATED so that the synthesis needs no sign
// flip: with p = q dn, (cos 2 pi p, sin 2 pi p) = (Re, Im) of exp(-i 2 pi q (r_m - r0)).  The amplitude follows
// from r_m = r0 - dn.
__device__ __forceinline__ int tab_index(int m, int row) { return (((m >> 2) * TILE_G) + row) * 4 + (m & 3); }

// The four formulations of beamforming.py:1515-1703 all have |h_m| = c_g * u(r_m) with ONE per-grid-point
// constant c_g and u = 1 (classic), r (inverse) or 1/r (true power / true location).  Because the map is
// quadratic in h, the kernels work with h' = u(r) exp(-i phi) and scale the finished power by c_g^2.
enum { AMP_ONE = 0, AMP_R = 1, AMP_RINV = 2 };
__host__ __device__ inline int amp_kind(int steering) {
    return steering == B2_STEER_CLASSIC ? AMP_ONE : (steering == B2_STEER_INVERSE ? AMP_R : AMP_RINV);
}

// h'[g, m] = us u(r_m) exp(-i 2 pi q (r_m - r0)) for TWO microphones at once, in packed fp32 arithmetic:
// dn = (r0 - r_A, r0 - r_B), q = f / c in cycles per metre; returns hr = (Re h'_A, Re h'_B), hi = (Im ..).  (The
// common phase exp(+i 2 pi q r0) of a grid point cancels in h^H C h.)  The phase is reduced to a fraction of a
// cycle before the MUFU sincos: |q dn| is at most a few hundred cycles (dn is bounded by the array aperture), so
// the fp32 product carries an absolute phase error of ~1e-5 rad per element; being uncorrelated over
// microphones it changes the map by < 1e-6 of its maximum.  (Feeding the unreduced phase to the MUFU instead
// was measured: 1.2e-5 abs error at 32 cycles, 1e-4 at 256 -- scripts/probe_mufu.cu.)
// The amplitude is u = 1 / rr (AMP_RINV) or rr (AMP_R) with rr = ak * dn + ac, where (ak, ac) fold the distance
// r_m = r0 - dn and the power-of-two scale `us` of the FP16 operands: (-1/us, r0/us) resp. (-us, r0 us).
template <int AMP>
__device__ __forceinline__ void steer2(float2 dn, float q, float ak, float ac, float2 &hr, float2 &hi) {
#ifdef B2_ABL_NOSYNTH
    hr = __fmul2_rn(dn, make_float2(q, q)); hi = __ffma2_rn(dn, make_float2(ak, ak), make_float2(ac, ac)); return;
#endif
    const float2 p = __fmul2_rn(make_float2(q, q), dn);
    const float2 big = make_float2(12582912.0f, 12582912.0f);            // 1.5 * 2^23: rounds to the nearest integer
    const float2 rn = __fadd2_rn(__fadd2_rn(p, big), make_float2(-12582912.0f, -12582912.0f));
    const float2 fr = __ffma2_rn(rn, make_float2(-1.0f, -1.0f), p);     // exact
    const float2 a = __fmul2_rn(fr, make_float2(6.28318530717958647692f, 6.28318530717958647692f));
    float2 sn, cs;
    __sincosf(a.x, &sn.x, &cs.x);
    __sincosf(a.y, &sn.y, &cs.y);
    if (AMP == AMP_ONE) { hr = cs; hi = sn; return; }
    float2 u = __ffma2_rn(dn, make_float2(ak, ak), make_float2(ac, ac));
    if (AMP == AMP_RINV) {
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(u.x) : "f"(u.x));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(u.y) : "f"(u.y));
    }
    hr = __fmul2_rn(u, cs);
    hi = __fmul2_rn(u, sn);
}

// top 11 significant bits by truncation: the residual x - trunc(x) is exact and is carried by the
// correction operand, so no rounding step is needed in the split modes
__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// ---------------------------------------------------------------- the kernel
// The tensor-core work of one bin for a PAIR of neighbouring column blocks (b0, b1 = b0 + 1): chunks are walked
// downwards from the last one (kc = NKC-1: the widest, its MMA overwrites the accumulator).  Chunks kc >= kcB feed both
// blocks, chunks kcA <= kc < kcB feed b0 only.  Stream 0 (issue warp 0) multiplies every chunk into the accumulator of
// b0, stream 1 (issue warp 1) the chunks kc >= kcB into the accumulator of b1: every steering tile is synthesised ONCE
// per bin although both blocks need it.  A lone block is the case kcB = NKC.
// work unit -> (contiguous range of bins, tile of 128 grid points); units are dealt round-robin and ordered
// bin-range-major: all CTAs work on the same range of bins at the same time, so the operand images of that
// range (sized to fit the L2) are fetched from HBM once per launch instead of once per wave of tiles
#define B2_UNIT_DECODE                                                        \
    const int upart = unit / ntiles, tile = unit - upart * ntiles;           \
    const int f_begin = int(int64_t(nbins) * upart / nsplit);                \
    const int f_end = int(int64_t(nbins) * (upart + 1) / nsplit);

// The chunk positions of ONE bin -- the same for every bin -- as a small program in shared memory, so that the
// synthesis warps do not re-derive the pair / block structure for every chunk:
//   bits 0-5 kc | bit 6: the chunk's own columns belong to a block of this pair (reduce them here) |
//   bits 7-9: accumulator of that block within the bin (other chunks: number of accumulators of the bin completed before them) | bits 10-12: (first column of the chunk in its block) / 16 |
//   bit 13: lowest chunk of its block (the accumulator has been read completely) | bit 14: last chunk of the bin |
//   bit 15: one of the last NG chunks of its block (the group that owns it will not read this accumulator again)
constexpr uint32_t PG_RED = 1u << 6, PG_LASTBLK = 1u << 13, PG_ENDBIN = 1u << 14, PG_LASTOWN = 1u << 15;

template <int NB, int PASSES>
__global__ void __launch_bounds__(TcCfg<NB, PASSES>::THREADS, 1)
das_map_tc_kernel(const unsigned char *__restrict__ op, float *__restrict__ scratch,
                  const float *__restrict__ bin_scale, const float *__restrict__ bin_kf, int M, int nbins,
                  const double *__restrict__ freqs, const double *__restrict__ weights, double inv_c,
                  const double *__restrict__ mic, double cx0, double cy0, double cz0,
                  const double *__restrict__ grid, int64_t G, int steering, int clamp_negative,
                  double *__restrict__ map_out, double *__restrict__ per_bin, int ntiles, int nsplit) {
    using CF = TcCfg<NB, PASSES>;
    constexpr int UMMA_N = CF::UMMA_N, NSA = CF::NSA, NSB = CF::NSB, NG = CF::NG, NGEN = CF::NGEN, DLAG = CF::DLAG;
    constexpr int NS_WARPS = CF::NS_WARPS, W_MMA = CF::W_MMA, PARTIAL_ROWS = CF::PARTIAL_ROWS;
    extern __shared__ __align__(1024) unsigned char smem[];
    float *partial = reinterpret_cast<float *>(smem + NSA * CF::A_STAGE + NSB * CF::B_STAGE);   // [PARTIAL_ROWS][128]
    uint16_t *prog = reinterpret_cast<uint16_t *>(partial + PARTIAL_ROWS * TILE_G);              // [PROG_MAX]
    uint64_t *bars = reinterpret_cast<uint64_t *>(prog + PROG_MAX);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + CF::NBARS);
    uint16_t *lastpos = reinterpret_cast<uint16_t *>(tmem_slot + 4);       // [8] position in the bin of each accumulator's lowest chunk
    // shared-space addresses of the barrier arrays (8 bytes per barrier)
    const uint32_t smem_base = smem_u32(smem);
    // B barriers: ring 0 (NSB0 slots) then ring 1 (NSB1 slots)
    const uint32_t fullA = smem_u32(bars), emptyA = fullA + 8 * NSA, fullB = emptyA + 8 * NSA, emptyB = fullB + 8 * NSB;
    const uint32_t tdone = emptyB + 8 * NSB, tempty = tdone + 8 * CF::TD;
    const uint32_t pfull = tempty + 8 * CF::NBUF, pfree = pfull + 32;              // one pair per quadrant of TMEM lanes
    const uint32_t b_ring = smem_base + NSA * CF::A_STAGE;

    // the warp index through a shuffle: the compiler then KNOWS it is warp-uniform and keeps the role loops' state in
    // uniform registers
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int NKC = (M + KCH - 1) / KCH;          // K chunks
    const int Mk = NKC * KCH;
    const int NBLK = (M + NB - 1) / NB;           // column blocks
    const int akind = amp_kind(steering);
    float *tab = scratch + size_t(blockIdx.x) * TILE_G * Mk;
    // column blocks are processed in pairs that share their A tiles (CF::SHARE), else one by one
    constexpr bool SHARE = CF::SHARE;
    const int NPAIR = SHARE ? (NBLK + 1) / 2 : NBLK;
    auto pair_b0 = [&](int p) { return SHARE ? 2 * p : p; };
    auto pair_b1 = [&](int p) { return (SHARE && 2 * p + 1 < NBLK) ? 2 * p + 1 : -1; };
    auto first_chunk = [&](int blk) { return (blk * NB) / KCH; };          // first K chunk that reaches block blk
    int cpb = 0;                                  // chunk positions per bin
    for (int p = 0; p < NPAIR; ++p) cpb += NKC - first_chunk(pair_b0(p));

    if (tid == 0) {
        if (smem_base & 1023u) __trap();          // SWIZZLE_128B operand tiles need a 1024-byte aligned base
        for (int s = 0; s < NSA; ++s) {
            mbar_init(fullA + 8 * s, 4);          // one arrival per warp of the group that owns the chunk
            mbar_init(emptyA + 8 * s, 2);         // both issue warps release every tile
        }
        for (int s = 0; s < CF::TD; ++s) mbar_init(tdone + 8 * s, 1);   // committed once per chunk by the warp that issues its last relevant step
        for (int s = 0; s < NSB; ++s) { mbar_init(fullB + 8 * s, 1); mbar_init(emptyB + 8 * s, 1); }
        for (int b = 0; b < CF::NBUF; ++b) mbar_init(tempty + 8 * b, NS_WARPS);
        for (int q = 0; q < 4; ++q) { mbar_init(pfull + 8 * q, NG > 1 ? NG - 1 : 1); mbar_init(pfree + 8 * q, 1); }
        fence_barrier_init();
        // the per-bin chunk program
        int n = 0, accbase = 0;
        for (int p = 0; p < NPAIR; ++p) {
            const int b0 = pair_b0(p), b1 = pair_b1(p);
            for (int kc = NKC - 1; kc >= first_chunk(b0); --kc, ++n) {
                const int blk = (kc * KCH) / NB;
                uint32_t e = uint32_t(kc);
                if (blk == b0 || blk == b1) {
                    const int accoff = accbase + ((blk == b1 || b1 < 0) ? 0 : 1);   // completion order: b1 (if any), then b0
                    e |= PG_RED | (uint32_t(accoff) << 7) | (uint32_t((kc * KCH - blk * NB) / KCH) << 10);
                    if (kc == first_chunk(blk)) { e |= PG_LASTBLK; lastpos[accoff] = uint16_t(n); }
                    if (kc - first_chunk(blk) < NG) e |= PG_LASTOWN;
                } else {
                    e |= uint32_t(accbase) << 7;          // a chunk above both blocks: the accumulators of the earlier pairs are complete
                }
                prog[n] = uint16_t(e);
            }
            accbase += b1 >= 0 ? 2 : 1;
        }
        prog[n - 1] = uint16_t(prog[n - 1] | PG_ENDBIN);
    }
    if (warp == W_MMA) tmem_alloc(smem_u32(tmem_slot), CF::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < NS_WARPS) {
        // =========================== steering synthesis -> A operand tiles, and the reduction of T ===========================
        if constexpr (CF::SETMAXNREG) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CF::REGS_S));
        const int r = tid & 127, grp = tid >> 7, quad = warp & 3;
        const uint32_t lane_base = uint32_t(quad * 32) << 16;
        const uint32_t row_off = uint32_t((r >> 3) * 1024 + (r & 7) * 128);     // this row inside a swizzled tile
        const uint32_t rx = uint32_t(r & 7);
        const float *tabr = tab + r * 4;
        uint32_t seq_base = 0;                               // ring sequence number of the unit's first chunk
        uint32_t accs = 0;                                   // accumulator uses handed out before this unit
        uint32_t arrived_acc = 0;                            // accumulator uses this group has released (tempty) so far
        PROF_DECL;
#ifdef B2_MAP_TRACE
        const bool prof = blockIdx.x == 0 && lane == 0 && quad == 0;     // one warp per group (the cycle sums then add up over the groups)
        const int prof_role = 2 + grp;
#else
        const bool prof = blockIdx.x == 0 && tid == 0;
#endif
        (void)prof;
#ifdef B2_MAP_PROFILE
        const long long t_role0 = clock64();
#endif
        uint32_t hpar = 0;                                   // parity of the bin-sum hand-off phase (one phase per bin)
        for (int unit = blockIdx.x; unit < ntiles * nsplit; unit += gridDim.x) {
            B2_UNIT_DECODE
            const int64_t g = int64_t(tile) * TILE_G + r;
            PROF_T0();
            // ---- per-unit geometry: this thread's share of the row's distances, amplitude constants of the row ----
            float ak = 0.0f, ac = 1.0f;                      // amplitude argument rr = ak * dn + ac
            double c2 = 0.0;
            {
                double s_part = 0.0, r0 = 1.0;
                const bool valid = g < G;
                double x = 0, y = 0, z = 0;
                if (valid) {
                    x = grid[3 * g]; y = grid[3 * g + 1]; z = grid[3 * g + 2];
                    r0 = das_distance(x, y, z, cx0, cy0, cz0);
                }
                const float r0f = float(r0);
                // every thread is done with the previous unit: its table (read by all groups) and the A ring (used as
                // scratch below; this thread has passed every chunk of that unit, hence all its MMAs have retired)
                asm volatile("bar.sync 1, %0;" ::"n"(NGEN) : "memory");
                for (int m4 = grp * 4; m4 < Mk; m4 += 4 * NG) {
                    float4 dn4;
                    float *dnp = reinterpret_cast<float *>(&dn4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int m = m4 + i;
                        // dn = r0 - r_m (beamforming.py:1557 etc. use exactly this difference in the phase): bounded by
                        // the array aperture however far the grid point is, so the fp32 phase product q*dn stays accurate;
                        // r_m itself (amplitude) is recovered as r0 - dn.  Padding: dn with r0 - dn = 1.
                        double dn = double(r0f) - 1.0;
                        if (valid && m < M) {
                            const double rm = das_distance(x, y, z, mic[3 * m], mic[3 * m + 1], mic[3 * m + 2]);
                            dn = double(r0f) - rm;
                            s_part += 1.0 / (rm * rm);
                        }
                        dnp[i] = float(dn);
                    }
                    *reinterpret_cast<float4 *>(tab + tab_index(m4, r)) = dn4;
                }
                // s = sum_m 1 / r_m^2 over ALL microphones: the NG threads of a row add their shares through shared
                // memory (the free A ring).  The barrier also orders the table writes before the other groups' reads.
                double s = s_part;
                double *sp = reinterpret_cast<double *>(smem);
                if (NG > 1) sp[grp * TILE_G + r] = s_part;
                __threadfence_block();
                asm volatile("bar.sync 1, %0;" ::"n"(NGEN) : "memory");
                if (NG > 1) {
                    s = 0.0;
#pragma unroll
                    for (int k = 0; k < NG; ++k) s += sp[k * TILE_G + r];
                    asm volatile("bar.sync 1, %0;" ::"n"(NGEN) : "memory");      // read before the first tile overwrites it
                }
                if (valid) {
                    double c = 0.0, us = 1.0;
                    switch (steering) {
                    case B2_STEER_CLASSIC: c = 1.0 / double(M); break;
                    case B2_STEER_INVERSE: c = 1.0 / (double(M) * r0); break;
                    case B2_STEER_TRUE_POWER: c = 1.0 / (r0 * s); break;
                    default: c = 1.0 / sqrt(double(M) * s); break;
                    }
                    c2 = c * c;
                    if (PASSES == 4 && steering != B2_STEER_CLASSIC && r0 > 0.0 && r0 < 1e30) {
                        // power of two that brings u(r) of this grid point near 1 (FP16 operands); both the operand
                        // and the kept h' carry it, so the finished power is divided by its square
                        int e = ilogb(r0);                                     // r0 = 2^e * [1, 2)
                        e = e < -60 ? -60 : (e > 60 ? 60 : e);
                        us = scalbn(1.0, steering == B2_STEER_INVERSE ? -e : e);
                        c2 /= us * us;
                    }
                    if (akind == AMP_RINV) { ak = float(-1.0 / us); ac = float(double(r0f) / us); }     // rr = r_m / us, u = 1 / rr
                    else if (akind == AMP_R) { ak = float(-us); ac = float(double(r0f) * us); }         // rr = r_m us,   u = rr
                }
            }
            const int total = (f_end - f_begin) * cpb;
            PROF_ADD(24 + 7, prof);
            // ---- cursors ----
            // synthesis: own chunks i = grp, grp + NG, ... of the unit; (s_fi, s_cib) = bin and position inside the bin
            int s_cib = grp, s_fi = f_begin;
            while (s_cib >= cpb) { s_cib -= cpb; ++s_fi; }
            uint32_t s_slot = (seq_base + uint32_t(grp)) % NSA, s_ph = ((seq_base + uint32_t(grp)) / NSA) & 1;
            // reduction: this group's own chunks in order, DLAG own chunks behind the synthesis cursor
            int r_cib = s_cib, r_fi = s_fi;
            uint32_t r_seq = seq_base + uint32_t(grp);
            int cur_bin = f_begin;                               // bin whose sums are being accumulated
            float bf = s_fi < f_end ? bin_kf[s_fi] : 0.0f;
            float2 hkeep[DLAG][KCH];                          // h' of the last DLAG own chunks: (Re, Im) per microphone (pairs)
            float2 pacc = make_float2(0.f, 0.f);              // sum of Re T Re h' / of (-Im T) Im h', current bin
            float2 pre = make_float2(0.f, 0.f);               // planar layout: (sum Re T Re h', sum Im T Im h')
            double acc_map = 0.0;

            // distances of the next own chunk (4 x 16 bytes), fetched one own-chunk ahead
            float4 dnext[4];
            auto prefetch = [&](int cib) {
                const int kc = int(prog[cib] & 0x3Fu);
#pragma unroll
                for (int gi = 0; gi < 4; ++gi)
                    dnext[gi] = __ldcg(reinterpret_cast<const float4 *>(tabr + (kc * 4 + gi) * (TILE_G * 4)));
            };
            if (grp < total) prefetch(s_cib);

            // group 0: the bin whose sums of the other groups are still awaited
            bool pending = false;
            float pend_p = 0.f;
            int pend_fi = 0;
            uint32_t pend_par = 0;
            auto finish_bin = [&](bool block) {
                double extra = 0.0;
                if (NG > 1) {
                    if (block) mbar_wait(pfull + 8 * quad, pend_par);
                    else if (!mbar_test(pfull + 8 * quad, pend_par)) return;
#pragma unroll
                    for (int k = 0; k < NG - 1; ++k) extra += double(partial[k * TILE_G + r]);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(pfree + 8 * quad);
                }
                double p = (double(pend_p) + extra) * c2;                              // |h|^2 scale of this grid point
                if (PASSES == 4) p /= double(bin_scale[pend_fi]);                      // power-of-two scale of L_f
                if (clamp_negative && p < 0.0) p = 0.0;                                // beamforming.py:871-872
                acc_map += weights[pend_fi] * p;                                       // :875-876
                if (per_bin != nullptr && g < G) per_bin[g * nbins + pend_fi] = p;
                pending = false;
            };

            // the bins [cur_bin, upto) are complete as far as this group is concerned: hand their sums over.  Groups 1..
            // write theirs to shared memory and signal `pfull`; group 0 finishes P[g, f] in fp64 -- not on the spot, it
            // keeps its own sum pending and completes the bin at a later chunk, when `pfull` has fired (finish_bin polls
            // without blocking), so that no group sits out the skew between the groups.
            auto flush_bins = [&](int upto) {
                while (cur_bin < upto) {
                    const float ptot = CF::PLANAR ? pre.x - pre.y : pacc.x + pacc.y;     // (zero for a bin without an own chunk)
                    pacc = make_float2(0.f, 0.f);
                    pre = make_float2(0.f, 0.f);
                    if (NG > 1 && grp > 0) {
                        mbar_wait(pfree + 8 * quad, hpar ^ 1);                  // group 0 has read the previous bin's sums
                        partial[(grp - 1) * TILE_G + r] = ptot;
                        __syncwarp();
                        if (lane == 0) mbar_arrive(pfull + 8 * quad);
                    } else {
                        if (pending) finish_bin(true);                          // (a whole bin went by without a poll succeeding)
                        pend_p = ptot; pend_fi = cur_bin; pend_par = hpar; pending = true;
                        finish_bin(NG == 1);
                    }
                    hpar ^= 1;
                    ++cur_bin;
                }
            };

            // Reduction of this group's next own chunk (DLAG own chunks behind the synthesis).  There is no walk over the
            // chunks of the other groups: everything that used to be an event of the chunk stream is derived from the own
            // chunk -- the bin it belongs to (a new bin flushes the sums of the old one), the accumulator it reads
            // (accumulator uses this group has NO chunk in are released first, after waiting for their lowest chunk: a
            // release must not be counted for the previous use of the TMEM buffer, scripts/sim_protocol.py), and whether
            // it is this group's last chunk of that accumulator (released right after the read).
            auto reduce_own = [&](const float2 (&hk)[KCH]) {
                if (NG > 1 && pending) finish_bin(false);
                const uint32_t e = prog[r_cib];
                const uint32_t accg = accs + uint32_t(r_fi - f_begin) * uint32_t(NBLK) + ((e >> 7) & 7u);
                const uint32_t buf = accg & uint32_t(CF::NBUF - 1);
                {
                    while (arrived_acc < accg) {                 // accumulator uses without a chunk of this group (blocks of < NG chunks)
                        const uint32_t u = arrived_acc - accs;
                        const uint32_t ls = seq_base + (u / uint32_t(NBLK)) * uint32_t(cpb) + lastpos[u % uint32_t(NBLK)];
                        mbar_wait(tdone + 8 * (ls & uint32_t(CF::TD - 1)), (ls / uint32_t(CF::TD)) & 1);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty + 8 * (arrived_acc & uint32_t(CF::NBUF - 1)));
                        ++arrived_acc;
                    }
                }
                if (e & PG_RED) {
                    PROF_T0();
                    mbar_wait(tdone + 8 * (r_seq & uint32_t(CF::TD - 1)), (r_seq / uint32_t(CF::TD)) & 1);
                    PROF_ADD(24 + 1, prof);
                    tc_fence_after();
                    flush_bins(r_fi);
                    if (CF::PLANAR) {
                        // accumulator of a block: Re T in columns [0, NB), Im T in [NB, 2 NB); hk = (Re h' of the 8
                        // microphone pairs, Im h' of them): P = sum Re T Re h' - Im T Im h'
                        const uint32_t tcol = tmem_base + lane_base + buf * UMMA_N + 16u * ((e >> 10) & 7u);
                        float tr[16], ti[16];                // Re T, Im T of the chunk's 16 microphones: both loads, one wait
                        tmem_ld16(tcol, tr);
                        tmem_ld16(tcol + NB, ti);
                        tmem_ld_wait();
                        float2 pi2 = make_float2(0.f, 0.f);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {        // padded microphones have all-zero operand rows: T = 0, no bounds test
                            pacc = __ffma2_rn(make_float2(tr[2 * j], tr[2 * j + 1]), hk[j], pacc);
                            pi2 = __ffma2_rn(make_float2(ti[2 * j], ti[2 * j + 1]), hk[8 + j], pi2);
                        }
                        pre.x += pacc.x + pacc.y;
                        pre.y += pi2.x + pi2.y;
                        pacc = make_float2(0.f, 0.f);
                    } else {
                        const uint32_t tcol = tmem_base + lane_base + buf * UMMA_N + 32u * ((e >> 10) & 7u);
#pragma unroll
                        for (int hlf = 0; hlf < 2; ++hlf) {
                            float t[16];                     // (Re T, -Im T) of 8 microphones
                            tmem_ld16(tcol + 16 * hlf, t);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 8; ++j)      // padded microphones have all-zero operand rows: T = 0, no bounds test
                                pacc = __ffma2_rn(make_float2(t[2 * j], t[2 * j + 1]), hk[8 * hlf + j], pacc);
                        }
                    }
                    PROF_ADD(24 + 2, prof);
                    if (e & PG_LASTOWN) {
                        // this group's last own chunk of the block: it will not read this accumulator again and says so
                        // now -- the issue warp that overwrites the buffer for the next bin then only waits for the group
                        // that reads the block's lowest chunk
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty + 8 * buf);
                        arrived_acc = accg + 1;
                    }
                }
                r_cib += NG;
                while (r_cib >= cpb) { r_cib -= cpb; ++r_fi; }
                r_seq += uint32_t(NG);
            };

            // synthesis of the own chunk at (s_fi, s_cib): 16 microphones of this row -> A tile + hk
            auto synth_chunk = [&](float2 (&hk)[KCH]) {
                float2 dn[KCH / 2];
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) {
                    dn[2 * gi] = make_float2(dnext[gi].x, dnext[gi].y);
                    dn[2 * gi + 1] = make_float2(dnext[gi].z, dnext[gi].w);
                }
                // the next own chunk: position and bin (its distances are fetched at the end of this synthesis, when the
                // registers of this chunk's distances are free again; the reduction and the waits in between hide the L2 latency)
                int n_cib = s_cib + NG, n_fi = s_fi;
                while (n_cib >= cpb) { n_cib -= cpb; ++n_fi; }
                PROF_T0();
                mbar_wait(emptyA + 8 * s_slot, s_ph ^ 1);
                PROF_ADD(24 + 3, prof);
                const uint32_t ahi = smem_base + s_slot * CF::A_STAGE + row_off;
                const uint32_t alo = ahi + (CF::NSPLIT - 1) * A_TILE;
                auto synth = [&](auto amp_tag) {
                    constexpr int AMP = decltype(amp_tag)::value;
                    if constexpr (CF::PLANAR) {
                        // row = [f1 Re (16 mics) | f1 Im | f2 Re | f2 Im], 32 bytes each; 8 microphones = 16 bytes of every part
#pragma unroll
                        for (int g8 = 0; g8 < 2; ++g8) {
                            uint4 f1r, f1i, f2r, f2i;
                            uint32_t *q1r = reinterpret_cast<uint32_t *>(&f1r), *q1i = reinterpret_cast<uint32_t *>(&f1i);
                            uint32_t *q2r = reinterpret_cast<uint32_t *>(&f2r), *q2i = reinterpret_cast<uint32_t *>(&f2i);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {      // microphone pairs
                                float2 hr, hi;
                                steer2<AMP>(dn[4 * g8 + j], bf, ak, ac, hr, hi);
                                hk[4 * g8 + j] = hr;
                                hk[8 + 4 * g8 + j] = hi;
                                const __half2 ar = __floats2half2_rn(hr.x, hr.y), ai = __floats2half2_rn(hi.x, hi.y);
                                const float2 rr = __ffma2_rn(__half22float2(ar), make_float2(-1.0f, -1.0f), hr);   // exact residuals
                                const float2 ri = __ffma2_rn(__half22float2(ai), make_float2(-1.0f, -1.0f), hi);
                                q1r[j] = pack_f16(ar);
                                q1i[j] = pack_f16(ai);
                                q2r[j] = pack_f16(__floats2half2_rn(rr.x, rr.y));
                                q2i[j] = pack_f16(__floats2half2_rn(ri.x, ri.y));
                            }
                            sts128u(ahi + ((uint32_t(0 + g8) ^ rx) << 4), f1r);
                            sts128u(ahi + ((uint32_t(2 + g8) ^ rx) << 4), f1i);
                            sts128u(ahi + ((uint32_t(4 + g8) ^ rx) << 4), f2r);
                            sts128u(ahi + ((uint32_t(6 + g8) ^ rx) << 4), f2i);
                        }
                        return;
                    }
#pragma unroll
                    for (int gi = 0; gi < 4; ++gi) {           // 4 microphones = 16 bytes of each 16-bit image
                        float2 hr0, hi0, hr1, hi1;
                        steer2<AMP>(dn[2 * gi], bf, ak, ac, hr0, hi0);
                        steer2<AMP>(dn[2 * gi + 1], bf, ak, ac, hr1, hi1);
                        const float2 h[4] = {make_float2(hr0.x, hi0.x), make_float2(hr0.y, hi0.y),
                                             make_float2(hr1.x, hi1.x), make_float2(hr1.y, hi1.y)};
#pragma unroll
                        for (int j = 0; j < 4; ++j) hk[4 * gi + j] = h[j];
                        {
                            // TF32 image: 2 microphones per 16 bytes
                            float4 hi4[2], lo4[2];
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const float2 h0 = h[2 * j], h1 = h[2 * j + 1];
                                if (PASSES == 1) hi4[j] = make_float4(to_tf32(h0.x), to_tf32(h0.y), to_tf32(h1.x), to_tf32(h1.y));
                                else hi4[j] = make_float4(trunc_tf32(h0.x), trunc_tf32(h0.y), trunc_tf32(h1.x), trunc_tf32(h1.y));
                                lo4[j] = make_float4(h0.x - hi4[j].x, h0.y - hi4[j].y, h1.x - hi4[j].z, h1.y - hi4[j].w);
                                sts128(ahi + ((uint32_t(2 * gi + j) ^ rx) << 4), hi4[j]);
                                // residuals are exact in fp32; the tensor core keeps their top 11 bits
                                if (PASSES == 3) sts128(alo + ((uint32_t(2 * gi + j) ^ rx) << 4), lo4[j]);
                            }
                            if (PASSES == 2) {
                                // 16-bit correction image row = [bf16 residuals of the 16 mics | bf16 values]
                                uint4 i1, i2;
                                i1.x = pack_bf16(lo4[0].x, lo4[0].y); i1.y = pack_bf16(lo4[0].z, lo4[0].w);
                                i1.z = pack_bf16(lo4[1].x, lo4[1].y); i1.w = pack_bf16(lo4[1].z, lo4[1].w);
                                i2.x = pack_bf16(h[0].x, h[0].y); i2.y = pack_bf16(h[1].x, h[1].y);
                                i2.z = pack_bf16(h[2].x, h[2].y); i2.w = pack_bf16(h[3].x, h[3].y);
                                sts128u(alo + ((uint32_t(gi) ^ rx) << 4), i1);
                                sts128u(alo + ((uint32_t(4 + gi) ^ rx) << 4), i2);
                            }
                        }
                    }
                };
                if (akind == AMP_RINV) synth(std::integral_constant<int, AMP_RINV>{});
                else if (akind == AMP_R) synth(std::integral_constant<int, AMP_R>{});
                else synth(std::integral_constant<int, AMP_ONE>{});
                PROF_ADD(24 + 4, prof);
                if (n_fi < f_end) prefetch(n_cib);
                // generic-proxy writes -> visible to the MMA (async proxy); one arrival per warp
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(fullA + 8 * s_slot);
                PROF_ADD(24 + 5, prof);
                PROF_CNT(24 + 6, prof);
                s_slot += NG;
                if (s_slot >= uint32_t(NSA)) { s_slot -= NSA; s_ph ^= 1; }
                s_cib = n_cib;
                if (n_fi != s_fi) { s_fi = n_fi; if (s_fi < f_end) bf = bin_kf[s_fi]; }
            };

            // own chunk i is synthesised after the own chunk i - DLAG NG has been reduced (its h' sits in the register
            // slot that chunk i then takes); the tail reduces this group's last DLAG own chunks
#pragma unroll 1
            for (int i0 = grp; i0 < total + DLAG * NG; i0 += DLAG * NG) {
#pragma unroll
                for (int d = 0; d < DLAG; ++d) {
                    const int i = i0 + d * NG;
                    if (i < total + DLAG * NG) {
                        if (i - DLAG * NG >= 0) reduce_own(hkeep[d]);
                        if (i < total) synth_chunk(hkeep[d]);
                    }
                }
            }
            // uses of the unit's last accumulators without a chunk of this group, then the last bins
            {
                const uint32_t acc_end = accs + uint32_t(f_end - f_begin) * uint32_t(NBLK);
                while (arrived_acc < acc_end) {
                    const uint32_t u = arrived_acc - accs;
                    const uint32_t ls = seq_base + (u / uint32_t(NBLK)) * uint32_t(cpb) + lastpos[u % uint32_t(NBLK)];
                    mbar_wait(tdone + 8 * (ls & uint32_t(CF::TD - 1)), (ls / uint32_t(CF::TD)) & 1);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty + 8 * (arrived_acc & uint32_t(CF::NBUF - 1)));
                    ++arrived_acc;
                }
                accs = acc_end;
            }
            flush_bins(f_end);
            if (pending) finish_bin(true);
            seq_base += uint32_t(total);
            // the bins of a tile may be split over several work units: the map is zeroed by the
            // launcher and every unit adds its weighted partial sum
            if (grp == 0 && g < G) atomicAdd(map_out + g, acc_map);
        }
#ifdef B2_MAP_PROFILE
        if (prof) atomicAdd(&g_map_prof[24], (unsigned long long)(clock64() - t_role0));
#endif
        // (the two register-budget regions do not join again: each side has its own copy of the epilogue)
        tc_fence_before();
        __syncthreads();
        return;
    } else {
      if constexpr (CF::SETMAXNREG) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CF::REGS_IO));   // the whole warpgroup, one instruction
      if (warp == W_MMA || warp == W_MMA + 1) {
        // =========================== MMA issue ===========================
        // tcgen05.mma issue is back-pressured by the tensor pipe (the issuing thread advances at the pipe's rate), and
        // every step also costs its issuing thread a few hundred cycles the pipe cannot use (mbarrier waits, elect,
        // commits): the pipe only stays busy while the OTHER warp is issuing during those gaps.  Warp `my` issues
        // stream `my`; the streams are independent of each other except through the A tiles they share.
        // mbarrier waits only see the PARITY of a phase, so a warp must observe EVERY use of a barrier it ever waits
        // on: both warps release every A tile (emptyA counts 2) -- with tcgen05.commit if they multiply it, else
        // (stream 1, chunks below kcB) by waiting for fullA and arriving directly.
        const int my = warp - W_MMA;
        // The whole warp walks the loop, so every address, descriptor and flag is warp-uniform and lives in
        // uniform registers; one elected lane issues the tcgen05 instructions.
        constexpr int NSBm_max = CF::NSB0 > CF::NSB1 ? CF::NSB0 : CF::NSB1;
        (void)NSBm_max;
        const int nsb_my = my ? CF::NSB1 : CF::NSB0;                            // depth of this stream's B ring
        const uint32_t slot0 = my ? uint32_t(CF::NSB0) : 0u;                     // its first slot / barrier
        const uint32_t a0 = desc_lo(smem_base), b0d = desc_lo(b_ring + slot0 * CF::B_STAGE);
        const uint32_t fullBm = fullB + 8 * slot0, emptyBm = emptyB + 8 * slot0;
        uint32_t seqA = 0;                                   // A tiles consumed so far (ring sequence number)
        uint32_t sb = 0, phb = 0;                            // this stream's B ring cursor
        uint32_t acc_it = 0;
        PROF_DECL;
        const bool prof = blockIdx.x == 0 && lane == 0;
        (void)prof;
#ifdef B2_MAP_TRACE
        const int prof_role = my;
#endif
#ifdef B2_MAP_PROFILE
        const long long t_role0 = clock64();
#endif
        for (int unit = blockIdx.x; unit < ntiles * nsplit; unit += gridDim.x) {
            B2_UNIT_DECODE
            (void)tile;
            for (int fi = f_begin; fi < f_end; ++fi) {
                for (int p = 0; p < NPAIR; ++p) {
                    const int blk0 = pair_b0(p), blk1 = pair_b1(p);
                    const int kcA = first_chunk(blk0), kcB = blk1 >= 0 ? first_chunk(blk1) : NKC;
                    // accumulators are handed out in the order b1 (if any), then b0
                    const uint32_t acc = my ? acc_it : acc_it + (blk1 >= 0 ? 1u : 0u);
                    acc_it += blk1 >= 0 ? 2u : 1u;
                    const uint32_t buf = acc & uint32_t(CF::NBUF - 1);
                    const uint32_t d_tmem = tmem_base + buf * UMMA_N;
                    const int kbase = my ? kcB : kcA;
                    const int myblk = my ? blk1 : blk0;
                    // the tcgen05 instructions and commits of ONE chunk (elected lane only)
                    auto issue_chunk = [&](int kc, uint32_t seq, uint32_t sa, uint32_t sbk) {
                        // columns of this block that chunk kc can reach (lower triangle): w
                        const int w = min(NB, KCH * (kc - kbase + 1));
                        // Every chunk signals tdone exactly once per use of its ring slot: on the step that makes its
                        // own columns final (the block that contains it), or -- for a chunk above both blocks of the
                        // pair, whose columns belong to another pair -- on the step of stream 0.
                        const int cblk = (kc * KCH) / NB;
                        const bool signal_done = cblk == myblk || (my == 0 && cblk != blk0 && cblk != blk1);
                        // low words of the operand descriptors: first image, second (correction) image
                        const uint32_t a1 = a0 + sa * (CF::A_STAGE >> 4), b1 = b0d + sbk * (CF::B_STAGE >> 4);
                        const uint32_t a2 = a1 + (CF::NSPLIT - 1) * (A_TILE >> 4), b2 = b1 + (CF::NSPLIT - 1) * (CF::B_TILE >> 4);
                        const uint32_t first = kc != NKC - 1 ? 1u : 0u;    // 0: the first MMA overwrites the accumulator
#ifdef B2_ABL_NOMMA
                        if (false) {
#else
                        if (PASSES == 4) {
#endif
                            // rows = [x1 Re | x1 Im | x2 Re | x2 Im] of the chunk's 16 microphones, 32 B (K = 16 halfs)
                            // each; x = f (A: steering) resp. g (B: matrix); three products f1 g1 + f1 g2 + f2 g1:
                            //   Re T += fr gr^T + fi gi^T,   Im T += fr gi^T + (-fi) gr^T
                            // The A slice of a pair of MMAs is read from shared memory once (collector).
                            const uint32_t idp = idesc_f16(TILE_G, w), idn = idp | IDESC_NEGATE_A;
                            const uint32_t d_re = d_tmem, d_im = d_tmem + NB;
                            const uint64_t f1r = desc64(a1), f1i = desc64(a1 + 2), f2r = desc64(a1 + 4), f2i = desc64(a1 + 6);
                            const uint64_t g1r = desc64(b1), g1i = desc64(b1 + 2), g2r = desc64(b1 + 4), g2i = desc64(b1 + 6);
#ifdef B2_CHAIN4
                            umma_f16_afill(d_re, f1r, g1r, idp, first);
                            umma_f16_ause(d_re, f1r, g2r, idp, 1u);
                            umma_f16_ause(d_im, f1r, g1i, idp, first);
                            umma_f16_alastuse(d_im, f1r, g2i, idp, 1u);
#else
                            umma_f16_afill(d_re, f1r, g1r, idp, first);
                            umma_f16_alastuse(d_re, f1r, g2r, idp, 1u);
                            umma_f16_afill(d_im, f1r, g1i, idp, first);
                            umma_f16_alastuse(d_im, f1r, g2i, idp, 1u);
#endif
                            umma_f16_afill(d_re, f1i, g1i, idp, 1u);
                            umma_f16_alastuse(d_re, f1i, g2i, idp, 1u);
                            umma_f16_afill(d_im, f1i, g1r, idn, 1u);
                            umma_f16_alastuse(d_im, f1i, g2r, idn, 1u);
                            umma_f16_afill(d_re, f2r, g1r, idp, 1u);
                            umma_f16_alastuse(d_im, f2r, g1i, idp, 1u);
                            umma_bf16(d_re, f2i, g1i, idp, 1u);
                            umma_bf16(d_im, f2i, g1r, idn, 1u);
#ifdef B2_ABL_NOMMA
                        } else if (false) {
#else
                        } else {
#endif
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
                        if (signal_done) umma_commit(tdone + 8 * (seq & uint32_t(CF::TD - 1)));   // columns of chunk kc are final -> reduction
                        umma_commit(emptyA + 8 * sa);                      // this warp is done with the A tile
                        umma_commit(emptyBm + 8 * sbk);
                    };
                    // Up to TWO chunks per step: every step costs its warp ~450 cycles that the tensor pipe cannot use
                    // (two barrier round trips of ~150 cycles even when the tiles are there, fences, elect, descriptors,
                    // commits -- role profile r2) and the tcgen05 queue is too shallow to cover them; with the barriers of
                    // two chunks looked at in ONE round trip and 24 MMAs behind them, that cost is paid once per two chunks.
#pragma unroll 1
                    for (int kc = NKC - 1; kc >= kcA;) {
                        PROF_T0();
                        const uint32_t sa = seqA % NSA, pha = (seqA / NSA) & 1;
                        if (my == 1 && kc < kcB) {
                            // a tile only block b0 uses: released after its phase is complete, so that the arrival
                            // cannot land in an earlier phase
                            mbar_wait(fullA + 8 * sa, pha);
                            if (elect_one_sync()) mbar_arrive(emptyA + 8 * sa);
                            __syncwarp();
                            PROF_ADD(8 * my + 5, prof);
                            --kc; ++seqA;
                            continue;
                        }
                        // first MMA of this accumulator: have the synthesis warps read its previous contents?
                        if (kc == NKC - 1) mbar_wait(tempty + 8 * buf, ((acc / uint32_t(CF::NBUF)) & 1) ^ 1);
                        PROF_ADD(8 * my + 1, prof);
#ifndef B2_CHUNKS_PER_STEP
#define B2_CHUNKS_PER_STEP 2
#endif
                        // chunks of this step: the next ones are this warp's next multiply steps too, and the B ring keeps
                        // at least one stage for the loader to prefetch into
                        // (measured: 256 mics 12.95 / 12.3 / 12.6 ms for 1 / 2 / 3 chunks per step; 64 mics -- 4 chunks
                        // per bin -- 12.6 / 12.9 / 14.1 ms: short bins keep one chunk per step)
                        int adv = kc - kbase + 1;
                        constexpr int MAXADV = NB >= 128 ? B2_CHUNKS_PER_STEP : 1;     // (compile-time for the small-array kernels)
                        if (adv > MAXADV) adv = MAXADV;
                        if (NKC < 8) adv = 1;
                        if (adv > nsb_my - 1) adv = nsb_my - 1 > 0 ? nsb_my - 1 : 1;
                        const uint32_t sa2 = (seqA + 1) % NSA, pha2 = ((seqA + 1) / NSA) & 1;
                        const uint32_t sa3 = (seqA + 2) % NSA, pha3 = ((seqA + 2) / NSA) & 1;
                        const uint32_t sb2 = sb + 1 == uint32_t(nsb_my) ? 0u : sb + 1, phb2 = sb2 ? phb : phb ^ 1u;
                        const uint32_t sb3 = sb2 + 1 == uint32_t(nsb_my) ? 0u : sb2 + 1, phb3 = sb3 ? phb2 : phb2 ^ 1u;
                        if (adv >= 3) {
                            mbar_wait4(fullBm + 8 * sb, phb, fullA + 8 * sa, pha, fullBm + 8 * sb2, phb2, fullA + 8 * sa2, pha2);
                            mbar_wait2(fullBm + 8 * sb3, phb3, fullA + 8 * sa3, pha3);
                        } else if (adv == 2) {
                            mbar_wait4(fullBm + 8 * sb, phb, fullA + 8 * sa, pha, fullBm + 8 * sb2, phb2, fullA + 8 * sa2, pha2);
                        } else {
                            mbar_wait2(fullBm + 8 * sb, phb, fullA + 8 * sa, pha);
                        }
                        PROF_ADD(8 * my + 3, prof);
                        tc_fence_after();
                        if (elect_one_sync()) {
                            issue_chunk(kc, seqA, sa, sb);
                            if (adv >= 2) issue_chunk(kc - 1, seqA + 1, sa2, sb2);
                            if (adv >= 3) issue_chunk(kc - 2, seqA + 2, sa3, sb3);
                        }
                        __syncwarp();
                        PROF_ADD(8 * my + 4, prof);
                        PROF_CNT(8 * my + 6, prof);
                        kc -= adv; seqA += uint32_t(adv);
                        for (int q = 0; q < adv; ++q)
                            if (++sb == uint32_t(nsb_my)) { sb = 0; phb ^= 1; }
                    }
                }
            }
        }
#ifdef B2_MAP_PROFILE
        if (prof) atomicAdd(&g_map_prof[8 * my], (unsigned long long)(clock64() - t_role0));
#endif
    } else {
        // =========================== B operand loaders (one thread per stream) ===========================
        const int st = warp - (W_MMA + 2);                   // stream 0: blocks b0, stream 1: blocks b1
        const int nsb_my = st ? CF::NSB1 : CF::NSB0;
        if (lane == 0 && nsb_my > 0) {
            const uint32_t slot0 = st ? uint32_t(CF::NSB0) : 0u;
            const uint32_t ring = b_ring + slot0 * CF::B_STAGE, fullBm = fullB + 8 * slot0, emptyBm = emptyB + 8 * slot0;
            uint32_t sb = 0, phb = 0;
            const size_t bin_bytes = size_t(NBLK) * NKC * CF::NSPLIT * CF::B_TILE;
            for (int unit = blockIdx.x; unit < ntiles * nsplit; unit += gridDim.x) {
                B2_UNIT_DECODE
                (void)tile;
                for (int fi = f_begin; fi < f_end; ++fi) {
                    for (int p = 0; p < NPAIR; ++p) {
                        const int blk = st ? pair_b1(p) : pair_b0(p);
                        if (blk < 0) continue;
                        const int kbase = first_chunk(blk);
                        // the tiles are consumed in exactly the order of the issue warp's walk
                        for (int kc = NKC - 1; kc >= kbase; --kc) {
                            const int w = min(NB, KCH * (kc - kbase + 1));
                            const uint32_t bytes = uint32_t(CF::PLANAR ? w : 2 * w) * 128;   // the rows of the first w columns of each image
                            mbar_wait(emptyBm + 8 * sb, phb ^ 1);
                            const unsigned char *src = op + size_t(fi) * bin_bytes + (size_t(blk) * NKC + kc) * (CF::NSPLIT * CF::B_TILE);
#ifdef B2_ABL_NOLOAD
                            (void)src; (void)bytes; (void)ring;
                            mbar_arrive(fullBm + 8 * sb);
#else
                            mbar_arrive_expect_tx(fullBm + 8 * sb, CF::NSPLIT * bytes);
                            bulk_g2s(ring + sb * CF::B_STAGE, src, bytes, fullBm + 8 * sb);
                            if (CF::NSPLIT == 2) bulk_g2s(ring + sb * CF::B_STAGE + CF::B_TILE, src + CF::B_TILE, bytes, fullBm + 8 * sb);
#endif
                            if (++sb == uint32_t(nsb_my)) { sb = 0; phb ^= 1; }
                        }
                    }
                }
            }
        }
      }
    }

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

// One thread per row of an operand image: (bin, column block, K chunk, row).
template <int NB, int PASSES>
__global__ void das_prepare_tc_kernel(const double2 *__restrict__ csm, int M, int bin0, double gain, int remove_diag,
                                      unsigned char *__restrict__ op, int NKC, int NBLK,
                                      const float *__restrict__ bin_scale) {
    using CF = TcCfg<NB, PASSES>;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;       // 0 .. B_ROWS-1
    if (row >= CF::B_ROWS) return;
    const int kc = blockIdx.y % NKC, nbk = blockIdx.y / NKC;
    const int fi = blockIdx.z;
    unsigned char *tile = op + ((size_t(fi) * NBLK + nbk) * NKC + kc) * size_t(CF::NSPLIT * CF::B_TILE);
    if (PASSES == 4) gain *= double(bin_scale[fi]);
    // L[m, n] = 2 gain C[m, n] below the diagonal, gain C[n, n] (or 0) on it
    auto lval = [&](int m, int n, double &lr, double &li) {
        lr = 0.0; li = 0.0;
        if (m < M && n < M && n <= m) {
            const double2 c = csm[(int64_t(bin0 + fi) * M + m) * M + n];
            if (n == m) { lr = remove_diag ? 0.0 : gain * c.x; }
            else { lr = 2.0 * gain * c.x; li = 2.0 * gain * c.y; }
        }
    };
    if constexpr (CF::PLANAR) {
        // one row per matrix column n: [g1 Re (16 mics) | g1 Im | g2 Re | g2 Im], 32 bytes each (FP16 halves g1 + g2)
        const int n = nbk * NB + row;
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {                          // 8 microphones = 16 bytes of every part
            uint4 g1r, g1i, g2r, g2i;
            uint32_t *q1r = reinterpret_cast<uint32_t *>(&g1r), *q1i = reinterpret_cast<uint32_t *>(&g1i);
            uint32_t *q2r = reinterpret_cast<uint32_t *>(&g2r), *q2i = reinterpret_cast<uint32_t *>(&g2i);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double lr0, li0, lr1, li1;
                lval(kc * KCH + 8 * h8 + 2 * j, n, lr0, li0);
                lval(kc * KCH + 8 * h8 + 2 * j + 1, n, lr1, li1);
                const float vr0 = float(lr0), vr1 = float(lr1), vi0 = float(li0), vi1 = float(li1);
                const __half2 ar = __floats2half2_rn(vr0, vr1), ai = __floats2half2_rn(vi0, vi1);
                const float2 fr = __half22float2(ar), fim = __half22float2(ai);
                q1r[j] = pack_f16(ar);
                q1i[j] = pack_f16(ai);
                q2r[j] = pack_f16(__floats2half2_rn(vr0 - fr.x, vr1 - fr.y));
                q2i[j] = pack_f16(__floats2half2_rn(vi0 - fim.x, vi1 - fim.y));
            }
            *reinterpret_cast<uint4 *>(tile + sw128_offset(row, 0 + h8)) = g1r;
            *reinterpret_cast<uint4 *>(tile + sw128_offset(row, 2 + h8)) = g1i;
            *reinterpret_cast<uint4 *>(tile + sw128_offset(row, 4 + h8)) = g2r;
            *reinterpret_cast<uint4 *>(tile + sw128_offset(row, 6 + h8)) = g2i;
        }
        return;
    }
    const bool imag_row = row & 1;                                // rows alternate Re T / -Im T of column n
    const int n = nbk * NB + (row >> 1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float v[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double lr, li;
            lval(kc * KCH + 2 * j + e, n, lr, li);
            // real rows: (Re L, Im L) -> Re T ; imaginary rows: (-Im L, Re L) -> -Im T
            v[2 * e] = float(imag_row ? -li : lr);
            v[2 * e + 1] = float(imag_row ? lr : li);
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
    const int64_t btile = (PASSES == 4 ? 1 : 2) * nb * 128;       // planar FP16x3 image: one row per column
    return int64_t(nbins) * NBLK * NKC * ((PASSES == 1 || PASSES == 4) ? 1 : 2) * btile;
}
// f / c of every bin as fp32 (rounded once from the fp64 product): the synthesis warps read it when they enter a bin --
// with the fp64 multiply in their loop (predicated, but issued with every chunk) 4 % of all warp samples of the kernel sat
// on that one DMUL (ncu source page, r2)
__global__ void das_binfreq_kernel(const double *__restrict__ freqs, double inv_c, int nbins, float *__restrict__ out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < nbins) out[f] = float(freqs[f] * inv_c);
}

// operand buffer = [images][one distance table of 128 x Mk floats per persistent CTA][per-bin scales][per-bin f / c]
inline int64_t tables_bytes(int M) { return int64_t(MAX_CTAS) * TILE_G * ((M + KCH - 1) / KCH) * KCH * 4; }
template <int PASSES> float *tables_ptr(const void *op, int M, int nbins) {
    return reinterpret_cast<float *>((unsigned char *)op + images_bytes<PASSES>(M, nbins));
}
template <int PASSES> float *scales_ptr(const void *op, int M, int nbins) {
    return reinterpret_cast<float *>((unsigned char *)op + images_bytes<PASSES>(M, nbins) + tables_bytes(M));
}

template <int PASSES> float *binkf_ptr(const void *op, int M, int nbins) {
    return scales_ptr<PASSES>(op, M, nbins) + ((int64_t(nbins) * 4 + 255) & ~int64_t(255)) / 4;
}

template <int NB, int PASSES>
int launch_prepare(const void *csm, int M, int bin0, int nbins, int remove_diag, void *op, cudaStream_t st) {
    using CF = TcCfg<NB, PASSES>;
    const int NKC = (M + KCH - 1) / KCH, NBLK = (M + NB - 1) / NB;
    const double gain = remove_diag ? double(M) / double(M - 1) : 1.0;
    dim3 grid((CF::B_ROWS + 127) / 128, NKC * NBLK, nbins);
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
    float *bin_kf = binkf_ptr<PASSES>(op, M, nbins);
    das_binfreq_kernel<<<(nbins + 255) / 256, 256, 0, st>>>((const double *)freqs, 1.0 / c, nbins, bin_kf);
    B2_LAUNCH_OK("b2_das_map(bin frequencies)");
    kern<<<nblocks, CF::THREADS, CF::SMEM, st>>>((const unsigned char *)op, scratch, scales_ptr<PASSES>(op, M, nbins), bin_kf, M, nbins,
                                                (const double *)freqs,
                                                (const double *)weights, 1.0 / c, (const double *)mic, center[0],
                                                center[1], center[2], (const double *)grid, G, steering,
                                                clamp_negative, (double *)map_out, (double *)per_bin, int(ntiles), nsplit);
    B2_LAUNCH_OK("b2_das_map(tc)");
    return B2_OK;
}

template <int PASSES> int64_t operand_bytes(int M, int nbins) {
    return images_bytes<PASSES>(M, nbins) + tables_bytes(M) + 2 * ((int64_t(nbins) * 4 + 255) & ~int64_t(255));
}

}  // namespace

#ifdef B2_MAP_TRACE
extern "C" int b2_debug_map_trace(long long *out) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_map_trace, sizeof(long long) * 5 * TRACE_LEN * 2);
    return 0;
}
#endif
#ifdef B2_MAP_PROFILE
extern "C" int b2_debug_map_profile(unsigned long long *out) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_map_prof, sizeof(unsigned long long) * 32);
    unsigned long long z[32] = {0};
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
