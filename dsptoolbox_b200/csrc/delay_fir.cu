// Source simulation in front of stage 1: the signals S monopole sources leave on M
// microphones, mixed, written once.
//
//   out[n][m] = sum_s sum_t taps[s][t][m] * x_s[n - shift[s][m] - t]        n < L, 0 outside a source's samples
//
// replaces `MonopoleSource.get_signals_on_array` + `mix_sources_on_array` (beamforming/beamforming.py:1427-1511),
// which delay the emitted signal per microphone with `fractional_delay(..., keep_length=True)`
// (standard/latency_delay.py:159-257: FFT overlap-add with a 31-tap windowed sinc, then an integer shift), divide
// by 1 + distance, append channel after channel and add source after source.  The host computes the taps (the
// 1 / (1 + r) gain folded in) and the integer shifts -- M x 31 numbers per source -- and this kernel does the
// arithmetic: one thread per (microphone, run of consecutive samples), the taps of its microphone in registers,
// a sliding window over the emitted signal (46 loads per 16 outputs instead of 31 per output; the signal is a few
// MB and stays in L2), coalesced (time, channel) stores.  The bound is the HBM write of the result, 8 L M bytes.
#include "common.cuh"

namespace {

constexpr int NT = 31;        // filter_order 30 (the default MonopoleSource uses) + 1
constexpr int R = 16; // outputs per register tile
constexpr int TILES = 8; // tiles per thread: 128 consecutive samples of one microphone

__global__ void __launch_bounds__(128, 3)
delay_mix_kernel(const double *__restrict__ x, const int64_t *__restrict__ x_off, const int64_t *__restrict__ x_len,
                 int n_src, const double *__restrict__ taps, const int32_t *__restrict__ shift, int M, int64_t L,
                 double *__restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int m = (int)(idx % M);
    const int64_t n_begin = (idx / M) * (R * TILES);
    if (n_begin >= L) return;
    for (int tile = 0; tile < TILES; ++tile) {
        const int64_t n0 = n_begin + (int64_t)tile * R;
        if (n0 >= L) break;
        double acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.0;
        for (int s = 0; s < n_src; ++s) {
            double h[NT];
            const double *hp = taps + ((size_t)s * NT) * M + m;
#pragma unroll
            for (int t = 0; t < NT; ++t) h[t] = hp[(size_t)t * M];
            const double *xs = x + x_off[s];
            const int64_t len = x_len[s];
            // x index of window element j: base + j; output r takes tap t = r + NT - 1 - j
            const int64_t base = n0 - shift[(size_t)s * M + m] - (NT - 1);
#pragma unroll
            for (int j = 0; j < R + NT - 1; ++j) {
                if (j % 8 == 0) asm volatile("" ::: "memory");      // keeps ptxas from hoisting all window loads (spills)
                const int64_t i = base + j;
                const double xv = (i >= 0 && i < len) ? xs[i] : 0.0;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int t = r + NT - 1 - j;
                    if (t >= 0 && t < NT) acc[r] = fma(h[t], xv, acc[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (n0 + r < L) out[(size_t)(n0 + r) * M + m] = acc[r];
    }
}

}  // namespace

extern "C" int b2_delay_mix(const void *x_dev, const void *x_offset_dev, const void *x_len_dev, int n_src,
                            const void *taps_dev, const void *shift_dev, int n_taps, int M, int64_t L, void *out_dev,
                            void *stream) {
    B2_REQUIRE(n_taps == NT, B2_EUNSUPPORTED, "b2_delay_mix: %d taps (only filter order 30 = 31 taps is on the GPU path)", n_taps);
    B2_REQUIRE(n_src >= 1 && M >= 1 && L >= 0, B2_EINVAL, "b2_delay_mix: bad sizes (sources %d, M %d, L %lld)", n_src, M, (long long)L);
    if (L == 0) return B2_OK;
    B2_REQUIRE(x_dev && x_offset_dev && x_len_dev && taps_dev && shift_dev && out_dev, B2_EINVAL, "b2_delay_mix: null pointer");
    const int64_t strips = (L + R * TILES - 1) / (R * TILES);
    const int64_t threads = strips * M;
    const int64_t blocks = (threads + 127) / 128;
    B2_REQUIRE(blocks < (1ll << 31), B2_EINVAL, "b2_delay_mix: %lld blocks", (long long)blocks);
    delay_mix_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(
        (const double *)x_dev, (const int64_t *)x_offset_dev, (const int64_t *)x_len_dev, n_src, (const double *)taps_dev,
        (const int32_t *)shift_dev, M, L, (double *)out_dev);
    B2_LAUNCH_OK("delay_mix_kernel");
    return B2_OK;
}
