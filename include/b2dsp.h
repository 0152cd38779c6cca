/*
 * b2dsp.h -- C ABI of the B200 (sm_100a) Welch-CSM + delay-and-sum-map path.
 *
 * The reference (nico-franco-gomez/dsptoolbox v0.8) is pure Python and has no FFI of its
 * own: its boundary for this path is Python-object level.  Each entry point below therefore
 * cites the reference *function* whose arithmetic it replaces (paths relative to
 * /root/reference/dsptoolbox).  The Python host side in dsptoolbox_b200/ binds these with
 * ctypes and mirrors the reference's classes (Signal, MicArray, Grid,
 * BeamformerDASFrequency); INTEGRATION.md shows the monkey-patch a reference maintainer adds.
 *
 * Conventions
 *   - every function returns 0 on success, a negative B2_E* code otherwise; it never throws,
 *     never exits, and stores a human-readable message retrievable with b2_last_error()
 *     (thread-local).
 *   - all *_dev pointers are DEVICE pointers on the current CUDA device; the library does no
 *     allocation and no synchronisation: kernels are enqueued on `stream` (a cudaStream_t
 *     passed as void*, 0 = default stream) and the caller owns every buffer.
 *   - re-entrant across streams and devices: per-device kernel attributes are set up on the
 *     first call on each device (keyed by cudaGetDevice()); the only shared state is the
 *     thread-local error string.
 *   - precision: B2_F64 = strict (fp64 FFT, fp64 tensor-core DMMA accumulation);
 *                B2_F32 = fast   (fp32 FFT, fp32 spectra).
 *   - complex numbers are interleaved (re, im).
 */
#ifndef B2DSP_H
#define B2DSP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_OK 0
#define B2_EINVAL (-1)     /* bad argument (shape, alignment, enum)            */
#define B2_EUNSUPPORTED (-2) /* valid in the reference but outside the GPU path */
#define B2_ECUDA (-3)      /* CUDA runtime / launch error                       */

#define B2_F64 0
#define B2_F32 1

/* steering-vector formulations, beamforming/enums.py:4-18 */
#define B2_STEER_CLASSIC 0
#define B2_STEER_INVERSE 1
#define B2_STEER_TRUE_POWER 2
#define B2_STEER_TRUE_LOCATION 3

/* map arithmetic */
#define B2_MAP_SIMT_F64 0  /* fp64 CUDA-core kernel (strict checker-grade)                 */
#define B2_MAP_SIMT_F32 1  /* fp32 CUDA-core kernel                                         */
#define B2_MAP_TC_TF32 2   /* tcgen05 tensor cores, single-pass TF32, fp32 accumulate       */
#define B2_MAP_TC_TF32X3 3 /* tcgen05 tensor cores, split-TF32 (3 products), fp32 accumulate */
#define B2_MAP_TC_MIXED 4  /* tcgen05: TF32 main product + one BF16 product for both corrections */
#define B2_MAP_TC_F16X3 5  /* tcgen05: operands split into two FP16 halves (22 bits), 3 FP16 products,
                              per-bin / per-grid-point power-of-two scaling, fp32 accumulate     */

int b2_version(void);
const char *b2_last_error(void);

/* Number of Welch frames and zero-padding for a signal of L samples:
 * helpers/other.py:181-213 (_compute_number_frames, zero_padding=True):
 * frames = ceil(L/step), pad = wl - L % step. */
int b2_welch_frames(int64_t L, int wl, int step, int64_t *frames, int64_t *pad);

/* Frames-per-CTA granularity of b2_welch_rfft for (wl, precision); the frame capacity Kp
 * of the spectra buffer must be a multiple of max(16, this). */
int b2_welch_frame_tile(int wl, int precision);

/* Fill the length-wl twiddle table exp(-2*pi*i*t/wl), t = 0..wl-1, in the precision's
 * complex type (replaces pocketfft's plan inside np.fft.rfft, _spectral_methods.py:142). */
int b2_fft_twiddles(int wl, int precision, void *twiddle_dev, void *stream);

/* K0: gather the caller's (L, M) recording (any strides, float32 or float64) into the
 * channel-major, zero-padded device layout the FFT kernel reads:
 *   out[m][t] = x[(t0+t)*stride_t + m*stride_m]  for t < nt, 0 for nt <= t < Lp.
 * Replaces the zero-padding of helpers/other.py:216-259 (_pad_trim) and the float64 upcast
 * of _framed_signal_representation.py:57-59.  in_dtype: B2_F64 / B2_F32; strides in elements. */
int b2_prepare_channels(const void *x_dev, int in_dtype, int64_t stride_t, int64_t stride_m,
                        int64_t t0, int64_t nt, int M, void *out_dev, int64_t Lp,
                        int precision, void *stream);

/* K1: framing + window + (optional) mean removal + real FFT, all frames of all channels.
 * Replaces _get_framed_signal (_framed_signal_representation.py:50-65) and the
 * window / detrend / np.fft.rfft lines of _welch (_spectral_methods.py:126-148).
 *   xc_dev      (M, Lp) channel-major samples from b2_prepare_channels, Lp >= (Kp-1)*step+wl
 *   window_dev  wl window values in the precision's real type
 *   fft_scale   1, 1/wl or 1/sqrt(wl): the rfft norm of enums.py:53-75
 *   X_out_dev   spectra, b2_welch_spectra_elems(wl, M, Kp) complex values laid out
 *               [Kp/16 frame chunks][G slot groups][M channels][8 frame pairs][4 slots][2 frames],
 *               G = wl/8 + 1; bin f lives in slot b2_welch_bin_slot(wl, f) (bins 0..wl/4-1 ascending, then
 *               bins wl/2..wl/4+1 descending, then bin wl/4: the order in which the real-FFT untangling
 *               produces them, so that every 32-byte store of the kernel is a quarter of a fully written 128-byte
 *               line); frames >= nframes are written as zeros so the rank-K update may run over whole chunks. */
int b2_welch_rfft(const void *xc_dev, int64_t Lp, int M, int wl, int step, int64_t nframes,
                  const void *window_dev, const void *twiddle_dev, int detrend,
                  double fft_scale, void *X_out_dev, int64_t Kp, int precision, void *stream);

/* Size (complex elements) of the spectra buffer of b2_welch_rfft, and the slot of a bin in it (-1 on bad arguments). */
int64_t b2_welch_spectra_elems(int wl, int M, int64_t Kp);
int b2_welch_bin_slot(int wl, int bin);

/* K2: per-bin Hermitian rank-K accumulation S_f (+)= X_f X_f^H of the bins [bin0, bin0 + nbins).
 * Replaces the M(M+1)/2 `_welch` calls of _csm_welch (_spectral_methods.py:351-369) and the
 * frame sum of np.mean (:152); the division by the frame count happens in b2_csm_finalize
 * so partial sums of frame shards can be all-reduced first -- bin range by bin range, which is
 * what lets the reduction of one range overlap the accumulation of the next.
 *   X_dev   spectra in the layout written by b2_welch_rfft (nb = wl/2 + 1)
 *   S_dev   PACKED lower triangle, [nb][M(M+1)/2] complex128: S_f[j][i] (j >= i) at
 *           f*M(M+1)/2 + j(j+1)/2 + i.  The pointer is the base of the whole [nb] buffer.
 *   accumulate  0: overwrite, 1: add to the existing sums */
int b2_csm_herk(const void *X_dev, int M, int64_t Kp, int nb, int bin0, int nbins, void *S_dev,
                int accumulate, int precision, void *stream);

/* K2e: averaging, physical scaling, DC/Nyquist halving, element-wise principal complex
 * sqrt, Hermitian mirroring.  Replaces _welch (_spectral_methods.py:152,164-171) and
 * _csm_welch (:355-369).  S_dev: packed sums as above; csm_out_dev: [nb][M][M] complex128,
 * fully populated. */
int b2_csm_finalize(const void *S_dev, int M, int nb, double inv_frames, double factor,
                    int halve_edges, int amplitude_sqrt, void *csm_out_dev, void *stream);

/* Auto-spectra only (the diagonal of the CSM): P[f][m] (+)= sum_k |X_f[m,k]|^2 as float64 [nb][M], and
 * its finalisation (mean, factor, DC/Nyquist halving, sqrt for amplitude scalings).  Replaces
 * `_welch(x, None)` (_spectral_methods.py:146-152, 164-171) behind Signal.get_spectrum's Welch branch
 * (classes/signal.py:886-898). */
int b2_welch_autopower(const void *X_dev, int M, int64_t Kp, int nb, void *P_dev, int accumulate,
                       int precision, void *stream);
int b2_auto_finalize(const void *P_dev, int M, int nb, double inv_frames, double factor, int halve_edges,
                     int amplitude_sqrt, void *out_dev, void *stream);

/* CLEAN-SC (float64), all bins of a band in lock-step; one iteration of `_clean_sc_deconvolve`
 * (beamforming/_beamforming.py:246-295) is these two launches plus the arg-max of the dirty map.
 *
 * b2_cleansc_component, one CTA per bin: stopping criterion ||D1||_1 >= ||D0||_1 (:259; a finished bin clears its
 * `active` flag and stays finished), steering vector w to the bin's peak point (distances by the expansion formula of
 * helpers/other.py:131-153, amplitudes of beamforming.py:1515-1703), d = D1 w / P_max, the 20 fixed-point steps
 * h = (d + |h|^2 w) / sqrt(1 + sum |h|^2 |w|^2) (:274-278), G = h h^H P_max (diagonal removed on request, :280-283),
 * D0 <- D1, D1 <- D1 - safety_factor G (:291-294).
 *   D0_dev, D1_dev   [nbins][M][M] complex128, updated in place
 *   peak_dev         [nbins] int64 grid-point index of each bin's maximum;  pmax_dev [nbins] float64 its power
 *   active_dev       [nbins] bytes (in / out)
 *   h_out_dev        [nbins][M] complex128 (zeros for finished bins);  scale_out_dev [nbins] float64: P_max or 0
 *
 * b2_das_rank1_update: out[g][f] += alpha * scale_f * (|sum_m conj(h_gm) v_fm|^2 - [remove_diag] sum_m |h_gm|^2 |v_fm|^2)
 * with the steering vectors h_g of the map kernels -- Re(h_g^H A_f h_g) of the rank-one matrices A_f = scale_f v_f v_f^H
 * in O(G F M), which is what :285-289 subtracts from the dirty map (alpha = -safety_factor).
 *   vectors_dev [nbins][M] complex128, scale_dev [nbins] float64 (0 skips the bin), out_dev [G][nbins] float64 */
int b2_cleansc_component(void *D0_dev, void *D1_dev, int M, int nbins, const void *peak_dev, const void *pmax_dev,
                         void *active_dev, const void *wave_numbers_dev, const void *mic_xyz_dev,
                         const double *center_xyz, const void *grid_xyz_dev, int steering, double safety_factor,
                         int remove_diag, void *h_out_dev, void *scale_out_dev, void *stream);
int b2_das_rank1_update(const void *vectors_dev, const void *scale_dev, int M, int nbins, const void *freqs_dev,
                        double c, const void *mic_xyz_dev, const double *center_xyz, const void *grid_xyz_dev,
                        int64_t G, int steering, int remove_diag, double alpha, void *out_dev, void *stream);

/* Batched Hermitian eigen-decomposition of small matrices, one CTA per matrix (parallel two-sided Jacobi in shared
 * memory, float64).  Replaces the per-bin `np.linalg.eigh(csm)` of BeamformerOrthogonal (beamforming/beamforming.py:
 * 1107) and the `np.linalg.svd(csm)` of BeamformerFunctional (:1201; for a Hermitian matrix the singular
 * triplets are (|lambda|, sign(lambda) v, v)).
 *   A_dev        [nb][M][M] complex128, Hermitian: the lower triangle (and the real part of the diagonal) is read
 *   w_out_dev    [nb][M] float64 eigenvalues, NOT sorted
 *   V_out_dev    [nb][M][M] complex128, column j = eigenvector of w[j]
 *   sweeps_out_dev  optional [nb] int32: Jacobi sweeps each matrix took
 * M <= b2_heig_max_order() (the matrix and the vectors stay in shared memory), else B2_EUNSUPPORTED. */
int b2_heig_max_order(void);
int b2_heig_jacobi(const void *A_dev, int M, int nb, void *w_out_dev, void *V_out_dev, void *sweeps_out_dev,
                   void *stream);

/* Large / arbitrary-length real FFT along the time axis, batched over rows (float64 arithmetic).  Replaces
 * `np.fft.rfft(time_data, axis=0, n=fft_length)` of the single-FFT spectrum (classes/signal.py:903-911;
 * fft_length = scipy.fft.next_fast_len(L, True) or L itself -- not a power of two) and
 * `np.fft.rfft(frames, axis=0, n=fft_length_samples)` of `_stft` (_spectral_methods.py:263) when the FFT length differs
 * from the window length or exceeds the shared-memory kernel.  Power-of-two lengths run as global-memory Stockham
 * passes, every other length as Bluestein's chirp-z transform on top of them.
 *   plan      b2_fft_plan_bytes(n) bytes, filled once per length by b2_fft_plan (needs b2_fft_work_bytes(n, 1) of work)
 *   work      b2_fft_work_bytes(n, batch) bytes of scratch
 *   input     row b starts at element (b / rows_per_group) * stride_group + (b % rows_per_group) * stride_b and
 *             advances by stride_t per sample; n_in samples are read (cropped to n or zero-padded to n: numpy's
 *             rfft(x, n)); optional window (n_in float64) and mean removal of the windowed row (detrend)
 *   output    complex128, bin k of row b at k * out_stride_k + (b / rows_per_group) * out_stride_group
 *             + (b % rows_per_group) * out_stride_b, k = 0 .. n/2, multiplied by `scale` (the rfft norm). */
int64_t b2_fft_plan_bytes(int64_t n);
int64_t b2_fft_work_bytes(int64_t n, int batch);
int b2_fft_plan(int64_t n, void *plan_dev, void *work_dev, void *stream);
int b2_rfft_any(const void *x_dev, int in_dtype, int64_t stride_t, int64_t stride_b, int64_t stride_group,
                int rows_per_group, int64_t n_in, int64_t n, int batch, const void *window_dev, int detrend,
                const void *plan_dev, void *work_dev, double scale, void *out_dev, int64_t out_stride_k,
                int64_t out_stride_b, int64_t out_stride_group, void *stream);

/* Cross-spectral matrix of ONE spectrum per channel (no averaging): csm[f][j][i] = X[f][j] conj(X[f][i]), fully
 * populated [nb][M][M] complex128.  Replaces `_csm_fft` (_spectral_methods.py:374-443): for scalings other than
 * FFTBackward the first and last bin are halved, everything is multiplied by `factor` and the amplitude scalings
 * take the element-wise principal square root (:430-443).  spectrum_dev: [nb][M] complex128. */
int b2_csm_rank1(const void *spectrum_dev, int M, int64_t nb, double factor, int halve_edges, int amplitude_sqrt,
                 void *csm_out_dev, void *stream);

/* Source simulation in front of stage 1: the mixed signals n_src monopole sources leave
 * on M microphones,
 *     out[n][m] = sum_s sum_t taps[s][t][m] * x_s[n - shift[s][m] - t],   n < L,  x_s = 0 outside its x_len[s] samples.
 * Replaces the per-microphone loop of `MonopoleSource.get_signals_on_array` (beamforming/beamforming.py:1427-1458:
 * `fractional_delay(..., keep_length=True)` of standard/latency_delay.py:159-257, gain 1 / (1 + r), append) and the
 * per-source sum of `mix_sources_on_array` (:1461-1511).  The host passes the windowed-sinc taps of
 * `_fractional_delay_filter` (standard/_standard_backend.py:430-493) with the gain folded in, and its integer delay.
 *   x_dev [sum of lengths] float64, source s at x_offset[s] (int64, device) with x_len[s] samples (int64, device);
 *   taps_dev [n_src][n_taps][M] float64;  shift_dev [n_src][M] int32 (may be negative);  out_dev [L][M] float64.
 * n_taps must be 31 (filter order 30, the only one the simulator uses); other orders -> B2_EUNSUPPORTED. */
int b2_delay_mix(const void *x_dev, const void *x_offset_dev, const void *x_len_dev, int n_src, const void *taps_dev,
                 const void *shift_dev, int n_taps, int M, int64_t L, void *out_dev, void *stream);

/* K3a: operand preparation for the map: diagonal removal + M/(M-1) gain
 * (beamforming/beamforming.py:841-845), bin selection (:853-855) and the split of the
 * Hermitian matrix into L = strict lower triangle + diag/2 so that h^H C h = 2 Re(h^H L h).
 *   csm_dev   [nb][M][M] complex128;  bins [bin0, bin0+nbins)
 *   op_dev    operand buffer of b2_das_operand_bytes(M, nbins, mode) bytes */
int64_t b2_das_operand_bytes(int M, int nbins, int mode);
int b2_das_prepare(const void *csm_dev, int M, int nb, int bin0, int nbins, int remove_diag,
                   void *op_dev, int mode, void *stream);

/* K3: delay-and-sum map of grid points [0, G): steering vectors synthesised on the fly
 * (beamforming.py:1515-1703), P[g,f] = Re(h^H C_f h) (:863-868), clamp at 0 when
 * clamp_negative (:871-872), weighted sum over the bins with `weights` = Simpson weights
 * (:875-876) or {1.0} for a single bin (:878).
 *   grid_xyz_dev  (G, 3) float64     mic_xyz_dev (M, 3) float64
 *   center_xyz    reference mic (beamforming.py:517-541), host pointer to 3 doubles
 *   freqs_dev     nbins float64 bin frequencies in Hz
 *   weights_dev   nbins float64
 *   map_out_dev   G float64
 *   per_bin_out_dev  optional (G, nbins) float64 of the clamped per-bin powers, or NULL
 * Two properties of the tensor-core modes a caller must know:
 *   - op_dev also holds the kernel's per-CTA distance tables (written during the launch): one operand
 *     buffer must not be shared by two b2_das_map launches that are in flight at the same time
 *     (consecutive launches on one stream are fine).
 *   - when there are more tiles of 128 grid points than SMs, the bins are split into ranges (sized to
 *     keep the operand images of a range L2-resident) and the weighted partial sums of the ranges are
 *     combined with fp64 atomicAdd: the map is reproducible only up to the order of at most
 *     (number of ranges) fp64 additions per grid point (~1e-16 relative, tests/test_gpu_fullsize.py);
 *     the per-bin output and the CUDA-core modes are bit-reproducible. */
int b2_das_map(const void *op_dev, int M, int nbins, const void *freqs_dev,
               const void *weights_dev, double c, const void *mic_xyz_dev,
               const double *center_xyz, const void *grid_xyz_dev, int64_t G, int steering,
               int clamp_negative, void *map_out_dev, void *per_bin_out_dev, int mode,
               void *stream);

/* Steering vectors materialised like the reference's SteeringVector.get_vector
 * (beamforming.py:1515-1703): h_out_dev complex128 [nbins][M][G].  Small problems only. */
int b2_steering_vector(const void *wave_numbers_dev, int nbins, const void *mic_xyz_dev, int M,
                       const double *center_xyz, const void *grid_xyz_dev, int64_t G,
                       int steering, void *h_out_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B2DSP_H */
