/*
 * detprocess_b200 -- C ABI of the B200-native optimal-filter feature-extraction path.
 *
 * This is the drop-in boundary: plain pointers and sizes, int error codes, no
 * exceptions, no torch types.  Device pointers are CUDA device pointers of the
 * current device; `stream` is a cudaStream_t passed as void* (NULL = default
 * stream).  All hot calls are asynchronous on `stream` and allocate nothing.
 *
 * The reference (spice-herald/detprocess, pure Python) has no FFI; each entry point
 * names the reference interface it replaces (paths relative to the reference tree).
 * INTEGRATION.md shows the ctypes binding a maintainer would add on the
 * reference side.
 */
#ifndef DETPROCESS_B200_H
#define DETPROCESS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* error codes */
#define DP_OK 0
#define DP_ERR_INVALID 1   /* bad argument / configuration (reference raises ValueError) */
#define DP_ERR_CUDA 2      /* CUDA runtime failure */
#define DP_ERR_STATE 3     /* call order (e.g. plan not finalized) */
#define DP_ERR_UNSUPPORTED 4

/* arithmetic precision of the fused OF kernel */
#define DP_PREC_F64 0      /* parity mode: 1e-9 rel vs the float64 reference path */
#define DP_PREC_F32 1      /* fast mode:   1e-5 amp / 1e-4 chi2 */

/* sample type of the trace buffer */
#define DP_IN_F64 0
#define DP_IN_F32 1
#define DP_IN_I16 2

/* window-reduction ops */
#define DP_OP_BASELINE 0   /* np.mean(trace[a:b])        detprocess/core/algorithms.py:651-704 */
#define DP_OP_INTEGRAL 1   /* np.trapz(trace[a:b]) / fs  detprocess/core/algorithms.py:709-765 */
#define DP_OP_MAXIMUM 2    /* np.amax(trace[a:b])        detprocess/core/algorithms.py:771-824 */
#define DP_OP_MINIMUM 3    /* np.amin(trace[a:b])        detprocess/core/algorithms.py:830-885 */

/* doubles written per OF fit: amp, rolled delay index, chi2, lowchi2, timeres */
#define DP_FIT_NOUT 5

const char* dp_last_error(void);         /* thread-local message of the last failing call */
int dp_version(void);
int dp_device_count(int* count);

/* ------------------------------------------------------------------ OF1x1 plan
 * One plan == one qp.OFBase object of the reference, i.e. one
 * (nb_samples, nb_pretrigger, csd_tag/coupling) key built once in
 * ProcessingData.instantiate_OF_base (detprocess/process/processing_data.py:155-433):
 *   dp_of_plan_create        <- qp.OFBase(sample_rate)                       :278
 *   dp_of_plan_set_psd       <- OFBase.set_csd(chan, csd, coupling=...)      :321-326
 *   dp_of_plan_add_template  <- OFBase.add_template(chan, template, tag,
 *                               pretrigger_samples, integralnorm)            :369-376
 *                               + OFBase.calc_phi(chan, tag)                 :379-381
 *   dp_of_plan_add_fit       <- one YAML OF algorithm block: the delay-search
 *                               window qp.OF1x1.calc(...) receives from
 *                               FeatureExtractors.of1x1_nodelay/unconstrained/
 *                               constrained (detprocess/core/algorithms.py:278-570)
 * Channels are dense indices 0..n_chan-1; row r of a trace batch belongs to channel
 * r % n_chan (layout [n_events][n_chan][nb_samples]).
 */
typedef struct dp_of_plan dp_of_plan;

int dp_of_plan_create(dp_of_plan** plan, int nb_samples, double sample_rate, int n_chan, int precision);
void dp_of_plan_destroy(dp_of_plan* plan);

/* psd: two-sided PSD [nb_samples], fftfreq order, A^2/Hz.  coupling_ac != 0 sets psd[0] = inf. */
int dp_of_plan_set_psd(dp_of_plan* plan, int chan, const double* psd, int coupling_ac);

/* returns the template index (>= 0) within the channel in *templ_index */
int dp_of_plan_add_template(dp_of_plan* plan, int chan, const double* templ, int pretrigger_samples, int integralnorm,
                            int* templ_index);

/* Candidate delays are rolled indices [window_lo, window_hi) (zero delay sits at the
 * template's pretrigger_samples); outside != 0 searches the complement.  A no-delay
 * fit is the window [pretrigger, pretrigger+1).  Returns the fit slot in *fit_index. */
int dp_of_plan_add_fit(dp_of_plan* plan, int chan, int templ_index, int window_lo, int window_hi, int outside,
                       int* fit_index);

/* the same with the fit's own lowchi2_fcutoff (Hz; < 0: the plan's default): every YAML block carries its own
 * `lowchi2_fcutoff` kwarg (algorithms.py:280, 357, 438), two blocks of one OFBase may differ */
int dp_of_plan_add_fit_ex(dp_of_plan* plan, int chan, int templ_index, int window_lo, int window_hi, int outside,
                          double lowchi2_fcutoff_hz, int* fit_index);

/* plan default of lowchi2_fcutoff of qp.OF1x1.calc (10000 Hz, algorithms.py:280) */
int dp_of_plan_set_lowchi2_fcutoff(dp_of_plan* plan, double fcutoff_hz);

/* DP_IN_I16 traces are raw ADC counts: sample = adc * gain + offset, converted in the kernel's load (float64 mode: one
 * fused multiply-add in float64; float32 mode: the offset cancels with the AC coupling).  Replaces the host-side
 * conversion of pytesio H5Reader.read_single_event(..., adctoamp=True) (processing_data.py:674-684) so that a
 * quarter of the bytes cross PCIe.  Default gain 1, offset 0.  Before finalize. */
int dp_of_plan_set_adc_conversion(dp_of_plan* plan, int chan, double gain, double offset);

/* interpolate_t0 of qp.OF1x1.calc (parabolic refinement of the best delay, detprocess/core/algorithms.py:415, 543): with
 * on != 0 every fit also reports the amplitude ONE SAMPLE BEFORE and AFTER its best delay (NaN at the ends of the trace),
 * two doubles per fit behind the channel's regular block, at dp_of_plan_neighbour_offset; the three-point parabola itself
 * is host arithmetic.  nb_samples 16384 / 32768 / 65536 or a non power of two.  Before finalize. */
int dp_of_plan_set_neighbours(dp_of_plan* plan, int on);
int dp_of_plan_neighbour_offset(const dp_of_plan* plan, int chan, int fit_index, int* offset);

/* builds the device tables on `device`; the plan is immutable afterwards */
int dp_of_plan_finalize(dp_of_plan* plan, int device);

/* introspection (what OFBase.phi()/norm and OF1x1.get_energy_resolution() return) */
int dp_of_plan_n_out(const dp_of_plan* plan, int* n_out);            /* doubles per event */
int dp_of_plan_fit_offset(const dp_of_plan* plan, int chan, int fit_index, int* offset);
int dp_of_plan_chi0_offset(const dp_of_plan* plan, int chan, int* offset);
int dp_of_plan_get_phi(const dp_of_plan* plan, int chan, int templ_index, double* phi_re_im /* [nb_samples][2] */);
int dp_of_plan_get_norm(const dp_of_plan* plan, int chan, int templ_index, double* norm);
int dp_of_plan_get_template_fft(const dp_of_plan* plan, int chan, int templ_index, double* s_re_im);

/* ----------------------------------------------------------------- OF1x1 batch
 * Replaces, for a whole batch of events, the per-event
 *   OFBase.clear_signal / update_signal(calc_fft=True) / calc_signal_filt /
 *   calc_signal_filt_td      (detprocess/process/processing_data.py:712-772)
 * and every qp.OF1x1(...).calc + get_result_* + get_chisq_nopulse the enabled
 * extractors would run (detprocess/core/algorithms.py:331-341, 410-421, 533-558).
 *
 * traces_dev : [n_events][n_chan][row_stride >= nb_samples] samples of `in_dtype`
 * out_dev    : [n_events][n_out] float64; per channel: chi0 (= chi2 no pulse), then per
 *              fit DP_FIT_NOUT doubles {amp, rolled index of best delay, chi2, lowchi2,
 *              time resolution}.  t0 = (index - pretrigger) / sample_rate.
 */
int dp_of1x1_batch(dp_of_plan* plan, const void* traces_dev, int in_dtype, long long n_events, long long row_stride,
                   double* out_dev, void* stream);

/* Host-buffer convenience used by the plugin layer and the end-to-end benchmark:
 * copies traces_host (pageable or pinned) to the device in chunks on two streams,
 * runs the batch and copies the feature table back.  Synchronous. */
/* Window mode: event i is the nb_samples window of ONE continuous float64 stream that starts at sample
 * start_index_dev[i] -- what ProcessingData.read_next_event does with a trigger dataframe row
 * (detprocess/process/processing_data.py:643-688: read_single_event(trigger_index, trace_length_samples,
 * pretrigger_length_samples)), fused into the kernel's trace load so the [n_events][nb_samples] staging copy
 * never exists.  start = trigger_index - nb_pretrigger_samples; windows that leave [0, n_stream_samples) get
 * -999999.0 in every output column.  Single channel, nb_samples 16384 / 32768 / 65536.
 */
int dp_of1x1_windows(dp_of_plan* plan, const double* stream_dev, long long n_stream_samples,
                     const long long* start_index_dev, long long n_events, double* out_dev, void* stream);

/* General input layout: the first sample of (event ev, plan channel c) is element
 *     (start_index_dev ? start_index_dev[ev] : ev * event_stride) + (chan_offsets ? chan_offsets[c] : c * chan_stride)
 * of base_dev (in_dtype DP_IN_F64 / F32 / I16).
 *   - a reader batch [B][n_file_chan][N] of which the plan fits some channels, consumed in place (no gather / stack
 *     copy): event_stride = n_file_chan * N, chan_offsets[c] = file_channel(c) * N  (host array [n_chan]);
 *   - window mode (start_index_dev != NULL, device int64 [n_events]): every channel is a continuous stream of
 *     n_stream_samples samples (stream c starts at chan_offsets[c] or c * chan_stride) and event ev is the window that
 *     starts at sample start_index_dev[ev] = trigger_index - nb_pretrigger_samples of all of them -- what
 *     ProcessingData.read_next_event does with a trigger dataframe row through H5Reader.read_single_event(trigger_index,
 *     trace_length_samples, pretrigger_length_samples) (detprocess/process/processing_data.py:643-688), fused into the
 *     kernel's trace load.  Windows that leave [0, n_stream_samples) get -999999.0 in every column.
 * Rows that are not aligned to a sample pair (windows, odd strides) need nb_samples 16384 / 32768 / 65536. */
int dp_of1x1_batch_ex(dp_of_plan* plan, const void* base_dev, int in_dtype, long long n_events, long long event_stride,
                      const long long* chan_offsets, long long chan_stride, const long long* start_index_dev,
                      long long n_stream_samples, double* out_dev, void* stream);

int dp_of1x1_batch_host(dp_of_plan* plan, const void* traces_host, int in_dtype, long long n_events,
                        long long row_stride, double* out_host);

/* measured duration (ms) of the last dp_of1x1_batch launch on this plan, from CUDA
 * events recorded on the launch stream; synchronises on the end event. */
int dp_of_plan_last_kernel_ms(dp_of_plan* plan, float* ms);
int dp_of_plan_launch_count(const dp_of_plan* plan, long long* n);

/* --------------------------------------------------------- window reductions
 * dp_reduce_plan_add <- one YAML block of base_algorithm baseline / integral /
 * maximum / minimum with its resolved [window_min_index, window_max_index) slice
 * (FeatureProcessing._get_window_indices, detprocess/process/features.py:1243-1344).
 * Results are bit-identical to numpy float64 (pairwise summation order reproduced).
 */
typedef struct dp_reduce_plan dp_reduce_plan;
int dp_reduce_plan_create(dp_reduce_plan** plan, int nb_samples, double sample_rate, int n_chan);
void dp_reduce_plan_destroy(dp_reduce_plan* plan);
/* returns the feature's index within its channel; its output column is
 * dp_reduce_plan_column(plan, chan, feat_index) (channels first, then add order) */
int dp_reduce_plan_add(dp_reduce_plan* plan, int chan, int op, int window_lo, int window_hi, int* feat_index);
int dp_reduce_plan_column(const dp_reduce_plan* plan, int chan, int feat_index, int* column);
int dp_reduce_plan_finalize(dp_reduce_plan* plan, int device);
int dp_reduce_plan_n_out(const dp_reduce_plan* plan, int* n_out);
/* traces_dev: float64 [n_events][n_chan][row_stride]; out_dev: float64 [n_events][n_out] */
int dp_window_reduce_batch(dp_reduce_plan* plan, const double* traces_dev, long long n_events, long long row_stride,
                           double* out_dev, void* stream);
/* int16 traces are raw ADC counts: sample = adc * gain + offset, converted in the load with numpy's two roundings
 * (adc.astype(float64) * gain + offset), so that every reduction is bit-identical to numpy on the trace the reference's
 * reader converts on the host (H5Reader adctoamp=True, processing_data.py:674-684).  Before finalize. */
int dp_reduce_plan_set_adc_conversion(dp_reduce_plan* plan, int chan, double gain, double offset);
/* in_dtype DP_IN_F64 or DP_IN_I16 */
int dp_window_reduce_batch_raw(dp_reduce_plan* plan, const void* traces_dev, int in_dtype, long long n_events,
                               long long row_stride, double* out_dev, void* stream);
/* the input layouts of dp_of1x1_batch_ex (reader batch consumed in place; windows of continuous streams at
 * start_index_dev, out-of-range windows -> -999999.0) */
int dp_window_reduce_batch_ex(dp_reduce_plan* plan, const void* base_dev, int in_dtype, long long n_events,
                              long long event_stride, const long long* chan_offsets, long long chan_stride,
                              const long long* start_index_dev, long long n_stream_samples, double* out_dev, void* stream);
int dp_reduce_plan_last_kernel_ms(dp_reduce_plan* plan, float* ms);

/* ------------------------------------------------------------ channel algebra
 * ProcessingData.get_channel_trace for "a+b" / "a-b" channels (detprocess/process/processing_data.py:1033-1047):
 * out[ev][j][i] = sum_t weights[j][t] * x(ev, offsets[j][t] + i), j < n_out <= 8, t < n_terms[j] <= 4, every product and
 * sum rounded separately in that order (bit-identical to numpy; a subtraction is a negative weight; weighted[j] == 0: the
 * weights are +-1 signs only).  x is read from base_dev (element ev * event_stride + offset + i) as in_dtype; int16 ADC
 * counts become adc * adc_gain[j][t] + adc_offset[j][t] first.  Arrays [n_out][4] on the host; out_dev float64
 * [n_events][n_out][nb_samples]. */
int dp_channel_combine(const void* base_dev, int in_dtype, long long n_events, long long event_stride, int nb_samples,
                       int n_out, const int* n_terms, const long long* offsets, const double* weights, const int* weighted,
                       const double* adc_gain, const double* adc_offset, double* out_dev, void* stream);

/* ------------------------------------------------------------------ noise PSD
 * Replaces qp.calc_psd(traces[cut], fs, folded_over=False) as called by Noise.calc_psd
 * (detprocess/core/noise.py:344): the plan accumulates sum_traces |fft(x)_k|^2 for
 * k = 0..N/2 over any number of dp_psd_accumulate calls (float64, float32 or int16 traces; mask_dev selects the traces that
 * passed the cut, noise.py:331; NULL = all).  dp_psd_get_sums returns the per-GPU sums and
 * the number of accepted traces; the host layer all-reduces both over NCCL and forms
 * psd[k] = sum[k] / (count * N * fs), mirrored to the two-sided layout.
 */
typedef struct dp_psd_plan dp_psd_plan;
int dp_psd_plan_create(dp_psd_plan** plan, int nb_samples, double sample_rate, int precision, int device);
void dp_psd_plan_destroy(dp_psd_plan* plan);
int dp_psd_plan_set_scale(dp_psd_plan* plan, double typical_rms);   /* fp32 mode: sample scale hint */
int dp_psd_reset(dp_psd_plan* plan, void* stream);
int dp_psd_accumulate(dp_psd_plan* plan, const void* traces_dev, int in_dtype, long long n_traces,
                      long long row_stride, const unsigned char* mask_dev, void* stream);
int dp_psd_get_sums(dp_psd_plan* plan, double* sums_dev /* [N/2+1] */, unsigned long long* count_dev /* [1] */,
                    void* stream);
int dp_psd_plan_last_kernel_ms(dp_psd_plan* plan, float* ms);

/* ------------------------------------------------------- continuous-stream OF trigger
 * Replaces OptimumFilterTrigger.update_trace + find_triggers_once for one trigger channel and
 * one amplitude (detprocess/core/oftrigger.py:588-679, 884-1034).  phi_td is the reference's
 * `self._phi_td[0,0]` (real(ifft(phi_fd with the DC bin zeroed)), oftrigger.py:491-493), iw / w its
 * 1x1 `_iw_matrix` / `_w_matrix`:  filtered = iw * oaconvolve(trace, phi_td, 'same'),
 * delta_chi2 = filtered^2 * w.  dp_trigger_run filters the stream (overlap-save FFT chunks in shared
 * memory), thresholds delta_chi2 > chi2_threshold (host computes it from sigma, oftrigger.py:961-965),
 * merges indices closer than pileup_window_samples and returns, per group, the first arg-max:
 * trig_index = argmax + index_shift (pretrigger - nb_samples/2, oftrigger.py:456,1005), the
 * filtered amplitude and delta_chi2 there.  *n_triggers_dev receives the number of groups found
 * (only the first max_triggers are stored).  padding != 0 zeroes the edges like the reference.
 */
typedef struct dp_trigger_plan dp_trigger_plan;
int dp_trigger_plan_create(dp_trigger_plan** plan, const double* phi_td, int nb_filter, double iw, double w,
                           int precision, long long max_samples, int device);
void dp_trigger_plan_destroy(dp_trigger_plan* plan);
int dp_trigger_plan_set_scale(dp_trigger_plan* plan, double typical_rms);   /* fp32 mode: sample scale hint */
int dp_trigger_plan_geometry(const dp_trigger_plan* plan, int* fft_size, int* hop);
int dp_trigger_run(dp_trigger_plan* plan, const double* trace_dev, long long n_samples, double chi2_threshold,
                   long long pileup_window_samples, long long index_shift, int padding, long long* trig_index_dev,
                   double* trig_amp_dev, double* trig_dchi2_dev, int max_triggers, int* n_triggers_dev, void* stream);
/* same on a stream of in_dtype samples (DP_IN_F64 / DP_IN_F32 / DP_IN_I16, e.g. raw ADC counts) */
int dp_trigger_run_raw(dp_trigger_plan* plan, const void* trace_dev, int in_dtype, long long n_samples, double chi2_threshold,
                       long long pileup_window_samples, long long index_shift, int padding, long long* trig_index_dev,
                       double* trig_amp_dev, double* trig_dchi2_dev, int max_triggers, int* n_triggers_dev, void* stream);
int dp_trigger_plan_last_kernel_ms(dp_trigger_plan* plan, float* filter_ms, float* group_ms);

/* Follow-up calls on the candidate list of the last dp_trigger_run(_raw) of the plan (the samples above threshold, ordered by
 * stream index) -- what OptimumFilterTrigger.find_triggers adds around find_triggers_once (oftrigger.py:682-845):
 *
 * dp_trigger_candidates      the list as flat arrays: unshifted stream index, filtered amplitude and (dchi2_dev may be NULL)
 *                            delta chi2 -- amp^2 w, or the residual delta chi2 once dp_trigger_residual_run has run.
 *                            *n_candidates_dev = length of the list (only the first max_candidates entries are stored).
 *                            The dynamic pile-up window (oftrigger.py:78-141, a Python callable per candidate) groups it on the host.
 * dp_trigger_filtered_at     filtered = iw * oaconvolve(trace, phi_td, 'same') at n_idx arbitrary stream indices (direct sums):
 *                            the residual pass takes the pulse amplitude at the SHIFTED trigger index (oftrigger.py:794).
 * dp_trigger_residual_run    the second pass of residual=True (oftrigger.py:772-824): every listed pulse subtracts
 *                            pulse_amp2[i] * shape[t - pulse_start[i]] (shape = delta chi2 trace of a unit pulse, pulse_start ascending,
 *                            subtraction in list order) from the candidates it covers; candidates whose residual still exceeds
 *                            chi2_threshold are kept (the list is compacted in place), grouped by pileup_window_samples and
 *                            reported like dp_trigger_run (trig_dchi2 = the residual delta chi2 at the arg-max). */
int dp_trigger_candidates(dp_trigger_plan* plan, long long* idx_dev, double* amp_dev, double* dchi2_dev, long long max_candidates,
                          long long* n_candidates_dev, void* stream);
int dp_trigger_filtered_at(dp_trigger_plan* plan, const void* trace_dev, int in_dtype, long long n_samples, const long long* idx_dev,
                           int n_idx, double* filtered_dev, void* stream);
int dp_trigger_residual_run(dp_trigger_plan* plan, const long long* pulse_start_dev, const double* pulse_amp2_dev, int n_pulses,
                            const double* shape_dev, int n_shape, double chi2_threshold, long long pileup_window_samples,
                            long long index_shift, long long* trig_index_dev, double* trig_amp_dev, double* trig_dchi2_dev,
                            int max_triggers, int* n_triggers_dev, void* stream);

/* ------------------------------------------------------------------ band amplitudes of the event spectrum
 * FeatureExtractors.psd_amp (detprocess/core/algorithms.py:953-1042) and the average_range / single-frequency case of psd_peaks
 * (:1045-1150): out[event][b] = mean over the one-sided bins k in [bin_lo[b], bin_hi[b]) of sqrt(psd_fold[k]), with
 * psd = |fft(x) / N / df|^2 * N / fs folded to one side (every bin but DC and Nyquist doubled).  The host turns f_lims into bin
 * ranges like utils.get_ind_freq_ranges (+1: the reference drops the DC bin first).  One channel per call: base_dev = first sample of
 * event 0 of that channel, event_stride in elements; int16 samples are converted as adc * gain + offset. */
typedef struct dp_band_plan dp_band_plan;
int dp_band_plan_create(dp_band_plan** plan, int nb_samples, double sample_rate, const int* bin_lo, const int* bin_hi, int n_bands,
                        int device);
void dp_band_plan_destroy(dp_band_plan* plan);
int dp_band_amplitudes(dp_band_plan* plan, const void* base_dev, int in_dtype, long long n_events, long long event_stride, double adc_gain,
                       double adc_offset, double* out_dev /* [n_events][n_bands] */, void* stream);

/* ------------------------------------------------------------------ NxM optimal filter
 * Replaces qp.OFnxm(of_base, channels, template_tag).calc() + get_fit_withdelay(window...) + get_fit_nodelay() as driven
 * per event by FeatureExtractors.ofnxm (reference detprocess/core/algorithms.py:141-274): n channels with an n x n
 * cross-spectral density, m templates that share one time delay.  Output row per event (float64, n_out = 4 + 2 m):
 *   chi0, chi2_constrained, index_constrained (rolled: zero delay = pretrigger_samples), amps_constrained[m],
 *   chi2_nodelay, amps_nodelay[m].           t0 = (index - pretrigger_samples) / sample_rate. */
typedef struct dp_nxm_plan dp_nxm_plan;
int dp_nxm_plan_create(dp_nxm_plan** plan, int nb_samples, double sample_rate, int n_chan, int n_templ, int precision);
void dp_nxm_plan_destroy(dp_nxm_plan* plan);
/* templates: [n_chan][n_templ][nb_samples] float64 (of_base.template(channel, tag), algorithms.py:196);
 * csd: [n_chan][n_chan][nb_samples] complex128 as (re, im) pairs, two-sided, fftfreq order (filterdata.py:380 get_csd);
 * coupling_ac != 0 drops the DC bin as OFBase.set_csd(coupling='AC') does (processing_data.py:321-326) */
int dp_nxm_plan_set_filter(dp_nxm_plan* plan, const double* templates, const double* csd, int pretrigger_samples,
                           int coupling_ac);
/* delay window [window_lo, window_hi) in rolled indices, outside != 0 searches the complement
 * (window_min_index / window_max_index / lgc_outside_window of algorithms.py:146-150); default: every delay.
 * May be changed between batches. */
int dp_nxm_plan_set_window(dp_nxm_plan* plan, int window_lo, int window_hi, int outside);
int dp_nxm_plan_finalize(dp_nxm_plan* plan, int device);
int dp_nxm_plan_n_out(const dp_nxm_plan* plan, int* n_out);
/* the m x m template matrix P (row major) and its inverse (qp.OFBase.calc_p_and_p_inverse) */
int dp_nxm_plan_get_p_matrix(const dp_nxm_plan* plan, double* p_matrix, double* p_inverse);
/* traces_dev: float64 [n_events][n_chan][nb_samples] with element strides event_stride / chan_stride;
 * out_dev: [n_events][n_out] */
int dp_ofnxm_batch(dp_nxm_plan* plan, const double* traces_dev, long long n_events, long long event_stride,
                   long long chan_stride, double* out_dev, void* stream);
int dp_nxm_plan_last_kernel_ms(dp_nxm_plan* plan, float* ms);

/* ------------------------------------------------------------------ noise cross-spectral density
 * Replaces qp.calc_csd(traces[cut], fs=fs, folded_over=False) of Noise.calc_csd (reference detprocess/core/noise.py:374-470;
 * mask = the per-channel autocuts AND-ed at :431-445).  Accumulates sum_e X_a[k] conj(X_b[k]) on the one-sided bins
 * k = 0..N/2 over any number of dp_csd_accumulate calls.  dp_csd_get_sums writes [n_chan * n_chan][N/2 + 1] float64:
 * rows 0..n-1 the diagonal sums |X_a|^2, then for every pair a < b (lexicographic) a row of real parts and a row of
 * imaginary parts of X_a conj(X_b); csd[a][b][k] = sums / (count * N * fs), csd[b][a] = conj, csd[..][N - k] = conj. */
typedef struct dp_csd_plan dp_csd_plan;
int dp_csd_plan_create(dp_csd_plan** plan, int nb_samples, double sample_rate, int n_chan, int precision, int device);
void dp_csd_plan_destroy(dp_csd_plan* plan);
int dp_csd_plan_set_scale(dp_csd_plan* plan, double typical_rms);
int dp_csd_reset(dp_csd_plan* plan, void* stream);
int dp_csd_accumulate(dp_csd_plan* plan, const double* traces_dev, long long n_events, long long event_stride,
                      long long chan_stride, const unsigned char* mask_dev, void* stream);
int dp_csd_get_sums(dp_csd_plan* plan, double* sums_dev, unsigned long long* count_dev, void* stream);
int dp_csd_plan_last_kernel_ms(dp_csd_plan* plan, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* DETPROCESS_B200_H */
