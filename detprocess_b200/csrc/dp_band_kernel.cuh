// Band amplitudes of the event spectrum -- FeatureExtractors.psd_amp (reference detprocess/core/algorithms.py:953-1042) and the
// average_range / single-frequency case of psd_peaks (:1045-1150):
//     psd      = |fft(x) / N / df|^2 * N / fs            (:1006-1016, trace_fft = of_base.signal_fft)
//     psd_fold = one-sided, every bin but DC (and Nyquist, N even) doubled   (qp.utils.fold_spectrum)
//     out[b]   = mean over the one-sided bins k in [bin_lo[b], bin_hi[b]) of sqrt(psd_fold[k])          (:1019-1040)
// The bands are a handful of bins (45-65 Hz, 120-130 Hz lines ...), so the bins are evaluated directly: X[k] = sum_n x[n]
// W^(k n), four bins per sweep over the trace (coalesced loads, the trace of a group re-read from L2), the twiddle of a thread
// advanced by recurrence over its strided samples and re-seeded from the exact table every 32 steps.  One CTA per event.
#pragma once
#include "dp_platform.cuh"

struct DpBandParams {
    const void* base;        // first sample of event 0 of this channel
    int in_dtype;            // DP_IN_F64 / DP_IN_F32 / DP_IN_I16
    long long n_events;
    long long event_stride;  // elements
    int N;
    double gain, offset;     // int16: x = adc * gain + offset
    const int* bin_lo;       // [n_bands] one-sided bin ranges [lo, hi), 1 <= lo < hi <= N/2 + 1
    const int* bin_hi;
    int n_bands;
    const double2* roots;    // [N] exp(-2 pi i j / N)
    double norm;             // N / fs^3
    double* out;             // [n_events][n_bands]
};

#ifdef DP_HOST_EMU
static inline double dp_band_adc(double a, double g, double o) { return a * g + o; }   // built with -ffp-contract=off
#else
__device__ __forceinline__ double dp_band_adc(double a, double g, double o) { return __dadd_rn(__dmul_rn(a, g), o); }
#endif
DP_DEV double dp_band_sample(const DpBandParams& p, const unsigned char* row, int n) {
    if (p.in_dtype == 0) return reinterpret_cast<const double*>(row)[n];
    if (p.in_dtype == 1) return (double)reinterpret_cast<const float*>(row)[n];
    return dp_band_adc((double)reinterpret_cast<const short*>(row)[n], p.gain, p.offset);
}

// the CTA's loop over events; s_re / s_im: [KB][NT / 32] warp partial sums, s_sum: [1] (shared memory)
constexpr int DP_BAND_NT = 256, DP_BAND_KB = 4;
DP_DEV void dp_band_rows(const DpBandParams& p, double* s_re, double* s_im, double* s_sum) {
    constexpr int NT = DP_BAND_NT, KB = DP_BAND_KB, RESEED = 32, NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t esz = p.in_dtype == 0 ? 8 : (p.in_dtype == 1 ? 4 : 2);
    for (long long ev = blockIdx.x; ev < p.n_events; ev += gridDim.x) {
        const unsigned char* row = reinterpret_cast<const unsigned char*>(p.base) + (size_t)ev * (size_t)p.event_stride * esz;
        for (int b = 0; b < p.n_bands; ++b) {
            const int lo = p.bin_lo[b], hi = p.bin_hi[b];
            if (tid == 0) *s_sum = 0.0;
            for (int k0 = lo; k0 < hi; k0 += KB) {
                double ar[KB], ai[KB], wr[KB], wi[KB], sr[KB], si[KB];
#pragma unroll
                for (int j = 0; j < KB; ++j) {
                    const long long k = (k0 + j < hi) ? k0 + j : k0;  // padding bins repeat k0 (discarded below)
                    const double2 st = p.roots[(k * NT) % p.N];       // step of the recurrence: W^(k NT)
                    sr[j] = st.x, si[j] = st.y;
                    ar[j] = ai[j] = 0.0;
                    wr[j] = 1.0, wi[j] = 0.0;
                }
                int step = 0;
                for (int n = tid; n < p.N; n += NT, ++step) {
                    if ((step & (RESEED - 1)) == 0) {
#pragma unroll
                        for (int j = 0; j < KB; ++j) {
                            const long long k = (k0 + j < hi) ? k0 + j : k0;
                            const double2 w = p.roots[(k * n) % p.N];
                            wr[j] = w.x, wi[j] = w.y;
                        }
                    }
                    const double x = dp_band_sample(p, row, n);
#pragma unroll
                    for (int j = 0; j < KB; ++j) {
                        ar[j] = fma(x, wr[j], ar[j]);
                        ai[j] = fma(x, wi[j], ai[j]);
                        const double t = fma(wr[j], sr[j], -(wi[j] * si[j]));
                        wi[j] = fma(wr[j], si[j], wi[j] * sr[j]);
                        wr[j] = t;
                    }
                }
                // every thread takes part in the shuffles (threads past the end of a short trace carry zeros)
#pragma unroll
                for (int j = 0; j < KB; ++j) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        ar[j] += __shfl_down_sync(0xffffffffu, ar[j], o);
                        ai[j] += __shfl_down_sync(0xffffffffu, ai[j], o);
                    }
                    if (lane == 0) s_re[j * NW + warp] = ar[j], s_im[j * NW + warp] = ai[j];
                }
                __syncthreads();
                if (tid == 0) {
                    for (int j = 0; j < KB && k0 + j < hi; ++j) {
                        double re = 0.0, im = 0.0;
                        for (int w = 0; w < NW; ++w) re += s_re[j * NW + w], im += s_im[j * NW + w];
                        const int k = k0 + j;
                        const double fold = (k == 0 || 2 * k == p.N) ? 1.0 : 2.0;
                        *s_sum += sqrt(fold * (re * re + im * im) * p.norm);
                    }
                }
                __syncthreads();
            }
            if (tid == 0) p.out[ev * p.n_bands + b] = *s_sum / (double)(hi - lo);
            __syncthreads();
        }
    }
}

#ifdef DP_BAND_DEFINE_KERNEL   // exactly one translation unit (dp_band_inst.cu)
__global__ void __launch_bounds__(DP_BAND_NT) dp_band_kernel(const DpBandParams p) {
    __shared__ double s_re[DP_BAND_KB * DP_BAND_NT / 32], s_im[DP_BAND_KB * DP_BAND_NT / 32];
    __shared__ double s_sum;
    dp_band_rows(p, s_re, s_im, &s_sum);
}
#endif
