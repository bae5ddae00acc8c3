// Weighted channel algebra of ProcessingData.get_channel_trace (reference detprocess/process/processing_data.py:1033-1047):
// a YAML channel "a+b" / "a-b" (optionally with weights) is the float64 trace  w_a * a (+|-) w_b * b  of two (or more)
// stored channels.  One launch forms every combined channel of a batch straight from the reader's buffer -- int16 ADC
// counts are converted on the way (adc * gain + offset, numpy's two roundings) -- instead of a chain of elementwise
// PyTorch kernels with a temporary per operator.  Every product and sum is rounded separately and in the reference's
// order, (w_a * a) + (w_b * b) + ..., so the result is bit-identical to numpy on the host-converted traces; a
// subtraction is the sum with the negated weight ((-w) * b == -(w * b) exactly).  HBM bound: terms read, result written once.
#pragma once
#include "dp_platform.cuh"
#include "dp_reduce_kernel.cuh"   // dp_add_rn / dp_mul_rn

#define DP_COMBINE_MAX_OUT 8
#define DP_COMBINE_MAX_TERMS 4

struct DpCombineParams {
    const void* base;         // reader batch, element (event, input row) at event * event_stride + off
    long long event_stride;   // elements
    long long n_events;
    int nb_samples;
    int n_out;
    int n_terms[DP_COMBINE_MAX_OUT];
    long long off[DP_COMBINE_MAX_OUT][DP_COMBINE_MAX_TERMS];
    double w[DP_COMBINE_MAX_OUT][DP_COMBINE_MAX_TERMS];
    double gain[DP_COMBINE_MAX_OUT][DP_COMBINE_MAX_TERMS];   // int16 input only
    double offs[DP_COMBINE_MAX_OUT][DP_COMBINE_MAX_TERMS];
    int weighted[DP_COMBINE_MAX_OUT];   // 0: plain a (+|-) b (weights are +-1 and not multiplied in)
    double* out;              // [n_events][n_out][nb_samples]
};

template <int IN> DP_DEV double dp_combine_sample(const void* base, long long i, double gain, double offs) {
    if constexpr (IN == 0) {
        return __ldg(reinterpret_cast<const double*>(base) + i);
    } else if constexpr (IN == 1) {
        return (double)__ldg(reinterpret_cast<const float*>(base) + i);
    } else {
        return dp_add_rn(dp_mul_rn((double)__ldg(reinterpret_cast<const short*>(base) + i), gain), offs);
    }
}

#ifndef DP_HOST_EMU
template <int IN> __global__ void __launch_bounds__(256) dp_combine_kernel(const DpCombineParams prm) {
    const long long per_event = (long long)prm.n_out * prm.nb_samples;
    const long long total = prm.n_events * per_event;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const long long ev = g / per_event;
        const int r = (int)(g - ev * per_event);
        const int j = r / prm.nb_samples, i = r - j * prm.nb_samples;
        const long long eb = ev * prm.event_stride + i;
        double acc = 0.0;
        for (int t = 0; t < prm.n_terms[j]; ++t) {
            const double x = dp_combine_sample<IN>(prm.base, eb + prm.off[j][t], prm.gain[j][t], prm.offs[j][t]);
            double term;
            if (prm.weighted[j])
                term = dp_mul_rn(prm.w[j][t], x);
            else
                term = prm.w[j][t] < 0 ? -x : x;
            acc = (t == 0) ? term : dp_add_rn(acc, term);
        }
        prm.out[g] = acc;
    }
}
#endif
