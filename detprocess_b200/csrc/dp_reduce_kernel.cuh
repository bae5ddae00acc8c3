// Windowed trace reductions, bit-exact with numpy float64:
//   mean  -> np.mean(trace[a:b])            reference detprocess/core/algorithms.py:698
//   trapz -> np.trapz(trace[a:b]) / fs      reference detprocess/core/algorithms.py:759
//   max   -> np.amax(trace[a:b])            reference detprocess/core/algorithms.py:818
//   min   -> np.amin(trace[a:b])            reference detprocess/core/algorithms.py:879
//
// numpy sums float64 with pairwise_sum_DOUBLE: recursive halving (left half rounded
// down to a multiple of 8) until <= 128 elements, then 8 strided accumulators
// combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail.  The host
// plan (dp_reduce_plan.hpp) flattens that recursion into leaves + a combine tree;
// here 8 lanes play the 8 accumulators of one leaf and the tree is evaluated level
// by level, so every addition happens in numpy's order with numpy's operands.
// HBM-bound: each window sample is read once, coalesced 64 B per leaf row.
#pragma once
#include "dp_platform.cuh"

enum { DP_OP_MEAN = 0, DP_OP_TRAPZ = 1, DP_OP_MAX = 2, DP_OP_MIN = 3 };

struct DpLeaf {
    int off;   // first element (absolute sample index)
    int len;   // number of summed elements
    int node;  // node slot receiving the leaf sum
    int trapz; // element i is (y[i+1] + y[i]) / 2
};
struct DpNode {
    int left, right, out, pad;
};
struct DpRedFeat {
    int op, lo, hi, root;  // root node slot (sum ops)
    int out;               // column in the event's output row
    int n;                 // divisor for mean
    int pad0, pad1;
};
struct DpRedChan {
    int leaf_begin, leaf_end;
    int level_begin, level_end;  // range in level_off[] (level l spans nodes level_off[l]..level_off[l+1])
    int feat_begin, feat_end;
    int pad0, pad1;
};
struct DpReduceParams {
    const void* traces;     // float64, or int16 ADC counts (sample = adc * gain + offset, rounded like numpy: product, then sum)
    const double* adc;      // [n_chan][2] gain, offset (int16 input only)
    // first sample of (event ev, channel c) = element (row_start ? row_start[ev] : ev * event_stride) +
    // (chan_offset ? chan_offset[c] : c * chan_stride) -- the layouts of dp_of1x1_batch_ex (Dp2Params)
    long long event_stride, chan_stride;
    const long long* chan_offset;  // [n_chan] or null
    const long long* row_start;    // [n_events] or null: windows of continuous streams of stream_len samples
    long long stream_len;
    int nb_samples;
    int n_rows, n_chan;
    const DpRedChan* chans;
    const DpLeaf* leaves;
    const DpNode* nodes;
    const int* level_off;
    const DpRedFeat* feats;
    double* out;  // [n_events][n_out]
    int n_out;
    double fs;
    int max_nodes;  // smem doubles
};

#ifndef DP_HOST_EMU
DP_DEV double dp_add_rn(double a, double b) { return __dadd_rn(a, b); }
#else
DP_DEV double dp_add_rn(double a, double b) { return a + b; }
#endif

#ifndef DP_HOST_EMU
DP_DEV double dp_mul_rn(double a, double b) { return __dmul_rn(a, b); }
#else
DP_DEV double dp_mul_rn(double a, double b) { return a * b; }
#endif

// sample i of the row: IN = 0 float64, IN = 2 int16 ADC counts converted as numpy does (adc.astype(float64) * gain + offset:
// two roundings, never contracted into an FMA), so that every reduction stays bit-identical to numpy on the converted trace
template <int IN> struct DpRedRow {
    const void* x;
    double gain, offs;
    DP_DEV double operator[](int i) const {
        if constexpr (IN == 0) {
            return __ldg(reinterpret_cast<const double*>(x) + i);
        } else {
            const short v = __ldg(reinterpret_cast<const short*>(x) + i);
            return dp_add_rn(dp_mul_rn((double)v, gain), offs);
        }
    }
};

template <int IN> DP_DEV double dp_red_elem(const DpRedRow<IN>& x, int i, int trapz) {
    const double a = x[i];
    if (!trapz) return a;
    const double b = x[i + 1];
    return dp_add_rn(b, a) * 0.5;  // exact halving == numpy's / 2.0
}

template <int NT, int IN = 0> DP_DEV void dp_reduce_rows(const DpReduceParams& prm, double* nodeval, double* red) {
    const int tid = threadIdx.x;
    const int lane8 = tid & 7;
    const int grp = tid >> 3;
    constexpr int NG = NT / 8;
    for (int row = blockIdx.x; row < prm.n_rows; row += gridDim.x) {
        const int chan = row % prm.n_chan;
        const int ev = row / prm.n_chan;
        const DpRedChan ch = prm.chans[chan];
        DpRedRow<IN> x;
        long long base = (long long)ev * prm.event_stride;
        if (prm.row_start != nullptr) {
            base = prm.row_start[ev];
            if (base < 0 || base + prm.nb_samples > prm.stream_len) {  // CTA-uniform: the window leaves the stream
                for (int f = ch.feat_begin + tid; f < ch.feat_end; f += NT) prm.out[(long long)ev * prm.n_out + prm.feats[f].out] = -999999.0;
                continue;
            }
        }
        const long long first = base + (prm.chan_offset != nullptr ? prm.chan_offset[chan] : (long long)chan * prm.chan_stride);
        x.x = reinterpret_cast<const unsigned char*>(prm.traces) + (size_t)first * (IN == 0 ? 8 : 2);
        x.gain = IN == 0 ? 1.0 : prm.adc[2 * chan];
        x.offs = IN == 0 ? 0.0 : prm.adc[2 * chan + 1];
        // ---- leaves: 8 lanes = numpy's 8 accumulators -------------------------------
        // (loop bounds are CTA-uniform so the shuffles below are never divergent)
        for (int L0 = ch.leaf_begin; L0 < ch.leaf_end; L0 += NG) {
            const int L = L0 + grp;
            const bool valid = L < ch.leaf_end;
            DpLeaf lf = DpLeaf{0, 0, 0, 0};
            if (valid) lf = prm.leaves[L];
            const bool small = lf.len < 8;
            const int nfull = small ? 0 : lf.len - (lf.len & 7);
            // a leaf has at most 128 elements = 16 rows of 8: all 16 loads of the lane are issued before the first
            // addition (the additions keep numpy's order; one load in flight per thread left the kernel latency bound)
            double v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (8 * j < nfull) ? x[lf.off + 8 * j + lane8] : 0.0;
            if (lf.trapz) {  // element i is (y[i+1] + y[i]) / 2 (exact halving == numpy's / 2.0)
                double w[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) w[j] = (8 * j < nfull) ? x[lf.off + 8 * j + lane8 + 1] : 0.0;
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = dp_add_rn(w[j], v[j]) * 0.5;
            }
            double r = v[0];
#pragma unroll
            for (int j = 1; j < 16; ++j)
                if (8 * j < nfull) r = dp_add_rn(r, v[j]);
            r = dp_add_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
            r = dp_add_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
            r = dp_add_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
            double res = small ? 0.0 : r;
            for (int i = nfull; i < lf.len; ++i) res = dp_add_rn(res, dp_red_elem(x, lf.off + i, lf.trapz));
            if (valid && lane8 == 0) nodeval[lf.node] = res;
        }
        __syncthreads();
        // ---- combine tree, one level per barrier --------------------------------------
        for (int l = ch.level_begin; l < ch.level_end; ++l) {
            const int b = prm.level_off[l], e = prm.level_off[l + 1];
            for (int i = b + tid; i < e; i += NT) {
                const DpNode nd = prm.nodes[i];
                nodeval[nd.out] = dp_add_rn(nodeval[nd.left], nodeval[nd.right]);
            }
            __syncthreads();
        }
        // ---- features ------------------------------------------------------------------
        for (int f = ch.feat_begin; f < ch.feat_end; ++f) {
            const DpRedFeat ft = prm.feats[f];
            double* o = prm.out + (long long)ev * prm.n_out + ft.out;
            if (ft.op == DP_OP_MEAN) {
                if (tid == 0) *o = nodeval[ft.root] / (double)ft.n;
            } else if (ft.op == DP_OP_TRAPZ) {
                if (tid == 0) *o = nodeval[ft.root] / prm.fs;
            } else {
                // max / min with numpy NaN propagation.  A maximum and a minimum over the same window (the usual YAML
                // pair) share one pass over the samples: the first of them in feature order produces the others.
                bool done_earlier = false;
                for (int g = ch.feat_begin; g < f; ++g) {
                    const DpRedFeat fg = prm.feats[g];
                    if ((fg.op == DP_OP_MAX || fg.op == DP_OP_MIN) && fg.lo == ft.lo && fg.hi == ft.hi) done_earlier = true;
                }
                if (done_earlier) continue;  // CTA-uniform
                double mx = -INFINITY, mn = INFINITY;
                int has_nan = 0;
#pragma unroll 4
                for (int i = ft.lo + tid; i < ft.hi; i += NT) {
                    const double v = x[i];
                    if (v != v) has_nan = 1;
                    mx = fmax(mx, v);
                    mn = fmin(mn, v);
                }
#pragma unroll
                for (int o2 = 16; o2 > 0; o2 >>= 1) {
                    has_nan |= __shfl_xor_sync(0xffffffffu, has_nan, o2);
                    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o2));
                    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o2));
                }
                if ((tid & 31) == 0) {
                    red[tid >> 5] = mx;
                    red[32 + (tid >> 5)] = (double)has_nan;
                    red[64 + (tid >> 5)] = mn;
                }
                __syncthreads();
                if (tid == 0) {
                    double amx = red[0], amn = red[64];
                    double nn = red[32];
                    for (int w = 1; w < NT / 32; ++w) {
                        amx = fmax(amx, red[w]);
                        amn = fmin(amn, red[64 + w]);
                        nn += red[32 + w];
                    }
                    const bool is_max = ft.op == DP_OP_MAX;
                    *o = (nn > 0.0) ? NAN : (is_max ? amx : amn);
                    for (int g = f + 1; g < ch.feat_end; ++g) {  // every other extremum over the same window
                        const DpRedFeat fg = prm.feats[g];
                        if ((fg.op == DP_OP_MAX || fg.op == DP_OP_MIN) && fg.lo == ft.lo && fg.hi == ft.hi)
                            prm.out[(long long)ev * prm.n_out + fg.out] = (nn > 0.0) ? NAN : (fg.op == DP_OP_MAX ? amx : amn);
                    }
                }
                __syncthreads();
            }
        }
        __syncthreads();  // nodeval reuse by the next row
    }
}

#ifndef DP_HOST_EMU
template <int NT, int IN = 0> __global__ void __launch_bounds__(NT) dp_reduce_kernel(const DpReduceParams prm) {
    extern __shared__ __align__(16) unsigned char dp_red_smem[];
    double* nodeval = reinterpret_cast<double*>(dp_red_smem);
    double* red = nodeval + prm.max_nodes;
    dp_reduce_rows<NT, IN>(prm, nodeval, red);
}
#endif
