// Fused per-event OF1x1 kernel: real FFT -> x phi -> inverse real FFT -> windowed
// argmax, chi-square by Parseval, low-frequency chi-square; one HBM read per trace.
//
// Replaces, per event: qp.OFBase.update_signal(calc_fft=True) + calc_signal_filt +
// calc_signal_filt_td (reference detprocess/process/processing_data.py:763-772) and
// qp.OF1x1.calc / get_result_* (reference detprocess/core/algorithms.py:331-341,
// 410-421, 533-558).
//
// Algorithm (all indices refer to the packed complex sequence c[m] = x[2m] + i x[2m+1],
// m < M = N/2; for P = 2 the complex FFT is split once more, z_p[m'] = c[2m'+p], and
// the outer radix-2 is merged into the point-wise stage):
//   M' = M/P = 512*R1 = R1 * 32 * 16      (R1 = 2..32),  NT = 16*R1 threads per CTA
//   fwd pass 1  radix R1, stride 512 : global -> registers -> DFT -> twiddle -> smem
//   fwd pass 2  radix 32, stride 16  : smem -> DFT -> twiddle -> smem (in place)
//   fwd pass 3  radix 16, stride 1   : smem -> DFT -> registers.  Each thread owns the
//               two butterflies that hold Z[k] and Z[M'-k] for 16 k's, so the
//               real-FFT untangle, the filter multiply, chi0 and the inverse
//               untangle are thread-local (no exchange, no mirror tables).
//   inverse     mirrored passes 3',2',1'; the last pass leaves the amplitude-vs-delay
//               samples in registers where the windowed arg-max runs.
// Shared memory holds ONE complex array of M' elements (+1/16 padding): element p
// lives at p + (p >> 4), which makes every access pattern above bank-conflict free
// for 8- and 16-byte elements.
#pragma once
#include "dp_fft.cuh"

#define DP_MAX_TEMPLATES 4
#define DP_MAX_SLOTS 8   // fits per channel
#define DP_MAX_TSLOTS 4  // fits per template
#define DP_NLOW_MAX 1024
#define DP_SLOT_NOUT 5  // amp, ind, chi2, lowchi2, timeres

struct DpSlot {
    int templ;    // template index within the channel
    int lo, hi;   // candidate rolled delay indices [lo, hi)
    int outside;  // 1: candidates are the complement of [lo, hi)
    int nlow;     // lowchi2 sums the bins k < nlow of this fit (its own lowchi2_fcutoff; <= the plan's nlow)
};

template <class T> struct DpTemplDev {
    const cx<T>* phi;       // [32*P][NT] thread-order filter
    const cx<T>* phi_self;  // [17][2*P] filter at the bins of the self-paired butterflies
    const cx<T>* s_low;  // [nlow] scaled template spectrum, natural order
    double norm;         // QETpy OF norm
    double tsum;         // sum((2 pi f)^2 |s|^2 / J) * df
    int pretrigger;      // roll applied to the delay axis
    int pad_;
};

template <class T> struct DpChanDev {
    const T* wj;       // [32*P][NT] thread-order chi0 weights
    const T* wj_self;  // [17][2*P] weights at the bins of the self-paired butterflies (0 for duplicates)
    const T* wj_low;   // [nlow] natural order
    double adc_gain;   // int16 traces (raw ADC counts): sample = adc * adc_gain + adc_offset
    double adc_offset;
    int n_templ;
    int n_slots;
    int out_base;  // offset (doubles) of this channel's block in an event's output row
    int pad_;
    DpTemplDev<T> templ[DP_MAX_TEMPLATES];
    DpSlot slots[DP_MAX_SLOTS];
};

template <class T> struct DpOfParams {
    const void* traces;      // samples (in_dtype); first sample of (event ev, channel c) = element
                             // ev * event_stride + (chan_offset ? chan_offset[c] : c * chan_stride)
    long long event_stride;  // elements
    long long chan_stride;   // elements
    const long long* chan_offset;  // [n_chan] or null
    int n_rows;              // events * n_chan; row r belongs to channel r % n_chan
    int n_chan;
    const DpChanDev<T>* chans;
    const cx<T>* tw1;   // [512]  exp(-2 pi i m / M')
    const cx<T>* tw2;   // [16]   exp(-2 pi i m2 / 512)
    const cx<T>* twn;   // [NT]   exp(-2 pi i K12(t) / N)   thread order
    const cx<T>* twp;   // [NT]   exp(-2 pi i K12(t) / M)   thread order (P = 2 only)
    cx<T>* scratch;     // [grid][scratch_per_cta] thread-private spill (P = 2 / multi-template)
    long long scratch_per_cta;
    double* out;        // [n_events][n_out]
    int n_out;
    int nlow;           // number of low-frequency bins (k < nlow) in lowchi2
    double scale;       // power-of-two pre-scale applied to the trace (fp32 range safety)
    int subtract_first; // 1: subtract sample 0 before conversion (fp32 mode, AC coupling)
    int in_dtype;       // 0: f64, 1: f32, 2: i16
};

template <class T> DP_DEV long long dp_first_sample(const DpOfParams<T>& prm, int row) {
    const int ev = row / prm.n_chan, chan = row % prm.n_chan;
    return (long long)ev * prm.event_stride + (prm.chan_offset != nullptr ? prm.chan_offset[chan] : (long long)chan * prm.chan_stride);
}

// ------------------------------------------------------------------- geometry
template <int R1> struct DpGeom {
    static constexpr int MS = 512 * R1;  // sub-FFT size M'
    static constexpr int NT = 16 * R1;   // threads per CTA
    static constexpr int NB1 = 32 / R1;  // pass-1 butterflies per thread
    static constexpr int KQ = 32 * R1;   // MS/16: k-stride between elements of a pass-3 butterfly
    static DP_HD int phys(int p) { return p + (p >> 4); }
    static constexpr int SMEM_ELEMS = MS + MS / 16;
    // thread t -> (K12, butterflies A and B).  A holds k = K12 + KQ*r, B holds
    // k = K12' + KQ*r with K12' = KQ - K12, so A[r] pairs with B[15-r].
    // t == 0 owns the two self-paired butterflies K12 = 0 and K12 = KQ/2.
    static DP_HD void map(int t, int& K12, int& bA, int& bB) {
        const int k1 = t >> 4, k2 = t & 15;
        K12 = k1 + R1 * k2;
        bA = k1 * 32 + k2;
        int k1p, k2p;
        if (k1 != 0) {
            k1p = R1 - k1;
            k2p = 31 - k2;
        } else if (k2 != 0) {
            k1p = 0;
            k2p = 32 - k2;
        } else {
            k1p = 0;
            k2p = 16;
        }
        bB = k1p * 32 + k2p;
    }
};

// ------------------------------------------------------------------ reductions
DP_DEV double dp_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the CTA; result returned to every thread.  red: >= 33 doubles of smem.
template <int NT> DP_DEV double dp_block_sum(double v, double* red) {
    constexpr int NW = (NT + 31) / 32;
    v = dp_warp_sum(v);
    const int tid = threadIdx.x;
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
        double s = (tid < NW) ? red[tid] : 0.0;
        s = dp_warp_sum(s);
        if (tid == 0) red[32] = s;
    }
    __syncthreads();
    const double r = red[32];
    __syncthreads();
    return r;
}

// arg-max of |val| with smallest-index tie break.  idx < 0 means "no candidate".
template <class T> struct DpBest {
    T val;
    int idx;
};
DP_DEV float dp_abs(float a) { return fabsf(a); }
DP_DEV double dp_abs(double a) { return fabs(a); }
template <class T> DP_DEV void dp_best_merge(DpBest<T>& a, const DpBest<T>& b) {
    const T ka = dp_abs(a.val), kb = dp_abs(b.val);
    const bool take = (b.idx >= 0) && (a.idx < 0 || kb > ka || (kb == ka && b.idx < a.idx));
    if (take) a = b;
}
template <class T> DP_DEV DpBest<T> dp_warp_best(DpBest<T> a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        DpBest<T> b;
        b.val = __shfl_xor_sync(0xffffffffu, a.val, o);
        b.idx = __shfl_xor_sync(0xffffffffu, a.idx, o);
        dp_best_merge(a, b);
    }
    return a;
}

DP_DEV void dp_prefetch_l2(const void* p) {
#ifndef DP_HOST_EMU
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// ------------------------------------------------------------------ trace loads
// Element j of the packed complex sequence is the sample pair (x[2j], x[2j+1]).  The
// input type is a template parameter so that the R1 loads of a butterfly are issued
// back to back (a runtime switch made ptxas serialise them: one load in flight).
template <int IN> struct DpRaw;
template <> struct DpRaw<0> { using type = double2; using scalar = double; };
template <> struct DpRaw<1> { using type = float2; using scalar = float; };
template <> struct DpRaw<2> { using type = short2; using scalar = short; };
template <> struct DpRaw<3> { using type = double2; using scalar = double; };  // float64, rows only 8-byte aligned
template <> struct DpRaw<4> { using type = float2; using scalar = float; };    // float32, rows only 4-byte aligned
template <> struct DpRaw<5> { using type = short2; using scalar = short; };    // int16, rows only 2-byte aligned
// IN 3..5 = IN - 3 with element-aligned rows (windows of a continuous stream start at any sample)
DP_HD constexpr bool dp_in_is_adc(int in) { return in == 2 || in == 5; }

template <int IN> DP_DEV typename DpRaw<IN>::type dp_load_raw(const void* row, long long j) {
    if constexpr (IN >= 3) {
        const typename DpRaw<IN>::scalar* p = reinterpret_cast<const typename DpRaw<IN>::scalar*>(row) + 2 * j;
        typename DpRaw<IN>::type v;
        v.x = __ldg(p);
        v.y = __ldg(p + 1);
        return v;
    } else {
        return __ldg(reinterpret_cast<const typename DpRaw<IN>::type*>(row) + j);
    }
}
template <class T, int IN> DP_DEV cx<T> dp_convert_raw(typename DpRaw<IN>::type d, double x0, double sc) {
    return cx<T>{(T)(((double)d.x - x0) * sc), (T)(((double)d.y - x0) * sc)};
}
template <int IN> DP_DEV double dp_load_first(const void* row) {
    return (double)__ldg(reinterpret_cast<const typename DpRaw<IN>::scalar*>(row));
}

// Conversion of a raw sample to the kernel's working value: (raw - x0) * sc.  Float traces: x0 = first sample (fp32 mode,
// AC coupling) or 0, sc = the plan's power-of-two scale.  int16 ADC counts (IN == 2): sample = adc * gain + offset
// = (adc + offset / gain) * gain, so x0 = -offset / gain (or the first count when it is subtracted anyway) and
// sc = gain * scale -- the same conversion pytesio applies at read time (adctoamp=True, processing_data.py:675-684).
template <class T, int IN> DP_DEV void dp_row_conversion(const DpOfParams<T>& prm, int chan, const void* xrow, double& x0, double& sc) {
    x0 = prm.subtract_first ? dp_load_first<IN>(xrow) : 0.0;
    sc = prm.scale;
    if constexpr (IN == 2) {
        const double gain = dp_ldg(&prm.chans[chan].adc_gain), offs = dp_ldg(&prm.chans[chan].adc_offset);
        if (!prm.subtract_first) x0 = -offs / gain;
        sc = gain * prm.scale;
    }
}

// ------------------------------------------------------------- forward sub-FFT
// Loads sub-sequence z[j] = c[P*j + p] from global, runs passes 1..3 and leaves
// Z[K12 + KQ*r] in za[r], Z[K12' + KQ*r] in zb[r].
template <class T, int R1, int P, int IN>
DP_DEV void dp_fwd_subfft(const void* row, int p, double x0, double sc, cx<T>* buf, const cx<T>* DP_RESTRICT tw1,
                          const cx<T>* DP_RESTRICT tw2, int bA, int bB, cx<T> (&za)[16], cx<T> (&zb)[16]) {
    using G = DpGeom<R1>;
    const int tid = threadIdx.x;
    // pass 1: radix R1 over n1 (stride 512); loads issued in batches of <= 16 (64 registers of raw f64)
#pragma unroll
    for (int i = 0; i < G::NB1; ++i) {
        const int m = tid + i * G::NT;
        cx<T> v[R1];
        constexpr int BATCH = R1 < 16 ? R1 : 16;
#pragma unroll
        for (int n0 = 0; n0 < R1; n0 += BATCH) {
            typename DpRaw<IN>::type raw[BATCH];
#pragma unroll
            for (int n = 0; n < BATCH; ++n) raw[n] = dp_load_raw<IN>(row, (long long)P * (m + (n0 + n) * 512) + p);
#pragma unroll
            for (int n = 0; n < BATCH; ++n) v[n0 + n] = dp_convert_raw<T, IN>(raw[n], x0, sc);
        }
        dp_dft<R1, -1, T>::run(v);
        dp_twiddle<R1, false, T>(v, dp_ldg(tw1 + m));
#pragma unroll
        for (int k = 0; k < R1; ++k) buf[G::phys(k * 512 + m)] = v[k];
    }
    __syncthreads();
    // pass 2: radix 32 over n2 (stride 16) inside block k1
    {
        const int base = (tid >> 4) * 512 + (tid & 15);
        cx<T> v[32];
#pragma unroll
        for (int n = 0; n < 32; ++n) v[n] = buf[G::phys(base + n * 16)];
        dp_dft<32, -1, T>::run(v);
        dp_twiddle<32, false, T>(v, dp_ldg(tw2 + (tid & 15)));
#pragma unroll
        for (int k = 0; k < 32; ++k) buf[G::phys(base + k * 16)] = v[k];
    }
    __syncthreads();
    // pass 3: radix 16 over n3 (stride 1), two butterflies per thread, kept in registers
#pragma unroll
    for (int n = 0; n < 16; ++n) {
        za[n] = buf[G::phys(16 * bA + n)];
        zb[n] = buf[G::phys(16 * bB + n)];
    }
    dp_dft<16, -1, T>::run(za);
    dp_dft<16, -1, T>::run(zb);
}

// ------------------------------------------------------------- inverse sub-FFT
// Consumes Z' in (za, zb); leaves z'[j], j = tid + i*NT + 512*n, in y[i*R1 + n].
// Starts with a barrier-free in-place store (own positions) and ends without a
// barrier: the caller must __syncthreads() before the next store into buf.
template <class T, int R1>
DP_DEV void dp_inv_subfft(cx<T>* buf, const cx<T>* DP_RESTRICT tw1, const cx<T>* DP_RESTRICT tw2, int bA, int bB,
                          cx<T> (&za)[16], cx<T> (&zb)[16], cx<T> (&y)[32]) {
    using G = DpGeom<R1>;
    const int tid = threadIdx.x;
    dp_dft<16, +1, T>::run(za);
    dp_dft<16, +1, T>::run(zb);
#pragma unroll
    for (int n = 0; n < 16; ++n) {
        buf[G::phys(16 * bA + n)] = za[n];
        buf[G::phys(16 * bB + n)] = zb[n];
    }
    __syncthreads();
    {
        const int base = (tid >> 4) * 512 + (tid & 15);
#pragma unroll
        for (int k = 0; k < 32; ++k) y[k] = buf[G::phys(base + k * 16)];
        dp_twiddle<32, true, T>(y, dp_ldg(tw2 + (tid & 15)));
        dp_dft<32, +1, T>::run(y);
#pragma unroll
        for (int n = 0; n < 32; ++n) buf[G::phys(base + n * 16)] = y[n];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < G::NB1; ++i) {
        const int m = tid + i * G::NT;
        cx<T> v[R1];
#pragma unroll
        for (int k = 0; k < R1; ++k) v[k] = buf[G::phys(k * 512 + m)];
        dp_twiddle<R1, true, T>(v, dp_ldg(tw1 + m));
        dp_dft<R1, +1, T>::run(v);
#pragma unroll
        for (int n = 0; n < R1; ++n) y[i * R1 + n] = v[n];
    }
}

// ------------------------------------------------------- windowed arg-max scans
// y[i*R1 + n] holds the complex sample j = tid + i*NT + 512*n of the sub-sequence,
// i.e. the two real amplitude samples with rolled delay index
//     r = RS*j + RO (+ RS/2 for the imaginary part),   RS = 2*P, RO = 2*p
// (the roll by `pretrigger` is folded into phi on the host, so the output index IS
// the rolled index).  Scan order is ascending r, so a strict '>' keeps the first
// (smallest-index) maximum exactly like numpy's argmin on the chi-square.
template <class T, int R1, int P> struct DpScan {
    static constexpr int NT = 16 * R1, NB1 = 32 / R1, RS = 2 * P;
    // every sample is a candidate
    static DP_DEV void full(const cx<T> (&y)[32], int tid, int p, DpBest<T>& b) {
        T vb = y[0].re;
        int pb = 0;
#pragma unroll
        for (int n = 0; n < R1; ++n) {
#pragma unroll
            for (int i = 0; i < NB1; ++i) {
                const int e = i * R1 + n;
                const int off = RS * (i * NT + 512 * n);
                if (e != 0) {
                    if (dp_abs(y[e].re) > dp_abs(vb)) {
                        vb = y[e].re;
                        pb = off;
                    }
                }
                if (dp_abs(y[e].im) > dp_abs(vb)) {
                    vb = y[e].im;
                    pb = off + 1;
                }
            }
        }
        DpBest<T> c{vb, RS * tid + 2 * p + pb};
        dp_best_merge(b, c);
    }
    // candidates: r in [lo, lo+len) (outside: the complement)
    static DP_DEV void window(const cx<T> (&y)[32], int tid, int p, int lo, unsigned len, bool outside, DpBest<T>& b) {
        T kb = (T)-1, vb = (T)0;
        int pb = -1;
        const int rb = RS * tid + 2 * p;
#pragma unroll
        for (int n = 0; n < R1; ++n) {
            // CTA-uniform skip: block n covers r in [RS*512*n, RS*512*(n+1))
            const int blo = RS * 512 * n, bhi = RS * 512 * (n + 1);
            const bool hit = outside ? !(lo <= blo && (long long)bhi <= (long long)lo + (long long)len)
                                     : (lo < bhi && (long long)lo + (long long)len > (long long)blo);
            if (hit) {
#pragma unroll
                for (int i = 0; i < NB1; ++i) {
                    const int e = i * R1 + n;
                    const int r0 = rb + RS * (i * NT + 512 * n);
                    const int r1 = r0 + 1;
                    const bool in0 = ((unsigned)(r0 - lo) < len) != outside;
                    const bool in1 = ((unsigned)(r1 - lo) < len) != outside;
                    if (in0 && dp_abs(y[e].re) > kb) {
                        kb = dp_abs(y[e].re);
                        vb = y[e].re;
                        pb = r0;
                    }
                    if (in1 && dp_abs(y[e].im) > kb) {
                        kb = dp_abs(y[e].im);
                        vb = y[e].im;
                        pb = r1;
                    }
                }
            }
        }
        DpBest<T> c{vb, pb};
        dp_best_merge(b, c);
    }
};

// --------------------------------------------------------- point-wise helpers
// Real-FFT untangle of the pair (k, M-k):  Zk = C[k], Zm = C[M-k], w = exp(-2 pi i k/N).
// Returns 2*X[k] and 2*X[M-k] (the factor 2 is folded into the tables).
template <class T> DP_DEV void dp_untangle(cx<T> Zk, cx<T> Zm, cx<T> w, cx<T>& Xk, cx<T>& Xm) {
    const cx<T> E2 = cx<T>{Zk.re + Zm.re, Zk.im - Zm.im};  // Zk + conj(Zm)
    const cx<T> D = cx<T>{Zk.re - Zm.re, Zk.im + Zm.im};   // Zk - conj(Zm)
    const cx<T> G = cmuli(cmul(w, D));                      // i*w*D
    Xk = csub(E2, G);
    Xm = cconj(cadd(E2, G));
}
// Inverse: from F[k], F[M-k] to C'[k], C'[M-k].
template <class T> DP_DEV void dp_retangle(cx<T> Fk, cx<T> Fm, cx<T> w, cx<T>& Ck, cx<T>& Cm) {
    const cx<T> E = cx<T>{Fk.re + Fm.re, Fk.im - Fm.im};  // Fk + conj(Fm)
    const cx<T> D = cx<T>{Fk.re - Fm.re, Fk.im + Fm.im};  // Fk - conj(Fm)
    const cx<T> O = cmuli(cmulc(D, w));                    // i * D * conj(w)
    Ck = cadd(E, O);
    Cm = cconj(csub(E, O));
}

// ======================================================================== kernel
// Self-paired butterflies.  Thread 0 owns K12 = 0 (k' = KQ*r: r pairs with 16-r; r = 0
// holds DC/Nyquist, r = 8 is its own mirror) and K12 = KQ/2 (k' = KQ/2 + KQ*r: r pairs
// with 15-r).  Their 17 pairs are processed by lanes 0..16 of warp 0 through a small
// shared scratch, so warp 0 does not run a second unrolled copy of the point-wise code;
// bins and duplicate handling live in the host-built phi_self / wj_self tables.
template <class T> struct DpSelfLane {
    int ek, em;  // element slots (0..15 = A, 16..31 = B) of k' and of its mirror
    cx<T> w;     // exp(-2 pi i k' / N)
    cx<T> u;     // exp(-2 pi i k' / M)   (P = 2 only)
};
template <class T, int P> DP_DEV DpSelfLane<T> dp_self_lane(int lane) {
    DpSelfLane<T> sp;
    double s_, c_;
    int num;  // k' / (KQ/2)
    if (lane < 9) {
        sp.ek = lane;
        sp.em = (16 - lane) & 15;
        num = 2 * lane;
    } else {
        const int r = (lane - 9) & 7;
        sp.ek = 16 + r;
        sp.em = 16 + 15 - r;
        num = 1 + 2 * r;
    }
    // KQ/2 / N = 1/(64*P),  KQ/2 / M = 1/(32*P)
    sincospi(-2.0 * (double)num / (64.0 * P), &s_, &c_);
    sp.w = cx<T>{(T)c_, (T)s_};
    sincospi(-2.0 * (double)num / (32.0 * P), &s_, &c_);
    sp.u = cx<T>{(T)c_, (T)s_};
    return sp;
}

// P = 2 quad: Z_p at k' (k) and at k'' = M' - k' (m);  u = exp(-2 pi i k'/M), w1 = exp(-2 pi i k'/N).
// X[0..3] = 2*X at bins k', M-k', k'+M', M'-k'.
template <class T>
DP_DEV void dp_quad_x(cx<T> Z0k, cx<T> Z1k, cx<T> Z0m, cx<T> Z1m, cx<T> u, cx<T> w1, cx<T> (&X)[4]) {
    const cx<T> t = cmul(u, Z1k), q = cmulc(Z1m, u);
    dp_untangle(cadd(Z0k, t), cadd(Z0m, q), w1, X[0], X[1]);
    dp_untangle(csub(Z0k, t), csub(Z0m, q), cmulni(w1), X[2], X[3]);
}
template <class T>
DP_DEV void dp_quad_z(const cx<T> (&F)[4], cx<T> u, cx<T> w1, cx<T>& Z0k, cx<T>& Z1k, cx<T>& Z0m, cx<T>& Z1m) {
    cx<T> Ck, Cmp, Ckp, Cm;
    dp_retangle(F[0], F[1], w1, Ck, Cmp);
    dp_retangle(F[2], F[3], cmulni(w1), Ckp, Cm);
    Z0k = cadd(Ck, Ckp);
    Z1k = cmulc(csub(Ck, Ckp), u);
    Z0m = cadd(Cm, Cmp);
    const cx<T> d = csub(Cmp, Cm);  // (Cm - Cmp) * (-u)
    Z1m = cmul(d, u);
}

template <class T, int R1, int P, int IN> struct DpOfKernel {
    using G = DpGeom<R1>;
    static constexpr int NT = G::NT;
    static constexpr int NW = (NT + 31) / 32;
    static constexpr int NE = 32 * P;  // table entries per thread
    static constexpr int N = 2 * P * G::MS;
    static constexpr int RED_DOUBLES = (DP_MAX_TSLOTS + 1) * 32;
    static constexpr int BEST_ELEMS = DP_MAX_TSLOTS * 32;
    static constexpr int SP_ELEMS = 64;  // thread 0's butterflies: [0..31] sub-sequence 1 (or the only one), [32..63] sub-sequence 0
    static constexpr size_t SMEM_BYTES = sizeof(cx<T>) * (G::SMEM_ELEMS + DP_NLOW_MAX + SP_ELEMS) +
                                         2 * (sizeof(double) * RED_DOUBLES + sizeof(DpBest<T>) * BEST_ELEMS +
                                              sizeof(int) * DP_MAX_TSLOTS) + 64;

    struct Smem {
        cx<T>* buf;
        cx<T>* stash;
        cx<T>* sp;
        // double-buffered epilogue arrays (buffer = parity of the template iteration)
        double* red0;      // [2][DP_MAX_TSLOTS + 1][32] per-warp partial sums
        DpBest<T>* best0;  // [2][DP_MAX_TSLOTS][32] per-warp arg-max
        int* slot0;        // [2][DP_MAX_TSLOTS]
        DP_DEV double* red(int par) const { return red0 + par * RED_DOUBLES; }
        DP_DEV DpBest<T>* best(int par) const { return best0 + par * BEST_ELEMS; }
        DP_DEV int* slot_id(int par) const { return slot0 + par * DP_MAX_TSLOTS; }
    };
    static DP_DEV Smem carve(unsigned char* raw) {
        Smem s;
        s.buf = reinterpret_cast<cx<T>*>(raw);
        s.stash = s.buf + G::SMEM_ELEMS;
        s.sp = s.stash + DP_NLOW_MAX;
        s.red0 = reinterpret_cast<double*>(s.sp + SP_ELEMS);
        s.best0 = reinterpret_cast<DpBest<T>*>(s.red0 + 2 * RED_DOUBLES);
        s.slot0 = reinterpret_cast<int*>(s.best0 + 2 * BEST_ELEMS);
        return s;
    }

    // ---- windowed arg-max of one inverse sub-FFT: one pass over y per fit of template `it`
    // Per-warp winners go to best[q][warp]; `merge` folds them into the previous round
    // (P = 2: the two sub-sequences are scanned one after the other).
    static DP_DEV int scan_slots(const cx<T> (&y)[32], const DpChanDev<T>& ch, int it, int p, bool merge,
                                 DpBest<T>* best, int* slot_id) {
        const int tid = threadIdx.x;
        int nts = 0;
        for (int s = 0; s < ch.n_slots; ++s) {
            const DpSlot sl = ch.slots[s];
            if (sl.templ != it) continue;
            DpBest<T> b{(T)0, -1};
            if (sl.lo == 0 && sl.hi == N && !sl.outside)
                DpScan<T, R1, P>::full(y, tid, p, b);
            else
                DpScan<T, R1, P>::window(y, tid, p, sl.lo, (unsigned)(sl.hi - sl.lo), sl.outside != 0, b);
            b = dp_warp_best(b);
            if ((tid & 31) == 0) {
                if (merge) dp_best_merge(b, best[nts * 32 + (tid >> 5)]);
                best[nts * 32 + (tid >> 5)] = b;
            }
            if (tid == 0) slot_id[nts] = s;
            ++nts;
        }
        return nts;
    }

    // ---- low-frequency chi2 at each fit's (amp, delay), chi0, outputs.  Two barriers; the
    // small smem arrays are double buffered so no trailing barrier is needed.
    static DP_DEV void epilogue(const DpOfParams<T>& prm, const Smem& sm, const DpChanDev<T>& ch,
                                const DpTemplDev<T>& tp, int it, int nts, int par, int ev, T chi, double& chi0_keep) {
        const int tid = threadIdx.x;
        DpBest<T>* best = sm.best(par);
        double* red = sm.red(par);
        __syncthreads();  // per-warp winners + (first template) the lowchi2 stash are visible
        double part[DP_MAX_TSLOTS + 1];
#pragma unroll
        for (int q = 0; q < DP_MAX_TSLOTS; ++q) {
            part[q] = 0.0;
            const int nlow_q = q < nts ? ch.slots[sm.slot_id(par)[q]].nlow : 0;
            if (q < nts && tid < nlow_q) {
                DpBest<T> b = best[q * 32];
                for (int w = 1; w < NW; ++w) dp_best_merge(b, best[q * 32 + w]);
                const int d = b.idx - tp.pretrigger;
                for (int k = tid; k < nlow_q; k += NT) {
                    const int ph = (int)((((long long)k * (long long)d) % N + N) % N);  // exp(-2 pi i k d / N)
                    T sn, cs;
                    if constexpr (sizeof(T) == 8) {
                        double s_, c_;
                        sincospi(2.0 * (double)ph / (double)N, &s_, &c_);
                        sn = (T)s_;
                        cs = (T)c_;
                    } else {
                        float s_, c_;
                        sincospif(2.0f * (float)ph / (float)N, &s_, &c_);
                        sn = (T)s_;
                        cs = (T)c_;
                    }
                    const cx<T> mdl = cmul(cx<T>{cs, -sn}, dp_ldg(tp.s_low + k));
                    const cx<T> X = sm.stash[k];
                    const cx<T> R = cx<T>{dp_fma(-b.val, mdl.re, X.re), dp_fma(-b.val, mdl.im, X.im)};
                    part[q] += (double)(dp_ldg(ch.wj_low + k) * cnorm2(R));
                }
            }
        }
        part[DP_MAX_TSLOTS] = (it == 0) ? (double)chi : 0.0;
#pragma unroll
        for (int q = 0; q <= DP_MAX_TSLOTS; ++q) {
            if (q < nts || (q == DP_MAX_TSLOTS && it == 0)) {
                const double v = dp_warp_sum(part[q]);
                if ((tid & 31) == 0) red[q * 32 + (tid >> 5)] = v;
            }
        }
        __syncthreads();
        if (tid == 0) {
            double* o = prm.out + (long long)ev * prm.n_out + ch.out_base;
            if (it == 0) {
                double c0 = 0.0;
                for (int w = 0; w < NW; ++w) c0 += red[DP_MAX_TSLOTS * 32 + w];
                chi0_keep = c0;
                o[0] = c0;
            }
            const double chi0 = chi0_keep;
            for (int q = 0; q < nts; ++q) {
                double low = 0.0;
                for (int w = 0; w < NW; ++w) low += red[q * 32 + w];
                DpBest<T> b = best[q * 32];
                for (int w = 1; w < NW; ++w) dp_best_merge(b, best[q * 32 + w]);
                double* os = o + 1 + sm.slot_id(par)[q] * DP_SLOT_NOUT;
                const double amp = (double)b.val;
                os[0] = amp;
                os[1] = (double)b.idx;
                os[2] = chi0 - amp * amp * tp.norm;
                os[3] = low;
                os[4] = 1.0 / sqrt(amp * amp * tp.tsum);
            }
        }
    }

    static DP_DEV void prefetch_next(const DpOfParams<T>& prm, int row) {
        constexpr size_t ESZ = sizeof(typename DpRaw<IN>::scalar);
        const int nrow = row + gridDim.x;
        if (nrow < prm.n_rows) {
            const unsigned char* nx = reinterpret_cast<const unsigned char*>(prm.traces) + (size_t)dp_first_sample(prm, nrow) * ESZ;
            constexpr int nlines = (int)((size_t)N * ESZ / 128);
            for (int l = threadIdx.x; l < nlines; l += NT) dp_prefetch_l2(nx + (size_t)l * 128);
        }
    }

    static DP_DEV void run(const DpOfParams<T>& prm, unsigned char* smem_raw) {
        if constexpr (P == 1)
            run_p1(prm, smem_raw);
        else
            run_p2(prm, smem_raw);
    }
    static DP_DEV void run_p1(const DpOfParams<T>& prm, unsigned char* smem_raw);
    static DP_DEV void run_p2(const DpOfParams<T>& prm, unsigned char* smem_raw);
};

// ------------------------------------------------------------------------- P = 1
template <class T, int R1, int P, int IN>
DP_DEV void DpOfKernel<T, R1, P, IN>::run_p1(const DpOfParams<T>& prm, unsigned char* smem_raw) {
    const Smem sm = carve(smem_raw);
    const int tid = threadIdx.x;
    int K12, bA, bB;
    G::map(tid, K12, bA, bB);
    const cx<T> wn = dp_ldg(prm.twn + tid);  // exp(-2 pi i K12 / N)
    cx<T>* scr = prm.scratch + (long long)blockIdx.x * prm.scratch_per_cta;
    constexpr size_t ESZ = sizeof(typename DpRaw<IN>::scalar);
    const DpSelfLane<T> sp = dp_self_lane<T, 1>(tid & 31);
    int par = 0;
    double chi0_keep = 0.0;

    for (int row = blockIdx.x; row < prm.n_rows; row += gridDim.x) {
        const int chan = row % prm.n_chan;
        const int ev = row / prm.n_chan;
        const DpChanDev<T>& ch = prm.chans[chan];
        const void* xrow = reinterpret_cast<const unsigned char*>(prm.traces) + (size_t)dp_first_sample(prm, row) * ESZ;
        double x0, xsc;
        dp_row_conversion<T, IN>(prm, chan, xrow, x0, xsc);

        cx<T> za[16], zb[16];
        T chi = (T)0;
        dp_fwd_subfft<T, R1, 1, IN>(xrow, 0, x0, xsc, sm.buf, prm.tw1, prm.tw2, bA, bB, za, zb);

        // ---- self-paired butterflies of thread 0, cooperatively in warp 0 ------------------
        cx<T> sXk = cx<T>{(T)0, (T)0}, sXm = sXk;  // lanes 0..16 of warp 0: 2*X of their pair
        if (tid < 32) {
            if (tid == 0) {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    sm.sp[r] = za[r];
                    sm.sp[16 + r] = zb[r];
                }
            }
            __syncwarp();
            if (tid < 17) {
                dp_untangle(sm.sp[sp.ek], sm.sp[sp.em], sp.w, sXk, sXm);
                chi = dp_fma(dp_ldg(ch.wj_self + 2 * tid), cnorm2(sXk), chi);
                chi = dp_fma(dp_ldg(ch.wj_self + 2 * tid + 1), cnorm2(sXm), chi);
                if (tid == 0) sm.stash[0] = sXk;                          // X at DC, first lowchi2 bin
                if (tid == 9 && G::KQ / 2 < prm.nlow) sm.stash[G::KQ / 2] = sXk;  // k = KQ/2
            }
            __syncwarp();
        }
        // ---- untangle: (za, zb) <- 2*X   (thread 0 computes throw-away values) -------------
        {
            T chin = (T)0;
            // A[r] (k = K12 + KQ r) pairs with B[15-r] (k' = M - k);  w = wn * W_32^r
#define DP_XP(r)                                                                                         \
    {                                                                                                    \
        cx<T> Xk, Xm;                                                                                    \
        dp_untangle(za[r], zb[15 - r], cmul(wn, dp_w64<T, 2 * r, -1>()), Xk, Xm);                        \
        chin = dp_fma(dp_ldg(ch.wj + r * NT + tid), cnorm2(Xk), chin);                                   \
        chin = dp_fma(dp_ldg(ch.wj + (16 + 15 - r) * NT + tid), cnorm2(Xm), chin);                       \
        za[r] = Xk;                                                                                      \
        zb[15 - r] = Xm;                                                                                 \
    }
            DP_XP(0) DP_XP(1) DP_XP(2) DP_XP(3) DP_XP(4) DP_XP(5) DP_XP(6) DP_XP(7)
            DP_XP(8) DP_XP(9) DP_XP(10) DP_XP(11) DP_XP(12) DP_XP(13) DP_XP(14) DP_XP(15)
#undef DP_XP
            if (tid != 0) {
                chi += chin;
                // low-frequency bins for lowchi2: k = K12 lives in A[0], k = KQ - K12 in B[0]
                if (K12 < prm.nlow) sm.stash[K12] = za[0];
                if (G::KQ - K12 < prm.nlow) sm.stash[G::KQ - K12] = zb[0];
            }
        }
        prefetch_next(prm, row);

        // multi-template: X must survive the in-place inverse of the previous template
        const bool spill_x = ch.n_templ > 1;
        if (spill_x) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                scr[r * NT + tid] = za[r];
                scr[(16 + r) * NT + tid] = zb[r];
            }
        }

        for (int it = 0; it < ch.n_templ; ++it) {
            const DpTemplDev<T>& tp = ch.templ[it];
            if (it > 0) {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    za[r] = scr[r * NT + tid];
                    zb[r] = scr[(16 + r) * NT + tid];
                }
            }
            // ---- filter + inverse untangle: (za, zb) <- Z' -----------------------------
            if (tid < 17) {
                const cx<T> Fk = cmul(dp_ldg(tp.phi_self + 2 * tid), sXk);
                const cx<T> Fm = cmul(dp_ldg(tp.phi_self + 2 * tid + 1), sXm);
                cx<T> Ck, Cm;
                dp_retangle(Fk, Fm, sp.w, Ck, Cm);
                sm.sp[sp.ek] = Ck;
                if (sp.ek != sp.em) sm.sp[sp.em] = Cm;
            }
#define DP_FP(r)                                                                                         \
    {                                                                                                    \
        const cx<T> Fk = cmul(dp_ldg(tp.phi + r * NT + tid), za[r]);                                     \
        const cx<T> Fm = cmul(dp_ldg(tp.phi + (16 + 15 - r) * NT + tid), zb[15 - r]);                    \
        dp_retangle(Fk, Fm, cmul(wn, dp_w64<T, 2 * r, -1>()), za[r], zb[15 - r]);                        \
    }
            DP_FP(0) DP_FP(1) DP_FP(2) DP_FP(3) DP_FP(4) DP_FP(5) DP_FP(6) DP_FP(7)
            DP_FP(8) DP_FP(9) DP_FP(10) DP_FP(11) DP_FP(12) DP_FP(13) DP_FP(14) DP_FP(15)
#undef DP_FP
            if (tid < 32) {
                __syncwarp();
                if (tid == 0) {
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        za[r] = sm.sp[r];
                        zb[r] = sm.sp[16 + r];
                    }
                }
                __syncwarp();
            }
            // ---- inverse, windowed arg-max, outputs -----------------------------------------
            cx<T> y[32];
            dp_inv_subfft<T, R1>(sm.buf, prm.tw1, prm.tw2, bA, bB, za, zb, y);
            const int nts = scan_slots(y, ch, it, 0, false, sm.best(par), sm.slot_id(par));
            epilogue(prm, sm, ch, tp, it, nts, par, ev, chi, chi0_keep);
            par ^= 1;
        }
    }
}

// ------------------------------------------------------------------------- P = 2
// N = 4*M': the packed complex sequence c[m] is split once more, z_p[m'] = c[2m'+p].  Z_0
// is parked in a thread-private global scratch column (L2 resident) while Z_1 is
// transformed; the outer radix-2, the real-FFT untangle, the filter and their inverses
// are one thread-local "quad" step.
template <class T, int R1, int P, int IN>
DP_DEV void DpOfKernel<T, R1, P, IN>::run_p2(const DpOfParams<T>& prm, unsigned char* smem_raw) {
    const Smem sm = carve(smem_raw);
    const int tid = threadIdx.x;
    int K12, bA, bB;
    G::map(tid, K12, bA, bB);
    const cx<T> wn = dp_ldg(prm.twn + tid);  // exp(-2 pi i K12 / N)
    const cx<T> wp = dp_ldg(prm.twp + tid);  // exp(-2 pi i K12 / M)
    cx<T>* scr0 = prm.scratch + (long long)blockIdx.x * prm.scratch_per_cta;  // Z_0   [32][NT]
    cx<T>* scr1 = scr0 + 32 * NT;                                             // Z_1   [32][NT] (multi-template)
    cx<T>* scrz = scr1 + 32 * NT;                                             // Z'_0  [32][NT]
    constexpr size_t ESZ = sizeof(typename DpRaw<IN>::scalar);
    const DpSelfLane<T> sp = dp_self_lane<T, 2>(tid & 31);
    int par = 0;
    double chi0_keep = 0.0;

    for (int row = blockIdx.x; row < prm.n_rows; row += gridDim.x) {
        const int chan = row % prm.n_chan;
        const int ev = row / prm.n_chan;
        const DpChanDev<T>& ch = prm.chans[chan];
        const void* xrow = reinterpret_cast<const unsigned char*>(prm.traces) + (size_t)dp_first_sample(prm, row) * ESZ;
        double x0, xsc;
        dp_row_conversion<T, IN>(prm, chan, xrow, x0, xsc);

        cx<T> za[16], zb[16];
        T chi = (T)0;
        // ---- forward: sub-sequence 0 -> scratch, sub-sequence 1 -> registers --------------
        // (a real loop: ONE copy of the sub-FFT code; instruction fetch is what limits
        //  the 8-warp fp64 CTA, see profiles/)
#pragma unroll 1
        for (int p = 0; p < 2; ++p) {
            dp_fwd_subfft<T, R1, 2, IN>(xrow, p, x0, xsc, sm.buf, prm.tw1, prm.tw2, bA, bB, za, zb);
            if (p == 0) {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    scr0[r * NT + tid] = za[r];
                    scr0[(16 + r) * NT + tid] = zb[r];
                }
                if (tid == 0) {
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        sm.sp[32 + r] = za[r];
                        sm.sp[32 + 16 + r] = zb[r];
                    }
                }
                __syncthreads();  // pass-3 reads of buf are done before the next pass-1 stores
            }
        }
        prefetch_next(prm, row);
        const bool multi = ch.n_templ > 1;
        if (multi) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                scr1[r * NT + tid] = za[r];
                scr1[(16 + r) * NT + tid] = zb[r];
            }
        }
        // self-paired butterflies: Z of thread 0 into shared scratch, X of each lane's quad in registers
        cx<T> sX[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) sX[j] = cx<T>{(T)0, (T)0};
        if (tid < 32) {
            if (tid == 0) {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    sm.sp[r] = za[r];
                    sm.sp[16 + r] = zb[r];
                }
            }
            __syncwarp();
            if (tid < 17) {
                dp_quad_x(sm.sp[32 + sp.ek], sm.sp[sp.ek], sm.sp[32 + sp.em], sm.sp[sp.em], sp.u, sp.w, sX);
#pragma unroll
                for (int j = 0; j < 4; ++j) chi = dp_fma(dp_ldg(ch.wj_self + 4 * tid + j), cnorm2(sX[j]), chi);
                if (tid == 0) sm.stash[0] = sX[0];
                if (tid == 9 && G::KQ / 2 < prm.nlow) sm.stash[G::KQ / 2] = sX[0];
            }
            __syncwarp();
        }

        for (int it = 0; it < ch.n_templ; ++it) {
            const DpTemplDev<T>& tp = ch.templ[it];
            if (it > 0) {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    za[r] = scr1[r * NT + tid];
                    zb[r] = scr1[(16 + r) * NT + tid];
                }
            }
            // ---- quads: outer radix-2 + untangle + filter + retangle + inverse radix-2 ------
            if (tid < 17) {
                cx<T> F[4], Z0k, Z1k, Z0m, Z1m;
#pragma unroll
                for (int j = 0; j < 4; ++j) F[j] = cmul(dp_ldg(tp.phi_self + 4 * tid + j), sX[j]);
                dp_quad_z(F, sp.u, sp.w, Z0k, Z1k, Z0m, Z1m);
                sm.sp[32 + sp.ek] = Z0k;
                sm.sp[sp.ek] = Z1k;
                if (sp.ek != sp.em) {
                    sm.sp[32 + sp.em] = Z0m;
                    sm.sp[sp.em] = Z1m;
                }
            }
            {
                T chin = (T)0;
                const bool first = it == 0;
#define DP_QD(r)                                                                                         \
    {                                                                                                    \
        const cx<T> u = cmul(wp, dp_w64<T, 2 * r, -1>());                                                \
        const cx<T> w1 = cmul(wn, dp_w64<T, r, -1>());                                                   \
        cx<T> X[4], F[4], Z0k, Z0m;                                                                      \
        dp_quad_x(scr0[r * NT + tid], za[r], scr0[(16 + 15 - r) * NT + tid], zb[15 - r], u, w1, X);      \
        if (first) {                                                                                     \
            chin = dp_fma(dp_ldg(ch.wj + (4 * r + 0) * NT + tid), cnorm2(X[0]), chin);                   \
            chin = dp_fma(dp_ldg(ch.wj + (4 * r + 1) * NT + tid), cnorm2(X[1]), chin);                   \
            chin = dp_fma(dp_ldg(ch.wj + (4 * r + 2) * NT + tid), cnorm2(X[2]), chin);                   \
            chin = dp_fma(dp_ldg(ch.wj + (4 * r + 3) * NT + tid), cnorm2(X[3]), chin);                   \
            if (r == 0 && tid != 0 && K12 < prm.nlow) sm.stash[K12] = X[0];                              \
            if (r == 15 && tid != 0 && G::KQ - K12 < prm.nlow) sm.stash[G::KQ - K12] = X[3];             \
        }                                                                                                \
        F[0] = cmul(dp_ldg(tp.phi + (4 * r + 0) * NT + tid), X[0]);                                      \
        F[1] = cmul(dp_ldg(tp.phi + (4 * r + 1) * NT + tid), X[1]);                                      \
        F[2] = cmul(dp_ldg(tp.phi + (4 * r + 2) * NT + tid), X[2]);                                      \
        F[3] = cmul(dp_ldg(tp.phi + (4 * r + 3) * NT + tid), X[3]);                                      \
        dp_quad_z(F, u, w1, Z0k, za[r], Z0m, zb[15 - r]);                                                \
        scrz[r * NT + tid] = Z0k;                                                                        \
        scrz[(16 + 15 - r) * NT + tid] = Z0m;                                                            \
    }
                DP_QD(0) DP_QD(1) DP_QD(2) DP_QD(3) DP_QD(4) DP_QD(5) DP_QD(6) DP_QD(7)
                DP_QD(8) DP_QD(9) DP_QD(10) DP_QD(11) DP_QD(12) DP_QD(13) DP_QD(14) DP_QD(15)
#undef DP_QD
                if (first && tid != 0) chi += chin;
            }
            if (tid < 32) {
                __syncwarp();
                if (tid == 0) {
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        za[r] = sm.sp[r];
                        zb[r] = sm.sp[16 + r];
                        scrz[r * NT] = sm.sp[32 + r];
                        scrz[(16 + r) * NT] = sm.sp[32 + 16 + r];
                    }
                }
                __syncwarp();
            }
            // ---- inverse of sub-sequence 1, then 0; arg-max over both (one code copy) ---------
            int nts = 0;
#pragma unroll 1
            for (int p = 1; p >= 0; --p) {
                if (p == 0) {
                    __syncthreads();  // pass-1' reads of buf are done before the next pass-3' stores
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        za[r] = scrz[r * NT + tid];
                        zb[r] = scrz[(16 + r) * NT + tid];
                    }
                }
                cx<T> y[32];
                dp_inv_subfft<T, R1>(sm.buf, prm.tw1, prm.tw2, bA, bB, za, zb, y);
                nts = scan_slots(y, ch, it, p, p == 0, sm.best(par), sm.slot_id(par));
            }
            epilogue(prm, sm, ch, tp, it, nts, par, ev, chi, chi0_keep);
            par ^= 1;
        }
    }
}

#ifndef DP_HOST_EMU
template <class T, int R1, int P, int IN>
__global__ void __launch_bounds__(DpGeom<R1>::NT, 1) dp_of_kernel(const DpOfParams<T> prm) {
    extern __shared__ __align__(16) unsigned char dp_smem_raw[];
    DpOfKernel<T, R1, P, IN>::run(prm, dp_smem_raw);
}
#endif
