// Fused OF1x1 kernel for trace lengths that are NOT 2^k: the reference's own example configuration processes
// 20 ms / 10 ms at 1.25 MHz = 25000 / 12500 samples (reference examples/processing/process_example.yaml:93-94; the
// length must equal the filter's, detprocess/process/processing_data.py:312, 351).  Same maths, outputs and tables
// conventions as dp_of2_kernel.cuh, on a mixed-radix transform:
//   * the real FFT of N samples is the complex FFT of M = N / 2 points, M = r_1 r_2 ... r_p with r_j in {2, 3, 4, 5}
//     (25000 -> 12500 = 5^5 * 4, 12500 -> 6250 = 5^5 * 2);
//   * one CTA per event, the M points in shared memory (float64: 16 M bytes <= 208 KB), in place: decimation-in-
//     frequency passes forward (natural order in, digit-reversed out), the point-wise stage on the digit-reversed
//     positions (host-built pair tables), the mirrored decimation-in-time passes back (natural order out);
//   * every template re-runs the forward transform (no second buffer fits next to a 200 KB event): the lengths this
//     kernel serves are a convenience of the reference's examples, not the benchmark shapes;
//   * samples are loaded one by one, so any input layout of dp_of1x1_batch_ex works, windows of streams included.
// Replaces the same reference calls as dp_of2_kernel.cuh (processing_data.py:763-772, algorithms.py:331-341, 410-421,
// 533-558).
#pragma once
#include "dp_of_kernel.cuh"

#define DPG_MAX_PASSES 16
#define DPG_NT 512

template <class T> struct DpGenTemplDev {
    const cx<T>* phi_k;   // [n_pairs] filter at bin k of the pair
    const cx<T>* phi_m;   // [n_pairs] filter at bin M - k (the Nyquist bin M for the DC pair)
    const cx<T>* s_low;   // [nlow] scaled template spectrum, natural order
    double norm, tsum;
    int pretrigger, pad_;
};
template <class T> struct DpGenChanDev {
    const T* wj_k;        // [n_pairs] chi0 weights
    const T* wj_m;
    const T* wj_low;      // [nlow]
    double adc_gain, adc_offset;
    int n_templ, n_slots, out_base, pad_;
    DpGenTemplDev<T> templ[DP_MAX_TEMPLATES];
    DpSlot slots[DP_MAX_SLOTS];
};
template <class T> struct DpGenParams {
    const void* traces;
    int in_dtype;            // DP_IN_F64 / F32 / I16 (element-wise loads: no alignment requirement)
    long long event_stride, chan_stride;
    const long long* chan_offset;
    const long long* row_start;
    long long stream_len;
    int n_rows, n_chan;
    const DpGenChanDev<T>* chans;
    int M;                   // complex points
    int n_pass;
    int radix[DPG_MAX_PASSES];
    const cx<T>* tw;         // [M] exp(-2 pi i j / M)
    const int* pos_k;        // [n_pairs] position (digit-reversed order) of bin k of pair p: k = p, p = 0 .. M/2
    const int* pos_m;        // [n_pairs] position of bin M - k (== pos_k for the two self pairs)
    const cx<T>* wn;         // [n_pairs] exp(-2 pi i k / N)
    int n_pairs;             // M / 2 + 1
    double* out;
    int n_out, nlow;
    double scale;
    int subtract_first;
    int neighbours;          // also report the amplitude one sample before / after each fit's best delay
};

// ---- small DFTs.  SIGN = -1 forward, +1 inverse.
template <int SIGN, class T> DP_DEV void dpg_dft2(cx<T>* x) {
    const cx<T> a = x[0], b = x[1];
    x[0] = cadd(a, b);
    x[1] = csub(a, b);
}
template <int SIGN, class T> DP_DEV void dpg_dft3(cx<T>* x) {
    const T h = (T)0.86602540378443864676;  // sin(2 pi / 3)
    const cx<T> t = cadd(x[1], x[2]);
    const cx<T> m = cx<T>{x[0].re - (T)0.5 * t.re, x[0].im - (T)0.5 * t.im};
    const cx<T> d = cx<T>{h * (x[1].re - x[2].re), h * (x[1].im - x[2].im)};
    x[0] = cadd(x[0], t);
    // forward: y1 = m - i d, y2 = m + i d
    const cx<T> id = cx<T>{-d.im, d.re};
    if (SIGN < 0) {
        x[1] = csub(m, id);
        x[2] = cadd(m, id);
    } else {
        x[1] = cadd(m, id);
        x[2] = csub(m, id);
    }
}
template <int SIGN, class T> DP_DEV void dpg_dft4(cx<T>* x) {
    const cx<T> a = cadd(x[0], x[2]), b = csub(x[0], x[2]), c = cadd(x[1], x[3]), d = csub(x[1], x[3]);
    const cx<T> jd = (SIGN < 0) ? cmulni(d) : cmuli(d);  // (-i) d forward, (+i) d inverse
    x[0] = cadd(a, c);
    x[2] = csub(a, c);
    x[1] = cadd(b, jd);
    x[3] = csub(b, jd);
}
template <int SIGN, class T> DP_DEV void dpg_dft5(cx<T>* x) {
    const T c1 = (T)0.30901699437494742410, c2 = (T)-0.80901699437494742410;  // cos(2 pi/5), cos(4 pi/5)
    const T s1 = (T)0.95105651629515357212, s2 = (T)0.58778525229247312917;   // sin(2 pi/5), sin(4 pi/5)
    const cx<T> t1 = cadd(x[1], x[4]), t2 = cadd(x[2], x[3]), t3 = csub(x[1], x[4]), t4 = csub(x[2], x[3]);
    const cx<T> m1 = cx<T>{x[0].re + c1 * t1.re + c2 * t2.re, x[0].im + c1 * t1.im + c2 * t2.im};
    const cx<T> m2 = cx<T>{x[0].re + c2 * t1.re + c1 * t2.re, x[0].im + c2 * t1.im + c1 * t2.im};
    const cx<T> u1 = cx<T>{s1 * t3.re + s2 * t4.re, s1 * t3.im + s2 * t4.im};
    const cx<T> u2 = cx<T>{s2 * t3.re - s1 * t4.re, s2 * t3.im - s1 * t4.im};
    const cx<T> iu1 = cx<T>{-u1.im, u1.re}, iu2 = cx<T>{-u2.im, u2.re};
    x[0] = cadd(x[0], cadd(t1, t2));
    if (SIGN < 0) {  // y1 = m1 - i u1, y4 = m1 + i u1, y2 = m2 - i u2, y3 = m2 + i u2
        x[1] = csub(m1, iu1);
        x[4] = cadd(m1, iu1);
        x[2] = csub(m2, iu2);
        x[3] = cadd(m2, iu2);
    } else {
        x[1] = cadd(m1, iu1);
        x[4] = csub(m1, iu1);
        x[2] = cadd(m2, iu2);
        x[3] = csub(m2, iu2);
    }
}
template <int R, int SIGN, class T> DP_DEV void dpg_dft(cx<T>* x) {
    if constexpr (R == 2) dpg_dft2<SIGN>(x);
    if constexpr (R == 3) dpg_dft3<SIGN>(x);
    if constexpr (R == 4) dpg_dft4<SIGN>(x);
    if constexpr (R == 5) dpg_dft5<SIGN>(x);
}

// one in-place pass over the M points: blocks of `nblk` points, radix R, sub-block L = nblk / R.
// forward (decimation in frequency): v = DFT_R(u) then v_k *= W_nblk^(i k);  inverse (decimation in time): u_k *=
// conj(W_nblk^(i k)) then v = IDFT_R(u).  W_nblk^j = tw[j * (M / nblk)].
template <int R, bool FWD, class T> DP_DEV void dpg_pass(cx<T>* buf, const cx<T>* DP_RESTRICT tw, int M, int nblk) {
    const int L = nblk / R, step = M / nblk, nbf = M / R;
    for (int g = threadIdx.x; g < nbf; g += DPG_NT) {
        const int b = g / L, i = g - b * L;
        cx<T>* p = buf + b * nblk + i;
        cx<T> x[R];
#pragma unroll
        for (int q = 0; q < R; ++q) x[q] = p[q * L];
        cx<T> w1 = dp_ldg(tw + i * step);
        if constexpr (!FWD) {
            w1.im = -w1.im;
            cx<T> w = w1;
#pragma unroll
            for (int k = 1; k < R; ++k) {
                x[k] = cmul(x[k], w);
                if (k + 1 < R) w = cmul(w, w1);
            }
        }
        dpg_dft<R, FWD ? -1 : +1, T>(x);
        if constexpr (FWD) {
            cx<T> w = w1;
#pragma unroll
            for (int k = 1; k < R; ++k) {
                x[k] = cmul(x[k], w);
                if (k + 1 < R) w = cmul(w, w1);
            }
        }
#pragma unroll
        for (int q = 0; q < R; ++q) p[q * L] = x[q];
    }
}
template <bool FWD, class T> DP_DEV void dpg_pass_any(int r, cx<T>* buf, const cx<T>* DP_RESTRICT tw, int M, int nblk) {
    switch (r) {
        case 2: dpg_pass<2, FWD, T>(buf, tw, M, nblk); break;
        case 3: dpg_pass<3, FWD, T>(buf, tw, M, nblk); break;
        case 4: dpg_pass<4, FWD, T>(buf, tw, M, nblk); break;
        default: dpg_pass<5, FWD, T>(buf, tw, M, nblk); break;
    }
}

DP_DEV double dpg_sample(const void* base, int in_dtype, long long i) {
    if (in_dtype == 0) return __ldg(reinterpret_cast<const double*>(base) + i);
    if (in_dtype == 1) return (double)__ldg(reinterpret_cast<const float*>(base) + i);
    return (double)__ldg(reinterpret_cast<const short*>(base) + i);
}

template <class T> struct DpGenKernel {
    static constexpr int NT = DPG_NT, NW = NT / 32;
    static constexpr int RED_DOUBLES = (DP_MAX_TSLOTS + 1) * 32;
    static constexpr int BEST_ELEMS = DP_MAX_TSLOTS * 32;
    static DP_HD size_t smem_bytes(int M) {
        return sizeof(cx<T>) * (size_t)M + sizeof(cx<T>) * DP_NLOW_MAX + sizeof(double) * RED_DOUBLES + sizeof(DpBest<T>) * BEST_ELEMS + 64;
    }

    static DP_DEV void run(const DpGenParams<T>& prm, unsigned char* smem_raw) {
        cx<T>* buf = reinterpret_cast<cx<T>*>(smem_raw);
        cx<T>* stash = buf + prm.M;
        double* red = reinterpret_cast<double*>(stash + DP_NLOW_MAX);
        DpBest<T>* best = reinterpret_cast<DpBest<T>*>(red + RED_DOUBLES);
        const int tid = threadIdx.x, M = prm.M, N = 2 * M;
        const size_t esz = prm.in_dtype == 0 ? 8 : (prm.in_dtype == 1 ? 4 : 2);
        (void)esz;
        for (int row = blockIdx.x; row < prm.n_rows; row += gridDim.x) {
            const int chan = row % prm.n_chan, ev = row / prm.n_chan;
            const DpGenChanDev<T>& ch = prm.chans[chan];
            long long base = (long long)ev * prm.event_stride;
            if (prm.row_start != nullptr) {
                base = prm.row_start[ev];
                if (base < 0 || base + N > prm.stream_len) {  // CTA-uniform: the window leaves the stream
                    const int nb = 1 + (DP_SLOT_NOUT + (prm.neighbours ? 2 : 0)) * ch.n_slots;
                    for (int o = tid; o < nb; o += NT) prm.out[(long long)ev * prm.n_out + ch.out_base + o] = -999999.0;
                    continue;
                }
            }
            const long long first = base + (prm.chan_offset != nullptr ? prm.chan_offset[chan] : (long long)chan * prm.chan_stride);
            // (raw - x0) * sc: fp32 mode removes the first sample (AC coupling) / the ADC offset in float64 first
            double x0 = prm.subtract_first ? dpg_sample(prm.traces, prm.in_dtype, first) : 0.0, sc = prm.scale;
            double gain = 1.0, offs = 0.0;
            if (prm.in_dtype == 2) {
                gain = ch.adc_gain;
                offs = ch.adc_offset;
            }
            double chi0 = 0.0;
            for (int it = 0; it < ch.n_templ; ++it) {
                const DpGenTemplDev<T>& tp = ch.templ[it];
                // ---- load: c[n] = x[2n] + i x[2n+1]
                for (int n = tid; n < M; n += NT) {
                    double a = dpg_sample(prm.traces, prm.in_dtype, first + 2 * n), b = dpg_sample(prm.traces, prm.in_dtype, first + 2 * n + 1);
                    if (prm.in_dtype == 2) {
                        if (prm.subtract_first) {       // (adc - adc0) * gain: the offset cancels
                            a = (a - x0) * gain;
                            b = (b - x0) * gain;
                        } else {
                            a = dp_fma(a, gain, offs);
                            b = dp_fma(b, gain, offs);
                        }
                    } else {
                        a -= x0;
                        b -= x0;
                    }
                    buf[n] = cx<T>{(T)(a * sc), (T)(b * sc)};
                }
                __syncthreads();
                // ---- forward passes
                int nblk = M;
                for (int j = 0; j < prm.n_pass; ++j) {
                    dpg_pass_any<true, T>(prm.radix[j], buf, prm.tw, M, nblk);
                    nblk /= prm.radix[j];
                    __syncthreads();
                }
                // ---- point-wise: untangle the (k, M - k) pairs, chi0 (first template), filter, retangle
                T chi = (T)0;
                for (int p = tid; p < prm.n_pairs; p += NT) {
                    const int pk = prm.pos_k[p], pm = prm.pos_m[p];
                    const cx<T> w = dp_ldg(prm.wn + p);
                    cx<T> Xk, Xm;
                    dp_untangle(buf[pk], buf[pm], w, Xk, Xm);
                    if (it == 0) {
                        chi = dp_fma(dp_ldg(ch.wj_k + p), cnorm2(Xk), chi);
                        chi = dp_fma(dp_ldg(ch.wj_m + p), cnorm2(Xm), chi);
                        if (p < prm.nlow) stash[p] = Xk;     // bin k = p
                    }
                    const cx<T> Fk = cmul(dp_ldg(tp.phi_k + p), Xk), Fm = cmul(dp_ldg(tp.phi_m + p), Xm);
                    cx<T> Ck, Cm;
                    dp_retangle(Fk, Fm, w, Ck, Cm);
                    buf[pk] = Ck;
                    if (pm != pk) buf[pm] = Cm;
                }
                __syncthreads();
                // ---- inverse passes (mirror order)
                for (int j = prm.n_pass - 1; j >= 0; --j) {
                    nblk *= prm.radix[j];
                    dpg_pass_any<false, T>(prm.radix[j], buf, prm.tw, M, nblk);
                    __syncthreads();
                }
                // ---- windowed arg-max of |amplitude| per fit of this template (first maximum, numpy argmin on chi2)
                int slot_of[DP_MAX_TSLOTS], nts = 0;
#pragma unroll
                for (int q = 0; q < DP_MAX_TSLOTS; ++q) slot_of[q] = -1;
                for (int s = 0; s < ch.n_slots; ++s)
                    if (ch.slots[s].templ == it) {
#pragma unroll
                        for (int q = 0; q < DP_MAX_TSLOTS; ++q)
                            if (q == nts) slot_of[q] = s;
                        ++nts;
                    }
#pragma unroll
                for (int q = 0; q < DP_MAX_TSLOTS; ++q) {
                    if (q >= nts) continue;
                    const DpSlot sl = ch.slots[slot_of[q]];
                    DpBest<T> b{(T)0, -1};
                    for (int n = tid; n < M; n += NT) {
                        const cx<T> v = buf[n];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int r = 2 * n + h;
                            const T a = h ? v.im : v.re;
                            const bool in = ((r >= sl.lo && r < sl.hi) != (sl.outside != 0));
                            if (in && (b.idx < 0 || dp_abs(a) > dp_abs(b.val))) b = DpBest<T>{a, r};   // ascending r per thread
                        }
                    }
                    b = dp_warp_best(b);
                    if ((tid & 31) == 0) best[q * 32 + (tid >> 5)] = b;
                }
                __syncthreads();
                // ---- lowchi2 at each fit's (amp, delay); chi0
                double part[DP_MAX_TSLOTS + 1];
#pragma unroll
                for (int q = 0; q < DP_MAX_TSLOTS; ++q) {
                    part[q] = 0.0;
                    const int nlow_q = q < nts ? ch.slots[slot_of[q]].nlow : 0;
                    if (q < nts && tid < nlow_q) {
                        DpBest<T> b = best[q * 32];
                        for (int w = 1; w < NW; ++w) dp_best_merge(b, best[q * 32 + w]);
                        const int d = b.idx - tp.pretrigger;
                        for (int k = tid; k < nlow_q; k += NT) {
                            const int ph = (int)((((long long)k * (long long)d) % N + N) % N);
                            double s_, c_;
                            sincospi(2.0 * (double)ph / (double)N, &s_, &c_);
                            const cx<T> mdl = cmul(cx<T>{(T)c_, (T)(-s_)}, dp_ldg(tp.s_low + k));
                            const cx<T> X = stash[k];
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
                        chi0 = c0;
                        o[0] = c0;
                    }
                    for (int q = 0; q < nts; ++q) {
                        double low = 0.0;
                        for (int w = 0; w < NW; ++w) low += red[q * 32 + w];
                        DpBest<T> b = best[q * 32];
                        for (int w = 1; w < NW; ++w) dp_best_merge(b, best[q * 32 + w]);
                        double* os = o + 1 + slot_of[q] * DP_SLOT_NOUT;
                        const double amp = (double)b.val;
                        os[0] = amp;
                        os[1] = (double)b.idx;
                        os[2] = chi0 - amp * amp * tp.norm;
                        os[3] = low;
                        os[4] = 1.0 / sqrt(amp * amp * tp.tsum);
                        if (prm.neighbours) {   // the whole amplitude series is still in shared memory
                            double* on = o + 1 + DP_SLOT_NOUT * ch.n_slots + 2 * slot_of[q];
                            const int r0 = b.idx - 1, r1 = b.idx + 1;
                            on[0] = (b.idx < 0 || r0 < 0) ? (double)NAN : (double)((r0 & 1) ? buf[r0 >> 1].im : buf[r0 >> 1].re);
                            on[1] = (b.idx < 0 || r1 >= N) ? (double)NAN : (double)((r1 & 1) ? buf[r1 >> 1].im : buf[r1 >> 1].re);
                        }
                    }
                }
                __syncthreads();  // red / best / buf are reused by the next template / event
            }
        }
    }
};

#ifndef DP_HOST_EMU
template <class T> __global__ void __launch_bounds__(DPG_NT, 1) dp_ofg_kernel(const DpGenParams<T> prm) {
    extern __shared__ __align__(16) unsigned char dpg_smem_raw[];
    DpGenKernel<T>::run(prm, dpg_smem_raw);
}
#endif
