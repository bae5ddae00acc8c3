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
#define DP_NLOW_MAX 512
#define DP_SLOT_NOUT 5  // amp, ind, chi2, lowchi2, timeres

struct DpSlot {
    int templ;    // template index within the channel
    int lo, hi;   // candidate rolled delay indices [lo, hi)
    int outside;  // 1: candidates are the complement of [lo, hi)
};

template <class T> struct DpTemplDev {
    const cx<T>* phi;    // [32*P][NT] thread-order filter
    cx<T> phi_nyq;       // filter at k = N/2
    const cx<T>* s_low;  // [nlow] scaled template spectrum, natural order
    double norm;         // QETpy OF norm
    double tsum;         // sum((2 pi f)^2 |s|^2 / J) * df
    int pretrigger;      // roll applied to the delay axis
    int pad_;
};

template <class T> struct DpChanDev {
    const T* wj;      // [32*P][NT] thread-order chi0 weights
    const T* wj_low;  // [nlow] natural order
    T wj_nyq;
    int n_templ;
    int n_slots;
    int out_base;  // offset (doubles) of this channel's block in an event's output row
    DpTemplDev<T> templ[DP_MAX_TEMPLATES];
    DpSlot slots[DP_MAX_SLOTS];
};

template <class T> struct DpOfParams {
    const void* traces;      // [n_rows][row_stride] samples (in_dtype)
    long long row_stride;    // elements
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

template <int IN> DP_DEV typename DpRaw<IN>::type dp_load_raw(const void* row, long long j) {
    return __ldg(reinterpret_cast<const typename DpRaw<IN>::type*>(row) + j);
}
template <class T, int IN> DP_DEV cx<T> dp_convert_raw(typename DpRaw<IN>::type d, double x0, double sc) {
    return cx<T>{(T)(((double)d.x - x0) * sc), (T)(((double)d.y - x0) * sc)};
}
template <int IN> DP_DEV double dp_load_first(const void* row) {
    return (double)__ldg(reinterpret_cast<const typename DpRaw<IN>::scalar*>(row));
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
// Self-paired butterflies.  Thread 0 owns K12 = 0 (k = KQ*r: pairs r <-> 16-r, r = 0 is
// DC/Nyquist, r = 8 is k = M/2) and K12 = KQ/2 (k = KQ/2 + KQ*r: pairs r <-> 15-r).
// Their 17 pairs are processed by lanes 0..16 of warp 0 through a 32-element shared
// scratch, so warp 0 does not execute a second unrolled copy of the point-wise code.
template <class T> struct DpSelfPair {
    int ek, em;   // element slots (0..15 = A, 16..31 = B) of k and its mirror
    cx<T> w;      // exp(-2 pi i k / N) of the pair
    bool dc;      // the (DC, Nyquist) pair
};
template <class T> DP_DEV DpSelfPair<T> dp_self_pair(int lane) {
    DpSelfPair<T> sp;
    double s_, c_;
    if (lane < 9) {
        const int r = lane;
        sp.ek = r;
        sp.em = (16 - r) & 15;
        sincospi(-2.0 * (double)r / 32.0, &s_, &c_);
    } else {
        const int r = lane - 9;
        sp.ek = 16 + r;
        sp.em = 16 + 15 - r;
        sincospi(-2.0 * (double)(1 + 2 * r) / 64.0, &s_, &c_);
    }
    sp.w = cx<T>{(T)c_, (T)s_};
    sp.dc = lane == 0;
    return sp;
}

template <class T, int R1, int P, int IN> struct DpOfKernel {
    using G = DpGeom<R1>;
    static constexpr int NT = G::NT;
    static constexpr int NE = 32 * P;  // table entries (and X values) per thread
    static constexpr int N = 2 * P * G::MS;
    static constexpr int RED_DOUBLES = (DP_MAX_TSLOTS + 1) * 32 + 8;
    static constexpr int BEST_ELEMS = DP_MAX_TSLOTS * 32 + DP_MAX_TSLOTS;
    static constexpr int SP_ELEMS = 33 + 32;  // self-pair X (+ Nyquist) and Z'
    static constexpr size_t SMEM_BYTES = sizeof(cx<T>) * (G::SMEM_ELEMS + DP_NLOW_MAX + SP_ELEMS + 1) +
                                         sizeof(double) * RED_DOUBLES + sizeof(DpBest<T>) * BEST_ELEMS +
                                         sizeof(int) * DP_MAX_TSLOTS + 64;

    struct Smem {
        cx<T>* buf;
        cx<T>* stash;
        cx<T>* spx;  // [33] X of thread 0's butterflies (A: 0..15, B: 16..31), [32] = X at Nyquist
        cx<T>* spz;  // [32] Z' of the same
        double* red;
        DpBest<T>* best;  // [DP_MAX_TSLOTS][32] per-warp + [DP_MAX_TSLOTS] final
        int* slot_id;     // [DP_MAX_TSLOTS]
    };
    static DP_DEV Smem carve(unsigned char* raw) {
        Smem s;
        s.buf = reinterpret_cast<cx<T>*>(raw);
        s.stash = s.buf + G::SMEM_ELEMS;
        s.spx = s.stash + DP_NLOW_MAX;
        s.spz = s.spx + 33;
        s.red = reinterpret_cast<double*>(s.spz + 33);
        s.best = reinterpret_cast<DpBest<T>*>(s.red + RED_DOUBLES);
        s.slot_id = reinterpret_cast<int*>(s.best + BEST_ELEMS);
        return s;
    }

    static DP_DEV void run(const DpOfParams<T>& prm, unsigned char* smem_raw);
};

// The body is long; keep it out of the class for readability.
template <class T, int R1, int P, int IN>
DP_DEV void DpOfKernel<T, R1, P, IN>::run(const DpOfParams<T>& prm, unsigned char* smem_raw) {
    static_assert(P == 1, "P = 2 (split) path is built separately");
    const Smem sm = carve(smem_raw);
    const int tid = threadIdx.x;
    int K12, bA, bB;
    G::map(tid, K12, bA, bB);
    const cx<T> wn = dp_ldg(prm.twn + tid);  // exp(-2 pi i K12 / N)
    cx<T>* scr = prm.scratch + (long long)blockIdx.x * prm.scratch_per_cta;
    constexpr size_t ESZ = sizeof(typename DpRaw<IN>::scalar);
    const DpSelfPair<T> sp = dp_self_pair<T>(tid & 31);

    for (int row = blockIdx.x; row < prm.n_rows; row += gridDim.x) {
        const int chan = row % prm.n_chan;
        const int ev = row / prm.n_chan;
        const DpChanDev<T>& ch = prm.chans[chan];
        const void* xrow = reinterpret_cast<const unsigned char*>(prm.traces) + (size_t)row * (size_t)prm.row_stride * ESZ;
        const double x0 = prm.subtract_first ? dp_load_first<IN>(xrow) : 0.0;
        const double sc = prm.scale;

        cx<T> za[16], zb[16];
        T chi = (T)0;

        dp_fwd_subfft<T, R1, 1, IN>(xrow, 0, x0, sc, sm.buf, prm.tw1, prm.tw2, bA, bB, za, zb);

        // ---- self-paired butterflies of thread 0, cooperatively in warp 0 ------------------
        if (tid < 32) {
            if (tid == 0) {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    sm.spx[r] = za[r];
                    sm.spx[16 + r] = zb[r];
                }
            }
            __syncwarp();
            if (tid < 17) {
                cx<T> Xk, Xm;
                dp_untangle(sm.spx[sp.ek], sm.spx[sp.em], sp.w, Xk, Xm);
                chi = dp_fma(dp_ldg(ch.wj + sp.ek * NT), cnorm2(Xk), chi);
                if (sp.dc)
                    chi = dp_fma(ch.wj_nyq, cnorm2(Xm), chi);
                else if (sp.ek != sp.em)
                    chi = dp_fma(dp_ldg(ch.wj + sp.em * NT), cnorm2(Xm), chi);
                sm.spx[sp.ek] = Xk;
                if (sp.ek != sp.em) sm.spx[sp.em] = Xm;
                if (sp.dc) {
                    sm.spx[32] = Xm;    // X at Nyquist
                    sm.stash[0] = Xk;   // X at DC, first lowchi2 bin
                }
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
                // low-frequency bins for lowchi2: k = K12 < nlow lives in A[0]
                if (K12 < prm.nlow) sm.stash[K12] = za[0];
            } else {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    za[r] = sm.spx[r];
                    zb[r] = sm.spx[16 + r];
                }
            }
        }

        // L2 prefetch of the trace this CTA processes next: the load phase of the next
        // event then overlaps with this event's compute instead of waiting on HBM.
        {
            const int nrow = row + gridDim.x;
            if (nrow < prm.n_rows) {
                const unsigned char* nx = reinterpret_cast<const unsigned char*>(prm.traces) + (size_t)nrow * (size_t)prm.row_stride * ESZ;
                constexpr int nlines = (int)((size_t)N * ESZ / 128);
                for (int l = tid; l < nlines; l += NT) dp_prefetch_l2(nx + (size_t)l * 128);
            }
        }

        // multi-template: X must survive the in-place inverse of the previous template
        const bool spill_x = ch.n_templ > 1;
        if (spill_x) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                scr[r * NT + tid] = za[r];
                scr[(16 + r) * NT + tid] = zb[r];
            }
        }
        __syncthreads();  // publishes the lowchi2 stash

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
                const cx<T> Fk = cmul(dp_ldg(tp.phi + sp.ek * NT), sm.spx[sp.ek]);
                const cx<T> Fm = sp.dc ? cmul(tp.phi_nyq, sm.spx[32]) : cmul(dp_ldg(tp.phi + sp.em * NT), sm.spx[sp.em]);
                cx<T> Ck, Cm;
                dp_retangle(Fk, Fm, sp.w, Ck, Cm);
                sm.spz[sp.ek] = Ck;
                if (sp.ek != sp.em) sm.spz[sp.em] = Cm;
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
                        za[r] = sm.spz[r];
                        zb[r] = sm.spz[16 + r];
                    }
                }
                __syncwarp();
            }

            // ---- inverse: amplitude-vs-delay samples into registers ------------------------
            cx<T> y[32];
            dp_inv_subfft<T, R1>(sm.buf, prm.tw1, prm.tw2, bA, bB, za, zb, y);

            // ---- windowed arg-max: one pass over y per fit of this template -------------
            int nts = 0;
            for (int s = 0; s < ch.n_slots; ++s) {
                const DpSlot sl = ch.slots[s];
                if (sl.templ != it) continue;
                DpBest<T> b{(T)0, -1};
                if (sl.lo == 0 && sl.hi == N && !sl.outside)
                    DpScan<T, R1, P>::full(y, tid, 0, b);
                else
                    DpScan<T, R1, P>::window(y, tid, 0, sl.lo, (unsigned)(sl.hi - sl.lo), sl.outside != 0, b);
                b = dp_warp_best(b);
                if ((tid & 31) == 0) sm.best[nts * 32 + (tid >> 5)] = b;
                if (tid == 0) sm.slot_id[nts] = s;
                ++nts;
            }
            __syncthreads();
            // warp q merges fit q
            {
                constexpr int NW = (NT + 31) / 32;
                const int l = tid & 31;
                for (int q = tid >> 5; q < nts; q += NW) {
                    DpBest<T> b{(T)0, -1};
                    if (l < NW) b = sm.best[q * 32 + l];
                    b = dp_warp_best(b);
                    if (l == 0) sm.best[DP_MAX_TSLOTS * 32 + q] = b;
                }
            }
            __syncthreads();
            // ---- low-frequency chi2 at each fit's (amp, delay); chi0 rides along -----------
            {
                double part[DP_MAX_TSLOTS + 1];
#pragma unroll
                for (int q = 0; q < DP_MAX_TSLOTS; ++q) {
                    part[q] = 0.0;
                    if (q < nts && tid < prm.nlow) {
                        const DpBest<T> b = sm.best[DP_MAX_TSLOTS * 32 + q];
                        const int k = tid;
                        const int d = b.idx - tp.pretrigger;
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
                        part[q] = (double)(dp_ldg(ch.wj_low + k) * cnorm2(R));
                    }
                }
                part[DP_MAX_TSLOTS] = (it == 0) ? (double)chi : 0.0;
#pragma unroll
                for (int q = 0; q <= DP_MAX_TSLOTS; ++q) {
                    if (q < nts || (q == DP_MAX_TSLOTS && it == 0)) {
                        const double v = dp_warp_sum(part[q]);
                        if ((tid & 31) == 0) sm.red[q * 32 + (tid >> 5)] = v;
                    }
                }
            }
            __syncthreads();
            if (tid == 0) {
                constexpr int NW = (NT + 31) / 32;
                double* o = prm.out + (long long)ev * prm.n_out + ch.out_base;
                if (it == 0) {
                    double c0 = 0.0;
                    for (int w = 0; w < NW; ++w) c0 += sm.red[DP_MAX_TSLOTS * 32 + w];
                    sm.red[(DP_MAX_TSLOTS + 1) * 32] = c0;
                    o[0] = c0;
                }
                const double chi0 = sm.red[(DP_MAX_TSLOTS + 1) * 32];
                for (int q = 0; q < nts; ++q) {
                    double low = 0.0;
                    for (int w = 0; w < NW; ++w) low += sm.red[q * 32 + w];
                    const DpBest<T> b = sm.best[DP_MAX_TSLOTS * 32 + q];
                    double* os = o + 1 + sm.slot_id[q] * DP_SLOT_NOUT;
                    const double amp = (double)b.val;
                    os[0] = amp;
                    os[1] = (double)b.idx;
                    os[2] = chi0 - amp * amp * tp.norm;
                    os[3] = low;
                    os[4] = 1.0 / sqrt(amp * amp * tp.tsum);
                }
            }
            __syncthreads();  // buf / best / red reuse
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
