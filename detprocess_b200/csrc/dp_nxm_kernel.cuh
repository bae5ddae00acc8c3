// Fused N-channel x M-template optimal filter (OFnxm) on the v2 FFT core (nb_samples 16384 / 32768 / 65536).
//
// One CTA fits one event: the n channel traces are transformed one after the other (forward half of
// dp_of2_kernel.cuh), their spectra X_a parked in a thread-private L2 column, and per template i
//     Q_i[k] = sum_a G_ia[k] X_a[k]          (G = conj(S)^T Sigma^-1, hermitian-folded, roll / scalings folded in)
// goes through the inverse passes; the time series q~_i(t) = q_i(t) / P_ii of the first m - 1 templates are parked
// too, and while the last one is in registers
//     dchi2(t) = sum_ij C_ij q~_i(t) q~_j(t)     (C = P^-1 scaled by P_ii P_jj)
// is scanned for its first maximum inside / outside the delay window.  chi0 is the CSD quadratic form
// sum_k X^H Sigma^-1 X on the same spectra (Parseval), chi2 = chi0 - dchi2(t_best), amps = P^-1 q(t_best); the
// no-delay fit reads q~(pretrigger).
//
// Replaces qp.OFnxm(of_base, channels, template_tag).calc() + get_fit_withdelay(...) + get_fit_nodelay() as driven by
// FeatureExtractors.ofnxm (reference detprocess/core/algorithms.py:141-274).
#pragma once
#include "dp_of2_kernel.cuh"

// DP_NXM_V3 (round 2): passes 3 / 4 / 4' / 3' warp-local like the OF kernel (0 = the lock-step passes of round 1)
#ifndef DP_NXM_V3
#define DP_NXM_V3 1
#endif
// DP_NXM_TMEM (round 2, at most two channels, 512-thread geometries): the X_a columns of the current phase (16 values per
// thread and channel, re-read by every template's filter pass) live in the thread's tensor-memory columns [64 a, 64 a + 64)
// instead of an L2 scratch column (profiles/r1_regions_nxm_*: the filter stage spent 25 % of its samples waiting on those loads)
#ifndef DP_NXM_TMEM
#ifdef DP_HOST_EMU
#define DP_NXM_TMEM 0
#else
#define DP_NXM_TMEM 1
#endif
#endif

#define DP_NXM_MAX_CHAN 4
#define DP_NXM_MAX_TEMPL 3
#define DP_NXM_MAX_PAIRS (DP_NXM_MAX_CHAN * (DP_NXM_MAX_CHAN - 1) / 2)

template <class T> struct DpNxmParams {
    using S = typename Dp2Traits<T>::S;
    const double* traces;    // [n_events][n_chan][N] float64
    long long ev_stride;     // elements between events
    long long chan_stride;   // elements between the channels of an event
    int n_events;
    int n_chan, n_templ;
    const cx<T>* tw1;
    const cx<T>* tw2;
    const cx<T>* tw3;
    const cx<S>* twn;
    const int2* groups;
    const int* chunk3;      // [NPH][NT] pass-3 chunk of the thread (warp-local passes)
    const cx<T>* g;       // [n_templ][n_chan][NPH][16][NT] thread-order filters (one array: no run-time indexed
    const cx<S>* g_self;  // [n_templ][n_chan][17][2]          kernel-parameter pointers)
    const T* wd[DP_NXM_MAX_CHAN];                            // chi0 weights, diagonal (real)
    const S* wd_self[DP_NXM_MAX_CHAN];
    const cx<T>* wo[DP_NXM_MAX_PAIRS];                       // chi0 weights, a < b: 2 * W_ab
    const cx<S>* wo_self[DP_NXM_MAX_PAIRS];
    double cmat[DP_NXM_MAX_TEMPL][DP_NXM_MAX_TEMPL];         // dchi2 = sum_ij cmat_ij q~_i q~_j
    double amat[DP_NXM_MAX_TEMPL][DP_NXM_MAX_TEMPL];         // amps_i = sum_j amat_ij q~_j
    int pretrigger;
    int lo, hi, outside;     // delay window in rolled indices
    cx<T>* scratch;
    long long scratch_per_cta;  // V units
    double* out;             // [n_events][n_out]: chi0, chi2, index, amps[m], chi2_nodelay, amps_nodelay[m]
    int n_out;
    double scale;
    int subtract_first;
    int prefetch;            // 0: no L2 prefetch of the next trace row
};

template <class T, int R1, int NCH> struct DpNxmKernel {
    using G = Dp2Geom<T, R1>;
    using S = typename G::S;
    using V = cx<T>;
    using Core = Dp2Core<T, R1, 0>;
    using OF = Dp2OfKernel<T, R1, 0>;
    static constexpr int NT = G::NT, VL = G::VL, NB = G::NB, NPH = G::NPH, VPB = G::VPB, GC = G::GC, NC = G::NC, N = G::N;
    static constexpr int NW = NT / 32;
    static constexpr int SX = 34;  // X of the 17 self pairs, per channel
    static constexpr long long GSTRIDE = (long long)NPH * 16 * NT;  // one thread-order table
    static constexpr size_t SMEM_BYTES = sizeof(V) * G::SMEM_V + sizeof(cx<S>) * (32 + SX * DP_NXM_MAX_CHAN) + 2 * sizeof(double) * 32 +
                                         2 * sizeof(DpBest<S>) * 32 + 64;
    // scratch per CTA (V units)
    static constexpr long long SCR_X = (long long)16 * NT;              // per channel: X of the current phase
    static constexpr long long SCR_PARK = (long long)(NPH - 1) * NB * VPB;  // per template: parked block results
    static constexpr long long SCR_Q = (long long)NC * R1 * NT;         // per template: q~_i(t), thread order
    static DP_HD long long scratch_v(int n_chan, int n_templ) { return SCR_X * n_chan + (SCR_PARK + SCR_Q) * n_templ; }

    // the real sample with rolled index r of a parked series
    static DP_DEV S sample_at(const V* q, int r) {
        const int part = r & 1, c2 = r >> 1;
        const int n1 = c2 / 4096, rem = c2 % 4096;
        const int lane = rem % VL, col = rem / VL;
        const int t = col % NT, ii = col / NT;
        const S* ps = reinterpret_cast<const S*>(q + (long long)(ii * R1 + n1) * NT + t);
#ifdef DP_HOST_EMU
        return ps[part * VL + lane];
#else
        return __ldcg(ps + part * VL + lane);  // written by other threads of the CTA: read at L2
#endif
    }

    static DP_DEV void retangle_all(V (&z)[16], cx<S> wn) {
        static_assert(VL == 2, "packed fp32 only");
#define DP2_FP(r)                                                                          \
    {                                                                                      \
        cx<S> Ck, Cm;                                                                      \
        dp_retangle(dp2_lane0(z[r]), dp2_lane1(z[15 - r]), cmul(wn, dp_w64<S, 2 * r, -1>()), Ck, Cm); \
        dp2_set0(z[r], Ck);                                                                \
        dp2_set1(z[15 - r], Cm);                                                           \
    }
        DP2_FP(0) DP2_FP(1) DP2_FP(2) DP2_FP(3) DP2_FP(4) DP2_FP(5) DP2_FP(6) DP2_FP(7)
        DP2_FP(8) DP2_FP(9) DP2_FP(10) DP2_FP(11) DP2_FP(12) DP2_FP(13) DP2_FP(14) DP2_FP(15)
#undef DP2_FP
    }

    static DP_DEV void run(const DpNxmParams<T>& prm, unsigned char* smem_raw) {
        V* const buf = reinterpret_cast<V*>(smem_raw);
        cx<S>* const sp = reinterpret_cast<cx<S>*>(buf + G::SMEM_V);  // [32] self-paired group values
        cx<S>* const sx = sp + 32;                                    // [n_chan][17][2]
        double* const red = reinterpret_cast<double*>(sx + SX * DP_NXM_MAX_CHAN);
        DpBest<S>* const best = reinterpret_cast<DpBest<S>*>(red + 64);  // red / best: [2][32], double buffered
        int par = 0;
        const int tid = threadIdx.x;
        constexpr int nch = NCH;  // compile-time channel count: the per-entry loops unroll without branches
        const int ntm = prm.n_templ;
        V* const scr_x = prm.scratch + (long long)blockIdx.x * prm.scratch_per_cta;
        V* const scr_park = scr_x + SCR_X * nch;
        V* const scr_q = scr_park + SCR_PARK * ntm;
        constexpr int NSPECIAL = (VL == 2) ? 1 : 2;
        const unsigned long long pol = dp2_policy_keep();
        constexpr bool TMN = DP_NXM_TMEM && NT == 512 && NCH * 64 <= 128;
        [[maybe_unused]] unsigned tm_thread = 0;
#ifndef DP_HOST_EMU
        if constexpr (TMN) {
            unsigned* slot = reinterpret_cast<unsigned*>(best + 64);
            if (tid < 32) dp_tmem_alloc512(slot);
            dp_tmem_fence_before();
            __syncthreads();
            dp_tmem_fence_after();
            tm_thread = Core::tm_thread_base(*slot);
        }
#endif

        for (int ev = blockIdx.x; ev < prm.n_events; ev += gridDim.x) {
            const double* xev = prm.traces + (long long)ev * prm.ev_stride;
            S chi = (S)0;
            DpBest<S> tb{(S)0, -1};

#pragma unroll 1
            for (int p = 0; p < NPH; ++p) {
                V z[16];
                [[maybe_unused]] V zm[VL == 1 ? 8 : 1];
                const int2 gg = prm.groups[p * NT + tid];
                const cx<S> wn = dp_ldg(prm.twn + p * NT + tid);
                const bool special = (p == 0) && (tid < NSPECIAL);
                [[maybe_unused]] int Gp = 0;
                if constexpr (VL == 1) Gp = __shfl_xor_sync(0xffffffffu, gg.x, 1);

                // ---------------- forward of phase p, channel by channel: X_a -> scratch column -------------
#pragma unroll 1
                for (int a = 0; a < nch; ++a) {
                    const double* xrow = xev + (long long)a * prm.chan_stride;
                    const double x0 = prm.subtract_first ? dp_load_first<0>(xrow) : 0.0;
                    Core::pass1_any(p, xrow, x0, prm.scale, buf, prm.tw1);
                    __syncthreads();
#ifndef DP_HOST_EMU
                    if (tid == 0) {
                        // rolling prefetch, one row ahead (one TMA bulk-prefetch instruction): the next channel of this
                        // event before its first read, the next event's first channel after this event's last read.
                        // A whole event ahead (n rows per CTA) does not fit in L2 next to the scratch columns.
                        const double* nx = nullptr;
                        if (p == 0 && a + 1 < nch)
                            nx = xrow + prm.chan_stride;
                        else if (p == NPH - 1 && a == nch - 1 && ev + (int)gridDim.x < prm.n_events)
                            nx = prm.traces + (long long)(ev + gridDim.x) * prm.ev_stride;
                        if (nx != nullptr && prm.prefetch)
                            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nx), "r"((unsigned)((size_t)N * sizeof(double)))
                                         : "memory");
                    }
#endif
#ifndef DP_HOST_EMU
                    // every second block starts its passes a little late (see dp_of2_kernel.cuh: the block sets' LDS / FP /
                    // STS phases interleave instead of hitting the same pipe at the same time)
                    if (DP2_SKEW_NS > 0 && ((tid / G::CV) & 1)) __nanosleep(DP2_SKEW_NS);
#endif
#if DP_NXM_V3
                    Core::fwd_2(buf, prm.tw2, z);
                    dp_bar_sync(G::bar_set_id(p, tid), G::bar_set_count(p, tid));  // pass 3 reads the chunks of the warp's block set
                    Core::fwd_3w(buf, prm.tw3, prm.chunk3[p * NT + tid], z);
                    __syncwarp();
                    Core::load_groups(buf, gg.x, gg.y, z);
                    __syncwarp();  // the point-wise stage rewrites the warp's group rows
                    dp_dft<16, -1, T>::run(z);
#else
                    Core::fwd_234(buf, prm.tw2, prm.tw3, gg.x, gg.y, z, p);
#endif
                    if (p == 0 && tid < 32) {
                        if constexpr (VL == 2) {
                            if (tid == 0) {
#pragma unroll
                                for (int r = 0; r < 16; ++r) {
                                    sp[r] = dp2_lane0(z[r]);
                                    sp[16 + r] = dp2_lane1(z[r]);
                                }
                            }
                        } else {
                            if (tid < 2) {
#pragma unroll
                                for (int r = 0; r < 16; ++r) sp[16 * tid + r] = z[r];
                            }
                        }
                        __syncwarp();
                        if (tid < 17) {
                            const DpSelfLane<S> sl = dp_self_lane<S, 1>(tid);
                            cx<S> sXk, sXm;
                            dp_untangle(sp[sl.ek], sp[sl.em], sl.w, sXk, sXm);
                            sx[a * SX + 2 * tid] = sXk;
                            sx[a * SX + 2 * tid + 1] = sXm;
                        }
                        __syncwarp();
                    }
                    V* dst = scr_x + SCR_X * a + tid;
                    if constexpr (VL == 2) {
                        (void)OF::template untangle_all<false>(buf, z, zm, nullptr, wn, gg.x, special);
                        if constexpr (TMN) {
#ifndef DP_HOST_EMU
#pragma unroll
                            for (int r = 0; r < 16; r += 2) dp_tmem_st2(tm_thread + (unsigned)(64 * a + 4 * r), z[r], z[r + 1]);
#endif
                        } else {
#pragma unroll
                            for (int r = 0; r < 16; ++r) dp2_st_keep(dst + r * NT, z[r], pol);
                        }
                    } else {
                        OF::pw_publish(buf, z, gg.x);
                        OF::pw_untangle(buf, z, wn, Gp, [&](int r, cx<S> Xk, cx<S> Xm) {
                            if constexpr (TMN) {
#ifndef DP_HOST_EMU
                                dp_tmem_st2(tm_thread + (unsigned)(64 * a + 8 * r), Xk, Xm);   // entry e at column 4 e
#endif
                            } else {
                                dp2_st_keep(dst + (2 * r) * NT, Xk, pol);
                                dp2_st_keep(dst + (2 * r + 1) * NT, Xm, pol);
                            }
                        });
                    }
#ifndef DP_HOST_EMU
                    if constexpr (TMN) dp_tmem_wait_st();
#endif
                    __syncthreads();  // group-row reads of this channel precede the next pass-1 stores
                }

                // ---------------- chi0 (self lanes; the regular bins are folded into the first template's filter pass) ----
                {
                    if (p == 0 && tid < 17) {
                        int pi = 0;
                        for (int a = 0; a < nch; ++a) {
                            const cx<S> Xk = sx[a * SX + 2 * tid], Xm = sx[a * SX + 2 * tid + 1];
                            chi = dp_fma(dp_ldg(prm.wd_self[a] + 2 * tid), cnorm2(Xk), chi);
                            chi = dp_fma(dp_ldg(prm.wd_self[a] + 2 * tid + 1), cnorm2(Xm), chi);
                            for (int b = a + 1; b < nch; ++b, ++pi) {
                                const cx<S> tk = cmul(dp_ldg(prm.wo_self[pi] + 2 * tid), sx[b * SX + 2 * tid]);
                                const cx<S> tm = cmul(dp_ldg(prm.wo_self[pi] + 2 * tid + 1), sx[b * SX + 2 * tid + 1]);
                                chi = dp_fma(Xk.re, tk.re, chi);
                                chi = dp_fma(Xk.im, tk.im, chi);
                                chi = dp_fma(Xm.re, tm.re, chi);
                                chi = dp_fma(Xm.im, tm.im, chi);
                            }
                        }
                    }
                }

                // ---------------- per template: Q_i = sum_a G_ia X_a, inverse passes 4' 3' 2' ---------------
#pragma unroll 1
                for (int it = 0; it < ntm; ++it) {
                    V* park = scr_park + SCR_PARK * it;
                    V* qout = scr_q + SCR_Q * it;
                    if (p == 0 && tid < 17) {
                        const DpSelfLane<S> sl = dp_self_lane<S, 1>(tid);
                        cx<S> Fk{(S)0, (S)0}, Fm{(S)0, (S)0};
                        for (int a = 0; a < nch; ++a) {
                            const cx<S>* gs = prm.g_self + (it * nch + a) * SX;
                            const cx<S> gk = dp_ldg(gs + 2 * tid), gm = dp_ldg(gs + 2 * tid + 1);
                            const cx<S> tk = cmul(gk, sx[a * SX + 2 * tid]), tm = cmul(gm, sx[a * SX + 2 * tid + 1]);
                            Fk.re += tk.re;
                            Fk.im += tk.im;
                            Fm.re += tm.re;
                            Fm.im += tm.im;
                        }
                        cx<S> Ck, Cm;
                        dp_retangle(Fk, Fm, sl.w, Ck, Cm);
                        sp[sl.ek] = Ck;
                        if (sl.ek != sl.em) sp[sl.em] = Cm;
                    }
                    const long long e0 = (long long)p * 16 * NT + tid;
                    const cx<T>* const git = prm.g + (long long)it * nch * GSTRIDE;
                    // X_a of one table entry -> registers; the first template's pass also takes the CSD quadratic form
                    // sum_ab conj(X_a) W_ab X_b (chi0) from them
                    T chi_acc = (T)0.0f;
                    // X_a of the entry pair (2 q, 2 q + 1): out of tensor memory (one 8-column load per channel) or the scratch columns
                    auto load_pair = [&](int q, V (&Xa)[NCH], V (&Xb)[NCH]) {
                        if constexpr (TMN) {
#ifndef DP_HOST_EMU
                            DpTmemRaw8 t[NCH];
#pragma unroll
                            for (int a = 0; a < NCH; ++a) t[a] = dp_tmem_ld8(tm_thread + (unsigned)(64 * a + 8 * q));
                            dp_tmem_wait_ld();
#pragma unroll
                            for (int a = 0; a < NCH; ++a) {
                                Xa[a] = dp_tmem_get<V>(t[a], 0);
                                Xb[a] = dp_tmem_get<V>(t[a], 1);
                            }
#endif
                        } else {
#pragma unroll
                            for (int a = 0; a < NCH; ++a) {
                                Xa[a] = dp2_ld_keep(scr_x + SCR_X * a + (2 * q) * NT + tid, pol);
                                Xb[a] = dp2_ld_keep(scr_x + SCR_X * a + (2 * q + 1) * NT + tid, pol);
                            }
                        }
                    };
                    auto entry = [&](int ent, const V (&X)[NCH]) -> V {
                        const long long e = e0 + (long long)ent * NT;
                        if (it == 0) {
#pragma unroll
                            for (int a = 0; a < NCH; ++a) {
                                chi_acc = dp_fma(dp_ldg(prm.wd[a] + e), cnorm2(X[a]), chi_acc);
#pragma unroll
                                for (int b = a + 1; b < NCH; ++b) {
                                    const int pi = a * NCH - a * (a + 1) / 2 + (b - a - 1);  // pair index of (a, b), a < b
                                    const V t = cmul(dp_ldg(prm.wo[pi] + e), X[b]);
                                    chi_acc = dp_fma(X[a].re, t.re, chi_acc);
                                    chi_acc = dp_fma(X[a].im, t.im, chi_acc);
                                }
                            }
                        }
                        V f = cmul(dp_ldg(git + e), X[0]);
#pragma unroll
                        for (int a = 1; a < NCH; ++a) {
                            const V t = cmul(dp_ldg(git + a * GSTRIDE + e), X[a]);
                            f.re = f.re + t.re;
                            f.im = f.im + t.im;
                        }
                        return f;
                    };
                    if constexpr (VL == 2) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            V Xa[NCH], Xb[NCH];
                            load_pair(q, Xa, Xb);
                            z[2 * q] = entry(2 * q, Xa);
                            z[2 * q + 1] = entry(2 * q + 1, Xb);
                        }
                        retangle_all(z, wn);
                    } else {
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            V Xa[NCH], Xb[NCH];
                            load_pair(r, Xa, Xb);
                            const V Fk = entry(2 * r, Xa);
                            const V Fm = entry(2 * r + 1, Xb);
                            cx<S> Ck, Cm;
                            dp_retangle(Fk, Fm, cmul(wn, dp_w64_rt<S>(2 * r)), Ck, Cm);
                            z[r] = Ck;
                            buf[Gp * 17 + 15 - r] = Cm;
                        }
                        OF::pw_collect(buf, z, gg.x);
                    }
                    if (it == 0 && !special) {
                        if constexpr (VL == 2) chi += chi_acc.x + chi_acc.y; else chi += chi_acc;
                    }
                    if (p == 0 && tid < 32) {
                        __syncwarp();
                        if constexpr (VL == 2) {
                            if (tid == 0) {
#pragma unroll
                                for (int r = 0; r < 16; ++r) z[r] = V{f2(sp[r].re, sp[16 + r].re), f2(sp[r].im, sp[16 + r].im)};
                            }
                        } else {
                            if (tid < 2) {
#pragma unroll
                                for (int r = 0; r < 16; ++r) z[r] = sp[16 * tid + r];
                            }
                        }
                        __syncwarp();
                    }
#if DP_NXM_V3
                    dp_dft<16, +1, T>::run(z);
                    Core::store_groups(buf, gg.x, gg.y, z);
                    __syncwarp();  // pass 3' reads the warp's own chunks
                    Core::inv_3w(buf, prm.tw3, prm.chunk3[p * NT + tid], z);
                    dp_bar_sync(G::bar_set_id(p, tid), G::bar_set_count(p, tid));  // pass 2' reads columns across the set's chunks
                    Core::inv_2(buf, prm.tw2, z);
#else
                    Core::inv_432(buf, prm.tw2, prm.tw3, gg.x, gg.y, z, p);
#endif
                    if (p < NPH - 1) {
                        Core::park_pass2(park, p, z);
                        __syncthreads();
                        continue;
                    }
                    // ------------ last phase: pass 1' over all blocks -> q~_i(t); last template: dchi2 scan ------
                    Core::store_pass2(buf, z);
                    __syncthreads();
                    const bool last = it == ntm - 1;
                    const bool everything = prm.outside || (prm.lo == 0 && prm.hi == N);
                    int nlo = everything ? 0 : (prm.lo >> 1), nhi = everything ? 0x7fffffff : ((prm.hi - 1) >> 1);
                    {   // the no-delay sample is always needed
                        const int np = prm.pretrigger >> 1;
                        nlo = np < nlo ? np : nlo;
                        nhi = np > nhi ? np : nhi;
                    }
#pragma unroll 1
                    for (int i0 = 0; i0 < NC; i0 += GC) {
                        V y[GC * R1];
#pragma unroll
                        for (int j = 0; j < GC * R1; ++j) y[j] = V{(T)0.0f, (T)0.0f};
                        const unsigned computed = Core::inv_pass1(buf, park, prm.tw1, i0, y, nlo, nhi);
                        if (computed == 0) continue;
#pragma unroll
                        for (int j = 0; j < GC * R1; ++j) dp2_st_keep(qout + (long long)(i0 * R1 + j) * NT + tid, y[j], pol);
                        if (!last) continue;
                        // dchi2 at the thread's samples (other templates from the thread's own parked values)
                        const int il = ntm - 1;
#pragma unroll
                        for (int j = 0; j < GC * R1; ++j) {
                            const V yl = y[j];
                            const T cll = (T)(S)prm.cmat[il][il];
                            T dre = cll * yl.re * yl.re, dim = cll * yl.im * yl.im;
                            V yo[DP_NXM_MAX_TEMPL - 1];
#pragma unroll
                            for (int i = 0; i < DP_NXM_MAX_TEMPL - 1; ++i)
                                if (i < il) yo[i] = dp2_ld_keep(scr_q + SCR_Q * i + (long long)(i0 * R1 + j) * NT + tid, pol);
#pragma unroll
                            for (int i = 0; i < DP_NXM_MAX_TEMPL - 1; ++i) {
                                if (i < il) {
                                    const T cil = (T)(S)(2.0 * prm.cmat[i][il]);
                                    dre = dp_fma(cil * yo[i].re, yl.re, dre);
                                    dim = dp_fma(cil * yo[i].im, yl.im, dim);
#pragma unroll
                                    for (int k = i; k < DP_NXM_MAX_TEMPL - 1; ++k) {
                                        if (k < il) {
                                            const T cik = (T)(S)((k == i ? 1.0 : 2.0) * prm.cmat[i][k]);
                                            dre = dp_fma(cik * yo[i].re, yo[k].re, dre);
                                            dim = dp_fma(cik * yo[i].im, yo[k].im, dim);
                                        }
                                    }
                                }
                            }
                            y[j] = V{dre, dim};
                        }
                        if (prm.lo == 0 && prm.hi == N && !prm.outside)
                            Dp2Scan<T, R1>::full(y, tid, i0, tb);
                        else
                            Dp2Scan<T, R1>::window(y, tid, i0, prm.lo, (unsigned)(prm.hi - prm.lo), prm.outside != 0, tb, computed);
                    }
                    if (!last) {
                        __syncthreads();  // pass-1' reads of buf precede the next template's group stores
                        continue;
                    }
                    // ------------ winner, chi0, outputs ---------------------------------------------------
                    // red / best are double buffered and warp 0 finishes the event alone, so the other warps start the
                    // next event's forward passes right away (they do not touch scr_q before its last phase)
                    DpBest<S>* const bestp = best + par * 32;
                    double* const redp = red + par * 32;
                    par ^= 1;
                    {
                        const DpBest<S> b = dp_warp_best(tb);
                        const double c = dp_warp_sum((double)chi);
                        if ((tid & 31) == 0) {
                            bestp[tid >> 5] = b;
                            redp[tid >> 5] = c;
                        }
                    }
                    __syncthreads();  // also orders the parked q~ values (global memory) within the CTA
                    if (tid < 32) {
                        const double chi0 = dp_warp_sum(tid < NW ? redp[tid] : 0.0);
                        DpBest<S> b = tid < NW ? bestp[tid] : DpBest<S>{(S)0, -1};
                        b = dp_warp_best(b);
                        // lane f * ntm + i reads q~_i at the delay of fit f (0: windowed arg-max, 1: no delay)
                        const int f_l = tid / ntm, i_l = tid % ntm;
                        const int idx_l = f_l == 0 ? b.idx : prm.pretrigger;
                        double ql = 0.0;
                        if (tid < 2 * ntm && idx_l >= 0) ql = (double)sample_at(scr_q + SCR_Q * i_l, idx_l);
                        double q[2][DP_NXM_MAX_TEMPL];
#pragma unroll
                        for (int f = 0; f < 2; ++f)
#pragma unroll
                            for (int i = 0; i < DP_NXM_MAX_TEMPL; ++i) q[f][i] = __shfl_sync(0xffffffffu, ql, (f * ntm + i) & 31);
                        if (tid == 0) {
                            double* o = prm.out + (long long)ev * prm.n_out;
                            o[0] = chi0;
#pragma unroll
                            for (int f = 0; f < 2; ++f) {
                                const int idx = f == 0 ? b.idx : prm.pretrigger;
                                double* of = o + 1 + f * (2 + ntm);
                                if (idx < 0) {  // empty window
                                    for (int i = 0; i < (f == 0 ? 2 : 1) + ntm; ++i) of[i] = -999999.0;
                                    continue;
                                }
                                double d = 0.0;
#pragma unroll
                                for (int i = 0; i < DP_NXM_MAX_TEMPL; ++i)
#pragma unroll
                                    for (int k = 0; k < DP_NXM_MAX_TEMPL; ++k)
                                        if (i < ntm && k < ntm) d += prm.cmat[i][k] * q[f][i] * q[f][k];
                                int oi = 0;
                                of[oi++] = chi0 - d;
                                if (f == 0) of[oi++] = (double)idx;
#pragma unroll
                                for (int i = 0; i < DP_NXM_MAX_TEMPL; ++i) {
                                    if (i < ntm) {
                                        double am = 0.0;
#pragma unroll
                                        for (int k = 0; k < DP_NXM_MAX_TEMPL; ++k)
                                            if (k < ntm) am += prm.amat[i][k] * q[f][k];
                                        of[oi++] = am;
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
#ifndef DP_HOST_EMU
        if constexpr (TMN) {
            dp_tmem_fence_before();
            __syncthreads();
            dp_tmem_fence_after();
            if (tid < 32) dp_tmem_dealloc512(*reinterpret_cast<unsigned*>(best + 64));
        }
#endif
    }
};

#ifndef DP_HOST_EMU
template <class T, int R1, int NCH>
__global__ void __launch_bounds__(Dp2Geom<T, R1>::NT, Dp2Geom<T, R1>::NT <= 256 ? 2 : 1) dp_nxm_kernel(const DpNxmParams<T> prm) {
    extern __shared__ __align__(16) unsigned char dp_smem_raw[];
    DpNxmKernel<T, R1, NCH>::run(prm, dp_smem_raw);
}
#endif
