// Noise cross-spectral density accumulation on the v2 FFT core (nb_samples 16384 / 32768 / 65536), NCH channels per
// event: the channels go through the forward half of the fused OF kernel one after the other (their spectra parked in a
// thread-private L2 column, as in dp_nxm_kernel.cuh), then the thread adds X_a conj(X_b) of its bins into per-CTA
// partial sums kept in thread order -- NCH real arrays (a == b) and a (re, im) pair of arrays per a < b; coalesced
// read-modify-write of thread-private slots, no atomics.  dp_csd_reduce_kernel folds the CTAs and maps thread order ->
// natural k; the per-GPU sums are all-reduced over NCCL by the host layer and turned into the two-sided [n, n, N] array.
//
// Replaces qp.calc_csd(traces[cut], fs, folded_over=False) as called from Noise.calc_csd
// (reference detprocess/core/noise.py:374-470; the cut of :431-450 enters as the event mask).
#pragma once
#include "dp_of2_kernel.cuh"

#ifndef DP_CSD_V3
#define DP_CSD_V3 1
#endif
#ifndef DP_CSD_TMEM
#ifdef DP_HOST_EMU
#define DP_CSD_TMEM 0
#else
#define DP_CSD_TMEM 1
#endif
#endif

#define DP_CSD_MAX_CHAN 4

template <class T> struct DpCsdParams {
    using S = typename Dp2Traits<T>::S;
    const double* traces;    // [n_events][n_chan][N] float64
    long long ev_stride;     // elements
    long long chan_stride;
    int n_events;
    const unsigned char* mask;  // [n_events] 1 = use the event (nullptr: all)
    const cx<T>* tw1;
    const cx<T>* tw2;
    const cx<T>* tw3;
    const cx<S>* twn;
    const int2* groups;
    const int* chunk3;          // [NPH][NT] pass-3 chunk of the thread (warp-local passes)
    cx<T>* scratch;
    long long scratch_per_cta;  // V units
    double* partial;            // [grid][PARTIAL slots][NCP components]
    long long partial_per_cta;
    unsigned long long* count;  // [grid] accepted events per CTA
    double scale;
    int subtract_first;
};

template <class T, int R1, int NCH> struct DpCsdKernel {
    using G = Dp2Geom<T, R1>;
    using S = typename G::S;
    using V = cx<T>;
    using Core = Dp2Core<T, R1, 0>;
    using OF = Dp2OfKernel<T, R1, 0>;
    static constexpr int NT = G::NT, VL = G::VL, NPH = G::NPH, N = G::N;
    static constexpr int SX = 34;
    static constexpr int NCOMP = NCH * NCH;  // NCH diagonal sums + (re, im) per pair a < b
    static constexpr int NCP = (NCOMP + 1) & ~1;  // components of one bin are adjacent (padded to 16-byte pairs)
    static constexpr size_t SMEM_BYTES = sizeof(V) * G::SMEM_V + sizeof(cx<S>) * (32 + SX * NCH) + sizeof(double) * NCH + 64;
    static constexpr long long PARTIAL = (long long)NPH * 16 * NT * VL + 17 * 2;  // slots of one component
    static constexpr long long SCR_X = (long long)16 * NT;
    static DP_HD long long scratch_v() { return SCR_X * NCH; }
    static DP_HD int pair_index(int a, int b) { return a * NCH - a * (a + 1) / 2 + (b - a - 1); }

    static DP_DEV void run(const DpCsdParams<T>& prm, unsigned char* smem_raw) {
        V* const buf = reinterpret_cast<V*>(smem_raw);
        cx<S>* const sp = reinterpret_cast<cx<S>*>(buf + G::SMEM_V);
        cx<S>* const sx = sp + 32;                                        // [NCH][17][2]
        double* const dcv = reinterpret_cast<double*>(sx + SX * NCH);     // [NCH] DC bin in double (thread 0)
        const int tid = threadIdx.x;
        constexpr int NSPECIAL = (VL == 2) ? 1 : 2;
        V* const scr_x = prm.scratch + (long long)blockIdx.x * prm.scratch_per_cta;
        double* const part = prm.partial + (long long)blockIdx.x * prm.partial_per_cta;
        const double inv_s2 = 1.0 / (4.0 * prm.scale * prm.scale);  // kernel values are 2*scale*X
        const unsigned long long pol = dp2_policy_keep();
        unsigned long long n_acc = 0;
        // round 2 (as in the PSD kernel): warp-local passes 3 / 4; with at most two channels the first pass of every channel
        // is computed once and phase 1 finds its blocks in the thread's TMEM columns (channel a: [a * TM_COLS, +TM_COLS));
        // the CTA's next event goes to L2 by bulk prefetch behind this event's last read
        constexpr bool TMC = DP_CSD_TMEM && Core::CAN_PARK && Core::PARK1_V == 0 && NCH * Core::TM_COLS <= 128;
        [[maybe_unused]] unsigned tm_thread = 0;
#ifndef DP_HOST_EMU
        if constexpr (TMC) {
            unsigned* slot = reinterpret_cast<unsigned*>(dcv + NCH);
            if (tid < 32) dp_tmem_alloc512(slot);
            dp_tmem_fence_before();
            __syncthreads();
            dp_tmem_fence_after();
            tm_thread = Core::tm_thread_base(*slot);
        }
#endif

        for (int ev = blockIdx.x; ev < prm.n_events; ev += gridDim.x) {
            if (prm.mask != nullptr && prm.mask[ev] == 0) continue;  // CTA-uniform
            ++n_acc;
            const double* xev = prm.traces + (long long)ev * prm.ev_stride;
            auto prefetch_next_event = [&]() {
#ifndef DP_HOST_EMU
                int nev = ev + gridDim.x;
                while (prm.mask != nullptr && nev < prm.n_events && prm.mask[nev] == 0) nev += gridDim.x;
                if (VL == 1 && tid < NCH && nev < prm.n_events) {
                    const double* nx = prm.traces + (long long)nev * prm.ev_stride + (long long)tid * prm.chan_stride;
                    if ((reinterpret_cast<unsigned long long>(nx) & 15ull) == 0)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nx), "r"((unsigned)(N * sizeof(double))) : "memory");
                }
#endif
            };
#pragma unroll 1
            for (int p = 0; p < NPH; ++p) {
                V z[16];
                [[maybe_unused]] V zm[VL == 1 ? 8 : 1];
                const int2 gg = prm.groups[p * NT + tid];
                const cx<S> wn = dp_ldg(prm.twn + p * NT + tid);
                const bool special = (p == 0) && (tid < NSPECIAL);
                [[maybe_unused]] int Gp = 0;
                if constexpr (VL == 1) Gp = __shfl_xor_sync(0xffffffffu, gg.x, 1);
#pragma unroll 1
                for (int a = 0; a < NCH; ++a) {
                    const double* xrow = xev + (long long)a * prm.chan_stride;
                    const double x0 = prm.subtract_first ? dp_load_first<0>(xrow) : 0.0;
                    if constexpr (TMC) {
                        if (p == 0)
                            Core::pass1_all(xrow, x0, prm.scale, buf, prm.tw1, tm_thread + (unsigned)(a * Core::TM_COLS), nullptr);
                        else
                            Core::pass1_fetch(p, buf, tm_thread + (unsigned)(a * Core::TM_COLS), nullptr);
                        if (p == 0 && a == NCH - 1) prefetch_next_event();
                    } else {
                        Core::pass1_any(p, xrow, x0, prm.scale, buf, prm.tw1);
                        if (p == NPH - 1 && a == NCH - 1) prefetch_next_event();
                    }
                    __syncthreads();
#ifndef DP_HOST_EMU
                    // every second block starts its passes a little late (see dp_of2_kernel.cuh: the block sets' LDS / FP /
                    // STS phases interleave instead of hitting the same pipe at the same time)
                    if (DP2_SKEW_NS > 0 && ((tid / G::CV) & 1)) __nanosleep(DP2_SKEW_NS);
#endif
                    // (packed fp32 keeps the lock-step passes and no prefetch: 2.67 M events/s against 2.46 M with them, 2 ch x 32768)
                    if constexpr (DP_CSD_V3 && VL == 1) {
                    Core::fwd_2(buf, prm.tw2, z);
                    dp_bar_sync(G::bar_set_id(p, tid), G::bar_set_count(p, tid));  // pass 3 reads the chunks of the warp's block set
                    Core::fwd_3w(buf, prm.tw3, prm.chunk3[p * NT + tid], z);
                    __syncwarp();
                    Core::load_groups(buf, gg.x, gg.y, z);
                    __syncwarp();  // the point-wise stage rewrites the warp's group rows
                    dp_dft<16, -1, T>::run(z);
                    } else {
                    Core::fwd_234(buf, prm.tw2, prm.tw3, gg.x, gg.y, z, p);
                    }
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
                            cx<S> Xk, Xm;
                            dp_untangle(sp[sl.ek], sp[sl.em], sl.w, Xk, Xm);
                            sx[a * SX + 2 * tid] = Xk;
                            sx[a * SX + 2 * tid + 1] = Xm;
                            // DC bin in double, with the subtracted first sample put back: X[0] += N*x0
                            if (tid == 0) dcv[a] = (double)Xk.re / (2.0 * prm.scale) + (double)N * x0;
                        }
                        __syncwarp();
                    }
                    V* dst = scr_x + SCR_X * a + tid;
                    if constexpr (VL == 2) {
                        (void)OF::template untangle_all<false>(buf, z, zm, nullptr, wn, gg.x, special);
#pragma unroll
                        for (int r = 0; r < 16; ++r) dp2_st_keep(dst + r * NT, z[r], pol);
                    } else {
                        OF::pw_publish(buf, z, gg.x);
                        OF::pw_untangle(buf, z, wn, Gp, [&](int r, cx<S> Xk, cx<S> Xm) {
                            dp2_st_keep(dst + (2 * r) * NT, Xk, pol);
                            dp2_st_keep(dst + (2 * r + 1) * NT, Xm, pol);
                        });
                    }
                    __syncthreads();  // group-row reads of this channel precede the next pass-1 stores
                }

                // ---------------- X_a conj(X_b) of the thread's bins -> its partial-sum slots ----------------
                if (!special) {
#pragma unroll 1
                    for (int r = 0; r < 16; ++r) {
                        V X[NCH];
#pragma unroll
                        for (int a = 0; a < NCH; ++a) X[a] = dp2_ld_keep(scr_x + SCR_X * a + r * NT + tid, pol);
                        T vals[NCP];
                        if constexpr (NCP != NCOMP) vals[NCP - 1] = (T)0.0f;
#pragma unroll
                        for (int a = 0; a < NCH; ++a) {
                            vals[a] = cnorm2(X[a]);
#pragma unroll
                            for (int b = a + 1; b < NCH; ++b) {
                                const int c = NCH + 2 * pair_index(a, b);
                                vals[c] = dp_fma(X[a].re, X[b].re, X[a].im * X[b].im);
                                vals[c + 1] = dp_fma(X[a].im, X[b].re, -(X[a].re * X[b].im));
                            }
                        }
                        // read-modify-write of the bin's component vector: 16-byte accesses kept in L2 (evict_last)
                        const long long slot = ((long long)p * 16 + r) * NT + tid;
#pragma unroll
                        for (int l = 0; l < VL; ++l) {
                            cx<double>* q = reinterpret_cast<cx<double>*>(part + (slot * VL + l) * NCP);
#pragma unroll
                            for (int c2 = 0; c2 < NCP / 2; ++c2) {
                                cx<double> acc = dp2_ld_keep(q + c2, pol);
                                double v0, v1;
                                if constexpr (VL == 2) {
                                    v0 = (double)(l == 0 ? vals[2 * c2].x : vals[2 * c2].y);
                                    v1 = (double)(l == 0 ? vals[2 * c2 + 1].x : vals[2 * c2 + 1].y);
                                } else {
                                    v0 = vals[2 * c2];
                                    v1 = vals[2 * c2 + 1];
                                }
                                acc.re += v0 * inv_s2;
                                acc.im += v1 * inv_s2;
                                dp2_st_keep(q + c2, acc, pol);
                            }
                        }
                    }
                }
                if (p == 0 && tid < 17) {
                    const long long base = ((long long)NPH * 16 * NT * VL + 2 * tid) * NCP;  // slot-major: [slot][component]
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const bool dc = (tid == 0 && j == 0);
#pragma unroll
                        for (int a = 0; a < NCH; ++a) {
                            const cx<S> Xa = sx[a * SX + 2 * tid + j];
                            part[base + j * NCP + a] += dc ? dcv[a] * dcv[a] : (double)cnorm2(Xa) * inv_s2;
#pragma unroll
                            for (int b = a + 1; b < NCH; ++b) {
                                const cx<S> Xb = sx[b * SX + 2 * tid + j];
                                const int c = NCH + 2 * pair_index(a, b);
                                const double re = (double)Xa.re * (double)Xb.re + (double)Xa.im * (double)Xb.im;
                                const double im = (double)Xa.im * (double)Xb.re - (double)Xa.re * (double)Xb.im;
                                part[base + j * NCP + c] += dc ? dcv[a] * dcv[b] : re * inv_s2;
                                part[base + j * NCP + c + 1] += dc ? 0.0 : im * inv_s2;
                            }
                        }
                    }
                }
                __syncthreads();  // sx / dcv / the scratch columns are rewritten by the next phase or event
            }
        }
        if (tid == 0) prm.count[blockIdx.x] += n_acc;
#ifndef DP_HOST_EMU
        if constexpr (TMC) {
            dp_tmem_fence_before();
            __syncthreads();
            dp_tmem_fence_after();
            if (tid < 32) dp_tmem_dealloc512(*reinterpret_cast<unsigned*>(dcv + NCH));
        }
#endif
    }
};

struct DpCsdReduceParams {
    const double* partial;
    long long partial_per_cta;
    int ncp;          // components per slot (padded)
    int grid;
    const int* loc;   // [nbins] natural bin k -> slot of one component
    int nbins;
    int ncomp;
    double* sum_out;  // [ncomp][nbins], zeroed by the caller
    const unsigned long long* count;
    unsigned long long* count_out;
};

#ifndef DP_HOST_EMU
template <class T, int R1, int NCH>
__global__ void __launch_bounds__(Dp2Geom<T, R1>::NT, Dp2Geom<T, R1>::NT <= 256 ? 2 : 1) dp_csd_kernel(const DpCsdParams<T> prm) {
    extern __shared__ __align__(16) unsigned char dp_smem_raw[];
    DpCsdKernel<T, R1, NCH>::run(prm, dp_smem_raw);
}
#endif
