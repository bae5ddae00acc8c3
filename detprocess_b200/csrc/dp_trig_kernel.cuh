// Continuous-stream optimal-filter trigger (BASELINE.json config C4).
//
// Replaces OptimumFilterTrigger.update_trace + find_triggers_once for the 1x1 case
// (reference detprocess/core/oftrigger.py:588-679, 884-1034):
//     V        = oaconvolve(raw, phi_td, 'same')           (FIR filter, Nt taps)
//     filtered = iW * V ;  delta_chi2 = filtered^2 * W     (:659-672)
//     delta_chi2[:Nt] = 0 ; delta_chi2[-Nt + (Nt+1)%2:] = 0 (:676-679)
//     mask = delta_chi2 > chi2_threshold ; groups of mask indices separated by gaps
//     > pileup_window (:29-74) ; per group the first arg-max of delta_chi2 (:1001-1005).
//
// Kernel 1 (dp_trig_filter_kernel): overlap-save on the v2 FFT core.  Chunk q is the F-point
// real FFT of samples [q*H - D, q*H - D + F) x the filter spectrum (alignment, iW and all
// scalings folded into the table, so output index r IS stream index q*H + r) and the inverse
// FFT, all in shared memory / registers: every stream sample is read from HBM once per chunk
// it belongs to (F/H ~ 2 reads for Nt = F/2).  The filtered samples never leave the SM: the
// threshold test runs on the registers that hold them, candidates are marked in a shared
// bit mask and written -- ordered by stream index -- into the chunk's slice of the candidate
// list.  Chunks without a candidate (the common case) stop after the mask.
// Kernel 2 (dp_trig_group_kernel): one CTA walks the ordered candidate list, splits it where
// the gap exceeds the pile-up window and emits the first arg-max of each group.
#pragma once
#include "dp_of2_kernel.cuh"

// DP_TRIG_V3 (round 2): passes 3 / 4 / 4' / 3' warp-local like the OF kernel -- between the set barrier after pass 2 and the
// one before pass 2' a warp only needs __syncwarp() (it owns whole chunks), so the warps overlap each other's shared-memory
// and FP phases.  0 = the lock-step passes of round 1.
#ifndef DP_TRIG_V3
#define DP_TRIG_V3 1
#endif
// DP_TRIG_TMEM (round 2, fp64): the first pass of a chunk is computed once; the later phases take their blocks from tensor
// memory / the CTA's L2-resident row (Dp2Core::pass1_all / pass1_fetch) instead of reading the chunk's samples in every phase.
#ifndef DP_TRIG_TMEM
#define DP_TRIG_TMEM 1
#endif

template <class T> struct DpTrigParams {
    using S = typename Dp2Traits<T>::S;
    const void* trace;      // continuous stream (in_dtype samples)
    long long n_samples;
    int n_chunks;
    int hop;                // H: outputs per chunk
    int lead;               // D: chunk q loads samples [q*H - D, q*H - D + F)
    long long valid_lo, valid_hi;  // only outputs i in [valid_lo, valid_hi) can trigger
    const cx<T>* tw1;
    const cx<T>* tw2;
    const cx<T>* tw3;
    const cx<S>* twn;
    const int2* groups;
    const int* chunk3;      // [NPH][NT] pass-3 chunk of the thread (warp-local passes)
    cx<T>* park1;           // [grid][Dp2Core::PARK1_V] first-pass outputs that do not fit into TMEM
    const cx<T>* phi;       // [NPH][16][NT] filter spectrum, thread order
    const cx<S>* phi_self;  // [17][2]
    cx<T>* scratch;         // [grid][(NPH-1)*NB*VPB] parked block results
    long long scratch_per_cta;
    double w;               // delta_chi2 = filtered^2 * w
    double thr;             // chi2 threshold
    double scale;
    int subtract_first;
    int* cand_idx;          // [n_chunks][hop] stream index relative to chunk start (r)
    double* cand_amp;       // [n_chunks][hop] filtered amplitude
    int* cand_count;        // [n_chunks]
};

template <class T, int R1, int IN> struct DpTrigKernel {
    using G = Dp2Geom<T, R1>;
    using S = typename G::S;
    using V = cx<T>;
    using Core = Dp2Core<T, R1, IN>;
    using OF = Dp2OfKernel<T, R1, IN>;
    static constexpr int NT = G::NT, VL = G::VL, NB = G::NB, NPH = G::NPH, VPB = G::VPB, GC = G::GC, NC = G::NC, F = G::N;
    static constexpr int NW = NT / 32;
    static constexpr int MASK_WORDS = F / 32;
    static constexpr size_t SMEM_BYTES = sizeof(V) * G::SMEM_V + sizeof(cx<S>) * (32 + 34) + sizeof(unsigned) * (2 * MASK_WORDS + 64) + 64;
    static constexpr long long SCR_PARK = (long long)(NPH - 1) * NB * VPB;

    // visit the candidate samples of one pass-1' column set: fn(r, value)
    template <class Fn> static DP_DEV void for_samples(const V (&y)[GC * R1], int tid, int i0, Fn&& fn) {
#pragma unroll
        for (int n = 0; n < R1; ++n)
#pragma unroll
            for (int i = 0; i < GC; ++i) {
                const int r0 = 2 * (n * 4096 + VL * (tid + (i0 + i) * NT));
                const V& v = y[i * R1 + n];
                fn(r0, Dp2Scan<T, R1>::template re_of<0>(v));
                fn(r0 + 1, Dp2Scan<T, R1>::template im_of<0>(v));
                if constexpr (VL == 2) {
                    fn(r0 + 2, Dp2Scan<T, R1>::template re_of<1>(v));
                    fn(r0 + 3, Dp2Scan<T, R1>::template im_of<1>(v));
                }
            }
    }

    static DP_DEV void run(const DpTrigParams<T>& prm, unsigned char* smem_raw) {
        V* buf = reinterpret_cast<V*>(smem_raw);
        cx<S>* sp = reinterpret_cast<cx<S>*>(buf + G::SMEM_V);
        cx<S>* sx = sp + 32;
        unsigned* mask = reinterpret_cast<unsigned*>(sx + 34);   // [MASK_WORDS] candidate bits
        unsigned* rank = mask + MASK_WORDS;                       // [MASK_WORDS] exclusive prefix of popc(mask)
        unsigned* wsum = rank + MASK_WORDS;                       // [33] warp totals
        const int tid = threadIdx.x;
        constexpr size_t ESZ = sizeof(typename DpRaw<IN>::scalar);
        constexpr int NSPECIAL = (VL == 2) ? 1 : 2;
        V* park = prm.scratch + (long long)blockIdx.x * prm.scratch_per_cta;
        const S wS = (S)prm.w, thrS = (S)prm.thr;
        const long long jmax = prm.n_samples;  // the clamped loader bounds by the sample count
        constexpr bool TM = DP_TRIG_TMEM && Core::CAN_PARK;
        [[maybe_unused]] unsigned tm_thread = 0;
        [[maybe_unused]] V* const park1 = prm.park1 + (long long)blockIdx.x * Core::PARK1_V;
        if constexpr (TM) {
            unsigned* slot = wsum + 48;
            if (tid < 32) dp_tmem_alloc512(slot);
            dp_tmem_fence_before();
            __syncthreads();
            dp_tmem_fence_after();
            tm_thread = Core::tm_thread_base(*slot);
        }

        for (int q = blockIdx.x; q < prm.n_chunks; q += gridDim.x) {
            const long long s0 = (long long)q * prm.hop - prm.lead;  // first sample of the chunk (even)
            const bool edge = s0 < 0 || s0 + F > prm.n_samples;
            const void* xrow = reinterpret_cast<const unsigned char*>(prm.trace) + (size_t)(edge ? 0 : s0) * ESZ;
            const long long jbase = s0 / 2;  // s0 is even, also when negative
            // fp32 mode removes the chunk's first sample before the conversion (the filter has no DC
            // gain); edge chunks are zero padded, so they keep their offset
            const double x0 = (prm.subtract_first && !edge) ? dp_load_first<IN>(xrow) : 0.0;
            for (int wd = tid; wd < MASK_WORDS; wd += NT) mask[wd] = 0u;

#pragma unroll 1
            for (int p = 0; p < NPH; ++p) {
                V z[16];
                V zm[VL == 1 ? 8 : 1];
                const int2 gg = prm.groups[p * NT + tid];
                const cx<S> wn = dp_ldg(prm.twn + p * NT + tid);
                const bool special = (p == 0) && (tid < NSPECIAL);
                if constexpr (TM) {
                    if (p > 0)
                        Core::pass1_fetch(p, buf, tm_thread, park1);
                    else if (edge)
                        Core::template pass1_all<true>(prm.trace, x0, prm.scale, buf, prm.tw1, tm_thread, park1, jbase, jmax);
                    else
                        Core::template pass1_all<false>(xrow, x0, prm.scale, buf, prm.tw1, tm_thread, park1);
                } else if (edge)
                    Core::template pass1_any<true>(p, prm.trace, x0, prm.scale, buf, prm.tw1, jbase, jmax);
                else
                    Core::template pass1_any<false>(p, xrow, x0, prm.scale, buf, prm.tw1);
                // (a bulk L2 prefetch of the CTA's next chunk was measured and not kept: 0.251 -> 0.265 ms fp64 -- the chunks overlap
                // and the neighbouring CTAs have just pulled most of those lines in)
                __syncthreads();
#ifndef DP_HOST_EMU
                if (DP2_SKEW_NS > 0 && ((tid / G::CV) & 1)) __nanosleep(DP2_SKEW_NS);  // see dp_of2_kernel.cuh
#endif
#if DP_TRIG_V3
                const int chunk = prm.chunk3[p * NT + tid];
                Core::fwd_2(buf, prm.tw2, z);
                dp_bar_sync(G::bar_set_id(p, tid), G::bar_set_count(p, tid));  // pass 3 reads the chunks of the warp's block set
                Core::fwd_3w(buf, prm.tw3, chunk, z);
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
                        cx<S> Xk, Xm;
                        dp_untangle(sp[sl.ek], sp[sl.em], sl.w, Xk, Xm);
                        sx[2 * tid] = Xk;
                        sx[2 * tid + 1] = Xm;
                    }
                    __syncwarp();
                }
                if constexpr (VL == 2) {
                    OF::template untangle_all<false>(buf, z, zm, nullptr, wn, gg.x, special);
                } else {
                    // fp64: untangle, filter and inverse untangle pair by pair
                    const int Gp = __shfl_xor_sync(0xffffffffu, gg.x, 1);
                    const V* phip = prm.phi + (long long)p * 16 * NT;
                    OF::pw_publish(buf, z, gg.x);
                    OF::pw_untangle(buf, z, wn, Gp, [&](int r, cx<S> Xk, cx<S> Xm) { OF::pw_filter_pair(buf, z, r, Xk, Xm, phip, wn, Gp); });
                }
                if (p == 0 && tid < 17) {
                    const DpSelfLane<S> sl = dp_self_lane<S, 1>(tid);
                    const cx<S> Fk = cmul(dp_ldg(prm.phi_self + 2 * tid), sx[2 * tid]);
                    const cx<S> Fm = cmul(dp_ldg(prm.phi_self + 2 * tid + 1), sx[2 * tid + 1]);
                    cx<S> Ck, Cm;
                    dp_retangle(Fk, Fm, sl.w, Ck, Cm);
                    sp[sl.ek] = Ck;
                    if (sl.ek != sl.em) sp[sl.em] = Cm;
                }
                if constexpr (VL == 2)
                    OF::filter_all(buf, z, zm, prm.phi + (long long)p * 16 * NT, wn, gg.x);
                else
                    OF::pw_collect(buf, z, gg.x);
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
#if DP_TRIG_V3
                dp_dft<16, +1, T>::run(z);
                Core::store_groups(buf, gg.x, gg.y, z);
                __syncwarp();  // pass 3' reads the warp's own chunks
                Core::inv_3w(buf, prm.tw3, chunk, z);
                dp_bar_sync(G::bar_set_id(p, tid), G::bar_set_count(p, tid));  // pass 2' reads columns across the set's chunks
                Core::inv_2(buf, prm.tw2, z);
#else
                Core::inv_432(buf, prm.tw2, prm.tw3, gg.x, gg.y, z, p);
#endif
                if (p < NPH - 1) {
                    Core::park_pass2(park, p, z);
                    __syncthreads();
                } else {
                    Core::store_pass2(buf, z);
                    __syncthreads();
                }
            }
            // ---- threshold on the filtered samples (registers) -> bit mask ----------------------
            const long long ibase = (long long)q * prm.hop;
            const long long lo = prm.valid_lo - ibase, hi = prm.valid_hi - ibase;  // candidate r range, clipped below
            const int rlo = (int)(lo < 0 ? 0 : (lo > prm.hop ? prm.hop : lo));
            const int rhi = (int)(hi < 0 ? 0 : (hi > prm.hop ? prm.hop : hi));
#pragma unroll 1
            for (int i0 = 0; i0 < NC; i0 += GC) {
                V y[GC * R1];
                Core::inv_pass1(buf, park, prm.tw1, i0, y);
                for_samples(y, tid, i0, [&](int r, S a) {
                    if (r >= rlo && r < rhi && a * a * wS > thrS) atomicOr(&mask[r >> 5], 1u << (r & 31));
                });
            }
            __syncthreads();
            // ---- exclusive prefix of the per-word candidate counts -------------------------------
            constexpr int WPT = (MASK_WORDS + NT - 1) / NT;  // words per thread (contiguous)
            unsigned cnt = 0;
#pragma unroll
            for (int j = 0; j < WPT; ++j) {
                const int wd = tid * WPT + j;
                if (wd < MASK_WORDS) cnt += __popc(mask[wd]);
            }
            unsigned inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                if ((tid & 31) >= o) inc += t;
            }
            if ((tid & 31) == 31) wsum[tid >> 5] = inc;
            __syncthreads();
            if (tid < 32) {
                unsigned v = tid < NW ? wsum[tid] : 0u;
                unsigned iv = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned t = __shfl_up_sync(0xffffffffu, iv, o);
                    if (tid >= o) iv += t;
                }
                wsum[tid] = iv - v;              // exclusive warp offsets
                if (tid == 31) wsum[32] = iv;    // total
            }
            __syncthreads();
            const unsigned total = wsum[32];
            if (tid == 0) prm.cand_count[q] = (int)total;
            if (total != 0) {  // CTA-uniform
                unsigned run = wsum[tid >> 5] + inc - cnt;
#pragma unroll
                for (int j = 0; j < WPT; ++j) {
                    const int wd = tid * WPT + j;
                    if (wd < MASK_WORDS) {
                        rank[wd] = run;
                        run += __popc(mask[wd]);
                    }
                }
                __syncthreads();
                // second look at the filtered samples: candidates go to their ordered slot
                int* ci = prm.cand_idx + (long long)q * prm.hop;
                double* ca = prm.cand_amp + (long long)q * prm.hop;
#pragma unroll 1
                for (int i0 = 0; i0 < NC; i0 += GC) {
                    V y[GC * R1];
                    Core::inv_pass1(buf, park, prm.tw1, i0, y);
                    for_samples(y, tid, i0, [&](int r, S a) {
                        if (r >= rlo && r < rhi && a * a * wS > thrS) {
                            const unsigned m = mask[r >> 5];
                            const unsigned slot = rank[r >> 5] + __popc(m & ((1u << (r & 31)) - 1u));
                            ci[slot] = r;
                            ca[slot] = (double)a;
                        }
                    });
                }
            }
            __syncthreads();  // mask / buf are rewritten by the next chunk
        }
        if constexpr (TM) {
            dp_tmem_fence_before();
            __syncthreads();
            dp_tmem_fence_after();
            if (tid < 32) dp_tmem_dealloc512(*(wsum + 48));
        }
    }
};

// --------------------------------------------------------------------- grouping
struct DpTrigGroupParams {
    const int* cand_idx;      // [n_chunks][hop]
    const double* cand_amp;
    const int* cand_count;    // [n_chunks]
    int n_chunks;
    int hop;
    long long pileup_window;  // gap (samples) above which a new group starts
    long long index_shift;    // added to the arg-max index (pretrigger - Nt/2, oftrigger.py:456)
    double w;
    long long* trig_index;    // [max_triggers]
    double* trig_amp;
    double* trig_dchi2;
    int max_triggers;
    int* n_triggers;          // [1] total number of groups found (may exceed max_triggers)
    long long* chunk_offset;  // [n_chunks + 1] scratch: exclusive prefix of cand_count
    // parallel grouping (dp_trig_par_* kernels) workspace
    int* tile_heads;               // [max_tiles + 1] group heads per tile of 1024 candidates -> exclusive prefix
    unsigned long long* best_key;  // [max_triggers] ordered bits of the largest |amp| of the group
    unsigned long long* best_g;    // [max_triggers] smallest candidate number attaining it
    // residual re-trigger (oftrigger.py:752-845): when set, [n_chunks][hop] delta-chi2 values (> 0) that rank the
    // candidates and are reported instead of amp^2 w (parallel grouping only)
    const double* cand_val;
};

// parameters of the follow-up kernels on the candidate list (residual re-trigger, flat list, filtered amplitude at an index)
struct DpTrigResidParams {
    int* cand_idx;            // [n_chunks][hop], compacted in place
    double* cand_amp;
    double* cand_val;         // [n_chunks][hop] residual delta chi2 of the survivors
    int* cand_count;          // [n_chunks], updated
    int n_chunks;
    int hop;
    double w, thr;
    const long long* pulse_start;  // [n_pulses] ascending: stream index of shape[0] for every subtracted pulse
    const double* pulse_a2;        // [n_pulses] squared filtered amplitude
    int n_pulses;
    const double* shape;           // [n_shape] D
    int n_shape;
};
struct DpTrigFlattenParams {
    const int* cand_idx;
    const double* cand_amp;
    const double* cand_val;   // may be null
    const int* cand_count;
    const long long* chunk_offset;
    int n_chunks;
    int hop;
    double w;
    long long* out_idx;
    double* out_amp;
    double* out_val;          // may be null
    long long max_out;
    long long* n_out;         // [1] total number of candidates (may exceed max_out)
};
struct DpTrigAtParams {
    const void* trace;
    int in_dtype;
    long long n_samples;
    const double* phi_td;
    int nt;
    double iw;
    const long long* idx;
    int n_idx;
    double* out;
};

#ifndef DP_HOST_EMU
#ifdef DP_TRIG_DEFINE_GROUP_KERNEL  // exactly one translation unit (dp_trig_inst.cu, float64)
// inclusive scan over the 1024 threads of the CTA (warp shuffles + one shared exchange)
__device__ __forceinline__ int dp_trig_scan1024(int v, int* s_warp /* [33] */, int& total) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();  // previous users of s_warp are done
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        const int x = s_warp[lane];
        int ix = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, ix, o);
            if (lane >= o) ix += t;
        }
        s_warp[lane] = ix - x;
        if (lane == 31) s_warp[32] = ix;
    }
    __syncthreads();
    total = s_warp[32];
    return inc + s_warp[w];
}

// single CTA of 1024 threads; candidates are globally ordered by stream index
__global__ void __launch_bounds__(1024, 1) dp_trig_group_kernel(const DpTrigGroupParams prm) {
    constexpr int NTG = 1024;
    constexpr int OFF_CACHE = 1024;                  // chunk offsets kept in shared memory (binary search)
    __shared__ long long s_coff[OFF_CACHE + 1];
    __shared__ long long s_off[NTG + 1];
    // slot 0 = group carried in from the previous tile, slots 1..1024 = groups that start in this tile
    __shared__ unsigned long long s_best[NTG + 1];   // ordered bits of |amp| per group
    __shared__ long long s_bidx[NTG + 1];            // smallest stream index attaining it
    __shared__ double s_bamp[NTG + 1];
    __shared__ int s_warp[33];
    __shared__ long long s_carry_idx, s_prev_idx, s_total;
    __shared__ unsigned long long s_carry_best;
    __shared__ double s_carry_amp;
    __shared__ int s_open, s_nout;
    const int tid = threadIdx.x;
    // 1) exclusive prefix of the chunk counts (tiles of 1024 chunks)
    if (tid == 0) s_total = 0;
    __syncthreads();
    for (int c0 = 0; c0 < prm.n_chunks; c0 += NTG) {
        const int c = c0 + tid;
        const int v = c < prm.n_chunks ? prm.cand_count[c] : 0;
        int tot;
        const int inc = dp_trig_scan1024(v, s_warp, tot);
        if (c < prm.n_chunks) {
            const long long o = s_total + inc - v;
            prm.chunk_offset[c] = o;
            if (c < OFF_CACHE) s_coff[c] = o;
        }
        __syncthreads();
        if (tid == 0) s_total += tot;
        __syncthreads();
    }
    if (tid == 0) {
        prm.chunk_offset[prm.n_chunks] = s_total;
        if (prm.n_chunks <= OFF_CACHE) s_coff[prm.n_chunks] = s_total;
        s_open = 0;
        s_nout = 0;
        s_prev_idx = 0;
    }
    __syncthreads();
    const long long K = s_total;
    const bool cached = prm.n_chunks <= OFF_CACHE;
    // 2) tiles of 1024 candidates
    for (long long g0 = 0; g0 < K; g0 += NTG) {
        const long long g = g0 + tid;
        long long idx = 0;
        double amp = 0.0;
        const bool have = g < K;
        if (have) {
            // chunk of candidate g: last chunk whose offset is <= g
            int lo = 0, hi = prm.n_chunks - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                const long long om = cached ? s_coff[mid] : prm.chunk_offset[mid];
                if (om <= g) lo = mid; else hi = mid - 1;
            }
            const long long pos = g - (cached ? s_coff[lo] : prm.chunk_offset[lo]);
            idx = (long long)lo * prm.hop + prm.cand_idx[(long long)lo * prm.hop + pos];
            amp = prm.cand_amp[(long long)lo * prm.hop + pos];
        }
        s_off[tid + 1] = idx;
        if (tid == 0) s_off[0] = s_prev_idx;
        __syncthreads();
        const bool head = have && ((g == 0) || (idx - s_off[tid] > prm.pileup_window));
        // tile-local group id = number of heads at or before this candidate (0: continues the carried group)
        int ngroups;
        const int gid = dp_trig_scan1024(head ? 1 : 0, s_warp, ngroups);
        s_best[tid] = 0ull;
        s_bidx[tid] = 0x7fffffffffffffffll;
        if (tid == 0) {
            s_best[NTG] = 0ull;
            s_bidx[NTG] = 0x7fffffffffffffffll;
            if (s_open) {  // slot 0 starts from the carried group
                s_best[0] = s_carry_best;
                s_bidx[0] = s_carry_idx;
                s_bamp[0] = s_carry_amp;
            }
        }
        __syncthreads();
        const unsigned long long key = have ? (unsigned long long)__double_as_longlong(fabs(amp)) : 0ull;
        // The candidates are ordered, so the lanes of a warp that belong to one group are contiguous: reduce them with a
        // segmented shuffle scan and let only the last lane of each segment touch shared memory (a pulse puts ~1000
        // candidates into ONE group -- 1024 same-address 64-bit atomics per tile otherwise).
        const int lane = tid & 31;
        const int seg = have ? gid : -1;
        const int seg_next = __shfl_down_sync(0xffffffffu, seg, 1);
        const bool seg_tail = have && (lane == 31 || seg_next != seg);
        {
            unsigned long long k = key;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long kk = __shfl_up_sync(0xffffffffu, k, off);
                const int ss = __shfl_up_sync(0xffffffffu, seg, off);
                if (lane >= off && ss == seg && kk > k) k = kk;
            }
            if (seg_tail) atomicMax(&s_best[gid], k);
        }
        __syncthreads();
        // a member beat the carried maximum: the carried index no longer counts
        if (tid == 0 && s_open && s_best[0] != s_carry_best) s_bidx[0] = 0x7fffffffffffffffll;
        __syncthreads();
        {
            unsigned long long m = (have && key == s_best[gid]) ? (unsigned long long)idx : 0xffffffffffffffffull;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long mm = __shfl_up_sync(0xffffffffu, m, off);
                const int ss = __shfl_up_sync(0xffffffffu, seg, off);
                if (lane >= off && ss == seg && mm < m) m = mm;
            }
            if (seg_tail && m != 0xffffffffffffffffull) atomicMin(reinterpret_cast<unsigned long long*>(&s_bidx[gid]), m);
        }
        __syncthreads();
        if (have && key == s_best[gid] && idx == s_bidx[gid]) s_bamp[gid] = amp;
        __syncthreads();
        // closed groups: slot 0 (if open and a head exists in this tile) and tile groups 1..ngroups-1;
        // the last group stays open.  With no head in the tile the carried group just grows.
        const int first_slot = s_open ? 0 : 1;
        const int n_closed = ngroups > 0 ? (ngroups - first_slot) : 0;  // slots first_slot .. ngroups-1
        if (tid < n_closed) {
            const int slot = first_slot + tid;
            const int o = s_nout + tid;
            if (o < prm.max_triggers) {
                const double a = s_bamp[slot];
                prm.trig_index[o] = s_bidx[slot] + prm.index_shift;
                prm.trig_amp[o] = a;
                prm.trig_dchi2[o] = a * a * prm.w;
            }
        }
        __syncthreads();
        if (tid == 0) {
            s_nout += n_closed;
            if (ngroups > 0 || s_open) {
                s_carry_best = s_best[ngroups];
                s_carry_idx = s_bidx[ngroups];
                s_carry_amp = s_bamp[ngroups];
                s_open = 1;
            }
            const long long nk = K - g0 < NTG ? K - g0 : NTG;
            s_prev_idx = s_off[nk];
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (s_open) {
            const int o = s_nout;
            if (o < prm.max_triggers) {
                prm.trig_index[o] = s_carry_idx + prm.index_shift;
                prm.trig_amp[o] = s_carry_amp;
                prm.trig_dchi2[o] = s_carry_amp * s_carry_amp * prm.w;
            }
            s_nout += 1;
        }
        *prm.n_triggers = s_nout;
    }
}

// ---- parallel grouping: the same result as dp_trig_group_kernel from five small multi-CTA kernels.
// A group is a maximal run of candidates whose neighbours are at most pileup_window apart; its id is the number of
// group heads before it.  (1) offsets: prefix of the chunk counts; (2) heads: head flags counted per tile of 1024
// candidates; (3) scan: prefix of the tile counts = first group id of every tile, total = number of triggers;
// (4) max: 64-bit atomicMax of the ordered |amp| bits per group (after a segmented warp reduction: a pulse puts ~1000
// candidates into one group); (5) arg: atomicMin of the candidate number among those that attain it (first arg-max);
// (6) emit.
struct DpTrigTile {
    bool have;
    long long g;     // candidate number
    long long idx;   // stream index
    double amp;
    double val;      // what the group maximises: |amp|, or the residual delta chi2
    bool head;
};
__device__ __forceinline__ void dp_trig_locate(const DpTrigGroupParams& prm, const long long* coff, long long g, long long& idx, double& amp,
                                               double& val) {
    int lo = 0, hi = prm.n_chunks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (coff[mid] <= g) lo = mid; else hi = mid - 1;
    }
    const long long pos = (long long)lo * prm.hop + (g - coff[lo]);
    idx = (long long)lo * prm.hop + prm.cand_idx[pos];
    amp = prm.cand_amp[pos];
    val = prm.cand_val ? prm.cand_val[pos] : fabs(amp);
}
// candidate tid of tile t, and whether it starts a group (needs its predecessor's index)
__device__ __forceinline__ DpTrigTile dp_trig_tile_load(const DpTrigGroupParams& prm, const long long* coff, long long K, long long t,
                                                        long long* s_idx /* [1025] */) {
    const int tid = threadIdx.x;
    DpTrigTile c;
    c.g = t * 1024 + tid;
    c.have = c.g < K;
    c.idx = 0;
    c.amp = 0.0;
    c.val = 0.0;
    if (c.have) dp_trig_locate(prm, coff, c.g, c.idx, c.amp, c.val);
    __syncthreads();  // previous users of s_idx are done
    s_idx[tid + 1] = c.idx;
    if (tid == 0) {
        long long pi = 0;
        double pa, pv;
        if (c.g > 0 && c.have) dp_trig_locate(prm, coff, c.g - 1, pi, pa, pv);
        s_idx[0] = pi;
    }
    __syncthreads();
    c.head = c.have && (c.g == 0 || c.idx - s_idx[tid] > prm.pileup_window);
    return c;
}
// chunk offsets -> shared memory when they fit (binary search per candidate), else read from global memory
#define DP_TRIG_OFFSETS()                                                                          \
    __shared__ long long s_coff[1024 + 1];                                                         \
    const bool cached = prm.n_chunks <= 1024;                                                      \
    if (cached)                                                                                    \
        for (int c = threadIdx.x; c <= prm.n_chunks; c += blockDim.x) s_coff[c] = prm.chunk_offset[c]; \
    __syncthreads();                                                                               \
    const long long* coff = cached ? s_coff : prm.chunk_offset;                                    \
    const long long K = prm.chunk_offset[prm.n_chunks];                                            \
    const long long n_tiles = (K + 1023) / 1024;

__global__ void __launch_bounds__(1024, 1) dp_trig_par_offsets_kernel(const DpTrigGroupParams prm) {
    __shared__ int s_warp[33];
    __shared__ long long s_total;
    const int tid = threadIdx.x;
    if (tid == 0) s_total = 0;
    __syncthreads();
    for (int c0 = 0; c0 < prm.n_chunks; c0 += 1024) {
        const int c = c0 + tid;
        const int v = c < prm.n_chunks ? prm.cand_count[c] : 0;
        int tot;
        const int inc = dp_trig_scan1024(v, s_warp, tot);
        if (c < prm.n_chunks) prm.chunk_offset[c] = s_total + inc - v;
        __syncthreads();
        if (tid == 0) s_total += tot;
        __syncthreads();
    }
    if (tid == 0) prm.chunk_offset[prm.n_chunks] = s_total;
}
__global__ void __launch_bounds__(1024, 1) dp_trig_par_heads_kernel(const DpTrigGroupParams prm) {
    DP_TRIG_OFFSETS()
    __shared__ long long s_idx[1025];
    __shared__ int s_warp[33];
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const DpTrigTile c = dp_trig_tile_load(prm, coff, K, t, s_idx);
        int tot;
        (void)dp_trig_scan1024(c.head ? 1 : 0, s_warp, tot);
        if (threadIdx.x == 0) prm.tile_heads[t] = tot;
    }
}
__global__ void __launch_bounds__(1024, 1) dp_trig_par_scan_kernel(const DpTrigGroupParams prm) {
    __shared__ int s_warp[33];
    __shared__ int s_total;
    const int tid = threadIdx.x;
    const long long K = prm.chunk_offset[prm.n_chunks];
    const long long n_tiles = (K + 1023) / 1024;
    if (tid == 0) s_total = 0;
    __syncthreads();
    for (long long t0 = 0; t0 < n_tiles; t0 += 1024) {
        const long long t = t0 + tid;
        const int v = t < n_tiles ? prm.tile_heads[t] : 0;
        int tot;
        const int inc = dp_trig_scan1024(v, s_warp, tot);
        if (t < n_tiles) prm.tile_heads[t] = s_total + inc - v;  // groups that start before tile t
        __syncthreads();
        if (tid == 0) s_total += tot;
        __syncthreads();
    }
    if (tid == 0) *prm.n_triggers = s_total;
}
template <int PHASE> __global__ void __launch_bounds__(1024, 1) dp_trig_par_best_kernel(const DpTrigGroupParams prm) {
    DP_TRIG_OFFSETS()
    __shared__ long long s_idx[1025];
    __shared__ int s_warp[33];
    const int lane = threadIdx.x & 31;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const DpTrigTile c = dp_trig_tile_load(prm, coff, K, t, s_idx);
        int tot;
        const int inc = dp_trig_scan1024(c.head ? 1 : 0, s_warp, tot);
        const long long gid = (long long)prm.tile_heads[t] + inc - 1;  // inc == 0: continues the previous tile's last group
        const bool use = c.have && gid < prm.max_triggers;
        const int seg = use ? (int)gid : -1;
        const int seg_next = __shfl_down_sync(0xffffffffu, seg, 1);
        const bool seg_tail = use && (lane == 31 || seg_next != seg);
        const unsigned long long key = (unsigned long long)__double_as_longlong(c.val);
        if (PHASE == 0) {
            unsigned long long k = use ? key : 0ull;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long kk = __shfl_up_sync(0xffffffffu, k, off);
                const int ss = __shfl_up_sync(0xffffffffu, seg, off);
                if (lane >= off && ss == seg && kk > k) k = kk;
            }
            if (seg_tail) atomicMax(&prm.best_key[gid], k);
        } else {
            unsigned long long m = (use && key == prm.best_key[gid]) ? (unsigned long long)c.g : 0xffffffffffffffffull;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long mm = __shfl_up_sync(0xffffffffu, m, off);
                const int ss = __shfl_up_sync(0xffffffffu, seg, off);
                if (lane >= off && ss == seg && mm < m) m = mm;
            }
            if (seg_tail && m != 0xffffffffffffffffull) atomicMin(&prm.best_g[gid], m);
        }
    }
}
__global__ void dp_trig_par_emit_kernel(const DpTrigGroupParams prm) {
    const int n = *prm.n_triggers < prm.max_triggers ? *prm.n_triggers : prm.max_triggers;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) {
        long long idx;
        double amp, val;
        dp_trig_locate(prm, prm.chunk_offset, (long long)prm.best_g[o], idx, amp, val);
        prm.trig_index[o] = idx + prm.index_shift;
        prm.trig_amp[o] = amp;
        prm.trig_dchi2[o] = prm.cand_val ? val : amp * amp * prm.w;
    }
}
// ---- residual re-trigger (oftrigger.py:752-845) on the candidate list.  A first-pass pulse of filtered amplitude A
// produces the delta-chi2 shape A^2 D[j] (D = w (iw oaconvolve(template, phi_td, 'same'))^2: the chi2 trace of a unit
// pulse); the reference subtracts that shape from the WHOLE delta-chi2 trace and thresholds again.  D >= 0, so the
// second-pass candidates are a subset of the first-pass ones: only the list is revisited -- every candidate gets its
// residual (the shapes of the pulses that cover it, subtracted in trigger order like the reference's loop), survivors
// are compacted in place, chunk by chunk, still ordered by stream index, and the grouping kernels run on them.
__global__ void __launch_bounds__(1024, 1) dp_trig_residual_kernel(const DpTrigResidParams prm) {
    __shared__ int s_warp[33];
    __shared__ int s_base;
    const int tid = threadIdx.x;
    for (int q = blockIdx.x; q < prm.n_chunks; q += gridDim.x) {
        const int cnt = prm.cand_count[q];
        const long long row = (long long)q * prm.hop;
        if (tid == 0) s_base = 0;
        __syncthreads();
        for (int j0 = 0; j0 < cnt; j0 += 1024) {
            const int j = j0 + tid;
            const bool have = j < cnt;
            int r = 0;
            double a = 0.0, val = 0.0;
            if (have) {
                r = prm.cand_idx[row + j];
                a = prm.cand_amp[row + j];
                val = a * a * prm.w;
                const long long t = row + r;
                // first pulse whose shape reaches t: start + n_shape > t
                int lo = 0, hi = prm.n_pulses;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (prm.pulse_start[mid] + prm.n_shape > t) hi = mid; else lo = mid + 1;
                }
                for (int i = lo; i < prm.n_pulses; ++i) {
                    const long long st = prm.pulse_start[i];
                    if (st > t) break;
                    val -= prm.pulse_a2[i] * prm.shape[t - st];
                }
            }
            const bool keep = have && val > prm.thr;
            int tot;
            const int inc = dp_trig_scan1024(keep ? 1 : 0, s_warp, tot);  // its barriers also order this tile's reads before the writes
            const int base = s_base;
            if (keep) {
                const long long o = row + base + inc - 1;
                prm.cand_idx[o] = r;
                prm.cand_amp[o] = a;
                prm.cand_val[o] = val;
            }
            __syncthreads();
            if (tid == 0) s_base = base + tot;
            __syncthreads();
        }
        if (tid == 0) prm.cand_count[q] = s_base;
        __syncthreads();
    }
}
// the ordered candidate list as flat arrays (stream index, filtered amplitude[, residual delta chi2])
__global__ void dp_trig_flatten_kernel(const DpTrigFlattenParams prm) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *prm.n_out = prm.chunk_offset[prm.n_chunks];
    for (int q = blockIdx.x; q < prm.n_chunks; q += gridDim.x) {
        const int cnt = prm.cand_count[q];
        const long long off = prm.chunk_offset[q], row = (long long)q * prm.hop;
        for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
            const long long g = off + j;
            if (g >= prm.max_out) break;
            prm.out_idx[g] = row + prm.cand_idx[row + j];
            prm.out_amp[g] = prm.cand_amp[row + j];
            if (prm.out_val) prm.out_val[g] = prm.cand_val ? prm.cand_val[row + j] : prm.cand_amp[row + j] * prm.cand_amp[row + j] * prm.w;
        }
    }
}
// filtered amplitude iw * oaconvolve(trace, phi_td, 'same')[t] at a handful of stream indices: direct sum, one CTA per index
// (the residual pass reads the filtered trace at the SHIFTED trigger indices, oftrigger.py:794, which need not be candidates)
__global__ void __launch_bounds__(256) dp_trig_filtered_at_kernel(const DpTrigAtParams prm) {
    __shared__ double s_part[8];
    const int tid = threadIdx.x;
    for (int b = blockIdx.x; b < prm.n_idx; b += gridDim.x) {
        const long long t = prm.idx[b];
        const long long top = t + (prm.nt - 1) / 2;  // sample that meets tap 0
        double acc = 0.0;
        for (int k = tid; k < prm.nt; k += 256) {
            const long long j = top - k;
            if (j >= 0 && j < prm.n_samples) {
                double x;
                if (prm.in_dtype == 0) x = reinterpret_cast<const double*>(prm.trace)[j];
                else if (prm.in_dtype == 1) x = (double)reinterpret_cast<const float*>(prm.trace)[j];
                else x = (double)reinterpret_cast<const short*>(prm.trace)[j];
                acc = fma(x, prm.phi_td[k], acc);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        __syncthreads();
        if ((tid & 31) == 0) s_part[tid >> 5] = acc;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int i = 0; i < 8; ++i) s += s_part[i];
            prm.out[b] = (t >= 0 && t < prm.n_samples) ? prm.iw * s : 0.0;
        }
    }
}
#undef DP_TRIG_OFFSETS
#endif  // DP_TRIG_DEFINE_GROUP_KERNEL

template <class T, int R1, int IN>
__global__ void __launch_bounds__(Dp2Geom<T, R1>::NT, Dp2Geom<T, R1>::NT <= 256 ? 2 : 1) dp_trig_filter_kernel(const DpTrigParams<T> prm) {
    extern __shared__ __align__(16) unsigned char dp_smem_raw[];
    DpTrigKernel<T, R1, IN>::run(prm, dp_smem_raw);
}
#endif
