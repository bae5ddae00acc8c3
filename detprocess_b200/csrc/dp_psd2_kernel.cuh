// Noise PSD accumulation on the v2 FFT core (nb_samples 16384 / 32768 / 65536): per trace the
// forward half of the fused OF kernel (dp_of2_kernel.cuh) and |X[k]|^2 added into a per-CTA
// partial-sum array kept in thread order (coalesced read-modify-write of thread-private
// slots, no atomics, L2 resident).  dp_psd_reduce_kernel (dp_of_inst.cu) folds the CTAs and
// maps thread order -> natural k; the per-GPU sums are all-reduced over NCCL by the host layer.
//
// Replaces qp.calc_psd(traces[cut], fs, folded_over=False) as called from Noise.calc_psd
// (reference detprocess/core/noise.py:344).
#pragma once
#include "dp_of2_kernel.cuh"

// DP_PSD_V3 (round 2): passes 3 / 4 warp-local like the OF kernel (one set barrier after pass 2 instead of a block and a set
// barrier; the warps drift through pass 3, pass 4 and the untangle), and the partial sums updated by fire-and-forget
// reductions (RED.ADD.F64 at the L2: the thread-private slots need no load, so no thread waits on an L2 round trip before
// its add).  0 = the round-1 kernel (lock-step passes, load / add / store).
#ifndef DP_PSD_V3
#define DP_PSD_V3 1
#endif
// DP_PSD_TMEM (round 2, fp64 at 32768 / 65536 samples): the first pass of every column is computed once per trace; the later
// phases find their blocks in tensor memory (and, at 65536 samples, the last phase in an L2-resident row) instead of reading
// the whole trace again -- Dp2Core::pass1_all / pass1_fetch.  Round 1 read the trace in every phase: 4 x 512 KB through the
// SM's L2 port at 65536 samples, 45 % of the kernel's samples waiting on those loads (profiles/r2_prof_psd2_f64_64k_before.txt).
#ifndef DP_PSD_TMEM
#define DP_PSD_TMEM 1
#endif

#if !defined(DP_HOST_EMU)
DP_DEV void dp_psd_acc(double* slot, double v) {
#if DP_PSD_V3
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(slot), "d"(v) : "memory");
#else
    *slot += v;
#endif
}
#else
static inline void dp_psd_acc(double* slot, double v) { *slot += v; }
#endif


template <class T> struct DpPsd2Params {
    using S = typename Dp2Traits<T>::S;
    const void* traces;
    long long row_stride;
    int n_rows;
    const unsigned char* mask;  // [n_rows] 1 = use the trace (nullptr: all)
    const cx<T>* tw1;
    const cx<T>* tw2;
    const cx<T>* tw3;
    const cx<S>* twn;
    const int2* groups;
    cx<T>* park1;               // [grid][Dp2Core::PARK1_V] first-pass outputs that do not fit into TMEM (L2 resident)
    const int* chunk3;          // [NPH][NT] pass-3 chunk of the thread (warp-local passes, dpplan2::build_chunks)
    double* partial;            // [grid][partial_per_cta]: [NPH][16][NT][VL] thread order, then [17][2] self lanes
    long long partial_per_cta;
    unsigned long long* count;  // [grid] accepted traces per CTA
    double scale;
    int subtract_first;
};

template <class T, int R1, int IN> struct DpPsd2Kernel {
    using G = Dp2Geom<T, R1>;
    using S = typename G::S;
    using V = cx<T>;
    using Core = Dp2Core<T, R1, IN>;
    using OF = Dp2OfKernel<T, R1, IN>;
    static constexpr int NT = G::NT, VL = G::VL, NPH = G::NPH, N = G::N;
    static constexpr size_t SMEM_BYTES = sizeof(V) * G::SMEM_V + sizeof(cx<S>) * 32 + 64;
    static constexpr long long PARTIAL = (long long)NPH * 16 * NT * VL + 17 * 2;

    static DP_DEV void run(const DpPsd2Params<T>& prm, unsigned char* smem_raw) {
        V* buf = reinterpret_cast<V*>(smem_raw);
        cx<S>* sp = reinterpret_cast<cx<S>*>(buf + G::SMEM_V);
        const int tid = threadIdx.x;
        constexpr size_t ESZ = sizeof(typename DpRaw<IN>::scalar);
        constexpr int NSPECIAL = (VL == 2) ? 1 : 2;
        double* part = prm.partial + (long long)blockIdx.x * prm.partial_per_cta;
        double* part_self = part + (long long)NPH * 16 * NT * VL;
        const double inv_s2 = 1.0 / (4.0 * prm.scale * prm.scale);  // kernel values are 2*scale*X
        unsigned long long n_acc = 0;
        [[maybe_unused]] V* const park1 = prm.park1 + (long long)blockIdx.x * Core::PARK1_V;  // first-pass outputs of the last phase
        // fp64 at 65536 samples: the first pass of phases 1 and 2 comes out of tensor memory (see above)
        constexpr bool TM = DP_PSD_TMEM && Core::CAN_PARK;
        [[maybe_unused]] unsigned tm_thread = 0;  // this thread's TMEM address: lane quarter of the warp, 128 columns of its own
        if constexpr (TM) {
            unsigned* slot = reinterpret_cast<unsigned*>(sp + 32);
            if (tid < 32) dp_tmem_alloc512(slot);
            dp_tmem_fence_before();
            __syncthreads();
            dp_tmem_fence_after();
            tm_thread = Core::tm_thread_base(*slot);
        }

        for (int row = blockIdx.x; row < prm.n_rows; row += gridDim.x) {
            if (prm.mask != nullptr && prm.mask[row] == 0) continue;  // CTA-uniform
            ++n_acc;
            const void* xrow = reinterpret_cast<const unsigned char*>(prm.traces) + (size_t)row * (size_t)prm.row_stride * ESZ;
            const double x0 = prm.subtract_first ? dp_load_first<IN>(xrow) : 0.0;
            // the CTA's next trace -> L2 (one bulk prefetch) once this trace's last read is done, so that the next first pass
            // finds it there (profiles/r2_prof_psd2_f64_64k.txt before this: 21 % of the samples waiting on the DRAM reads of pass 1)
            auto prefetch_next_row = [&]() {
#ifndef DP_HOST_EMU
                int nrow = row + gridDim.x;
                while (prm.mask != nullptr && nrow < prm.n_rows && prm.mask[nrow] == 0) nrow += gridDim.x;
                if (tid == 0 && nrow < prm.n_rows) {
                    const unsigned char* nx = reinterpret_cast<const unsigned char*>(prm.traces) + (size_t)nrow * (size_t)prm.row_stride * ESZ;
                    if ((reinterpret_cast<unsigned long long>(nx) & 15ull) == 0)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nx), "r"((unsigned)((size_t)N * ESZ)) : "memory");
                }
#endif
            };
#pragma unroll 1
            for (int p = 0; p < NPH; ++p) {
                V z[16];
                V zm[VL == 1 ? 8 : 1];
                const int2 gg = prm.groups[p * NT + tid];
                const cx<S> wn = dp_ldg(prm.twn + p * NT + tid);
                const bool special = (p == 0) && (tid < NSPECIAL);
                if constexpr (TM) {
                    if (p == 0) {
                        Core::pass1_all(xrow, x0, prm.scale, buf, prm.tw1, tm_thread, park1);
                        prefetch_next_row();
                    } else {
                        Core::pass1_fetch(p, buf, tm_thread, park1);
                    }
                } else {
                    Core::pass1_any(p, xrow, x0, prm.scale, buf, prm.tw1);
                    if (p == NPH - 1) prefetch_next_row();
                }
                __syncthreads();
#ifndef DP_HOST_EMU
                if (DP2_SKEW_NS > 0 && ((tid / G::CV) & 1)) __nanosleep(DP2_SKEW_NS);  // see dp_of2_kernel.cuh
#endif
#ifndef DP_PSD_V3_F32
#define DP_PSD_V3_F32 0
#endif
                if constexpr (DP_PSD_V3 && (VL == 1 || DP_PSD_V3_F32)) {
                Core::fwd_2(buf, prm.tw2, z);
                dp_bar_sync(G::bar_set_id(p, tid), G::bar_set_count(p, tid));  // pass 3 reads the chunks of the warp's block set
                Core::fwd_3w(buf, prm.tw3, prm.chunk3[p * NT + tid], z);
                __syncwarp();
                Core::load_groups(buf, gg.x, gg.y, z);
                __syncwarp();  // pw_publish rewrites the warp's group rows
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
                        double pk = (double)cnorm2(Xk) * inv_s2, pm = (double)cnorm2(Xm) * inv_s2;
                        if (tid == 0 && prm.subtract_first) {
                            // DC bin: put back the subtracted first sample, X[0] += N*x0 (in double)
                            const double dc = (double)Xk.re / (2.0 * prm.scale) + (double)N * x0;
                            pk = dc * dc;
                        }
                        dp_psd_acc(part_self + 2 * tid, pk);
                        dp_psd_acc(part_self + 2 * tid + 1, pm);
                    }
                    __syncwarp();
                }
                if constexpr (VL == 2) {
                    OF::template untangle_all<false>(buf, z, zm, nullptr, wn, gg.x, special);
                    if (!special) {
                        double2* dst = reinterpret_cast<double2*>(part) + (long long)p * 16 * NT + tid;
#pragma unroll
                        for (int r = 0; r < 16; ++r) {
                            const f2 pw = cnorm2(z[r]);
                            // (one 16-byte load / store per pair of bins; two 8-byte reductions each were slower: -5 %)
                            double2 a = dst[r * NT];
                            a.x += (double)pw.x * inv_s2;
                            a.y += (double)pw.y * inv_s2;
                            dst[r * NT] = a;
                        }
                    }
                } else {
                    const int Gp = __shfl_xor_sync(0xffffffffu, gg.x, 1);
                    double* dst = part + (long long)p * 16 * NT + tid;
                    OF::pw_publish(buf, z, gg.x);
                    OF::pw_untangle(buf, z, wn, Gp, [&](int r, cx<S> Xk, cx<S> Xm) {
                        if (!special) {
                            dp_psd_acc(dst + (2 * r) * NT, cnorm2(Xk) * inv_s2);
                            dp_psd_acc(dst + (2 * r + 1) * NT, cnorm2(Xm) * inv_s2);
                        }
                    });
                }
                __syncthreads();  // group rows of buf are rewritten by the next pass 1
            }
        }
        if (tid == 0) prm.count[blockIdx.x] += n_acc;
        if constexpr (TM) {
            dp_tmem_fence_before();
            __syncthreads();
            dp_tmem_fence_after();
            if (tid < 32) dp_tmem_dealloc512(*reinterpret_cast<unsigned*>(sp + 32));
        }
    }
};

#ifndef DP_HOST_EMU
template <class T, int R1, int IN>
__global__ void __launch_bounds__(Dp2Geom<T, R1>::NT, Dp2Geom<T, R1>::NT <= 256 ? 2 : 1) dp_psd2_kernel(const DpPsd2Params<T> prm) {
    extern __shared__ __align__(16) unsigned char dp_smem_raw[];
    DpPsd2Kernel<T, R1, IN>::run(prm, dp_smem_raw);
}
#endif
