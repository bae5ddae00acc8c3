// Noise PSD accumulation: sum over traces of |fft(x)|^2 per frequency bin.
//
// Replaces qp.calc_psd(traces[cut], fs, folded_over=False) as called from
// Noise.calc_psd (reference detprocess/core/noise.py:344): two-sided PSD =
// mean_traces |fft(x)|^2 / (N fs).  The kernel reuses the forward half of the fused OF
// kernel (registers <-> shared-memory sub-FFTs, thread-local real-FFT untangle) and adds
// |X[k]|^2 into a per-CTA partial-sum array kept in thread order (coalesced, no atomics, L2
// resident); dp_psd_reduce_kernel folds the CTAs and maps thread order -> natural k.  The
// per-GPU sums are then all-reduced over NCCL by the host layer.
#pragma once
#include "dp_of_kernel.cuh"

template <class T> struct DpPsdParams {
    const void* traces;
    long long row_stride;
    int n_rows;
    const unsigned char* mask;  // [n_rows] 1 = use the trace (nullptr: all)
    const cx<T>* tw1;
    const cx<T>* tw2;
    const cx<T>* twn;
    const cx<T>* twp;
    cx<T>* scratch;             // [grid][32*NT] (P = 2)
    long long scratch_per_cta;
    double* partial;            // [grid][partial_per_cta]: [32*P][NT] thread order, then [17][2*P] self lanes
    long long partial_per_cta;
    unsigned long long* count;  // [grid] accepted traces per CTA
    double scale;
    int subtract_first;
};

template <class T, int R1, int P, int IN> struct DpPsdKernel {
    using G = DpGeom<R1>;
    static constexpr int NT = G::NT;
    static constexpr int N = 2 * P * G::MS;
    static constexpr size_t SMEM_BYTES = sizeof(cx<T>) * (G::SMEM_ELEMS + 64) + 64;

    static DP_DEV void run(const DpPsdParams<T>& prm, unsigned char* smem_raw) {
        cx<T>* buf = reinterpret_cast<cx<T>*>(smem_raw);
        cx<T>* sp = buf + G::SMEM_ELEMS;
        const int tid = threadIdx.x;
        int K12, bA, bB;
        G::map(tid, K12, bA, bB);
        const cx<T> wn = dp_ldg(prm.twn + tid);
        const cx<T> wp = dp_ldg(prm.twp + tid);
        cx<T>* scr0 = prm.scratch + (long long)blockIdx.x * prm.scratch_per_cta;
        double* part = prm.partial + (long long)blockIdx.x * prm.partial_per_cta;
        double* part_self = part + 32 * P * NT;
        constexpr size_t ESZ = sizeof(typename DpRaw<IN>::scalar);
        const DpSelfLane<T> sl = dp_self_lane<T, P>(tid & 31);
        const double inv_s2 = 1.0 / (4.0 * prm.scale * prm.scale);  // kernel values are 2*scale*X
        unsigned long long n_acc = 0;

        for (int row = blockIdx.x; row < prm.n_rows; row += gridDim.x) {
            if (prm.mask != nullptr && prm.mask[row] == 0) continue;  // CTA-uniform
            ++n_acc;
            const void* xrow = reinterpret_cast<const unsigned char*>(prm.traces) + (size_t)row * (size_t)prm.row_stride * ESZ;
            const double x0 = prm.subtract_first ? dp_load_first<IN>(xrow) : 0.0;
            cx<T> za[16], zb[16];
            if constexpr (P == 1) {
                dp_fwd_subfft<T, R1, 1, IN>(xrow, 0, x0, prm.scale, buf, prm.tw1, prm.tw2, bA, bB, za, zb);
                if (tid < 32) {
                    if (tid == 0) {
#pragma unroll
                        for (int r = 0; r < 16; ++r) {
                            sp[r] = za[r];
                            sp[16 + r] = zb[r];
                        }
                    }
                    __syncwarp();
                    if (tid < 17) {
                        cx<T> Xk, Xm;
                        dp_untangle(sp[sl.ek], sp[sl.em], sl.w, Xk, Xm);
                        double pk = (double)cnorm2(Xk) * inv_s2, pm = (double)cnorm2(Xm) * inv_s2;
                        if (tid == 0 && prm.subtract_first) {
                            // DC bin: put back the subtracted first sample, X[0] += N*x0 (in double)
                            const double dc = (double)Xk.re / (2.0 * prm.scale) + (double)N * x0;
                            pk = dc * dc;
                        }
                        part_self[2 * tid] += pk;
                        part_self[2 * tid + 1] += pm;
                    }
                    __syncwarp();
                }
                if (tid != 0) {
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        cx<T> Xk, Xm;
                        dp_untangle(za[r], zb[15 - r], cmul(wn, dp_w64_rt<T>(2 * r)), Xk, Xm);
                        part[r * NT + tid] += (double)cnorm2(Xk) * inv_s2;
                        part[(16 + 15 - r) * NT + tid] += (double)cnorm2(Xm) * inv_s2;
                    }
                }
                __syncthreads();  // pass-3 reads of buf precede the next trace's pass-1 stores
            } else {
#pragma unroll 1
                for (int p = 0; p < 2; ++p) {
                    dp_fwd_subfft<T, R1, 2, IN>(xrow, p, x0, prm.scale, buf, prm.tw1, prm.tw2, bA, bB, za, zb);
                    if (p == 0) {
#pragma unroll
                        for (int r = 0; r < 16; ++r) {
                            scr0[r * NT + tid] = za[r];
                            scr0[(16 + r) * NT + tid] = zb[r];
                        }
                        if (tid == 0) {
#pragma unroll
                            for (int r = 0; r < 16; ++r) {
                                sp[32 + r] = za[r];
                                sp[32 + 16 + r] = zb[r];
                            }
                        }
                    }
                    __syncthreads();
                }
                if (tid < 32) {
                    if (tid == 0) {
#pragma unroll
                        for (int r = 0; r < 16; ++r) {
                            sp[r] = za[r];
                            sp[16 + r] = zb[r];
                        }
                    }
                    __syncwarp();
                    if (tid < 17) {
                        cx<T> X[4];
                        dp_quad_x(sp[32 + sl.ek], sp[sl.ek], sp[32 + sl.em], sp[sl.em], sl.u, sl.w, X);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            double pj = (double)cnorm2(X[j]) * inv_s2;
                            if (tid == 0 && j == 0 && prm.subtract_first) {
                                const double dc = (double)X[0].re / (2.0 * prm.scale) + (double)N * x0;
                                pj = dc * dc;
                            }
                            part_self[4 * tid + j] += pj;
                        }
                    }
                    __syncwarp();
                }
                if (tid != 0) {
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const cx<T> u = cmul(wp, dp_w64_rt<T>(2 * r));
                        const cx<T> w1 = cmul(wn, dp_w64_rt<T>(r));
                        cx<T> X[4];
                        dp_quad_x(scr0[r * NT + tid], za[r], scr0[(16 + 15 - r) * NT + tid], zb[15 - r], u, w1, X);
#pragma unroll
                        for (int j = 0; j < 4; ++j) part[(4 * r + j) * NT + tid] += (double)cnorm2(X[j]) * inv_s2;
                    }
                }
            }
        }
        if (tid == 0) prm.count[blockIdx.x] += n_acc;
    }
};

// sum[k] (k = 0..N/2) = sum over CTAs of the partial at loc[k]; also total count
struct DpPsdReduceParams {
    const double* partial;
    long long partial_per_cta;
    int grid;
    const int* loc;  // [N/2 + 1] index into one CTA's partial array
    int nbins;
    double* sum_out;                    // [nbins], accumulated (+=)
    const unsigned long long* count;    // [grid]
    unsigned long long* count_out;      // [1], accumulated (+=)
};

#ifndef DP_HOST_EMU
template <class T, int R1, int P, int IN>
__global__ void __launch_bounds__(DpGeom<R1>::NT, 1) dp_psd_kernel(const DpPsdParams<T> prm) {
    extern __shared__ __align__(16) unsigned char dp_smem_raw[];
    DpPsdKernel<T, R1, P, IN>::run(prm, dp_smem_raw);
}

#endif
