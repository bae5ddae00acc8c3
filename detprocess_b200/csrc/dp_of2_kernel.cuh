// v2 fused OF1x1 kernel for nb_samples = 16384 / 32768 / 65536 (the lengths BASELINE.json
// names): same maths as dp_of_kernel.cuh, re-laid out for the measured B200 pipes
// (tools/ubench/pipes.cu; profiles/README.md):
//   * fp32 mode computes on cx<f2>: every butterfly, twiddle and filter multiply is an
//     FFMA2 / FADD2 / FMUL2 on TWO complex points (adjacent elements in passes 1-3, the
//     (k, M-k) mirror groups in pass 4), halving the issue slots per event;
//   * fp64 mode uses the same code on cx<double> (DFMA issues at the FFMA2 rate on B200);
//   * all passes have radix <= 16 (16 register-resident points per thread, no spills),
//     512 threads per CTA;
//   * the complex FFT of M = N/2 = R1*4096 points is one radix-R1 pass followed by R1
//     independent 4096-point blocks (16 x 16 x 16).  Shared memory holds NB blocks at a time
//     (128 KB); when NB < R1 the blocks are processed in mirror-closed PHASES, each phase
//     re-reading the (L2-resident) trace, untangling its own (k, M-k) pairs, and only the
//     inverse needs one L2 round trip of the parked block results.
//
// Replaces, per event: qp.OFBase.update_signal / calc_signal_filt / calc_signal_filt_td
// (reference detprocess/process/processing_data.py:763-772) and qp.OF1x1.calc /
// get_result_* (reference detprocess/core/algorithms.py:331-341, 410-421, 533-558).
//
// Index conventions.  Packed complex sequence c[n] = x[2n] + i x[2n+1], n = n1*4096 + m,
// m = n2*256 + n3*16 + n4.  Decimation in frequency, in place: after pass j digit n_j holds
// k_j and the spectrum bin is k = k1 + R1*(k2 + 16 k3 + 256 k4).  A "group" is the 16
// elements (k4 = 0..15) of one (block, k2, k3); the mirror M - k of a group element lies in
// the mirror group with k4 -> 15 - k4 (except the two self-paired groups of block 0).
#pragma once
#include "dp_f2.cuh"
#include "dp_of_kernel.cuh"

#ifndef DP2_SKEW_NS
#define DP2_SKEW_NS 600
#endif

template <class T> struct Dp2Traits;
template <> struct Dp2Traits<double> {
    using S = double;
    static constexpr int VL = 1;
};
template <> struct Dp2Traits<f2> {
    using S = float;
    static constexpr int VL = 2;
};

// read-only 16-byte load of a packed pair of complex points
DP_DEV cx<f2> dp_ldg(const cx<f2>* p) {
#ifdef DP_HOST_EMU
    return *p;
#else
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    return cx<f2>{f2(v.x, v.y), f2(v.z, v.w)};
#endif
}
DP_DEV f2 dp_ldg(const f2* p) {
#ifdef DP_HOST_EMU
    return *p;
#else
    const float2 v = __ldg(reinterpret_cast<const float2*>(p));
    return f2(v.x, v.y);
#endif
}

// ------------------------------------------------------------------- geometry
template <class T, int R1_> struct Dp2Geom {
    using S = typename Dp2Traits<T>::S;
    static constexpr int R1 = R1_;
    static constexpr int VL = Dp2Traits<T>::VL;  // complex points per register vector V = cx<T>
    static constexpr int M = 4096 * R1, N = 2 * M;
#ifndef DP2_NBMAX_F32
#define DP2_NBMAX_F32 4
#endif
    // blocks per phase: fp64 2 (128 KB of shared memory, 512 threads, one CTA per SM);
    // packed fp32 2 (64 KB, 256 threads, TWO independent CTAs per SM whose load / butterfly /
    // store phases overlap) -- 4 would be one 512-thread CTA per SM
    static constexpr int NBMAX = (VL == 2) ? DP2_NBMAX_F32 : 2;
    static constexpr int NB = R1 < NBMAX ? R1 : NBMAX;    // blocks per phase
    static constexpr int NPH = R1 / NB;                   // phases
    static constexpr int VPB = 4096 / VL;                 // V's per block
    static constexpr int NV = NB * VPB;                   // V's per phase (<= 8192)
    static constexpr int NT = NV / 16;                    // threads per CTA
    static constexpr int GV = 16 / VL;                    // V's per group
    static constexpr int CV = 256 / VL;                   // V's per 256 elements
    static constexpr int NC = VPB / NT;                   // pass-1 columns per thread
    static constexpr int GC = (16 / R1 < NC) ? 16 / R1 : NC;  // pass-1' columns handled together (<= 16 V live)
    static constexpr int SMEM_V = NV + NV / GV;           // one pad V per group: conflict-free LDS.128 everywhere
    static constexpr int KQ = M / 16;                     // bin stride between group elements (k4)
    static DP_HD int phys(int v) { return v + v / GV; }
    // spectrum-block index k1 of local block b in phase p
    static DP_HD constexpr int k1_of(int p, int b) {
        if (NPH == 1) return b;
        if (NPH == 2) return p + 2 * b;
        // NPH == 4 (R1 = 8, NB = 2): {0,4} {2,6} {1,7} {3,5}
        const int a = (p == 0) ? 0 : (p == 1) ? 2 : (p == 2) ? 1 : 3;
        return b == 0 ? a : (a == 0 ? R1 / 2 : R1 - a);
    }
    // local index of the mirror block (R1 - k1) % R1
    static DP_HD constexpr int mirror_b(int p, int b) {
        const int km = (R1 - k1_of(p, b)) % R1;
        for (int c = 0; c < NB; ++c)
            if (k1_of(p, c) == km) return c;
        return -1;
    }
    // bin of element r of group G (phase-local group id b*256 + k2*16 + k3)
    static DP_HD int bin_of(int p, int G, int r) {
        const int b = G >> 8, k2 = (G >> 4) & 15, k3 = G & 15;
        return k1_of(p, b) + R1 * (k2 + 16 * k3 + 256 * r);
    }
    // Threads are tied to blocks: in passes 2, 3 (and 3', 2') thread t works on local block t / CV.  Pass 4 and
    // the point-wise stage work on (group, mirror group) pairs; pair -> thread assignment keeps every thread
    // inside its own mirror-closed block set {b, mirror_b(b)}, so between pass 1 and pass 1' the sets only
    // ever touch their own part of the shared buffer and synchronise with NAMED barriers (block / set) --
    // the sets drift apart and overlap each other's shared-memory, FP and global-latency phases.
    //
    // i-th pair of a self-mirrored block (128 pairs) / of a swapped block pair (256 pairs): (k2,k3) of the
    // group in the lower block and of its mirror.  Pair 0 of block k1 = 0 is the special one: both groups
    // ((0,0,0) and (0,0,8)) are self-paired.
    static DP_HD void self_pair(bool k1_zero, int i, int& gA, int& gB) {
        int k2, k3, k2m, k3m;
        if (!k1_zero) {
            k2 = i >> 4, k3 = i & 15, k2m = 15 - k2, k3m = 15 - k3;
        } else if (i == 0) {
            k2 = 0, k3 = 0, k2m = 0, k3m = 8;
        } else if (i < 8) {
            k2 = 0, k3 = i, k2m = 0, k3m = 16 - i;
        } else if (i < 16) {
            k2 = 8, k3 = i - 8, k2m = 8, k3m = 15 - k3;
        } else {
            k2 = i >> 4, k3 = i & 15, k2m = 16 - k2, k3m = 15 - k3;
        }
        gA = k2 * 16 + k3;
        gB = k2m * 16 + k3m;
    }
    // groups of thread t in phase p: VL == 2: (GA, GB) lanes; VL == 1: own group and the partner's (thread t ^ 1)
    static DP_HD void groups_of(int p, int t, int& Gown, int& Gother) {
        const int b = t / CV, u = t % CV, mb = mirror_b(p, b);
        int GA, GB, side = 0, i;
        if (mb == b) {
            i = (VL == 2) ? u : (u >> 1);
            side = (VL == 2) ? 0 : (u & 1);
            int gA, gB;
            self_pair(k1_of(p, b) == 0, i, gA, gB);
            GA = b * 256 + gA;
            GB = b * 256 + gB;
        } else {
            const int lo = b < mb ? b : mb, hi = b < mb ? mb : b;
            const int o = (b == lo ? 0 : CV) + u;  // ordinal inside the two-block set
            i = (VL == 2) ? o : (o >> 1);
            side = (VL == 2) ? 0 : (o & 1);
            const int k2 = i >> 4, k3 = i & 15;
            GA = lo * 256 + k2 * 16 + k3;
            GB = hi * 256 + (15 - k2) * 16 + (15 - k3);
        }
        Gown = side ? GB : GA;
        Gother = side ? GA : GB;
    }
    // named barriers of thread t in phase p: its block (CV threads) and its mirror-closed block set
    static DP_HD int bar_block_id(int t) { return 1 + t / CV; }
    static DP_HD int bar_set_id(int p, int t) {
        const int b = t / CV, mb = mirror_b(p, b);
        return 1 + NB + (b < mb ? b : mb);
    }
    static DP_HD int bar_set_count(int p, int t) { return mirror_b(p, t / CV) == t / CV ? CV : 2 * CV; }
};

// --------------------------------------------------------------- device tables
template <class T> struct Dp2TemplDev {
    using S = typename Dp2Traits<T>::S;
    const cx<T>* phi;       // [NPH][16][NT] thread-order filter (VL == 2: lanes = (group A, group B) bins)
    const cx<S>* phi_self;  // [17][2] filter at the bins of the self-paired groups
    const cx<S>* s_low;     // [nlow] scaled template spectrum, natural order
    double norm;
    double tsum;
    int pretrigger;
    int pad_;
};

template <class T> struct Dp2ChanDev {
    using S = typename Dp2Traits<T>::S;
    const T* wj;       // [NPH][16][NT] thread-order chi0 weights
    const S* wj_self;  // [17][2]
    const S* wj_low;   // [nlow]
    double adc_gain;   // int16 traces (raw ADC counts): sample = adc * adc_gain + adc_offset
    double adc_offset;
    int n_templ;
    int n_slots;
    int out_base;
    Dp2TemplDev<T> templ[DP_MAX_TEMPLATES];
    DpSlot slots[DP_MAX_SLOTS];
};

template <class T> struct Dp2Params {
    using S = typename Dp2Traits<T>::S;
    const void* traces;
    // Input layout.  The first sample of (event ev, plan channel c) is element
    //     (row_start ? row_start[ev] : ev * event_stride) + (chan_offset ? chan_offset[c] : c * chan_stride)
    // of `traces`.  Default batches [n_events][n_chan][row_stride]: event_stride = n_chan * row_stride, chan_stride =
    // row_stride.  A reader batch [B][n_file_chan][N] of which the plan uses some channels: chan_offset = file channel *
    // N.  Window mode (row_start, below): the channels are continuous streams [n_chan][stream_stride].
    long long event_stride;  // elements
    long long chan_stride;   // elements
    const long long* chan_offset;  // [n_chan] or null
    int n_rows;
    int n_chan;
    const Dp2ChanDev<T>* chans;
    const cx<T>* tw1;      // [VPB]  exp(-2 pi i m / M), lanes m = VL*c + lane
    const cx<T>* tw2;      // [CV]   exp(-2 pi i r / 4096)
    const cx<T>* tw3;      // [GV]   exp(-2 pi i n4 / 256)
    const cx<S>* twn;      // [NPH][NT] exp(-2 pi i k(own group, 0) / N)
    const int2* groups;    // [NPH][NT] (own / A group, other / B group)
    const int* chunk3;     // [NPH][NT] phase-local 256-element chunk (b*16 + k2) the thread transforms in passes 3 / 3'
    const uint4* zones;    // [NPH][NW] byte offsets (into the FFT buffer) of the warp's four 2048-byte landing pieces
    cx<T>* scratch;        // [grid][scratch_per_cta]
    long long scratch_per_cta;
    double* out;
    int n_out;
    int nlow;
    double scale;
    int subtract_first;
    int neighbours;  // also report the amplitude one sample before / after each fit's best delay (interpolate_t0)
    int skew_ns;  // start delay of every second block (DP2_SKEW_NS unless the plan overrides it)
    // window mode: event ev is the N-sample window that starts at sample row_start[ev] of every channel's continuous
    // stream (the step between the trigger and the features in the reference, processing_data.py:643-688); windows that
    // stick out of [0, stream_len) get the -999999 sentinels
    const long long* row_start;  // [n_events] or null
    long long stream_len;
};

// --------------------------------------------------------------------- helpers
template <class T> DP_DEV cx<T> dp2_csq(cx<T> a) { return cx<T>{dp_fma(a.re, a.re, -(a.im * a.im)), (a.re + a.re) * a.im}; }

// powers w^0..w^(R-1) (only the used ones survive dead-code elimination)
template <int R, class T> DP_DEV void dp2_powers(cx<T> w, cx<T> (&pw)[R]) {
    pw[0] = cx<T>{(T)1.0f, (T)0.0f};
    if constexpr (R > 1) pw[1] = w;
    if constexpr (R > 2) pw[2] = dp2_csq(w);
    if constexpr (R > 3) pw[3] = cmul(pw[2], w);
    if constexpr (R > 4) pw[4] = dp2_csq(pw[2]);
    if constexpr (R > 5) pw[5] = cmul(pw[4], w);
    if constexpr (R > 6) pw[6] = cmul(pw[4], pw[2]);
    if constexpr (R > 7) pw[7] = cmul(pw[4], pw[3]);
}

// ---- L2-resident scratch (X spill of multi-template plans, parked block results of multi-phase
// transforms): written and re-read within one event by the same CTA.  The streaming traces would
// otherwise evict it between the write and the read (ncu: every scratch byte went to DRAM and
// back); an evict_last cache policy on these 16-byte accesses keeps it in L2.
DP_DEV unsigned long long dp2_policy_keep() {
#ifdef DP_HOST_EMU
    return 0ull;
#else
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
#endif
}
template <class V> DP_DEV void dp2_st_keep(V* ptr, const V& v, unsigned long long pol) {
    static_assert(sizeof(V) == 16, "scratch vectors are 16 bytes");
#ifdef DP_HOST_EMU
    (void)pol;
    *ptr = v;
#else
    const uint4 u = *reinterpret_cast<const uint4*>(&v);
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(ptr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w), "l"(pol)
                 : "memory");
#endif
}
template <class V> DP_DEV V dp2_ld_keep(const V* ptr, unsigned long long pol) {
    static_assert(sizeof(V) == 16, "scratch vectors are 16 bytes");
#ifdef DP_HOST_EMU
    (void)pol;
    return *ptr;
#else
    uint4 u;
    asm volatile("ld.global.L2::cache_hint.v4.b32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(ptr), "l"(pol)
                 : "memory");
    V v;
    *reinterpret_cast<uint4*>(&v) = u;
    return v;
#endif
}



// 16-byte read-only load of a table row that is used once per event (filter rows that were not staged): no L1
// allocation, so that the small per-pass tables (twiddles, group ids) keep the ~60 KB of L1 the CTA leaves
template <class V> DP_DEV V dp2_ld_stream(const V* ptr) {
    static_assert(sizeof(V) == 16, "table vectors are 16 bytes");
#if defined(DP_HOST_EMU) || defined(DP2_TAB_ALLOC)
    return dp_ldg(ptr);
#else
    uint4 u;
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(ptr));
    V v;
    *reinterpret_cast<uint4*>(&v) = u;
    return v;
#endif
}

// ---- per-warp asynchronous staging (TMA bulk copies global -> shared, completion on an mbarrier).
// The filter / chi0-weight tables of the point-wise stage (and the X column of multi-template plans) are
// fetched by ONE lane per warp into shared memory the warp owns at that moment -- the group rows it has just
// emptied into registers (pass 4) plus a small private area -- while the warp computes its radix-16
// butterflies; the loads cost no registers, no issue slots and no scoreboard stall (round 1: 19 % of all
// warp-time in the point-wise stage was `long_scoreboard` on these tables).
#define DP2_PIECE 2048  // bytes per bulk copy
struct alignas(8) Dp2Mbar {
    unsigned long long v;
};
DP_DEV void dp2_mbar_init(Dp2Mbar* bar) {
#ifndef DP_HOST_EMU
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a) : "memory");
#else
    bar->v = 0;
#endif
}
DP_DEV void dp2_mbar_init_fence() {
#ifndef DP_HOST_EMU
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
#endif
}
// generic-proxy writes (X column stores) -> visible to later bulk copies of this CTA
DP_DEV void dp2_fence_async() {
#ifndef DP_HOST_EMU
    asm volatile("fence.proxy.async;" ::: "memory");
#endif
}
// one lane: announce `bytes` on the warp's barrier (its single arrival of this round)
DP_DEV void dp2_stage_begin(Dp2Mbar* bar, unsigned bytes) {
#ifndef DP_HOST_EMU
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
#else
    (void)bar;
    (void)bytes;
#endif
}
DP_DEV void dp2_stage_copy(void* dst_smem, const void* src_global, unsigned bytes, Dp2Mbar* bar) {
#ifndef DP_HOST_EMU
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src_global),
                 "r"(bytes), "r"(a)
                 : "memory");
#else
    (void)bar;
    std::memcpy(dst_smem, src_global, bytes);
#endif
}
// all lanes: wait for the round with parity `par` (emulation: the issuing lane copied synchronously)
DP_DEV void dp2_stage_wait(Dp2Mbar* bar, unsigned par) {
#ifndef DP_HOST_EMU
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "DP2_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra DP2_WAIT_%=;\n\t}" ::"r"(a),
        "r"(par)
        : "memory");
#else
    (void)bar;
    (void)par;
    __syncwarp();
#endif
}
// exchange with the partner lane (lane ^ 1)
DP_DEV cx<double> dp2_swap1(cx<double> v) {
    return cx<double>{__shfl_xor_sync(0xffffffffu, v.re, 1), __shfl_xor_sync(0xffffffffu, v.im, 1)};
}

// streaming policy for the LAST read of a trace (evict_first) / earlier reads (evict_normal)
DP_DEV unsigned long long dp2_policy_stream(bool last_read) {
#ifdef DP_HOST_EMU
    (void)last_read;
    return 0ull;
#else
    unsigned long long a, b;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(a));
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(b));
    return last_read ? a : b;
#endif
}
template <int IN> DP_DEV typename DpRaw<IN>::type dp2_load_pair(const void* row, long long j, unsigned long long pol) {
#ifdef DP_HOST_EMU
    (void)pol;
    return dp_load_raw<IN>(row, j);
#else
    typename DpRaw<IN>::type v;
    const typename DpRaw<IN>::type* ptr = reinterpret_cast<const typename DpRaw<IN>::type*>(row) + j;
    if constexpr (IN == 0) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(ptr), "l"(pol));
    } else if constexpr (IN == 3) {
        const double* p1 = reinterpret_cast<const double*>(row) + 2 * j;
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v.x) : "l"(p1), "l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v.y) : "l"(p1 + 1), "l"(pol));
    } else if constexpr (IN == 1) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(ptr), "l"(pol));
    } else if constexpr (IN == 4) {
        const float* p1 = reinterpret_cast<const float*>(row) + 2 * j;
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v.x) : "l"(p1), "l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v.y) : "l"(p1 + 1), "l"(pol));
    } else if constexpr (IN == 5) {
        const short* p1 = reinterpret_cast<const short*>(row) + 2 * j;
        short a, b;
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b16 %0, [%1], %2;" : "=h"(a) : "l"(p1), "l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b16 %0, [%1], %2;" : "=h"(b) : "l"(p1 + 1), "l"(pol));
        v.x = a;
        v.y = b;
    } else {
        unsigned u;
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(u) : "l"(ptr), "l"(pol));
        v.x = (short)(u & 0xffffu);
        v.y = (short)(u >> 16);
    }
    return v;
#endif
}

// lanes of a packed complex pair
DP_DEV cx<float> dp2_lane0(cx<f2> z) { return cx<float>{z.re.x, z.im.x}; }
DP_DEV cx<float> dp2_lane1(cx<f2> z) { return cx<float>{z.re.y, z.im.y}; }
DP_DEV void dp2_set0(cx<f2>& z, cx<float> v) { z.re.x = v.re, z.im.x = v.im; }
DP_DEV void dp2_set1(cx<f2>& z, cx<float> v) { z.re.y = v.re, z.im.y = v.im; }

// ---- trace loads: V (n1, c) = complex points n = n1*4096 + VL*c + lane = samples 2n, 2n+1
template <int IN, int VL> struct Dp2Raw {
    typename DpRaw<IN>::type q[VL];
};
template <int IN, int VL> DP_DEV Dp2Raw<IN, VL> dp2_load_raw(const void* row, int n1, int c, unsigned long long pol) {
    Dp2Raw<IN, VL> r;
    const long long j0 = (long long)n1 * 4096 + (long long)VL * c;
#pragma unroll
    for (int l = 0; l < VL; ++l) r.q[l] = dp2_load_pair<IN>(row, j0 + l, pol);
    return r;
}
// same, for a chunk of a continuous stream that may stick out of [0, n_pairs): `row` is the
// stream base, jbase the (possibly negative) pair index of the chunk start; out-of-range
// pairs read as zero (zero-padded linear convolution, like scipy's oaconvolve)
template <int IN, int VL>
DP_DEV Dp2Raw<IN, VL> dp2_load_raw_clamped(const void* row, long long jbase, long long n_samples, int n1, int c, unsigned long long pol) {
    Dp2Raw<IN, VL> r;
    const long long j0 = jbase + (long long)n1 * 4096 + (long long)VL * c;
#pragma unroll
    for (int l = 0; l < VL; ++l) {
        const long long j = j0 + l;
        typename DpRaw<IN>::type v;
        v.x = 0;
        v.y = 0;
        if (j >= 0 && 2 * j + 1 < n_samples) {
            v = dp2_load_pair<IN>(row, j, pol);
        } else if (j >= 0 && 2 * j < n_samples) {  // last sample of an odd-length stream
            v.x = reinterpret_cast<const typename DpRaw<IN>::scalar*>(row)[2 * j];
        }
        r.q[l] = v;
    }
    return r;
}
// fp64 mode computes on the samples as they are (the plans never scale or offset them); raw ADC counts (IN == 2)
// become adc * sc - x0 with the channel's conversion (sc = gain, x0 = -offset)
template <int IN> DP_DEV cx<double> dp2_convert(const Dp2Raw<IN, 1>& r, double x0, double sc) {
    if constexpr (dp_in_is_adc(IN)) {
        return cx<double>{dp_fma((double)r.q[0].x, sc, -x0), dp_fma((double)r.q[0].y, sc, -x0)};
    } else {
        (void)x0;
        (void)sc;
        return cx<double>{(double)r.q[0].x, (double)r.q[0].y};
    }
}
// fp32 mode: the first sample is removed in float64 (AC coupling; keeps the fp32 mantissa for the signal),
// the power-of-two scale is applied after the conversion as one packed multiply
template <int IN> DP_DEV cx<f2> dp2_convert(const Dp2Raw<IN, 2>& r, double x0, double sc) {
    const f2 s = f2((float)sc);
    return cx<f2>{f2((float)((double)r.q[0].x - x0), (float)((double)r.q[1].x - x0)) * s,
                  f2((float)((double)r.q[0].y - x0), (float)((double)r.q[1].y - x0)) * s};
}

// tie-aware running best (|val| larger, or equal and smaller index)
template <class S> DP_DEV void dp2_consider(DpBest<S>& b, S kabs, S v, int idx) {
    const S kb = dp_abs(b.val);  // (an empty best holds val = 0, idx = -1; a NaN candidate is never taken)
    if (kabs > kb || (kabs == kb && (b.idx < 0 || idx < b.idx))) {
        b.val = v;
        b.idx = idx;
    }
}


// exp(+2 pi i n / 16): output-row rotation of the pruned pass 2' (Dp2Core::inv_2_rows)
#ifdef DP_HOST_EMU
static const double dp2_w16_tab[16][2] = {
#else
static __constant__ double dp2_w16_tab[16][2] = {
#endif
    {1.0, 0.0},
    {0.92387953251128675613, 0.38268343236508977173},
    {0.70710678118654752440, 0.70710678118654752440},
    {0.38268343236508977173, 0.92387953251128675613},
    {0.0, 1.0},
    {-0.38268343236508977173, 0.92387953251128675613},
    {-0.70710678118654752440, 0.70710678118654752440},
    {-0.92387953251128675613, 0.38268343236508977173},
    {-1.0, 0.0},
    {-0.92387953251128675613, -0.38268343236508977173},
    {-0.70710678118654752440, -0.70710678118654752440},
    {-0.38268343236508977173, -0.92387953251128675613},
    {0.0, -1.0},
    {0.38268343236508977173, -0.92387953251128675613},
    {0.70710678118654752440, -0.70710678118654752440},
    {0.92387953251128675613, -0.38268343236508977173}};

// ---- tensor memory as a thread-private parking space (fp64 kernels whose transform is split into phases).
// Every phase of the split transform needs ALL samples of the trace for its radix-8 first pass; round 1 re-read the trace
// in every phase (4 x 512 KB through the SM's L2 port: 45 % of the kernel waiting on those loads, profiles/
// r2_prof_psd2_f64_64k_before.txt).  The first pass of a column is thread-private -- thread t handles columns t + i NT
// in every phase -- so phase 0 computes all eight outputs of its columns once, keeps its own two blocks in shared memory
// and parks the two blocks of each of the next two phases in TMEM: 256 KB per SM that this kernel (no tensor-core work)
// would otherwise leave idle, written at 256 B/clk and read back at 64 B/clk with ~12 clk latency.  A warp reaches the
// 32 lanes of its quarter (warp % 4); the four warps of a quarter take 128 columns each = 32 cx<double> per thread =
// 2 phases x 8 columns x 2 blocks.  The last phase (TMEM is full) still reads the trace.
#if !defined(DP_HOST_EMU)
DP_DEV void dp_tmem_alloc512(unsigned* smem_slot) {  // one whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((unsigned)__cvta_generic_to_shared(smem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
DP_DEV void dp_tmem_dealloc512(unsigned addr) {      // one whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(addr) : "memory");
}
DP_DEV void dp_tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DP_DEV void dp_tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
DP_DEV void dp_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
DP_DEV void dp_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// one register vector V (cx<double> or packed cx<f2>: 16 bytes) <-> four 32-bit TMEM columns
DP_DEV void dp_tmem_words(const cx<double>& a, int (&w)[4]) {
    w[0] = __double2loint(a.re), w[1] = __double2hiint(a.re), w[2] = __double2loint(a.im), w[3] = __double2hiint(a.im);
}
DP_DEV void dp_tmem_words(const cx<f2>& a, int (&w)[4]) {
    w[0] = __float_as_int(a.re.x), w[1] = __float_as_int(a.re.y), w[2] = __float_as_int(a.im.x), w[3] = __float_as_int(a.im.y);
}
// two vectors (8 x 32 bit) of this thread's lane at column `addr`
template <class V> DP_DEV void dp_tmem_st2(unsigned addr, const V& a, const V& b) {
    int x[4], y[4];
    dp_tmem_words(a, x);
    dp_tmem_words(b, y);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr), "r"(x[0]), "r"(x[1]), "r"(x[2]),
                 "r"(x[3]), "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3])
                 : "memory");
}
// the load is asynchronous: dp_tmem_wait_ld() before the first use of the registers
struct DpTmemRaw8 { int r[8]; };
DP_DEV DpTmemRaw8 dp_tmem_ld8(unsigned addr) {
    DpTmemRaw8 t;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(t.r[0]), "=r"(t.r[1]), "=r"(t.r[2]), "=r"(t.r[3]), "=r"(t.r[4]), "=r"(t.r[5]), "=r"(t.r[6]), "=r"(t.r[7])
                 : "r"(addr)
                 : "memory");
    return t;
}
DP_DEV cx<double> dp_tmem_cx(const DpTmemRaw8& t, int j) {
    return cx<double>{__hiloint2double(t.r[4 * j + 1], t.r[4 * j]), __hiloint2double(t.r[4 * j + 3], t.r[4 * j + 2])};
}
template <class V> DP_DEV V dp_tmem_get(const DpTmemRaw8& t, int j);
template <> DP_DEV cx<double> dp_tmem_get<cx<double>>(const DpTmemRaw8& t, int j) { return dp_tmem_cx(t, j); }
template <> DP_DEV cx<f2> dp_tmem_get<cx<f2>>(const DpTmemRaw8& t, int j) {
    return cx<f2>{f2(__int_as_float(t.r[4 * j]), __int_as_float(t.r[4 * j + 1])), f2(__int_as_float(t.r[4 * j + 2]), __int_as_float(t.r[4 * j + 3]))};
}
#endif


// ======================================================================== core
template <class T, int R1, int IN> struct Dp2Core {
    using G = Dp2Geom<T, R1>;
    using S = typename G::S;
    using V = cx<T>;
    static constexpr int NT = G::NT, VL = G::VL, NB = G::NB, NPH = G::NPH, VPB = G::VPB, GV = G::GV, CV = G::CV, NC = G::NC;
    // physical (padded) strides in V units: one group, 256 elements, one block
    static constexpr int PG = GV + 1, PC = CV + CV / GV, PB = VPB + VPB / GV;

    // ---- pass 1 of phase PH: global -> radix-R1 over n1 (only the phase's blocks) -> twiddle -> smem
    // CLAMP: `row` is a stream base of n_samples samples and jbase the (possibly negative) pair index of
    // the chunk start (dp2_load_raw_clamped)
    template <int PH, bool CLAMP = false>
    static DP_DEV void pass1(const void* row, double x0, double sc, V* buf, const V* DP_RESTRICT tw1, long long jbase = 0,
                             long long jmax = 0) {
        const int tid = threadIdx.x;
        const unsigned long long pol = dp2_policy_stream(PH == NPH - 1);  // the last phase's read is the trace's last use
        // columns loaded back to back: <= 64 registers of raw float64 samples in flight
        constexpr int CBW = ((VL == 2) ? 8 : 16) / R1;
        constexpr int CB = CBW < 1 ? 1 : (CBW > NC ? NC : CBW);
        // twiddle of column c = tid + j*NT: tw1[tid] * exp(-2 pi i j / (16*NPH)) -- ONE table value per thread and phase
        // (8 KB per CTA: stays in L1) times a compile-time constant instead of a 64 KB table streamed through L1 / L2.
        // Measured (C2, 32768 samples): packed fp32 +3 %, fp64 -2 % (four more FP64 instructions per column); the product
        // costs one more rounding of every pass-1 twiddle, which the fp32 mode cannot spare -> off.
#ifndef DP2_TW1_REG
#define DP2_TW1_REG 0
#endif
        constexpr bool TW1_REG = DP2_TW1_REG;
        [[maybe_unused]] V twb;
        if constexpr (TW1_REG) twb = dp_ldg(tw1 + tid);
#pragma unroll
        for (int i0 = 0; i0 < NC; i0 += CB) {
            Dp2Raw<IN, VL> raw[CB][R1];
#pragma unroll
            for (int i = 0; i < CB; ++i)
#pragma unroll
                for (int n = 0; n < R1; ++n)
                    raw[i][n] = CLAMP ? dp2_load_raw_clamped<IN, VL>(row, jbase, jmax, n, tid + (i0 + i) * NT, pol)
                                      : dp2_load_raw<IN, VL>(row, n, tid + (i0 + i) * NT, pol);
#pragma unroll
            for (int i = 0; i < CB; ++i) {
                const int c = tid + (i0 + i) * NT;
                V v[R1];
#pragma unroll
                for (int n = 0; n < R1; ++n) v[n] = dp2_convert<IN>(raw[i][n], x0, sc);
                dp_dft<R1, -1, T>::run(v);
                V pw[R1];
                if constexpr (TW1_REG) {
                    constexpr int STEP64 = 4 / NPH;   // exp(-2 pi i VL NT j / M) = W_64^(j * STEP64)
                    const int j = i0 + i;
                    const V wc = (j == 0) ? twb : cmul(twb, V{T((S)dp_cos64(j * STEP64)), T((S)(-dp_sin64(j * STEP64)))});
                    dp2_powers<R1, T>(wc, pw);
                } else {
                    dp2_powers<R1, T>(dp_ldg(tw1 + c), pw);
                }
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const int k1 = G::k1_of(PH, b);
                    buf[G::phys(c) + b * PB] = (k1 == 0) ? v[k1] : cmul(v[k1], pw[k1]);
                }
            }
        }
    }
    template <bool CLAMP = false>
    static DP_DEV void pass1_any(int p, const void* row, double x0, double sc, V* buf, const V* DP_RESTRICT tw1, long long jbase = 0,
                                 long long jmax = 0) {
        if constexpr (NPH == 1) {
            pass1<0, CLAMP>(row, x0, sc, buf, tw1, jbase, jmax);
        } else if constexpr (NPH == 2) {
            if (p == 0)
                pass1<0, CLAMP>(row, x0, sc, buf, tw1, jbase, jmax);
            else
                pass1<1, CLAMP>(row, x0, sc, buf, tw1, jbase, jmax);
        } else {
            switch (p) {
                case 0: pass1<0, CLAMP>(row, x0, sc, buf, tw1, jbase, jmax); break;
                case 1: pass1<1, CLAMP>(row, x0, sc, buf, tw1, jbase, jmax); break;
                case 2: pass1<2, CLAMP>(row, x0, sc, buf, tw1, jbase, jmax); break;
                default: pass1<3, CLAMP>(row, x0, sc, buf, tw1, jbase, jmax); break;
            }
        }
    }

    // ---- first pass computed ONCE per event (fp64, multi-phase geometries): thread t handles the same columns t + i NT in
    // every phase, so phase 0 runs the radix-R1 butterfly of its columns, keeps its own blocks in shared memory and parks
    // the blocks of the later phases where the thread finds them again: phases 1 .. TM_PHASES in its TMEM columns
    // [((ph-1) * NC + i) * 4 NB, +4 NB), a further phase (R1 = 8: phase 3) in an L2-resident scratch row of the CTA.
    // tm = this thread's TMEM address (dp_tmem_thread_base).  Replaces NPH reads of the trace + NPH butterflies by one.
    static constexpr bool CAN_PARK = (NPH > 1) && (NT == 512) && (NB % 2 == 0);   // fp64 at 32768 / 65536, packed fp32 at 65536 samples
    static constexpr int TM_PHASES = (NPH - 1) < 128 / (4 * NB * NC) ? (NPH - 1) : 128 / (4 * NB * NC);  // phases parked in TMEM
    static constexpr int TM_COLS = TM_PHASES * NC * NB * 4;             // TMEM columns used per thread (of 128)
    static constexpr long long PARK1_V = (NPH - 1 - TM_PHASES) > 0 ? (long long)(NPH - 1 - TM_PHASES) * NB * VPB : 0;
    template <bool CLAMP = false>
    static DP_DEV void pass1_all(const void* row, double x0, double sc, V* buf, const V* DP_RESTRICT tw1, unsigned tm, V* park,
                                 long long jbase = 0, long long jmax = 0) {
#ifndef DP_HOST_EMU
        if constexpr (CAN_PARK) {
            const int tid = threadIdx.x;
            const unsigned long long pol = dp2_policy_stream(true);  // the trace's only read
            [[maybe_unused]] const unsigned long long keep = dp2_policy_keep();
            constexpr int CBW = ((VL == 2) ? 8 : 16) / R1;
            constexpr int CB = CBW < 1 ? 1 : (CBW > NC ? NC : CBW);
#pragma unroll
            for (int i0 = 0; i0 < NC; i0 += CB) {
                Dp2Raw<IN, VL> raw[CB][R1];
#pragma unroll
                for (int i = 0; i < CB; ++i)
#pragma unroll
                    for (int n = 0; n < R1; ++n)
                        raw[i][n] = CLAMP ? dp2_load_raw_clamped<IN, VL>(row, jbase, jmax, n, tid + (i0 + i) * NT, pol)
                                          : dp2_load_raw<IN, VL>(row, n, tid + (i0 + i) * NT, pol);
#pragma unroll
                for (int i = 0; i < CB; ++i) {
                    const int c = tid + (i0 + i) * NT;
                    V v[R1];
#pragma unroll
                    for (int n = 0; n < R1; ++n) v[n] = dp2_convert<IN>(raw[i][n], x0, sc);
                    dp_dft<R1, -1, T>::run(v);
                    V pw[R1];
                    dp2_powers<R1, T>(dp_ldg(tw1 + c), pw);
#pragma unroll
                    for (int ph = 0; ph < NPH; ++ph) {
                        V o[NB];
#pragma unroll
                        for (int b = 0; b < NB; ++b) {
                            const int k1 = G::k1_of(ph, b);
                            o[b] = (k1 == 0) ? v[k1] : cmul(v[k1], pw[k1]);
                        }
                        if (ph == 0) {
#pragma unroll
                            for (int b = 0; b < NB; ++b) buf[G::phys(c) + b * PB] = o[b];
                        } else if (ph <= TM_PHASES) {
#pragma unroll
                            for (int b = 0; b < NB; b += 2)
                                dp_tmem_st2(tm + (unsigned)(((ph - 1) * NC + (i0 + i)) * 4 * NB + 4 * b), o[b], o[b + 1]);
                        } else {
#pragma unroll
                            for (int b = 0; b < NB; ++b) dp2_st_keep(park + (long long)((ph - 1 - TM_PHASES) * NB + b) * VPB + c, o[b], keep);
                        }
                    }
                }
            }
            dp_tmem_wait_st();
        }
#endif
    }
    // phase p >= 1: the parked first-pass outputs -> shared memory
    static DP_DEV void pass1_fetch(int p, V* buf, unsigned tm, const V* park) {
#ifndef DP_HOST_EMU
        if constexpr (CAN_PARK) {
            const int tid = threadIdx.x;
            if (p <= TM_PHASES) {
                constexpr int NL = NC * NB / 2;      // 8-column loads of this phase
                constexpr int LB = NL < 4 ? NL : 4;  // loads in flight before one wait (32 registers)
#pragma unroll
                for (int l0 = 0; l0 < NL; l0 += LB) {
                    DpTmemRaw8 t[LB];
#pragma unroll
                    for (int l = 0; l < LB; ++l) t[l] = dp_tmem_ld8(tm + (unsigned)((p - 1) * NC * 4 * NB + (l0 + l) * 8));
                    dp_tmem_wait_ld();
#pragma unroll
                    for (int l = 0; l < LB; ++l) {
                        const int i = (l0 + l) / (NB / 2), b = 2 * ((l0 + l) % (NB / 2));
                        const int c = tid + i * NT;
                        buf[G::phys(c) + b * PB] = dp_tmem_get<V>(t[l], 0);
                        buf[G::phys(c) + (b + 1) * PB] = dp_tmem_get<V>(t[l], 1);
                    }
                }
            } else {
                const unsigned long long keep = dp2_policy_keep();
                const V* src = park + (long long)(p - 1 - TM_PHASES) * NB * VPB;
                V t[NC * NB];
#pragma unroll
                for (int i = 0; i < NC; ++i)
#pragma unroll
                    for (int b = 0; b < NB; ++b) t[i * NB + b] = dp2_ld_keep(src + (long long)b * VPB + tid + i * NT, keep);
#pragma unroll
                for (int i = 0; i < NC; ++i)
#pragma unroll
                    for (int b = 0; b < NB; ++b) buf[G::phys(tid + i * NT) + b * PB] = t[i * NB + b];
            }
        }
#endif
    }
    // TMEM address of this thread's 128 private columns: the warp's lane quarter, one column range per warp of the quarter
    static DP_DEV unsigned tm_thread_base(unsigned tm_base) {
        const int warp = threadIdx.x >> 5;
        return tm_base + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)(128 * (warp >> 2));
    }

    // ---- passes 2..4 (caller has synchronised after pass 1); result: pass-4 outputs of the
    // thread's group(s) in z[16] (VL == 2: lane 0 = group GA, lane 1 = group GB)
    static DP_DEV void fwd_234(V* buf, const V* DP_RESTRICT tw2, const V* DP_RESTRICT tw3, int GA, int GB, V (&z)[16], int p) {
        const int tid = threadIdx.x;
        {
            const int b = tid / CV, cc = tid % CV;
            V* pb = buf + G::phys(b * VPB + cc);
#pragma unroll
            for (int n = 0; n < 16; ++n) z[n] = pb[n * PC];
            dp_dft<16, -1, T>::run(z);
            dp_twiddle<16, false, T>(z, dp_ldg(tw2 + cc));
#pragma unroll
            for (int k = 0; k < 16; ++k) pb[k * PC] = z[k];
        }
        dp_bar_sync(G::bar_block_id(tid), CV);  // pass 3 reads what the threads of this block wrote
        {
            const int q = tid % GV, k2 = (tid / GV) % 16, b = tid / CV;
            V* pb = buf + G::phys(b * VPB + k2 * CV + q);
#pragma unroll
            for (int n = 0; n < 16; ++n) z[n] = pb[n * PG];
            dp_dft<16, -1, T>::run(z);
            dp_twiddle<16, false, T>(z, dp_ldg(tw3 + q));
#pragma unroll
            for (int k = 0; k < 16; ++k) pb[k * PG] = z[k];
        }
        dp_bar_sync(G::bar_set_id(p, tid), G::bar_set_count(p, tid));  // pass 4 reads the thread's block and its mirror block
        load_groups(buf, GA, GB, z);
        dp_dft<16, -1, T>::run(z);
    }

    // group element r of GA (and GB) -> z[r] (lanes)
    static DP_DEV void load_groups(const V* buf, int GA, int GB, V (&z)[16]) {
        if constexpr (VL == 1) {
            (void)GB;
#pragma unroll
            for (int n = 0; n < 16; ++n) z[n] = buf[GA * PG + n];
        } else {
            // canonical layout packs adjacent elements (2j, 2j+1); pass 4 wants (A[r], B[r]) lanes
            const V* pa = buf + GA * PG;
            const V* pb = buf + GB * PG;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const V a = pa[j], b = pb[j];
                z[2 * j] = V{f2(a.re.x, b.re.x), f2(a.im.x, b.im.x)};
                z[2 * j + 1] = V{f2(a.re.y, b.re.y), f2(a.im.y, b.im.y)};
            }
        }
    }
    static DP_DEV void store_groups(V* buf, int GA, int GB, const V (&z)[16]) {
        if constexpr (VL == 1) {
            (void)GB;
#pragma unroll
            for (int n = 0; n < 16; ++n) buf[GA * PG + n] = z[n];
        } else {
            V* pa = buf + GA * PG;
            V* pb = buf + GB * PG;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                pa[j] = V{f2(z[2 * j].re.x, z[2 * j + 1].re.x), f2(z[2 * j].im.x, z[2 * j + 1].im.x)};
                pb[j] = V{f2(z[2 * j].re.y, z[2 * j + 1].re.y), f2(z[2 * j].im.y, z[2 * j + 1].im.y)};
            }
        }
    }

    // ---- inverse passes 4', 3', 2': consumes group values z; leaves the pass-2' outputs of
    // column (b, cc) in z[n2] (V index b*VPB + n2*CV + cc).  No trailing barrier.
    static DP_DEV void inv_432(V* buf, const V* DP_RESTRICT tw2, const V* DP_RESTRICT tw3, int GA, int GB, V (&z)[16], int p) {
        const int tid = threadIdx.x;
        dp_dft<16, +1, T>::run(z);
        store_groups(buf, GA, GB, z);
        dp_bar_sync(G::bar_set_id(p, tid), G::bar_set_count(p, tid));
        {
            const int q = tid % GV, k2 = (tid / GV) % 16, b = tid / CV;
            V* pb = buf + G::phys(b * VPB + k2 * CV + q);
#pragma unroll
            for (int k = 0; k < 16; ++k) z[k] = pb[k * PG];
            dp_twiddle<16, true, T>(z, dp_ldg(tw3 + q));
            dp_dft<16, +1, T>::run(z);
#pragma unroll
            for (int n = 0; n < 16; ++n) pb[n * PG] = z[n];
        }
        dp_bar_sync(G::bar_block_id(tid), CV);
        {
            const int b = tid / CV, cc = tid % CV;
            const V* pb = buf + G::phys(b * VPB + cc);
#pragma unroll
            for (int k = 0; k < 16; ++k) z[k] = pb[k * PC];
            dp_twiddle<16, true, T>(z, dp_ldg(tw2 + cc));
            dp_dft<16, +1, T>::run(z);
        }
    }

    // ---- warp-local variant of passes 2..4 / 4'..2' (OF kernel).  Pass 3 of a 256-element chunk (fixed block,
    // k2) is done by the warp that owns the chunk's 16 groups in pass 4 -- a warp owns whole chunks: a chunk
    // and its mirror chunk -- so pass 3 -> pass 4 and pass 4' -> pass 3' only need __syncwarp(), and between the
    // set barrier after pass 2 and the one before pass 2' every warp runs on its own (pass 3, pass 4, the
    // point-wise stage with its staged tables, pass 4', pass 3'): the warps drift apart and overlap each
    // other's shared-memory, FP and wait phases instead of meeting at a barrier after every pass.
    // `chunk` = phase-local chunk id (b*16 + k2) of this thread's sub-warp, from Dp2Params::chunk3.
    static DP_DEV void fwd_2(V* buf, const V* DP_RESTRICT tw2, V (&z)[16]) {
        const int tid = threadIdx.x;
        const int b = tid / CV, cc = tid % CV;
        V* pb = buf + G::phys(b * VPB + cc);
#pragma unroll
        for (int n = 0; n < 16; ++n) z[n] = pb[n * PC];
        dp_dft<16, -1, T>::run(z);
        dp_twiddle<16, false, T>(z, dp_ldg(tw2 + cc));
#pragma unroll
        for (int k = 0; k < 16; ++k) pb[k * PC] = z[k];
    }
    static DP_DEV void fwd_3w(V* buf, const V* DP_RESTRICT tw3, int chunk, V (&z)[16]) {
        const int q = threadIdx.x % GV;
        V* pb = buf + G::phys(chunk * CV + q);
#pragma unroll
        for (int n = 0; n < 16; ++n) z[n] = pb[n * PG];
        dp_dft<16, -1, T>::run(z);
        dp_twiddle<16, false, T>(z, dp_ldg(tw3 + q));
#pragma unroll
        for (int k = 0; k < 16; ++k) pb[k * PG] = z[k];
    }
    static DP_DEV void inv_3w(V* buf, const V* DP_RESTRICT tw3, int chunk, V (&z)[16]) {
        const int q = threadIdx.x % GV;
        V* pb = buf + G::phys(chunk * CV + q);
#pragma unroll
        for (int k = 0; k < 16; ++k) z[k] = pb[k * PG];
        dp_twiddle<16, true, T>(z, dp_ldg(tw3 + q));
        dp_dft<16, +1, T>::run(z);
#pragma unroll
        for (int n = 0; n < 16; ++n) pb[n * PG] = z[n];
    }
    static DP_DEV void inv_2(const V* buf, const V* DP_RESTRICT tw2, V (&z)[16]) {
        const int tid = threadIdx.x;
        const int b = tid / CV, cc = tid % CV;
        const V* pb = buf + G::phys(b * VPB + cc);
#pragma unroll
        for (int k = 0; k < 16; ++k) z[k] = pb[k * PC];
        dp_twiddle<16, true, T>(z, dp_ldg(tw2 + cc));
        dp_dft<16, +1, T>::run(z);
    }
    // pass 2' for the one or two output rows n2 a narrow delay window needs (rowmask, CTA-uniform): out[n] = sum_k in[k] u^k with
    // u = conj(w) exp(2 pi i n / 16), evaluated by Horner's rule straight from the 16 inputs (4 FMA per step instead of
    // the twiddle powers + radix-16 butterflies of the full pass), and stored where the full pass would have put it:
    // in place (last phase) or in the parked block results (PARK).
    template <bool PARK> static DP_DEV void inv_2_rows(V* buf, V* scr, int p, const V* DP_RESTRICT tw2, unsigned rowmask) {
        const int tid = threadIdx.x;
        const int b = tid / CV, cc = tid % CV;
        V* pb = buf + G::phys(b * VPB + cc);
        V in[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) in[k] = pb[k * PC];
        V w = dp_ldg(tw2 + cc);
        w.im = -w.im;
        [[maybe_unused]] V* dst = scr + (long long)(p * NB + b) * VPB + cc;
        [[maybe_unused]] const unsigned long long pol = dp2_policy_keep();
        auto root = [&](int n) {
            const T orr = (T)(S)dp2_w16_tab[n][0], oi = (T)(S)dp2_w16_tab[n][1];
            return V{dp_fma(w.re, orr, -(w.im * oi)), dp_fma(w.re, oi, w.im * orr)};
        };
        auto put = [&](int n, const V& acc) {
            if constexpr (PARK)
                dp2_st_keep(dst + n * CV, acc, pol);
            else
                pb[n * PC] = acc;
        };
        const int n0 = __ffs((int)rowmask) - 1;
        const unsigned rest = rowmask & (rowmask - 1);
#ifndef DP2_HORNER_SEQ
        if (rest != 0 && (rest & (rest - 1)) == 0) {
            // two rows (the usual case of a narrow window): the two Horner chains are independent -- interleaved they
            // hide each other's FMA latency
            const int n1 = __ffs((int)rest) - 1;
            const V u0 = root(n0), u1 = root(n1);
            V a0 = in[15], a1 = in[15];
#pragma unroll
            for (int k = 14; k >= 0; --k) {
                a0 = V{dp_fma(a0.re, u0.re, dp_fma(-a0.im, u0.im, in[k].re)), dp_fma(a0.re, u0.im, dp_fma(a0.im, u0.re, in[k].im))};
                a1 = V{dp_fma(a1.re, u1.re, dp_fma(-a1.im, u1.im, in[k].re)), dp_fma(a1.re, u1.im, dp_fma(a1.im, u1.re, in[k].im))};
            }
            put(n0, a0);
            put(n1, a1);
            return;
        }
#endif
#pragma unroll 1
        for (unsigned m = rowmask; m != 0; m &= m - 1) {
            const int n = __ffs((int)m) - 1;
            const V u = root(n);
            V acc = in[15];
#pragma unroll
            for (int k = 14; k >= 0; --k)
                acc = V{dp_fma(acc.re, u.re, dp_fma(-acc.im, u.im, in[k].re)), dp_fma(acc.re, u.im, dp_fma(acc.im, u.re, in[k].im))};
            put(n, acc);
        }
        (void)n0;
    }
    // pass-2' outputs -> smem (own positions, in place)
    static DP_DEV void store_pass2(V* buf, const V (&z)[16]) {
        const int tid = threadIdx.x;
        const int b = tid / CV, cc = tid % CV;
        V* pb = buf + G::phys(b * VPB + cc);
#pragma unroll
        for (int n = 0; n < 16; ++n) pb[n * PC] = z[n];
    }
    // pass-2' outputs of a non-final phase -> parked block results in scratch [p*NB + b][VPB]
    static DP_DEV void park_pass2(V* scr, int p, const V (&z)[16]) {
        const int tid = threadIdx.x;
        const int b = tid / CV, cc = tid % CV;
        V* dst = scr + (long long)(p * NB + b) * VPB + cc;
        const unsigned long long pol = dp2_policy_keep();
#pragma unroll
        for (int n = 0; n < 16; ++n) dp2_st_keep(dst + n * CV, z[n], pol);
    }
    // the same two stores restricted to the rows n2 (256 consecutive complex points each) that hold a candidate delay of
    // some fit (CTA-uniform mask; a +-500-sample window needs 2 of the 16 rows): the other rows are never read by pass 1'
    static DP_DEV void store_pass2(V* buf, const V (&z)[16], unsigned rowmask) {
        const int tid = threadIdx.x;
        const int b = tid / CV, cc = tid % CV;
        V* pb = buf + G::phys(b * VPB + cc);
#pragma unroll
        for (int n = 0; n < 16; ++n)
            if (rowmask & (1u << n)) pb[n * PC] = z[n];
    }
    static DP_DEV void park_pass2(V* scr, int p, const V (&z)[16], unsigned rowmask) {
        const int tid = threadIdx.x;
        const int b = tid / CV, cc = tid % CV;
        V* dst = scr + (long long)(p * NB + b) * VPB + cc;
        const unsigned long long pol = dp2_policy_keep();
#pragma unroll
        for (int n = 0; n < 16; ++n)
            if (rowmask & (1u << n)) dp2_st_keep(dst + n * CV, z[n], pol);
    }
    // rows of the complex points [nlo, nhi] (cyclic in the 4096-point block)
    static DP_DEV unsigned rows_of(int nlo, int nhi) {
        if (nhi < nlo) return 0u;
        if (nhi - nlo >= 4095) return 0xffffu;
        const int ra = (nlo & 4095) >> 8, rb = (nhi & 4095) >> 8;
        const unsigned upto_b = (2u << rb) - 1u, from_a = 0xffffu & ~((1u << ra) - 1u);
        return ((nlo & 4095) <= (nhi & 4095)) ? (upto_b & from_a) : (upto_b | from_a);
    }
    // ---- pass 1' for columns [i0, i0 + GC): y[i*R1 + n1] = c'[n1*4096 + VL*(tid + (i0+i)*NT) + lane].
    // Only columns that hold a complex point n in [nlo, nhi] are computed (constrained delay windows
    // touch ~1/8 of the columns); the others keep stale registers, which no window scan selects.
    // Returns the mask of the columns it computed.
    // columns (bit j = column tid + j*NT, j < NC) of this thread that hold a complex point of [nlo, nhi]: a column holds
    // the points n1*4096 + m, m = VL*c .. VL*c + VL - 1, of every n1, so it is needed when m falls into the cyclic
    // interval [nlo, nhi] mod 4096 (CTA-uniform bounds, two compares per column)
    static DP_DEV unsigned need_cols(int nlo, int nhi) {
        const int tid = threadIdx.x;
        if (nhi < nlo) return 0u;
        if (nhi - nlo >= 4095) return (1u << NC) - 1u;
        const int ma = nlo & 4095, mb = nhi & 4095;
        const bool wrap = ma > mb;
        unsigned m = 0;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const int c = tid + j * NT;
            const int m0 = VL * c, m1 = VL * c + VL - 1;
            const bool need = wrap ? (m1 >= ma || m0 <= mb) : (m1 >= ma && m0 <= mb);
            m |= need ? (1u << j) : 0u;
        }
        return m;
    }
    // pass 1' of the columns i0 + i, i < GC, whose bit i is set in `cols`; returns `cols`
    static DP_DEV unsigned inv_pass1m(const V* buf, const V* scr, const V* DP_RESTRICT tw1, int i0, V (&y)[G::GC * R1], unsigned cols) {
        const int tid = threadIdx.x;
        constexpr int LP = NPH - 1;
#pragma unroll
        for (int i = 0; i < G::GC; ++i) {
            if (!(cols & (1u << i))) continue;
            const int c = tid + (i0 + i) * NT;
            V u[R1];
#pragma unroll
            for (int b = 0; b < NB; ++b) u[G::k1_of(LP, b)] = buf[G::phys(c) + b * PB];
#pragma unroll
            for (int p = 0; p < LP; ++p)
#pragma unroll
                for (int b = 0; b < NB; ++b) u[G::k1_of(p, b)] = dp2_ld_keep(scr + (long long)(p * NB + b) * VPB + c, dp2_policy_keep());
            V w = dp_ldg(tw1 + c);
            w.im = -w.im;
            V pw[R1];
            dp2_powers<R1, T>(w, pw);
#pragma unroll
            for (int k = 1; k < R1; ++k) u[k] = cmul(u[k], pw[k]);
            dp_dft<R1, +1, T>::run(u);
#pragma unroll
            for (int n = 0; n < R1; ++n) y[i * R1 + n] = u[n];
        }
        return cols;
    }
    // pass 1' of ONE column c (any thread): u[n1] = c'[n1*4096 + VL*c + lane]
    static DP_DEV void inv_pass1_col(const V* buf, const V* scr, const V* DP_RESTRICT tw1, int c, V (&u)[R1]) {
        constexpr int LP = NPH - 1;
#pragma unroll
        for (int b = 0; b < NB; ++b) u[G::k1_of(LP, b)] = buf[G::phys(c) + b * PB];
#pragma unroll
        for (int p = 0; p < LP; ++p)
#pragma unroll
            for (int b = 0; b < NB; ++b) u[G::k1_of(p, b)] = dp2_ld_keep(scr + (long long)(p * NB + b) * VPB + c, dp2_policy_keep());
        V w = dp_ldg(tw1 + c);
        w.im = -w.im;
        V pw[R1];
        dp2_powers<R1, T>(w, pw);
#pragma unroll
        for (int k = 1; k < R1; ++k) u[k] = cmul(u[k], pw[k]);
        dp_dft<R1, +1, T>::run(u);
    }
    // amplitude sample with rolled index r from the pass-1' outputs of its column
    static DP_DEV S sample_of(const V (&u)[R1], int r) {
        const int n = r >> 1, n1 = n / 4096;
        S v = (S)0;
#pragma unroll
        for (int q = 0; q < R1; ++q)
            if (q == n1) {
                if constexpr (VL == 1) {
                    v = (r & 1) ? u[q].im : u[q].re;
                } else {
                    const bool hi = (n & 1) != 0;
                    v = (r & 1) ? (hi ? u[q].im.y : u[q].im.x) : (hi ? u[q].re.y : u[q].re.x);
                }
            }
        return v;
    }
    static DP_DEV unsigned inv_pass1(const V* buf, const V* scr, const V* DP_RESTRICT tw1, int i0, V (&y)[G::GC * R1], int nlo = 0,
                                     int nhi = 0x7fffffff) {
        return inv_pass1m(buf, scr, tw1, i0, y, (need_cols(nlo, nhi) >> i0) & ((1u << G::GC) - 1u));
    }
};

// ------------------------------------------------------- windowed arg-max scans
// y[i*R1 + n1], i < GC: amplitude samples with rolled delay index
//     r = 2*(n1*4096 + VL*(tid + (i0+i)*NT) + lane) + (0: real part, 1: imaginary part).
// Scanned in ascending r with a strict '>' (first maximum, like numpy argmin on chi2).
template <class T, int R1> struct Dp2Scan {
    using G = Dp2Geom<T, R1>;
    using S = typename G::S;
    static constexpr int VL = G::VL, NT = G::NT, GC = G::GC;

    template <int L> static DP_DEV S re_of(const cx<T>& v) {
        if constexpr (VL == 1) return v.re; else return L == 0 ? v.re.x : v.re.y;
    }
    template <int L> static DP_DEV S im_of(const cx<T>& v) {
        if constexpr (VL == 1) return v.im; else return L == 0 ? v.im.x : v.im.y;
    }
    static DP_DEV void full(const cx<T> (&y)[GC * R1], int tid, int i0, DpBest<S>& out) {
        S vb = (S)0, kb = (S)-1;
        int pb = 0;
#pragma unroll
        for (int n = 0; n < R1; ++n)
#pragma unroll
            for (int i = 0; i < GC; ++i) {
                const int off = 2 * (n * 4096 + VL * (i0 + i) * NT);
                const cx<T>& v = y[i * R1 + n];
                if (dp_abs(re_of<0>(v)) > kb) kb = dp_abs(re_of<0>(v)), vb = re_of<0>(v), pb = off;
                if (dp_abs(im_of<0>(v)) > kb) kb = dp_abs(im_of<0>(v)), vb = im_of<0>(v), pb = off + 1;
                if constexpr (VL == 2) {
                    if (dp_abs(re_of<1>(v)) > kb) kb = dp_abs(re_of<1>(v)), vb = re_of<1>(v), pb = off + 2;
                    if (dp_abs(im_of<1>(v)) > kb) kb = dp_abs(im_of<1>(v)), vb = im_of<1>(v), pb = off + 3;
                }
            }
        DpBest<S> c{vb, 2 * VL * tid + pb};
        dp_best_merge(out, c);
    }
    // `computed`: columns pass 1' produced (the others hold no candidate of any fit)
    static DP_DEV void window(const cx<T> (&y)[GC * R1], int tid, int i0, int lo, unsigned len, bool outside, DpBest<S>& out,
                              unsigned computed = 0xffffffffu) {
        S kb = (S)-1, vb = (S)0;
        int pb = -1;
        const int rb = 2 * VL * tid;
#pragma unroll
        for (int n = 0; n < R1; ++n) {
            // CTA-uniform skip: row n1 = n of this column set covers r in [blo, bhi)
            const int blo = 2 * (n * 4096 + VL * i0 * NT), bhi = 2 * (n * 4096 + VL * (i0 + GC) * NT);
            const bool hit = outside ? !(lo <= blo && (long long)bhi <= (long long)lo + (long long)len)
                                     : (lo < bhi && (long long)lo + (long long)len > (long long)blo);
            if (hit) {
#pragma unroll
                for (int i = 0; i < GC; ++i) {
                    if (!(computed & (1u << i))) continue;
                    const int r0 = rb + 2 * (n * 4096 + VL * (i0 + i) * NT);
                    const cx<T>& v = y[i * R1 + n];
#define DP2_CAND(val, rr)                                                  \
    {                                                                      \
        const S a_ = (val);                                                \
        const int r_ = (rr);                                               \
        const bool in_ = ((unsigned)(r_ - lo) < len) != outside;           \
        if (in_ && dp_abs(a_) > kb) kb = dp_abs(a_), vb = a_, pb = r_;     \
    }
                    DP2_CAND(re_of<0>(v), r0)
                    DP2_CAND(im_of<0>(v), r0 + 1)
                    if constexpr (VL == 2) {
                        DP2_CAND(re_of<1>(v), r0 + 2)
                        DP2_CAND(im_of<1>(v), r0 + 3)
                    }
#undef DP2_CAND
                }
            }
        }
        DpBest<S> c{vb, pb};
        dp_best_merge(out, c);
    }
};

// ====================================================================== kernel
template <class T, int R1, int IN> struct Dp2OfKernel {
    using G = Dp2Geom<T, R1>;
    using S = typename G::S;
    using V = cx<T>;
    using Core = Dp2Core<T, R1, IN>;
    static constexpr int NT = G::NT, VL = G::VL, NB = G::NB, NPH = G::NPH, VPB = G::VPB, GC = G::GC, NC = G::NC;
    static constexpr int NW = NT / 32;
    static constexpr int N = G::N;
    static constexpr int RED_DOUBLES = (DP_MAX_TSLOTS + 1) * 32;
    static constexpr int BEST_ELEMS = DP_MAX_TSLOTS * 32;
    static constexpr int SP_ELEMS = 32 + 34 + 2;  // self-paired group values, X of the 17 self pairs, chi0 (as one cx slot)
    static constexpr int CH_WORDS = (int)(sizeof(Dp2ChanDev<T>) / 4);  // channel descriptor, 32-bit words
    // staged tables: per warp four 2048-byte pieces inside the group rows it owns (its chunks, Dp2Params::zones) and
    // NSP private pieces.  Round A (fused untangle + first template): wj rows 0..15 | phi rows 0..7 in the owned rows,
    // phi rows 8.. in the private pieces; round B (further templates): X rows 0..15 in the owned rows, phi rows 0.. in
    // the private pieces.  Rows that do not fit are read through the read-only path when they are used.
#ifndef DP2_NSP
#define DP2_NSP 0
#endif
    static constexpr int NSP = (NT > 256) ? DP2_NSP : (DP2_NSP > 1 ? 1 : DP2_NSP);  // two 256-thread CTAs per SM leave room for one
    // two 256-thread CTAs per SM (packed fp32, 16384 samples) already overlap each other's table latency: there the
    // rows are read through the read-only path when they are used (staging cost 4 % there, measured)
#ifndef DP2_STAGE_MIN_NT
#define DP2_STAGE_MIN_NT 257
#endif
    static constexpr bool STAGE = NT >= DP2_STAGE_MIN_NT;
    static constexpr int PHI_ROWS_A = STAGE ? 8 + 4 * NSP : 0, PHI_ROWS_B = STAGE ? 4 * NSP : 0;
    static constexpr unsigned BYTES_A = 4 * DP2_PIECE + NSP * DP2_PIECE, BYTES_B = 4 * DP2_PIECE + NSP * DP2_PIECE;
    static_assert(sizeof(T) == 8 && sizeof(V) == 16, "table rows are 256 / 512 bytes per warp");
    static constexpr size_t SMEM_BYTES = sizeof(V) * G::SMEM_V + sizeof(cx<S>) * (DP_NLOW_MAX + SP_ELEMS) +
                                         2 * (sizeof(double) * RED_DOUBLES + sizeof(DpBest<S>) * BEST_ELEMS + sizeof(int) * DP_MAX_TSLOTS) +
                                         2 * sizeof(int) * CH_WORDS + 64 + sizeof(double) * 4 * DP_MAX_TSLOTS + (size_t)NW * NSP * DP2_PIECE + sizeof(Dp2Mbar) * NW + 16 + 16;
    // scratch per CTA (V units): X spill [NW][16][32] (multi-template; warp-sliced so that a warp's column is one
    // contiguous 8 KB block) + per template the parked block results of the non-final phases [(NPH-1)*NB][VPB]
    static constexpr long long SCR_X = (long long)16 * NT;
    static constexpr long long SCR_PARK = (long long)(NPH - 1) * NB * VPB;
    // (+ the first-pass outputs of the phase that does not fit into TMEM, Dp2Core::PARK1_V, in front)
#ifndef DP_OF_TMEM
#ifdef DP_HOST_EMU
#define DP_OF_TMEM 0   // the host-thread emulator has no tensor memory: it runs the per-phase first pass
#else
#define DP_OF_TMEM 1
#endif
#endif
    static constexpr bool TM = DP_OF_TMEM && Core::CAN_PARK;   // fp64, split transform: first pass computed once (pass1_all)
    static constexpr long long SCR_1 = TM ? Core::PARK1_V : 0;
    // multi-template plans, fp64: the thread's X column (16 values) lives in the TMEM columns the first pass leaves free, and
    // the rows the X column used to be copied back into take the next template's 16 filter rows instead (round 2: X went
    // to an L2 scratch column by 16-byte stores, came back by bulk copy behind a proxy fence, and the filter rows of the
    // second template were read through L2 when they were used)
#ifndef DP_OF_TMX
#define DP_OF_TMX 1
#endif
    static constexpr bool TMX = DP_OF_TMX && TM && STAGE && VL == 1 && (Core::TM_COLS + 64 <= 128);
    static constexpr unsigned XCOL = Core::TM_COLS;
    static DP_HD long long scratch_v(int n_templ) { return SCR_1 + SCR_X + SCR_PARK * n_templ; }
    // first element of warp w's rows in a thread-order table [NPH][NW][16][32]
    static DP_HD long long tab_block(int p, int w) { return ((long long)(p * NW + w) * 16) * 32; }

    struct Smem {
        V* buf;
        cx<S>* stash;
        cx<S>* sp;
        double* red0;
        DpBest<S>* best0;
        int* slot0;
        int* chs0;  // [2][CH_WORDS] this event's channel descriptor (table pointers are read from
                    // shared memory, not through a dependent global load)
        double* neigh0;        // [2][2 * DP_MAX_TSLOTS] amplitudes next to each fit's best delay (interpolate_t0)
        unsigned char* spare;  // [NW][NSP][2048] private landing pieces
        Dp2Mbar* mbar;         // [NW]
        DP_DEV double* red(int par) const { return red0 + par * RED_DOUBLES; }
        DP_DEV DpBest<S>* best(int par) const { return best0 + par * BEST_ELEMS; }
        DP_DEV int* slot_id(int par) const { return slot0 + par * DP_MAX_TSLOTS; }
        DP_DEV double* neigh(int par) const { return neigh0 + par * 2 * DP_MAX_TSLOTS; }
    };
    static DP_DEV Smem carve(unsigned char* raw) {
        Smem s;
        s.buf = reinterpret_cast<V*>(raw);
        s.stash = reinterpret_cast<cx<S>*>(s.buf + G::SMEM_V);
        s.sp = s.stash + DP_NLOW_MAX;
        s.red0 = reinterpret_cast<double*>(s.sp + SP_ELEMS);
        s.best0 = reinterpret_cast<DpBest<S>*>(s.red0 + 2 * RED_DOUBLES);
        s.slot0 = reinterpret_cast<int*>(s.best0 + 2 * BEST_ELEMS);
        s.chs0 = s.slot0 + 2 * DP_MAX_TSLOTS + ((2 * DP_MAX_TSLOTS) & 3 ? 4 - ((2 * DP_MAX_TSLOTS) & 3) : 0);
        unsigned char* e = reinterpret_cast<unsigned char*>(s.chs0 + 2 * CH_WORDS);
        e += (16 - (reinterpret_cast<unsigned long long>(e) & 15)) & 15;
        s.neigh0 = reinterpret_cast<double*>(e);
        e += sizeof(double) * 4 * DP_MAX_TSLOTS;
        s.spare = e;
        s.mbar = reinterpret_cast<Dp2Mbar*>(e + (size_t)NW * NSP * DP2_PIECE);
        return s;
    }

    // element index of the first sample of (event, channel); false: the window leaves the stream (window mode)
    static DP_DEV bool first_sample(const Dp2Params<T>& prm, int ev, int chan, long long& first) {
        long long base = (long long)ev * prm.event_stride;
        bool ok = true;
        if (prm.row_start != nullptr) {
            base = prm.row_start[ev];
            ok = base >= 0 && base + N <= prm.stream_len;
        }
        first = base + (prm.chan_offset != nullptr ? prm.chan_offset[chan] : (long long)chan * prm.chan_stride);
        return ok;
    }

    // next trace of this CTA -> L2, one bulk-prefetch instruction (TMA path, no LSU traffic) issued
    // by one thread at the start of the event; falls back to per-line prefetches when the row is
    // not 16-byte aligned
    static DP_DEV void prefetch_next(const Dp2Params<T>& prm, int row) {
        constexpr size_t ESZ = sizeof(typename DpRaw<IN>::scalar);
        const int nrow = row + gridDim.x;
        if (nrow < prm.n_rows) {
            long long first;
            if (!first_sample(prm, nrow / prm.n_chan, nrow % prm.n_chan, first)) return;
            const unsigned char* nx = reinterpret_cast<const unsigned char*>(prm.traces) + (size_t)first * ESZ;
#ifndef DP_HOST_EMU
            if ((reinterpret_cast<unsigned long long>(nx) & 15ull) == 0) {
                if (threadIdx.x == 0)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nx), "r"((unsigned)((size_t)N * ESZ)) : "memory");
                return;
            }
#endif
            constexpr int nlines = (int)((size_t)N * ESZ / 128);
            for (int l = threadIdx.x; l < nlines; l += NT) dp_prefetch_l2(nx + (size_t)l * 128);
        }
    }

    // ---- this lane's view of the warp's staged tables: byte offsets from the start of the dynamic shared memory
    // (32-bit each; the lane's own offset folded in)
    struct Staged {
        unsigned char* base;
        unsigned v[4];      // pieces inside the warp's own group rows (valid between pass 4 and pass 4')
        unsigned s[NSP > 0 ? NSP : 1];    // private pieces
        const V* phi_g;     // the same filter rows in global memory (rows that were not staged), this lane
        const T* wj_g;      // chi0 weights in global memory (!STAGE), this lane
        const V* x_g;       // X column in the scratch (!STAGE), this lane
        template <class Q> DP_DEV Q at(unsigned off) const { return *reinterpret_cast<const Q*>(base + off); }
        // round A
        DP_DEV T wj(int e) const {
            if constexpr (!STAGE) return dp_ldg(wj_g + e * 32);
            return at<T>(v[e >> 3] + (e & 7) * 256);
        }
        DP_DEV V phi_a(int e) const {
            if constexpr (!STAGE) return dp2_ld_stream(phi_g + e * 32);
            if (e < 8) return at<V>(v[2 + (e >> 2)] + (e & 3) * 512);
            if (e < PHI_ROWS_A) return at<V>(s[(e - 8) >> 2] + ((e - 8) & 3) * 512);
            return dp2_ld_stream(phi_g + e * 32);
        }
        // round B
        DP_DEV V x(int e) const {
            if constexpr (!STAGE) return dp2_ld_keep(x_g + e * 32, dp2_policy_keep());
            return at<V>(v[e >> 2] + (e & 3) * 512);
        }
        DP_DEV V phi_b(int e) const {
            if (e < PHI_ROWS_B) return at<V>(s[e >> 2] + (e & 3) * 512);
            return dp2_ld_stream(phi_g + e * 32);
        }
    };
    // lane offsets folded in: wj pieces (round A, v[0], v[1]) are 8-byte rows, everything else 16-byte rows
    static DP_DEV Staged staged_view(unsigned char* smem_raw, uint4 zo, unsigned spoff, int lane, bool round_a, const V* phi_lane,
                                     const T* wj_lane, const V* x_lane) {
        Staged t;
        t.wj_g = wj_lane;
        t.x_g = x_lane;
        const unsigned l16 = lane * 16, l8 = lane * 8;
        t.base = smem_raw;
        t.v[0] = zo.x + (round_a ? l8 : l16);
        t.v[1] = zo.y + (round_a ? l8 : l16);
        t.v[2] = zo.z + l16;
        t.v[3] = zo.w + l16;
#pragma unroll
        for (int j = 0; j < NSP; ++j) t.s[j] = spoff + j * DP2_PIECE + l16;
        t.phi_g = phi_lane;
        return t;
    }
    // one lane: round A = wj block (4096 B) + the first PHI_ROWS_A rows of the phi block
    static DP_DEV void issue_a(Dp2Mbar* bar, unsigned char* bufb, uint4 zo, unsigned char* spw, const T* wj_blk, const V* phi_blk) {
        const unsigned char* w = reinterpret_cast<const unsigned char*>(wj_blk);
        const unsigned char* f = reinterpret_cast<const unsigned char*>(phi_blk);
        dp2_stage_begin(bar, BYTES_A);
        dp2_stage_copy(bufb + zo.x, w, DP2_PIECE, bar);
        dp2_stage_copy(bufb + zo.y, w + DP2_PIECE, DP2_PIECE, bar);
        dp2_stage_copy(bufb + zo.z, f, DP2_PIECE, bar);
        dp2_stage_copy(bufb + zo.w, f + DP2_PIECE, DP2_PIECE, bar);
#pragma unroll
        for (int j = 0; j < NSP; ++j) dp2_stage_copy(spw + j * DP2_PIECE, f + (2 + j) * DP2_PIECE, DP2_PIECE, bar);
    }
    // one lane: round B = the warp's X column (8192 B) + the first PHI_ROWS_B rows of the phi block
    static DP_DEV void issue_b(Dp2Mbar* bar, unsigned char* bufb, uint4 zo, unsigned char* spw, const V* x_blk, const V* phi_blk) {
        const unsigned char* x = reinterpret_cast<const unsigned char*>(x_blk);
        const unsigned char* f = reinterpret_cast<const unsigned char*>(phi_blk);
        dp2_stage_begin(bar, BYTES_B);
        dp2_stage_copy(bufb + zo.x, x, DP2_PIECE, bar);
        dp2_stage_copy(bufb + zo.y, x + DP2_PIECE, DP2_PIECE, bar);
        dp2_stage_copy(bufb + zo.z, x + 2 * DP2_PIECE, DP2_PIECE, bar);
        dp2_stage_copy(bufb + zo.w, x + 3 * DP2_PIECE, DP2_PIECE, bar);
#pragma unroll
        for (int j = 0; j < NSP; ++j) dp2_stage_copy(spw + j * DP2_PIECE, f + j * DP2_PIECE, DP2_PIECE, bar);
    }

    // ---- packed fp32 (VL == 2) point-wise stage on whole register arrays: z[r] lanes = (group A, group B) elements.
    // forward half: z (pass-4 outputs) -> 2*X at the thread's bins, in place; returns the chi0 partial sum
    static DP_DEV S untangle_st(V (&z)[16], const Staged& tb, cx<S> wn) {
        S chi = (S)0;
        if constexpr (VL == 2) {
#define DP2_XP(r)                                                                         \
    {                                                                                     \
        cx<S> Xk, Xm;                                                                     \
        dp_untangle(dp2_lane0(z[r]), dp2_lane1(z[15 - r]), cmul(wn, dp_w64<S, 2 * r, -1>()), Xk, Xm); \
        dp2_set0(z[r], Xk);                                                               \
        dp2_set1(z[15 - r], Xm);                                                          \
    }
            DP2_XP(0) DP2_XP(1) DP2_XP(2) DP2_XP(3) DP2_XP(4) DP2_XP(5) DP2_XP(6) DP2_XP(7)
            DP2_XP(8) DP2_XP(9) DP2_XP(10) DP2_XP(11) DP2_XP(12) DP2_XP(13) DP2_XP(14) DP2_XP(15)
#undef DP2_XP
            f2 acc = f2(0.0f);
#pragma unroll
            for (int r = 0; r < 16; ++r) acc = dp_fma(tb.wj(r), cnorm2(z[r]), acc);
            chi = acc.x + acc.y;
        }
        return chi;
    }
    // filter multiply + inverse untangle, in place: z X -> Z' (group values for pass 4'); ROUND_A selects the table view
    template <bool ROUND_A> static DP_DEV void filter_st(V (&z)[16], const Staged& tb, cx<S> wn) {
        if constexpr (VL == 2) {
#pragma unroll
            for (int r = 0; r < 16; ++r) z[r] = cmul(ROUND_A ? tb.phi_a(r) : tb.phi_b(r), z[r]);
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
    }
    // ---- fp64 (VL == 1): a thread owns one group (elements 0..15), its partner lane (lane ^ 1) the mirror group;
    // pair r = (own[r], partner[15 - r]), r < 8.  The partner's element and the partner's half of the result travel
    // by warp shuffle, so the stage touches shared memory only for the staged tables.
    // filter + inverse untangle of pair r: C'[own r] -> z[r], C'[partner 15 - r] -> the partner's z[15 - r]
    static DP_DEV void pair_filter(V (&z)[16], int r, cx<S> Xk, cx<S> Xm, V phk, V phm, cx<S> w) {
        if constexpr (VL == 1) {
            const cx<S> Fk = cmul(phk, Xk);
            const cx<S> Fm = cmul(phm, Xm);
            cx<S> Ck, Cm;
            dp_retangle(Fk, Fm, w, Ck, Cm);
            z[r] = Ck;
            z[15 - r] = dp2_swap1(Cm);
        }
    }

    // ---- point-wise helpers of the PSD / CSD / NxM / trigger kernels (tables in [phase][16][NT] order read through the
    // read-only path, partner exchange through the idle group rows); the OF kernel itself uses the staged variants above
    // ---- point-wise stage, forward half: z (pass-4 outputs) -> 2*X at the thread's bins.
    // VL == 2: z[r] lanes = X at (bin of A[r], bin of B[r]).
    // (packed fp32 only; fp64 goes pair by pair through pw_untangle / pw_filter_pair below)
    // Returns the thread's chi0 partial sum (scalar type S).
    template <bool WITH_CHI = true>
    static DP_DEV S untangle_all(V* buf, V (&z)[16], V (&zm)[VL == 1 ? 8 : 1], const T* DP_RESTRICT wj, cx<S> wn, int Gown,
                                 bool special) {
        const int tid = threadIdx.x;
        S chi = (S)0;
        if constexpr (VL == 2) {
            (void)zm;
            (void)Gown;
            (void)buf;
#define DP2_XP(r)                                                                         \
    {                                                                                     \
        cx<S> Xk, Xm;                                                                     \
        dp_untangle(dp2_lane0(z[r]), dp2_lane1(z[15 - r]), cmul(wn, dp_w64<S, 2 * r, -1>()), Xk, Xm); \
        dp2_set0(z[r], Xk);                                                               \
        dp2_set1(z[15 - r], Xm);                                                          \
    }
            DP2_XP(0) DP2_XP(1) DP2_XP(2) DP2_XP(3) DP2_XP(4) DP2_XP(5) DP2_XP(6) DP2_XP(7)
            DP2_XP(8) DP2_XP(9) DP2_XP(10) DP2_XP(11) DP2_XP(12) DP2_XP(13) DP2_XP(14) DP2_XP(15)
#undef DP2_XP
            f2 acc = f2(0.0f);
            if constexpr (WITH_CHI) {
#pragma unroll
                for (int r = 0; r < 16; ++r) acc = dp_fma(dp_ldg(wj + r * NT + tid), cnorm2(z[r]), acc);
            }
            chi = acc.x + acc.y;
        } else {
            static_assert(VL == 2, "fp64 uses the pair-wise pw_* functions");
            (void)tid;
        }
        return special ? (S)0 : chi;
    }

    // filter multiply + inverse untangle, in place: (z, zm) X -> Z' (group values for pass 4')
    static DP_DEV void filter_all(V* buf, V (&z)[16], V (&zm)[VL == 1 ? 8 : 1], const V* DP_RESTRICT phi, cx<S> wn, int Gown) {
        const int tid = threadIdx.x;
        if constexpr (VL == 2) {
            (void)zm;
            (void)Gown;
            (void)buf;
#pragma unroll
            for (int r = 0; r < 16; ++r) z[r] = cmul(dp_ldg(phi + r * NT + tid), z[r]);
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
        } else {
            static_assert(VL == 2, "fp64 uses the pair-wise pw_* functions");
            (void)tid;
        }
    }

    // ---- fp64 (VL == 1) point-wise stage, one (k, M-k) pair at a time.  A thread owns group elements
    // 0..15; pair r = (own[r], partner[15 - r]), r < 8, where the partner thread (tid ^ 1) owns the mirror
    // group.  Nothing but z stays in registers: the partner's element is read from its group row when the
    // pair is processed and the pair's results go straight to their consumers (the former version kept
    // eight more vectors alive, which left ptxas three registers' worth of filter loads in flight).
    static DP_DEV void pw_publish(V* buf, const V (&z)[16], int Gown) {
#pragma unroll
        for (int j = 0; j < 8; ++j) buf[Gown * 17 + 8 + j] = z[8 + j];  // own elements 8..15 for the partner
        __syncwarp();
    }
    // fn(r, Xk, Xm): 2*X at bin(own[r]) and at M - bin(own[r])
    template <class Fn> static DP_DEV void pw_untangle(const V* buf, const V (&z)[16], cx<S> wn, int Gp, Fn&& fn) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const V Zm = buf[Gp * 17 + 15 - r];
            cx<S> Xk, Xm;
            dp_untangle(z[r], Zm, cmul(wn, dp_w64_rt<S>(2 * r)), Xk, Xm);
            fn(r, Xk, Xm);
        }
    }
    // filter + inverse untangle of pair r: C'[own r] -> z[r], C'[partner 15 - r] -> the partner's row
    static DP_DEV void pw_filter_pair(V* buf, V (&z)[16], int r, cx<S> Xk, cx<S> Xm, const V* DP_RESTRICT phi, cx<S> wn, int Gp) {
        const int tid = threadIdx.x;
        const cx<S> Fk = cmul(dp_ldg(phi + (2 * r) * NT + tid), Xk);
        const cx<S> Fm = cmul(dp_ldg(phi + (2 * r + 1) * NT + tid), Xm);
        cx<S> Ck, Cm;
        dp_retangle(Fk, Fm, cmul(wn, dp_w64_rt<S>(2 * r)), Ck, Cm);
        z[r] = Ck;
        buf[Gp * 17 + 15 - r] = Cm;
    }
    static DP_DEV void pw_collect(const V* buf, V (&z)[16], int Gown) {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) z[8 + j] = buf[Gown * 17 + 8 + j];  // written by the partner
    }

    // MULTI: some channel has more than one template (X goes through the warp's scratch column and comes back by
    // bulk copy for the second and later templates)
    // SCAN: how the delay search walks the pass-1' outputs -- 0: blocks of GC columns (wide windows), 1: column by column
    // (the host guarantees narrow windows: dp_capi.cu of2_narrow), 2: chosen per template at run time (emulator builds)
    template <bool MULTI, int SCAN = 2> static DP_DEV void run(const Dp2Params<T>& prm, unsigned char* smem_raw);
};

template <class T, int R1, int IN>
template <bool MULTI, int SCAN>
DP_DEV void Dp2OfKernel<T, R1, IN>::run(const Dp2Params<T>& prm, unsigned char* smem_raw) {
    const Smem sm = carve(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr size_t ESZ = sizeof(typename DpRaw<IN>::scalar);
    [[maybe_unused]] V* const scr_1 = prm.scratch + (long long)blockIdx.x * prm.scratch_per_cta;  // first-pass outputs parked in L2 (R1 = 8)
    V* scr_x = scr_1 + SCR_1;                                              // [NW][16][32] X of the current phase (multi-template)
    V* scr_park = scr_x + SCR_X;                                           // [n_templ][(NPH-1)*NB][VPB] parked block results
    V* const xs = scr_x + (long long)warp * 16 * 32 + lane;                // this lane's X column (rows 32 apart)
    // special threads: the self-paired groups (0,0,0) and (0,0,8) of block 0 (phase 0)
    constexpr int NSPECIAL = (VL == 2) ? 1 : 2;
    int par = 0;
    int evpar = 0;
    cx<S>* const sx = sm.sp + 32;                                        // [17][2] X of the self pairs
    double* const chi0_keep = reinterpret_cast<double*>(sm.sp + 32 + 34);  // chi0 of the current event (thread 0)
    unsigned char* const bufb = reinterpret_cast<unsigned char*>(sm.buf);
    unsigned char* const spw = sm.spare + (size_t)warp * NSP * DP2_PIECE;
    const unsigned spoff = (unsigned)(spw - bufb);  // the FFT buffer starts the dynamic shared memory
    Dp2Mbar* const bar = sm.mbar + warp;
    unsigned spar = 0;  // parity of the warp's next staging round
    if (lane == 0) dp2_mbar_init(bar);
    dp2_mbar_init_fence();
    [[maybe_unused]] unsigned* const tmslot = reinterpret_cast<unsigned*>(sm.mbar + NW);
    [[maybe_unused]] unsigned tm_thread = 0;  // this thread's 128 private TMEM columns
#ifndef DP_HOST_EMU
    if constexpr (TM) {
        if (warp == 0) dp_tmem_alloc512(tmslot);
        dp_tmem_fence_before();
    }
#endif
    __syncthreads();
#ifndef DP_HOST_EMU
    if constexpr (TM) {
        dp_tmem_fence_after();
        tm_thread = Core::tm_thread_base(*tmslot);
    }
#endif

    for (int row = blockIdx.x; row < prm.n_rows; row += gridDim.x) {
        const int chan = row % prm.n_chan;
        const int ev = row / prm.n_chan;
        // channel descriptor -> shared memory (double buffered: thread 0 may still be writing the
        // previous event's outputs); visible after the first barrier of the event
        int* chs = sm.chs0 + evpar * CH_WORDS;
        evpar ^= 1;
        if (tid < CH_WORDS) chs[tid] = reinterpret_cast<const int*>(prm.chans + chan)[tid];
        const Dp2ChanDev<T>& ch = *reinterpret_cast<const Dp2ChanDev<T>*>(chs);
        long long first;
        if (!first_sample(prm, ev, chan, first)) {  // CTA-uniform: the window leaves the stream
            const Dp2ChanDev<T>* chg = prm.chans + chan;
            const int nb = 1 + (DP_SLOT_NOUT + (prm.neighbours ? 2 : 0)) * chg->n_slots;
            for (int o = tid; o < nb; o += NT) prm.out[(long long)ev * prm.n_out + chg->out_base + o] = -999999.0;
            continue;
        }
        const void* xrow = reinterpret_cast<const unsigned char*>(prm.traces) + (size_t)first * ESZ;
        double x0 = prm.subtract_first ? dp_load_first<IN>(xrow) : 0.0;
        double xsc = prm.scale;
        if constexpr (dp_in_is_adc(IN)) {
            // ADC counts -> samples.  fp64: adc * gain + offset; fp32: (adc - x0) * (gain * scale) with x0 the first
            // count (AC coupling) or the count the offset cancels
            const double gain = dp_ldg(&prm.chans[chan].adc_gain), offs = dp_ldg(&prm.chans[chan].adc_offset);
            if constexpr (VL == 1) {
                x0 = -offs;
                xsc = gain;
            } else {
                if (!prm.subtract_first) x0 = -offs / gain;
                xsc = gain * prm.scale;
            }
        }
        S chi = (S)0;

#pragma unroll 1
        for (int p = 0; p < NPH; ++p) {
            // ---------------- forward of phase p: X at the phase's bins, in registers ----------
            V z[16];
            const int2 gg = prm.groups[p * NT + tid];
            const int chunk = prm.chunk3[p * NT + tid];
            const cx<S> wn = dp_ldg(prm.twn + p * NT + tid);
            const bool special = (p == 0) && (tid < NSPECIAL);
            if constexpr (TM) {
                // the trace is read once: phase 0 computes the first pass of all blocks, the later phases find theirs in TMEM / L2
                if (p == 0)
                    Core::pass1_all(xrow, x0, xsc, sm.buf, prm.tw1, tm_thread, scr_1);
                else
                    Core::pass1_fetch(p, sm.buf, tm_thread, scr_1);
            } else {
                Core::pass1_any(p, xrow, x0, xsc, sm.buf, prm.tw1);
            }
            __syncthreads();
#ifndef DP_HOST_EMU
            // every second block starts its passes a little late so that the block sets' LDS / FP / STS phases
            // interleave instead of hitting the same pipe at the same time (0 disables)
            if (prm.skew_ns > 0 && ((tid / G::CV) & 1)) __nanosleep(prm.skew_ns);
#endif
            Core::fwd_2(sm.buf, prm.tw2, z);
            dp_bar_sync(G::bar_set_id(p, tid), G::bar_set_count(p, tid));  // pass 3 reads the chunks of the warp's block set
            Core::fwd_3w(sm.buf, prm.tw3, chunk, z);
            __syncwarp();  // pass 4 reads the warp's own chunks
            Core::load_groups(sm.buf, gg.x, gg.y, z);
            __syncwarp();  // the group rows of this warp are now free: land the point-wise tables in them
            if (STAGE && lane == 0) issue_a(bar, bufb, prm.zones[p * NW + warp], spw, ch.wj + tab_block(p, warp), ch.templ[0].phi + tab_block(p, warp));
            dp_dft<16, -1, T>::run(z);
            if (p == 0 && tid < 32) {
                // self-paired groups -> 17 lanes of warp 0
                if constexpr (VL == 2) {
                    if (tid == 0) {
#pragma unroll
                        for (int r = 0; r < 16; ++r) {
                            sm.sp[r] = dp2_lane0(z[r]);
                            sm.sp[16 + r] = dp2_lane1(z[r]);
                        }
                    }
                } else {
                    if (tid < 2) {
#pragma unroll
                        for (int r = 0; r < 16; ++r) sm.sp[16 * tid + r] = z[r];
                    }
                }
                __syncwarp();
                if (tid < 17) {
                    const DpSelfLane<S> sp = dp_self_lane<S, 1>(tid);
                    cx<S> sXk, sXm;
                    dp_untangle(sm.sp[sp.ek], sm.sp[sp.em], sp.w, sXk, sXm);
                    sx[2 * tid] = sXk;
                    sx[2 * tid + 1] = sXm;
                    chi = dp_fma(dp_ldg(ch.wj_self + 2 * tid), cnorm2(sXk), chi);
                    chi = dp_fma(dp_ldg(ch.wj_self + 2 * tid + 1), cnorm2(sXm), chi);
                    if (tid == 0) sm.stash[0] = sXk;
                    if (tid == 9 && G::KQ / 2 < prm.nlow) sm.stash[G::KQ / 2] = sXk;
                }
                __syncwarp();
            }
            if constexpr (STAGE) {
                dp2_stage_wait(bar, spar);
                spar ^= 1;
            }
            {
                const Staged tb = staged_view(bufb, prm.zones[p * NW + warp], spoff, lane, true, ch.templ[0].phi + tab_block(p, warp) + lane,
                                              ch.wj + tab_block(p, warp) + lane, xs);
                if constexpr (VL == 2) {
                    const S c = untangle_st(z, tb, wn);
                    if (!special) {
                        chi += c;
                        // low-frequency bins for lowchi2: element 0 of each group (k4 = 0)
                        const int kA = G::bin_of(p, gg.x, 0), kB = G::bin_of(p, gg.y, 0);
                        if (kA < prm.nlow) sm.stash[kA] = dp2_lane0(z[0]);
                        if (kB < prm.nlow) sm.stash[kB] = dp2_lane1(z[0]);
                    }
                    if constexpr (MULTI) {
                        // X must survive the in-place inverse of the first template
                        const unsigned long long pol = dp2_policy_keep();
#pragma unroll
                        for (int r = 0; r < 16; ++r) dp2_st_keep(xs + r * 32, z[r], pol);
                    }
                    filter_st<true>(z, tb, wn);
                } else {
                    // fp64: pair by pair -- chi0, the lowchi2 stash, the X column (multi-template plans) and the first
                    // template's filter + inverse untangle while the pair is in registers
                    const int kA = G::bin_of(p, gg.x, 0);
                    const bool stash0 = !special && kA < prm.nlow;
                    [[maybe_unused]] const unsigned long long pol = dp2_policy_keep();
                    S chin = (S)0;
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
#ifndef DP2_PAIR_SPLIT
#define DP2_PAIR_SPLIT 4
#endif
                        // keeps the scheduler from hoisting all eight partner exchanges to the top (17 spilled doubles)
                        if (DP2_PAIR_SPLIT > 0 && r > 0 && r % (DP2_PAIR_SPLIT > 0 ? DP2_PAIR_SPLIT : 1) == 0) __syncwarp();
                        const V Zm = dp2_swap1(z[15 - r]);  // the partner's element 15 - r
                        const cx<S> w = cmul(wn, dp_w64_rt<S>(2 * r));
                        cx<S> Xk, Xm;
                        dp_untangle(z[r], Zm, w, Xk, Xm);
                        chin = dp_fma(tb.wj(2 * r), cnorm2(Xk), chin);
                        chin = dp_fma(tb.wj(2 * r + 1), cnorm2(Xm), chin);
                        if (r == 0 && stash0) sm.stash[kA] = Xk;
                        if constexpr (MULTI) {
                            if constexpr (TMX) {
#ifndef DP_HOST_EMU
                                dp_tmem_st2(tm_thread + XCOL + 8 * r, Xk, Xm);
#endif
                            } else {
                                dp2_st_keep(xs + (2 * r) * 32, Xk, pol);
                                dp2_st_keep(xs + (2 * r + 1) * 32, Xm, pol);
                            }
                        }
                        pair_filter(z, r, Xk, Xm, tb.phi_a(2 * r), tb.phi_a(2 * r + 1), w);
                    }
                    if (!special) chi += chin;
#ifndef DP_HOST_EMU
                    if constexpr (MULTI && TMX) dp_tmem_wait_st();  // the X column is read back for the next template
#endif
                }
            }
            __syncwarp();  // every lane is done with the staged tables before the rows take the pass-4' group values
            // next trace -> L2 once this event's last read of its own trace is done (a whole event of
            // lead time let the scratch / table traffic of 148 SMs evict the line before its use)
            if (p == NPH - 1) prefetch_next(prm, row);

            // ---------------- per template: inverse passes 4' 3' 2' (the first template's filter is done) -------
            const int n_templ = MULTI ? ch.n_templ : 1;
            int it = 0;
#pragma unroll 1
            for (;;) {
                const Dp2TemplDev<T>& tp = ch.templ[it];
                V* park = scr_park + (long long)it * SCR_PARK;
                const bool more = MULTI && (it + 1 < n_templ);
                if (p == 0 && tid < 32) {
                    if (tid < 17) {
                        const DpSelfLane<S> sp = dp_self_lane<S, 1>(tid);
                        const cx<S> Fk = cmul(dp_ldg(tp.phi_self + 2 * tid), sx[2 * tid]);
                        const cx<S> Fm = cmul(dp_ldg(tp.phi_self + 2 * tid + 1), sx[2 * tid + 1]);
                        cx<S> Ck, Cm;
                        dp_retangle(Fk, Fm, sp.w, Ck, Cm);
                        sm.sp[sp.ek] = Ck;
                        if (sp.ek != sp.em) sm.sp[sp.em] = Cm;
                    }
                    __syncwarp();
                    if constexpr (VL == 2) {
                        if (tid == 0) {
#pragma unroll
                            for (int r = 0; r < 16; ++r) z[r] = V{f2(sm.sp[r].re, sm.sp[16 + r].re), f2(sm.sp[r].im, sm.sp[16 + r].im)};
                        }
                    } else {
                        if (tid < 2) {
#pragma unroll
                            for (int r = 0; r < 16; ++r) z[r] = sm.sp[16 * tid + r];
                        }
                    }
                    __syncwarp();
                }
                dp_dft<16, +1, T>::run(z);
                Core::store_groups(sm.buf, gg.x, gg.y, z);
                __syncwarp();  // pass 3' reads the warp's own chunks
                Core::inv_3w(sm.buf, prm.tw3, chunk, z);
                // fits of this template (at most DP_MAX_TSLOTS) and the complex points n = r/2 that some of them can select
                int slot_of[DP_MAX_TSLOTS];
                int nts = 0;
#pragma unroll
                for (int q = 0; q < DP_MAX_TSLOTS; ++q) slot_of[q] = -1;
                int nlo = 0x7fffffff, nhi = -1;
                for (int s = 0; s < ch.n_slots; ++s) {
                    const DpSlot sl = ch.slots[s];
                    if (sl.templ == it) {
#pragma unroll
                        for (int q = 0; q < DP_MAX_TSLOTS; ++q)
                            if (q == nts) slot_of[q] = s;
                        ++nts;
                        const bool everything = sl.outside || (sl.lo == 0 && sl.hi == N);
                        const int a = everything ? 0 : (sl.lo >> 1), b = everything ? 0x7fffffff : ((sl.hi - 1) >> 1);
                        nlo = a < nlo ? a : nlo;
                        nhi = b > nhi ? b : nhi;
                    }
                }
                if (prm.neighbours && nhi >= nlo && nhi - nlo < 4095) {   // the samples next to a window edge must exist as well
                    nlo = nlo > 0 ? nlo - 1 : 0;
                    nhi = nhi < G::M - 1 ? nhi + 1 : G::M - 1;
                }
                const unsigned rowmask = Core::rows_of(nlo, nhi);
                const bool few_rows = __popc(rowmask) <= 2;  // narrow delay window(s): Horner evaluation of the rows it touches
                dp_bar_sync(G::bar_set_id(p, tid), G::bar_set_count(p, tid));  // pass 2' reads columns across the set's chunks
                if (p < NPH - 1) {
                    if (few_rows) {
                        Core::template inv_2_rows<true>(sm.buf, park, p, prm.tw2, rowmask);
                    } else {
                        Core::inv_2(sm.buf, prm.tw2, z);
                        Core::park_pass2(park, p, z, rowmask);
                    }
                    if (STAGE && MULTI && it == 0 && !TMX) dp2_fence_async();  // the X column (stored long ago: cheap by now) is read back by bulk copy
                    __syncthreads();  // pass-2' reads of buf precede the next group / pass-1 stores
                    if (STAGE && more && lane == 0)
                        issue_b(bar, bufb, prm.zones[p * NW + warp], spw, TMX ? ch.templ[it + 1].phi + tab_block(p, warp) : scr_x + (long long)warp * 16 * 32, ch.templ[it + 1].phi + tab_block(p, warp));
                } else {
                    // ------------ last phase: pass 1' over all blocks, arg-max, outputs --------------
                    if (few_rows) {
                        Core::template inv_2_rows<false>(sm.buf, nullptr, p, prm.tw2, rowmask);
                    } else {
                        Core::inv_2(sm.buf, prm.tw2, z);
                        Core::store_pass2(sm.buf, z, rowmask);
                    }
                    if (STAGE && MULTI && it == 0 && !TMX) dp2_fence_async();
                    __syncthreads();  // also orders the parked block results (global memory) within the CTA
                    DpBest<S> tb[DP_MAX_TSLOTS];
#pragma unroll
                    for (int q = 0; q < DP_MAX_TSLOTS; ++q) tb[q] = DpBest<S>{(S)0, -1};
                    const unsigned colneed = Core::need_cols(nlo, nhi);
                    // narrow delay windows (at most two columns per thread hold a candidate; CTA-uniform)
                    const bool narrow = SCAN == 1 || (SCAN == 2 && nhi >= nlo && (nhi - nlo) < 2 * NT * VL - 1);
                    if (narrow) {
                    // column by column: pass 1' of a column that holds a candidate delay, then its R1 * 2 * VL samples are
                    // offered to every fit of this template (tie-aware: the columns are not visited in ascending delay
                    // order).  A +-500-sample window leaves one column per thread.
                    bool whole[DP_MAX_TSLOTS];
                    int wlo[DP_MAX_TSLOTS];
                    unsigned wlen[DP_MAX_TSLOTS];
                    bool wout[DP_MAX_TSLOTS];
#pragma unroll
                    for (int q = 0; q < DP_MAX_TSLOTS; ++q) {
                        whole[q] = false, wlo[q] = 0, wlen[q] = 0u, wout[q] = false;
                        if (q < nts) {
                            const DpSlot sl = ch.slots[slot_of[q]];
                            whole[q] = sl.lo == 0 && sl.hi == N && !sl.outside;
                            wlo[q] = sl.lo, wlen[q] = (unsigned)(sl.hi - sl.lo), wout[q] = sl.outside != 0;
                        }
                    }
#pragma unroll 1
                    for (unsigned cm = colneed; cm != 0; cm &= cm - 1) {
                        const int c = tid + (__ffs((int)cm) - 1) * NT;
                        V u[R1];
                        Core::inv_pass1_col(sm.buf, park, prm.tw1, c, u);
#pragma unroll
                        for (int q = 0; q < DP_MAX_TSLOTS; ++q) {
                            if (q >= nts) continue;
#pragma unroll
                            for (int n = 0; n < R1; ++n) {
                                const int r0 = 2 * (n * 4096 + VL * c);
#define DP2_CAND(val, rr)                                                                          \
    {                                                                                              \
        const S a_ = (val);                                                                        \
        const int r_ = (rr);                                                                       \
        const bool in_ = whole[q] || ((((unsigned)(r_ - wlo[q]) < wlen[q])) != wout[q]);           \
        if (in_) dp2_consider(tb[q], dp_abs(a_), a_, r_);                                          \
    }
                                DP2_CAND((Dp2Scan<T, R1>::template re_of<0>(u[n])), r0)
                                DP2_CAND((Dp2Scan<T, R1>::template im_of<0>(u[n])), r0 + 1)
                                if constexpr (VL == 2) {
                                    DP2_CAND((Dp2Scan<T, R1>::template re_of<1>(u[n])), r0 + 2)
                                    DP2_CAND((Dp2Scan<T, R1>::template im_of<1>(u[n])), r0 + 3)
                                }
#undef DP2_CAND
                            }
                        }
                    }
                    } else {
#pragma unroll 1
                    for (int i0 = 0; i0 < NC; i0 += GC) {
                        const unsigned computed = (colneed >> i0) & ((1u << GC) - 1u);
                        if (computed == 0) continue;  // this thread holds no candidate delay of this template in these columns
                        // (all of y is written here, not only the columns pass 1' computes: registers that are not
                        // defined on every path stay live across the whole event loop -- 2.6 KB of spills)
                        V y[GC * R1];
#pragma unroll
                        for (int j = 0; j < GC * R1; ++j) y[j] = V{(T)0.0f, (T)0.0f};
                        Core::inv_pass1m(sm.buf, park, prm.tw1, i0, y, computed);
#pragma unroll
                        for (int q = 0; q < DP_MAX_TSLOTS; ++q) {
                            if (q < nts) {
                                const DpSlot sl = ch.slots[slot_of[q]];
                                if (sl.lo == 0 && sl.hi == N && !sl.outside)
                                    Dp2Scan<T, R1>::full(y, tid, i0, tb[q]);
                                else
                                    Dp2Scan<T, R1>::window(y, tid, i0, sl.lo, (unsigned)(sl.hi - sl.lo), sl.outside != 0, tb[q], computed);
                            }
                        }
                    }
                    }
                    DpBest<S>* best = sm.best(par);
                    double* red = sm.red(par);
#pragma unroll
                    for (int q = 0; q < DP_MAX_TSLOTS; ++q) {
                        if (q < nts) {
                            const DpBest<S> b = dp_warp_best(tb[q]);
                            if ((tid & 31) == 0) best[q * 32 + (tid >> 5)] = b;
                            if (tid == 0) sm.slot_id(par)[q] = slot_of[q];
                        }
                    }
                    __syncthreads();  // winners + the lowchi2 stash are visible; pass-1' reads of buf are done
                    if (prm.neighbours) {
                        // amplitude one sample before / after each fit's best delay: the thread that owns the sample's column
                        // runs pass 1' for it once more (the pass-2' rows are still in the buffer) and leaves the value where
                        // thread 0 forms the outputs
#pragma unroll
                        for (int q = 0; q < DP_MAX_TSLOTS; ++q) {
                            if (q >= nts) continue;
                            DpBest<S> bb = best[q * 32];
                            for (int w = 1; w < NW; ++w) dp_best_merge(bb, best[q * 32 + w]);
#pragma unroll
                            for (int side = 0; side < 2; ++side) {
                                const int r = bb.idx + (side ? 1 : -1);
                                const bool ok = bb.idx >= 0 && r >= 0 && r < N;
                                const int c = ok ? (((r >> 1) & 4095) / VL) : 0;
                                if (tid == (c % NT)) {
                                    double v = NAN;
                                    if (ok) {
                                        V u[R1];
                                        Core::inv_pass1_col(sm.buf, park, prm.tw1, c, u);
                                        v = (double)Core::sample_of(u, r);
                                    }
                                    sm.neigh(par)[2 * q + side] = v;
                                }
                            }
                        }
                        if (more) __syncthreads();  // these reads of the buffer precede the next template's bulk copies into it
                    }
                    // the FFT buffer is free: the next template's X column and filter rows land while the outputs are formed
                    if (STAGE && more && lane == 0)
                        issue_b(bar, bufb, prm.zones[p * NW + warp], spw, TMX ? ch.templ[it + 1].phi + tab_block(p, warp) : scr_x + (long long)warp * 16 * 32, ch.templ[it + 1].phi + tab_block(p, warp));
                    // ---- low-frequency chi2 at each fit's (amp, delay); chi0; outputs ----------------
                    double part[DP_MAX_TSLOTS + 1];
#pragma unroll
                    for (int q = 0; q < DP_MAX_TSLOTS; ++q) {
                        part[q] = 0.0;
                        const int nlow_q = q < nts ? ch.slots[slot_of[q]].nlow : 0;  // the fit's own lowchi2_fcutoff
                        if (q < nts && tid < nlow_q) {
                            DpBest<S> b = best[q * 32];
                            for (int w = 1; w < NW; ++w) dp_best_merge(b, best[q * 32 + w]);
                            const int d = b.idx - tp.pretrigger;
                            for (int k = tid; k < nlow_q; k += NT) {
                                const int ph = (int)((((long long)k * (long long)d) % N + N) % N);  // exp(-2 pi i k d / N)
                                S sn, cs;
                                if constexpr (sizeof(S) == 8) {
                                    double s_, c_;
                                    sincospi(2.0 * (double)ph / (double)N, &s_, &c_);
                                    sn = (S)s_;
                                    cs = (S)c_;
                                } else {
                                    float s_, c_;
                                    sincospif(2.0f * (float)ph / (float)N, &s_, &c_);
                                    sn = (S)s_;
                                    cs = (S)c_;
                                }
                                const cx<S> mdl = cmul(cx<S>{cs, -sn}, dp_ldg(tp.s_low + k));
                                const cx<S> X = sm.stash[k];
                                const cx<S> R = cx<S>{dp_fma(-b.val, mdl.re, X.re), dp_fma(-b.val, mdl.im, X.im)};
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
                    // only thread 0 needs the partial sums: warp 0 waits for them, the other warps check in and go on
                    // (next template / next event; red / best are double buffered by `par`)
                    if (warp == 0)
                        dp_bar_sync(15, NT);
                    else
                        dp_bar_arrive(15, NT);
                    if (tid == 0) {
                        double* o = prm.out + (long long)ev * prm.n_out + ch.out_base;
                        if (it == 0) {
                            double c0 = 0.0;
                            for (int w = 0; w < NW; ++w) c0 += red[DP_MAX_TSLOTS * 32 + w];
                            *chi0_keep = c0;
                            o[0] = c0;
                        }
                        const double chi0 = *chi0_keep;
                        for (int q = 0; q < nts; ++q) {
                            double low = 0.0;
                            for (int w = 0; w < NW; ++w) low += red[q * 32 + w];
                            DpBest<S> b = best[q * 32];
                            for (int w = 1; w < NW; ++w) dp_best_merge(b, best[q * 32 + w]);
                            double* os = o + 1 + sm.slot_id(par)[q] * DP_SLOT_NOUT;
                            const double amp = (double)b.val;
                            os[0] = amp;
                            os[1] = (double)b.idx;
                            os[2] = chi0 - amp * amp * tp.norm;
                            os[3] = low;
                            os[4] = 1.0 / sqrt(amp * amp * tp.tsum);
                            if (prm.neighbours) {
                                double* on = o + 1 + DP_SLOT_NOUT * ch.n_slots + 2 * sm.slot_id(par)[q];
                                on[0] = sm.neigh(par)[2 * q];
                                on[1] = sm.neigh(par)[2 * q + 1];
                            }
                        }
                    }
                    par ^= 1;
                }
                if (!more) break;
                ++it;
                // ---------------- next template: X (staged) times its filter, inverse untangle -> z ----------------
                if constexpr (STAGE) {
                    dp2_stage_wait(bar, spar);
                    spar ^= 1;
                }
                {
                    const Staged tb = staged_view(bufb, prm.zones[p * NW + warp], spoff, lane, false, ch.templ[it].phi + tab_block(p, warp) + lane,
                                                  ch.wj + tab_block(p, warp) + lane, xs);
                    if constexpr (VL == 2) {
#pragma unroll
                        for (int r = 0; r < 16; ++r) z[r] = tb.x(r);
                        cx<S> wn2 = wn;
#if !defined(DP2_WN_CSE) && !defined(DP_HOST_EMU)
                        asm volatile("" : "+f"(wn2.re), "+f"(wn2.im));   // see the fp64 branch
#endif
                        filter_st<false>(z, tb, wn2);
                    } else {
                        cx<S> wn2 = wn;
#if !defined(DP2_WN_CSE) && !defined(DP_HOST_EMU)
                        // the pair twiddles are recomputed from an opaque copy: common-subexpression elimination with the
                        // first template's round kept all eight of them alive across the inverse passes (14 spilled doubles)
                        asm volatile("" : "+d"(wn2.re), "+d"(wn2.im));
#endif
                        if constexpr (TMX) {
#ifndef DP_HOST_EMU
                            // X pairs out of TMEM, four at a time; the staged rows hold this template's filter
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                DpTmemRaw8 xr[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) xr[j] = dp_tmem_ld8(tm_thread + XCOL + 8 * (4 * h + j));
                                dp_tmem_wait_ld();
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const int r = 4 * h + j;
                                    const cx<S> w = cmul(wn2, dp_w64_rt<S>(2 * r));
                                    pair_filter(z, r, dp_tmem_cx(xr[j], 0), dp_tmem_cx(xr[j], 1), tb.x(2 * r), tb.x(2 * r + 1), w);
                                }
                            }
#endif
                        } else {
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            const cx<S> w = cmul(wn2, dp_w64_rt<S>(2 * r));
                            pair_filter(z, r, tb.x(2 * r), tb.x(2 * r + 1), tb.phi_b(2 * r), tb.phi_b(2 * r + 1), w);
                        }
                        }
                    }
                }
                __syncwarp();  // staged rows are consumed before pass 4' stores into them
            }
        }
    }
#ifndef DP_HOST_EMU
    if constexpr (TM) {
        dp_tmem_fence_before();
        __syncthreads();
        dp_tmem_fence_after();
        if (warp == 0) dp_tmem_dealloc512(*tmslot);
    }
#endif
}

#ifndef DP_HOST_EMU
template <class T, int R1, int IN, bool MULTI, bool NARROW = false>
__global__ void __launch_bounds__(Dp2Geom<T, R1>::NT, Dp2Geom<T, R1>::NT <= 256 ? 2 : 1) dp_of2_kernel(const Dp2Params<T> prm) {
    extern __shared__ __align__(16) unsigned char dp_smem_raw[];
    Dp2OfKernel<T, R1, IN>::template run<MULTI, NARROW ? 1 : 0>(prm, dp_smem_raw);
}
#endif
