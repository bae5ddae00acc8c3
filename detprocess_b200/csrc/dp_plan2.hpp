// Host-side table construction for the v2 kernels (dp_of2_kernel.cuh): twiddle bases in
// packed-lane order, the (group, mirror group) assignment of every thread and phase, and the
// filter / chi0-weight tables in thread order.  The optimal-filter maths (s, phi, norm, wJ,
// the hermitian-symmetrised one-sided filter) is shared with dp_plan.hpp and mirrors
// ProcessingData.instantiate_OF_base (reference detprocess/process/processing_data.py:275-381).
#pragma once
#include "dp_of2_kernel.cuh"
#include "dp_plan.hpp"
#include <map>

namespace dpplan2 {

using dpplan::cplx;

template <class T> struct Pack;
template <> struct Pack<double> {
    static cx<double> c(const cplx* z) { return cx<double>{z[0].real(), z[0].imag()}; }
    static double r(const double* v) { return v[0]; }
};
template <> struct Pack<f2> {
    static cx<f2> c(const cplx* z) {
        return cx<f2>{f2((float)z[0].real(), (float)z[1].real()), f2((float)z[0].imag(), (float)z[1].imag())};
    }
    static f2 r(const double* v) { return f2((float)v[0], (float)v[1]); }
};

template <class T> struct Tables2 {
    using S = typename Dp2Traits<T>::S;
    std::vector<cx<T>> tw1, tw2, tw3;
    std::vector<cx<S>> twn;
    std::vector<int2> groups;
    std::vector<int> chunk3;   // [NPH][NT] chunk (b*16 + k2) a thread transforms in the warp-local passes 3 / 3'
    std::vector<uint4> zones;  // [NPH][NW] byte offsets of the warp's four 2048-byte landing pieces in the FFT buffer
    struct Templ {
        std::vector<cx<T>> phi;
        std::vector<cx<S>> phi_self, s_low;
        double norm, tsum;
        int pretrigger;
    };
    struct Chan {
        std::vector<T> wj;
        std::vector<S> wj_self, wj_low;
        std::vector<Templ> templ;
    };
    std::vector<Chan> chans;
    int nlow = 0;
    double scale = 1.0;
};

// is nb_samples handled by the v2 kernels?
inline int r1_of(int N) { return (N == 16384) ? 2 : (N == 32768) ? 4 : (N == 65536) ? 8 : 0; }

// bins of self-pair lane l (0..16): (k, M-k) with the DC pair's mirror = Nyquist bin M
template <class G> inline void self_bins(int l, int* bins, bool* dup) {
    const int kp = (l < 9) ? G::KQ * l : G::KQ / 2 + G::KQ * (l - 9);
    bins[0] = kp;
    bins[1] = (kp == 0) ? G::M : G::M - kp;
    dup[0] = false;
    dup[1] = bins[1] == bins[0];
}

// bins (lane 0 / lane 1) of table entry e of thread t in phase p
template <class G> inline void entry_bins(int p, int t, int e, int* bins) {
    int Ga, Gb;
    G::groups_of(p, t, Ga, Gb);
    if (G::VL == 2) {
        bins[0] = G::bin_of(p, Ga, e);
        bins[1] = G::bin_of(p, Gb, e);
    } else {
        const int r = e >> 1;
        const int k = G::bin_of(p, Ga, r);
        bins[0] = (e & 1) ? (G::M - k) : k;
        bins[1] = bins[0];
    }
}

// every one-sided bin 0..M must be produced exactly once (regular entries + self lanes) and
// every regular pair must be a true (k, M-k) mirror pair
template <class G> inline void verify_geometry() {
    std::vector<int> seen(G::M + 1, 0);
    const int nspecial = (G::VL == 2) ? 1 : 2;
    for (int p = 0; p < G::NPH; ++p)
        for (int t = 0; t < G::NT; ++t) {
            if (p == 0 && t < nspecial) continue;
            int Ga, Gb;
            G::groups_of(p, t, Ga, Gb);
            if (Ga < 0 || Gb < 0) throw std::logic_error("v2 geometry: unassigned thread");
            for (int r = 0; r < 16; ++r)
                if (G::bin_of(p, Gb, 15 - r) != G::M - G::bin_of(p, Ga, r)) throw std::logic_error("v2 geometry: not a mirror pair");
            for (int e = 0; e < 16; ++e) {
                int b[2];
                entry_bins<G>(p, t, e, b);
                for (int l = 0; l < G::VL; ++l) {
                    if (b[l] < 0 || b[l] > G::M) throw std::logic_error("v2 geometry: bin out of range");
                    ++seen[b[l]];
                }
            }
        }
    for (int l = 0; l < 17; ++l) {
        int b[2];
        bool d[2];
        self_bins<G>(l, b, d);
        for (int j = 0; j < 2; ++j)
            if (!d[j]) ++seen[b[j]];
    }
    for (int k = 0; k <= G::M; ++k)
        if (seen[k] != 1) throw std::logic_error("v2 geometry: bin " + std::to_string(k) + " covered " + std::to_string(seen[k]) + " times");
}

// natural one-sided bin k -> slot in a CTA's thread-order partial-sum array (PSD / CSD accumulation kernels):
// regular entries [NPH][16][NT][VL], then the 17 self lanes [17][2]
template <class G> inline std::vector<int> partial_slot_of_bin() {
    std::vector<int> loc(G::M + 1, -1);
    const int nspecial = G::VL == 2 ? 1 : 2;
    for (int ph = 0; ph < G::NPH; ++ph)
        for (int t = 0; t < G::NT; ++t) {
            if (ph == 0 && t < nspecial) continue;
            for (int e = 0; e < 16; ++e) {
                int b[2];
                entry_bins<G>(ph, t, e, b);
                for (int l = 0; l < G::VL; ++l) loc[b[l]] = ((ph * 16 + e) * G::NT + t) * G::VL + l;
            }
        }
    for (int l = 0; l < 17; ++l) {
        int b[2];
        bool dup[2];
        self_bins<G>(l, b, dup);
        for (int j = 0; j < 2; ++j)
            if (!dup[j]) loc[b[j]] = G::NPH * 16 * G::NT * G::VL + 2 * l + j;
    }
    for (int k = 0; k <= G::M; ++k)
        if (loc[k] < 0) throw std::logic_error("v2 geometry: partial-sum bin map incomplete");
    return loc;
}


// position of (phase p, entry e, thread t) in the OF kernel's thread-order tables [NPH][NW][16][32]: the 16 rows of
// a warp are one contiguous block (4 KB of chi0 weights, 8 KB of filter values) that the warp fetches by bulk copy
template <class G> inline size_t ws_index(int p, int e, int t) { return ((((size_t)p * (G::NT / 32) + t / 32) * 16 + e) * 32) + (t % 32); }

// Warp-local passes 3 / 3' (Dp2Core::fwd_3w / inv_3w): every warp must own whole 256-element chunks in pass 4 (all
// 16 groups of the chunk, i.e. a chunk together with its mirror chunk); thread lane l of the warp then transforms
// chunk S[l / GV] in pass 3, S = the warp's chunks in ascending order.  The same chunks are the shared memory the
// warp may overwrite between pass 4 and pass 4' (its own group rows): cut into 2048-byte landing pieces.
template <class G> inline void build_chunks(std::vector<int>& chunk3, std::vector<uint4>& zones) {
    constexpr int NT = G::NT, NW = NT / 32, GV = G::GV, NPH = G::NPH;
    constexpr int CPW = 32 / GV;                        // chunks per warp
    constexpr unsigned CHUNK_BYTES = 16u * (GV + 1) * 16u;  // 16 padded group rows of GV + 1 vectors
    constexpr unsigned PIECE = 2048;
    static_assert(CPW * (CHUNK_BYTES / PIECE) == 4, "a warp's chunks must hold four landing pieces");
    chunk3.assign((size_t)NPH * NT, -1);
    zones.assign((size_t)NPH * NW, uint4{0, 0, 0, 0});
    for (int p = 0; p < NPH; ++p) {
        std::vector<int> owner(G::NB * 16, -1);
        for (int w = 0; w < NW; ++w) {
            std::map<int, int> cnt;
            for (int l = 0; l < 32; ++l) {
                int Ga, Gb;
                G::groups_of(p, w * 32 + l, Ga, Gb);
                ++cnt[Ga >> 4];
                if (G::VL == 2) ++cnt[Gb >> 4];
            }
            if ((int)cnt.size() != CPW) throw std::logic_error("v2 geometry: a warp does not own whole chunks");
            std::vector<int> S;
            for (const auto& kv : cnt) {
                if (kv.second != 16) throw std::logic_error("v2 geometry: chunk split between warps");
                if (owner[kv.first] >= 0) throw std::logic_error("v2 geometry: chunk owned twice");
                owner[kv.first] = w;
                S.push_back(kv.first);
            }
            for (int l = 0; l < 32; ++l) chunk3[(size_t)p * NT + w * 32 + l] = S[l / GV];
            unsigned off[4];
            int n = 0;
            for (int c : S)
                for (unsigned q = 0; q + PIECE <= CHUNK_BYTES; q += PIECE) off[n++] = (unsigned)c * CHUNK_BYTES + q;
            zones[(size_t)p * NW + w] = uint4{off[0], off[1], off[2], off[3]};
        }
        for (int c = 0; c < G::NB * 16; ++c)
            if (owner[c] < 0) throw std::logic_error("v2 geometry: chunk without a warp");
        // the set barrier between pass 2 and pass 3 must cover the blocks of every chunk a thread reads
        for (int t = 0; t < NT; ++t) {
            const int b = chunk3[(size_t)p * NT + t] >> 4, tb = t / G::CV;
            if (b != tb && b != G::mirror_b(p, tb)) throw std::logic_error("v2 geometry: chunk outside the thread's block set");
        }
    }
}

// a one-sided filter pe[0..M] (applied to 2*sc*fft(x), see dpplan::filter_onesided) in thread order
template <class T, int R1>
void pack_onesided(const std::vector<cplx>& pe, std::vector<cx<T>>& phi, std::vector<cx<typename Dp2Traits<T>::S>>& phi_self) {
    using G = Dp2Geom<T, R1>;
    using S = typename G::S;
    phi.resize((size_t)G::NPH * 16 * G::NT);
    for (int p = 0; p < G::NPH; ++p)
        for (int e = 0; e < 16; ++e)
            for (int t = 0; t < G::NT; ++t) {
                int b[2];
                entry_bins<G>(p, t, e, b);
                const cplx z[2] = {pe[b[0]], pe[b[1]]};
                phi[((size_t)p * 16 + e) * G::NT + t] = Pack<T>::c(z);
            }
    phi_self.resize(17 * 2);
    for (int l = 0; l < 17; ++l) {
        int b[2];
        bool dd[2];
        self_bins<G>(l, b, dd);
        for (int j = 0; j < 2; ++j) phi_self[l * 2 + j] = cx<S>{(S)pe[b[j]].real(), (S)pe[b[j]].imag()};
    }
}

template <class T, int R1>
Tables2<T> build_tables2(double fs, const std::vector<dpplan::Channel>& chans, double fcut, double scale) {
    using G = Dp2Geom<T, R1>;
    using S = typename G::S;
    constexpr int VL = G::VL, NT = G::NT, NPH = G::NPH, N = G::N, M = G::M;
    verify_geometry<G>();
    Tables2<T> dt;
    const double df = fs / N;
    dt.scale = scale;
    dt.nlow = dpplan::count_low_bins(N, fs, fcut);
    if (dt.nlow > 256 * R1 || dt.nlow > DP_NLOW_MAX)
        throw std::invalid_argument("lowchi2_fcutoff too high for the fused kernel (needs <= " +
                                    std::to_string(std::min(256 * R1, DP_NLOW_MAX)) + " bins)");
    auto lanes = [&](auto fn) {  // V from a per-lane complex value
        cplx z[2];
        for (int l = 0; l < VL; ++l) z[l] = fn(l);
        return Pack<T>::c(z);
    };
    auto cxS = [](cplx z) { return cx<S>{(S)z.real(), (S)z.imag()}; };
    dt.tw1.resize(G::VPB);
    for (int c = 0; c < G::VPB; ++c) dt.tw1[c] = lanes([&](int l) { return dpplan::unit_root(VL * c + l, M); });
    dt.tw2.resize(G::CV);
    for (int c = 0; c < G::CV; ++c) dt.tw2[c] = lanes([&](int l) { return dpplan::unit_root(VL * c + l, 4096); });
    dt.tw3.resize(G::GV);
    for (int c = 0; c < G::GV; ++c) dt.tw3[c] = lanes([&](int l) { return dpplan::unit_root(VL * c + l, 256); });
    dt.twn.resize((size_t)NPH * NT);
    dt.groups.resize((size_t)NPH * NT);
    for (int p = 0; p < NPH; ++p)
        for (int t = 0; t < NT; ++t) {
            int Ga, Gb;
            G::groups_of(p, t, Ga, Gb);
            dt.groups[(size_t)p * NT + t] = int2{Ga, Gb};
            dt.twn[(size_t)p * NT + t] = cxS(dpplan::unit_root(G::bin_of(p, Ga, 0), N));
        }
    build_chunks<G>(dt.chunk3, dt.zones);
    for (const auto& ch : chans) {
        typename Tables2<T>::Chan dc;
        if ((int)ch.J.size() != N) throw std::invalid_argument("psd not set for a channel");
        const std::vector<double> wJ = dpplan::chi0_weights(ch.J, fs, scale);
        dc.wj.resize((size_t)NPH * 16 * NT);
        for (int p = 0; p < NPH; ++p)
            for (int e = 0; e < 16; ++e)
                for (int t = 0; t < NT; ++t) {
                    int b[2];
                    entry_bins<G>(p, t, e, b);
                    const double v[2] = {wJ[b[0]], wJ[b[1]]};
                    dc.wj[ws_index<G>(p, e, t)] = Pack<T>::r(v);
                }
        dc.wj_self.resize(17 * 2);
        for (int l = 0; l < 17; ++l) {
            int b[2];
            bool d[2];
            self_bins<G>(l, b, d);
            for (int j = 0; j < 2; ++j) dc.wj_self[l * 2 + j] = d[j] ? (S)0 : (S)wJ[b[j]];
        }
        dc.wj_low.resize(dt.nlow);
        for (int k = 0; k < dt.nlow; ++k) dc.wj_low[k] = (S)wJ[k];
        for (const auto& tp : ch.templ) {
            typename Tables2<T>::Templ d;
            const std::vector<cplx> pe = dpplan::filter_onesided(tp, scale);
            d.phi.resize((size_t)NPH * 16 * NT);
            for (int p = 0; p < NPH; ++p)
                for (int e = 0; e < 16; ++e)
                    for (int t = 0; t < NT; ++t) {
                        int b[2];
                        entry_bins<G>(p, t, e, b);
                        const cplx z[2] = {pe[b[0]], pe[b[1]]};
                        d.phi[ws_index<G>(p, e, t)] = Pack<T>::c(z);
                    }
            d.phi_self.resize(17 * 2);
            for (int l = 0; l < 17; ++l) {
                int b[2];
                bool dd[2];
                self_bins<G>(l, b, dd);
                for (int j = 0; j < 2; ++j) d.phi_self[l * 2 + j] = cxS(pe[b[j]]);
            }
            d.s_low.resize(dt.nlow);
            for (int k = 0; k < dt.nlow; ++k) d.s_low[k] = cxS(tp.s[k] * ((double)N * df) * 2.0 * scale);
            d.norm = tp.norm;
            d.tsum = tp.tsum;
            d.pretrigger = tp.pretrigger;
            dc.templ.push_back(std::move(d));
        }
        dt.chans.push_back(std::move(dc));
    }
    return dt;
}

}  // namespace dpplan2
