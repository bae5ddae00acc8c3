// Host-side tables of the mixed-radix OF kernel (dp_ofg_kernel.cuh): radix list, twiddles, the digit-reversed
// positions of the (k, M - k) pairs and the filter / chi0-weight values in pair order.  The optimal-filter maths
// (s, phi, norm, wJ, the hermitian-symmetrised one-sided filter) is shared with dp_plan.hpp and mirrors
// ProcessingData.instantiate_OF_base (reference detprocess/process/processing_data.py:275-381).
#pragma once
#include "dp_ofg_kernel.cuh"
#include "dp_plan.hpp"

namespace dpgen {

using dpplan::cplx;

// radices of M (product == M), fives first: empty when M has another prime factor
inline std::vector<int> radices_of(int M) {
    std::vector<int> r;
    int m = M;
    while (m % 5 == 0) r.push_back(5), m /= 5;
    while (m % 4 == 0) r.push_back(4), m /= 4;
    while (m % 3 == 0) r.push_back(3), m /= 3;
    while (m % 2 == 0) r.push_back(2), m /= 2;
    if (m != 1 || (int)r.size() > DPG_MAX_PASSES) r.clear();
    return r;
}
// is nb_samples served by the mixed-radix kernel? (even, M = N/2 factorises, the event fits one CTA's shared memory)
inline bool supported(int N, bool f64) {
    if (N < 64 || (N & 1)) return false;
    const int M = N / 2;
    if (radices_of(M).empty()) return false;
    const size_t smem = f64 ? DpGenKernel<double>::smem_bytes(M) : DpGenKernel<float>::smem_bytes(M);
    return smem <= 227u * 1024u;
}
// position of spectrum bin k after the decimation-in-frequency passes with radices r_1 .. r_p:
// k = k_1 + r_1 (k_2 + r_2 (...)),  position = sum_j k_j * M / (r_1 ... r_j)
inline int position_of(int k, int M, const std::vector<int>& rad) {
    int pos = 0, L = M;
    for (int r : rad) {
        L /= r;
        pos += (k % r) * L;
        k /= r;
    }
    return pos;
}

template <class T> struct Tables {
    std::vector<int> radix;
    std::vector<cx<T>> tw, wn;
    std::vector<int> pos_k, pos_m;
    struct Templ {
        std::vector<cx<T>> phi_k, phi_m, s_low;
        double norm, tsum;
        int pretrigger;
    };
    struct Chan {
        std::vector<T> wj_k, wj_m, wj_low;
        std::vector<Templ> templ;
    };
    std::vector<Chan> chans;
    int M = 0, n_pairs = 0, nlow = 0;
    double scale = 1.0;
};

template <class T> Tables<T> build_tables(int N, double fs, const std::vector<dpplan::Channel>& chans, double fcut, double scale) {
    Tables<T> dt;
    const int M = N / 2;
    dt.M = M;
    dt.radix = radices_of(M);
    if (dt.radix.empty()) throw std::invalid_argument("nb_samples / 2 must factor into 2, 3, 4, 5");
    dt.scale = scale;
    dt.n_pairs = M / 2 + 1;
    dt.nlow = dpplan::count_low_bins(N, fs, fcut);
    if (dt.nlow > DP_NLOW_MAX || dt.nlow > dt.n_pairs - 1)
        throw std::invalid_argument("lowchi2_fcutoff too high for the fused kernel");
    const double df = fs / N;
    auto cxT = [](cplx z) { return cx<T>{(T)z.real(), (T)z.imag()}; };
    dt.tw.resize(M);
    for (int j = 0; j < M; ++j) dt.tw[j] = cxT(dpplan::unit_root(j, M));
    dt.pos_k.resize(dt.n_pairs);
    dt.pos_m.resize(dt.n_pairs);
    dt.wn.resize(dt.n_pairs);
    std::vector<int> seen(M, 0);
    for (int k = 0; k < dt.n_pairs; ++k) {
        dt.pos_k[k] = position_of(k, M, dt.radix);
        dt.pos_m[k] = position_of((M - k) % M, M, dt.radix);
        dt.wn[k] = cxT(dpplan::unit_root(k, N));
        ++seen[dt.pos_k[k]];
        if (dt.pos_m[k] != dt.pos_k[k]) ++seen[dt.pos_m[k]];
    }
    for (int p = 0; p < M; ++p)
        if (seen[p] != 1) throw std::logic_error("mixed-radix geometry: position covered " + std::to_string(seen[p]) + " times");
    for (const auto& ch : chans) {
        typename Tables<T>::Chan dc;
        if ((int)ch.J.size() != N) throw std::invalid_argument("psd not set for a channel");
        const std::vector<double> wJ = dpplan::chi0_weights(ch.J, fs, scale);
        dc.wj_k.resize(dt.n_pairs);
        dc.wj_m.resize(dt.n_pairs);
        for (int k = 0; k < dt.n_pairs; ++k) {
            dc.wj_k[k] = (T)wJ[k];
            // mirror bin of the pair: the Nyquist bin M for the DC pair, nothing for the self-mirrored bin M/2
            const int km = (k == 0) ? M : M - k;
            dc.wj_m[k] = (km == k) ? (T)0 : (T)wJ[km];
        }
        dc.wj_low.resize(dt.nlow);
        for (int k = 0; k < dt.nlow; ++k) dc.wj_low[k] = (T)wJ[k];
        for (const auto& tp : ch.templ) {
            typename Tables<T>::Templ d;
            const std::vector<cplx> pe = dpplan::filter_onesided(tp, scale);
            d.phi_k.resize(dt.n_pairs);
            d.phi_m.resize(dt.n_pairs);
            for (int k = 0; k < dt.n_pairs; ++k) {
                d.phi_k[k] = cxT(pe[k]);
                d.phi_m[k] = cxT(pe[(k == 0) ? M : M - k]);
            }
            d.s_low.resize(dt.nlow);
            for (int k = 0; k < dt.nlow; ++k) d.s_low[k] = cxT(tp.s[k] * ((double)N * df) * 2.0 * scale);
            d.norm = tp.norm;
            d.tsum = tp.tsum;
            d.pretrigger = tp.pretrigger;
            dc.templ.push_back(std::move(d));
        }
        dt.chans.push_back(std::move(dc));
    }
    return dt;
}

}  // namespace dpgen
