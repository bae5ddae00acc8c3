// Host-side setup of the fused NxM optimal filter (dp_nxm_kernel.cuh): what qp.OFBase.calc_phi_matrix /
// calc_weight_matrix / calc_p_and_p_inverse prepare for qp.OFnxm (driven from ProcessingData.instantiate_OF_base,
// reference detprocess/process/processing_data.py:275-381, and FeatureExtractors.ofnxm, algorithms.py:141-274):
//     S[a][i][k]   = fft(template of channel a, amplitude i) / N / df
//     iS[k]        = inverse of the n x n cross-spectral density at bin k (AC coupling: iS[0] = 0)
//     Phi[i][a][k] = sum_b conj(S[b][i][k]) iS[k][b][a]
//     P[i][j]      = Re sum_{a,k} Phi[i][a][k] S[a][j][k] df
// packed in the kernel's thread order with the hermitian fold, the roll to `pretrigger` and all scalings folded in.
#pragma once
#include "dp_nxm_kernel.cuh"
#include "dp_plan2.hpp"

namespace dpnxm {

using dpplan::cplx;

struct Setup {
    int N = 0, n = 0, m = 0, pretrigger = 0;
    double fs = 0;
    std::vector<std::vector<cplx>> phi;  // [m*n][N]   Phi[i][a][k]
    std::vector<std::vector<cplx>> isig; // [n*n][N]   iS[k][a][b]
    std::vector<double> P, Pinv;         // [m*m]
};

// in-place inverse of a small complex / real matrix (Gauss-Jordan with partial pivoting); false if (numerically) singular
template <class Z> inline bool invert(std::vector<Z>& a, int n) {
    std::vector<Z> inv((size_t)n * n, Z(0));
    for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = Z(1);
    // a pivot below 1e-12 of the largest entry of its (original) column counts as singular: rank-deficient csd,
    // degenerate templates
    std::vector<double> colmax(n, 0.0);
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) colmax[c] = std::max(colmax[c], (double)std::abs(a[(size_t)r * n + c]));
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r)
            if (std::abs(a[(size_t)r * n + c]) > std::abs(a[(size_t)piv * n + c])) piv = r;
        if (!(std::abs(a[(size_t)piv * n + c]) > 1e-12 * colmax[c])) return false;
        if (piv != c)
            for (int k = 0; k < n; ++k) {
                std::swap(a[(size_t)piv * n + k], a[(size_t)c * n + k]);
                std::swap(inv[(size_t)piv * n + k], inv[(size_t)c * n + k]);
            }
        const Z d = Z(1) / a[(size_t)c * n + c];
        for (int k = 0; k < n; ++k) {
            a[(size_t)c * n + k] *= d;
            inv[(size_t)c * n + k] *= d;
        }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const Z f = a[(size_t)r * n + c];
            if (f == Z(0)) continue;
            for (int k = 0; k < n; ++k) {
                a[(size_t)r * n + k] -= f * a[(size_t)c * n + k];
                inv[(size_t)r * n + k] -= f * inv[(size_t)c * n + k];
            }
        }
    }
    a.swap(inv);
    return true;
}

// templates: [n][m][N] float64; csd: [n][n][N] complex (re, im interleaved), two-sided, fftfreq order
inline Setup make_setup(int N, double fs, int n, int m, const double* templates, const double* csd, int pretrigger, bool coupling_ac) {
    Setup s;
    s.N = N;
    s.n = n;
    s.m = m;
    s.fs = fs;
    s.pretrigger = pretrigger;
    const double df = fs / N;
    std::vector<std::vector<cplx>> S((size_t)n * m);  // [a*m + i]
    for (int a = 0; a < n; ++a)
        for (int i = 0; i < m; ++i) {
            std::vector<cplx> t(N);
            const double* src = templates + ((size_t)a * m + i) * N;
            for (int k = 0; k < N; ++k) t[k] = cplx(src[k], 0.0);
            dpplan::fft_pow2(t);
            for (auto& v : t) v = v / (double)N / df;
            S[(size_t)a * m + i] = std::move(t);
        }
    s.isig.assign((size_t)n * n, std::vector<cplx>(N));
    std::vector<cplx> mat((size_t)n * n);
    for (int k = 0; k < N; ++k) {
        if (k == 0 && coupling_ac) {
            for (int ab = 0; ab < n * n; ++ab) s.isig[ab][0] = cplx(0, 0);
            continue;
        }
        for (int ab = 0; ab < n * n; ++ab) mat[ab] = cplx(csd[((size_t)ab * N + k) * 2], csd[((size_t)ab * N + k) * 2 + 1]);
        // a bin whose csd diagonal is not finite carries no weight (ignored_frequency_peaks of OFBase.set_csd,
        // reference processing_data.py:321-326: the host marks the peak bins with inf)
        bool ignored = false;
        for (int a = 0; a < n; ++a) ignored = ignored || !std::isfinite(mat[(size_t)a * n + a].real());
        if (ignored) {
            for (int ab = 0; ab < n * n; ++ab) s.isig[ab][k] = cplx(0, 0);
            continue;
        }
        if (!invert(mat, n)) throw std::invalid_argument("csd is singular at bin " + std::to_string(k));
        for (int ab = 0; ab < n * n; ++ab) s.isig[ab][k] = mat[ab];
    }
    s.phi.assign((size_t)m * n, std::vector<cplx>(N));
    for (int i = 0; i < m; ++i)
        for (int a = 0; a < n; ++a)
            for (int k = 0; k < N; ++k) {
                cplx acc(0, 0);
                for (int b = 0; b < n; ++b) acc += std::conj(S[(size_t)b * m + i][k]) * s.isig[(size_t)b * n + a][k];
                s.phi[(size_t)i * n + a][k] = acc;
            }
    s.P.assign((size_t)m * m, 0.0);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            cplx acc(0, 0);
            for (int a = 0; a < n; ++a)
                for (int k = 0; k < N; ++k) acc += s.phi[(size_t)i * n + a][k] * S[(size_t)a * m + j][k];
            s.P[(size_t)i * m + j] = acc.real() * df;
        }
    s.Pinv = s.P;
    if (!invert(s.Pinv, m)) throw std::invalid_argument("template matrix P is singular (degenerate templates)");
    for (int i = 0; i < m; ++i)
        if (!(s.P[(size_t)i * m + i] > 0)) throw std::invalid_argument("template matrix P is not positive");
    return s;
}

// typical sample rms from the csd diagonal (for the fp32 pre-scale)
inline double typical_rms(const Setup& s, const double* csd) {
    double sum = 0;
    long long cnt = 0;
    for (int a = 0; a < s.n; ++a)
        for (int k = 1; k < s.N; ++k) {
            const double v = csd[(((size_t)a * s.n + a) * s.N + k) * 2];
            if (std::isfinite(v)) {
                sum += v;
                ++cnt;
            }
        }
    return std::sqrt(std::max(sum / std::max<long long>(cnt, 1) * s.fs, 1e-300));
}

template <class T> struct Tables {
    using S = typename Dp2Traits<T>::S;
    std::vector<cx<T>> tw1, tw2, tw3;
    std::vector<cx<S>> twn;
    std::vector<int2> groups;
    std::vector<int> chunk3;   // per-thread pass-3 chunk ids (warp-local passes)
    std::vector<cx<T>> g;       // [m][n][NPH*16*NT]
    std::vector<cx<S>> g_self;  // [m][n][17*2]
    std::vector<std::vector<T>> wd;          // [n]
    std::vector<std::vector<S>> wd_self;
    std::vector<std::vector<cx<T>>> wo;      // [pairs a < b]
    std::vector<std::vector<cx<S>>> wo_self;
    double cmat[DP_NXM_MAX_TEMPL][DP_NXM_MAX_TEMPL] = {};
    double amat[DP_NXM_MAX_TEMPL][DP_NXM_MAX_TEMPL] = {};
};

template <class T, int R1> Tables<T> build_tables(const Setup& s, double scale) {
    using G = Dp2Geom<T, R1>;
    using S = typename G::S;
    constexpr int NT = G::NT, NPH = G::NPH, N = G::N, M = G::M;
    if (s.N != N) throw std::logic_error("nxm: geometry mismatch");
    Tables<T> dt;
    {   // twiddles and group assignment: the same tables as the 1x1 plans
        std::vector<dpplan::Channel> none;
        dpplan2::Tables2<T> base = dpplan2::build_tables2<T, R1>(s.fs, none, 0.0, scale);
        dt.tw1 = std::move(base.tw1);
        dt.tw2 = std::move(base.tw2);
        dt.tw3 = std::move(base.tw3);
        dt.twn = std::move(base.twn);
        dt.groups = std::move(base.groups);
        dt.chunk3 = std::move(base.chunk3);
    }
    const double df = s.fs / N;
    const int n = s.n, m = s.m;
    // filters: q~_i(t) = sum_a sum_k Phi_ia[k] fft(x_a)_k e^{2 pi i k t / N} / (N P_ii), kernel feeds 2*scale*fft(x)
    for (int i = 0; i < m; ++i)
        for (int a = 0; a < n; ++a) {
            const auto& ph = s.phi[(size_t)i * n + a];
            std::vector<cplx> pe(M + 1);
            for (int k = 0; k <= M; ++k) {
                const cplx u = ph[k], v = std::conj(ph[(N - k) % N]);
                const cplx roll = dpplan::unit_root(((long long)k * (long long)s.pretrigger) % N, N);
                pe[k] = 0.5 * (u + v) * roll / ((double)N * s.P[(size_t)i * m + i] * 2.0 * scale);
            }
            pe[0] = cplx(pe[0].real(), 0.0);
            pe[M] = cplx(pe[M].real(), 0.0);
            std::vector<cx<T>> g1;
            std::vector<cx<S>> gs1;
            dpplan2::pack_onesided<T, R1>(pe, g1, gs1);
            dt.g.insert(dt.g.end(), g1.begin(), g1.end());
            dt.g_self.insert(dt.g_self.end(), gs1.begin(), gs1.end());
        }
    // chi0 = sum_k X^H iS X df on the one-sided bins of X~ = 2*scale*fft(x)
    const double wsc = 1.0 / ((double)N * (double)N * df) / (4.0 * scale * scale);
    auto pack_real = [&](const std::vector<double>& w, std::vector<T>& out, std::vector<S>& self) {
        out.resize((size_t)NPH * 16 * NT);
        for (int p = 0; p < NPH; ++p)
            for (int e = 0; e < 16; ++e)
                for (int t = 0; t < NT; ++t) {
                    int b[2];
                    dpplan2::entry_bins<G>(p, t, e, b);
                    const double v[2] = {w[b[0]], w[b[1]]};
                    out[((size_t)p * 16 + e) * NT + t] = dpplan2::Pack<T>::r(v);
                }
        self.resize(17 * 2);
        for (int l = 0; l < 17; ++l) {
            int b[2];
            bool d[2];
            dpplan2::self_bins<G>(l, b, d);
            for (int j = 0; j < 2; ++j) self[l * 2 + j] = d[j] ? (S)0 : (S)w[b[j]];
        }
    };
    dt.wd.resize(n);
    dt.wd_self.resize(n);
    for (int a = 0; a < n; ++a) {
        const auto& is = s.isig[(size_t)a * n + a];
        std::vector<double> w(M + 1);
        for (int k = 0; k <= M; ++k) {
            double v = is[k].real();
            if (k != 0 && k != M) v += is[N - k].real();
            w[k] = v * wsc;
        }
        pack_real(w, dt.wd[a], dt.wd_self[a]);
    }
    for (int a = 0; a < n; ++a)
        for (int b = a + 1; b < n; ++b) {
            const auto& is = s.isig[(size_t)a * n + b];
            std::vector<cplx> w(M + 1);
            for (int k = 0; k <= M; ++k) {
                cplx v = is[k];
                if (k != 0 && k != M) v += std::conj(is[N - k]);
                w[k] = 2.0 * v * wsc;
            }
            std::vector<cx<T>> wo;
            std::vector<cx<S>> wos;
            dpplan2::pack_onesided<T, R1>(w, wo, wos);
            for (int l = 0; l < 17; ++l) {  // a self lane whose two bins coincide counts once
                int bins[2];
                bool d[2];
                dpplan2::self_bins<G>(l, bins, d);
                for (int j = 0; j < 2; ++j)
                    if (d[j]) wos[l * 2 + j] = cx<S>{(S)0, (S)0};
            }
            dt.wo.push_back(std::move(wo));
            dt.wo_self.push_back(std::move(wos));
        }
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            dt.cmat[i][j] = s.Pinv[(size_t)i * m + j] * s.P[(size_t)i * m + i] * s.P[(size_t)j * m + j];
            dt.amat[i][j] = s.Pinv[(size_t)i * m + j] * s.P[(size_t)j * m + j];
        }
    return dt;
}

}  // namespace dpnxm
