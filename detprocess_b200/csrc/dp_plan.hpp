// Host-side (one-time) construction of the optimal-filter tables: template FFT,
// phi = conj(s)/J, norm, chi0 weights and their thread-order layouts for the device
// kernel.  Plain C++ (no CUDA) so the same code serves the C-ABI library and the
// host emulator tests.
//
// Mirrors what the reference does once per (nb_samples, nb_pretrigger, tag) key in
// ProcessingData.instantiate_OF_base (detprocess/process/processing_data.py:275-381):
// OFBase.set_csd / add_template / calc_phi.
#pragma once
#include <cmath>
#include <complex>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "dp_of_kernel.cuh"

namespace dpplan {

using cplx = std::complex<double>;

// in-place iterative radix-2 FFT, n power of two, forward sign -
inline void fft_pow2(std::vector<cplx>& a) {
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        // twiddles from a directly evaluated table (no recurrence drift)
        std::vector<cplx> w(len / 2);
        for (size_t k = 0; k < len / 2; ++k) {
            const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)len;
            w[k] = cplx((double)cosl(ang), (double)sinl(ang));
        }
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const cplx u = a[i + k], v = a[i + k + len / 2] * w[k];
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
}


// forward DFT of any length: radix-2 when n is a power of two, Bluestein's chirp-z through fft_pow2 otherwise
// (one-time setup of a template spectrum; trace lengths such as 25000 = 2^3 5^5)
inline void fft_any(std::vector<cplx>& a) {
    const size_t n = a.size();
    if ((n & (n - 1)) == 0) {
        fft_pow2(a);
        return;
    }
    size_t f = 1;
    while (f < 2 * n - 1) f <<= 1;
    std::vector<cplx> chirp(n);   // exp(-i pi j^2 / n), j^2 reduced mod 2n in integers
    for (size_t j = 0; j < n; ++j) {
        const unsigned long long q = ((unsigned long long)j * j) % (2ull * n);
        const long double ang = -3.14159265358979323846264338327950288L * (long double)q / (long double)n;
        chirp[j] = cplx((double)cosl(ang), (double)sinl(ang));
    }
    std::vector<cplx> x(f, cplx(0, 0)), y(f, cplx(0, 0));
    for (size_t j = 0; j < n; ++j) x[j] = a[j] * chirp[j];
    y[0] = std::conj(chirp[0]);
    for (size_t j = 1; j < n; ++j) y[j] = y[f - j] = std::conj(chirp[j]);
    fft_pow2(x);
    fft_pow2(y);
    for (size_t j = 0; j < f; ++j) x[j] = std::conj(x[j] * y[j]);   // inverse transform as conj(fft(conj(.))) / f
    fft_pow2(x);
    for (size_t k = 0; k < n; ++k) a[k] = std::conj(x[k]) / (double)f * chirp[k];
}

struct Template {
    std::vector<double> trace;  // [N]
    int pretrigger = 0;
    bool integralnorm = false;
    // derived (QETpy conventions, see oracle/of1x1.py)
    std::vector<cplx> s;    // fft(template)/N/df  (two-sided, N)
    std::vector<cplx> phi;  // conj(s)/J
    double norm = 0.0;
    double tsum = 0.0;
};

struct Fit {
    int templ, lo, hi, outside;
    double fcut = -1.0;  // lowchi2_fcutoff of this fit (Hz); < 0: the plan's default
};

struct Channel {
    std::vector<double> J;  // two-sided PSD with coupling applied (inf allowed), empty = unset
    std::vector<Template> templ;
    std::vector<Fit> fits;
};

inline bool is_pow2(long long n) { return n > 0 && (n & (n - 1)) == 0; }

// geometry selection: N = 2 * P * 512 * R1
struct Geometry {
    int N = 0, P = 1, R1 = 0, NT = 0, MS = 0;
};
inline Geometry pick_geometry(int N, bool f64) {
    Geometry g;
    g.N = N;
    if (!is_pow2(N) || N < 2048 || N > 131072) throw std::invalid_argument("nb_samples must be a power of two in [2048, 131072]");
    const int M = N / 2;
    const int max_ms = f64 ? 8192 : 16384;  // one CTA's shared memory: 136 KB
    g.P = 1;
    while (M / g.P > max_ms) g.P *= 2;
    if (g.P > 2) throw std::invalid_argument("nb_samples too large for this precision");
    g.MS = M / g.P;
    g.R1 = g.MS / 512;
    g.NT = 16 * g.R1;
    return g;
}

// spectrum bin held by thread t, table entry e (thread-order tables).
//   P = 1: e in [0,32): e < 16 -> A[e] (k = K12 + KQ e), else B[e-16] (k = K12' + KQ (e-16))
//   P = 2: e = 4 r + j: quad r of the pair (A[r], B[15-r]) with k' = K12 + KQ r,
//          j = 0: k', 1: M - k', 2: k' + M', 3: M' - k'
// Entries of thread 0 are placeholders (its butterflies are self-paired, see self_bins).
inline int k_of(const Geometry& g, int t, int e) {
    const int R1 = g.R1, KQ = 32 * R1, MS = g.MS, M = g.N / 2;
    const int k1 = t >> 4, k2 = t & 15;
    const int K12 = k1 + R1 * k2;
    const int K12p = (t == 0) ? KQ / 2 : KQ - K12;
    if (g.P == 1) {
        if (e < 16) return K12 + KQ * e;
        return K12p + KQ * (e - 16);
    }
    const int r = e >> 2, j = e & 3;
    const int kp = K12 + KQ * r;
    switch (j) {
        case 0: return kp;
        case 1: return (M - kp) % M;
        case 2: return kp + MS;
        default: return (MS - kp + M) % M;
    }
}

// bins of self-pair lane l (0..16), 2*P entries: P = 1: (k, M-k); P = 2: (k', M-k', k'+M', M'-k').
// The DC pair's mirror is the Nyquist bin M.  dup[j] marks bins already listed by the lane
// (self-mirrored quads) whose chi0 weight must be zero.
inline void self_bins(const Geometry& g, int l, int* bins, bool* dup) {
    const int KQ = 32 * g.R1, MS = g.MS, M = g.N / 2;
    int kp;
    if (l < 9)
        kp = KQ * l;
    else
        kp = KQ / 2 + KQ * (l - 9);
    if (g.P == 1) {
        bins[0] = kp;
        bins[1] = (kp == 0) ? M : M - kp;
    } else {
        bins[0] = kp;
        bins[1] = (kp == 0) ? M : M - kp;
        bins[2] = kp + MS;
        bins[3] = MS - kp;
    }
    for (int j = 0; j < 2 * g.P; ++j) {
        dup[j] = false;
        for (int i = 0; i < j; ++i)
            if (bins[i] == bins[j]) dup[j] = true;
    }
}

template <class T> struct DeviceTables {
    // flat host images; the C-ABI layer copies them to the device
    std::vector<cx<T>> tw1, tw2, twn, twp;
    struct Templ {
        std::vector<cx<T>> phi, phi_self, s_low;
        double norm, tsum;
        int pretrigger;
    };
    struct Chan {
        std::vector<T> wj, wj_self, wj_low;
        std::vector<Templ> templ;
    };
    std::vector<Chan> chans;
    int nlow = 0;
    double scale = 1.0;
};

inline void finalize_template(Template& tp, const std::vector<double>& J, double fs) {
    const int N = (int)tp.trace.size();
    const double df = fs / N;
    std::vector<cplx> a(N);
    for (int i = 0; i < N; ++i) a[i] = cplx(tp.trace[i], 0.0);
    fft_any(a);
    tp.s.resize(N);
    for (int i = 0; i < N; ++i) tp.s[i] = a[i] / (double)N / df;
    if (tp.integralnorm) {
        const cplx s0 = tp.s[0];
        for (auto& v : tp.s) v /= s0;
    }
    tp.phi.resize(N);
    cplx acc(0, 0);
    double ts = 0.0;
    const double val = 1.0 / (N * (1.0 / fs));  // numpy.fft.fftfreq spacing
    for (int i = 0; i < N; ++i) {
        tp.phi[i] = std::conj(tp.s[i]) / J[i];
        acc += tp.phi[i] * tp.s[i];
        const double f = (i < (N + 1) / 2 ? i : i - N) * val;
        ts += (2 * M_PI * f) * (2 * M_PI * f) * std::norm(tp.s[i]) / J[i];
    }
    tp.norm = acc.real() * df;
    tp.tsum = ts * df;
}

// number of non-negative-frequency bins with f <= fcut (numpy fftfreq arithmetic)
inline int count_low_bins(int N, double fs, double fcut) {
    const double val = 1.0 / (N * (1.0 / fs));
    int n = 0;
    for (int k = 0; k <= N / 2; ++k)
        if (std::fabs(k * val) <= fcut) ++n; else break;
    return n;
}


// exp(-2 pi i num/den), evaluated directly in long double
inline cplx unit_root(long long num, long long den) {
    const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)num / (long double)den;
    return cplx((double)cosl(ang), (double)sinl(ang));
}

// chi0 weights on 2*sc*fft(x), one-sided bins k = 0..M:  wJ[k] = (1/J[k] + 1/J[N-k]) / (N^2 df) / (4 sc^2)
inline std::vector<double> chi0_weights(const std::vector<double>& J, double fs, double scale) {
    const int N = (int)J.size(), M = N / 2;
    const double df = fs / N;
    std::vector<double> wJ(M + 1);
    for (int k = 0; k <= M; ++k) {
        double w = 1.0 / J[k];
        if (k != 0 && k != M) w += 1.0 / J[N - k];
        wJ[k] = w / ((double)N * (double)N * df) / (4.0 * scale * scale);
    }
    return wJ;
}

// hermitian-symmetrised filter on the one-sided bins, all scalings folded in:
//   amps_td[n] = sum_k phi_k fft(x)_k e^{+2 pi i k n/N} / (N norm),  kernel feeds 2*sc*fft(x);
// the roll by `pretrigger` (zero delay at index pretrigger) is a phase ramp
inline std::vector<cplx> filter_onesided(const Template& tp, double scale) {
    const int N = (int)tp.phi.size(), M = N / 2;
    std::vector<cplx> pe(M + 1);
    for (int k = 0; k <= M; ++k) {
        const cplx a = tp.phi[k], b = std::conj(tp.phi[(N - k) % N]);
        const cplx roll = unit_root(((long long)k * (long long)tp.pretrigger) % N, N);
        pe[k] = 0.5 * (a + b) * roll / ((double)N * tp.norm * 2.0 * scale);
    }
    pe[0] = cplx(pe[0].real(), 0.0);
    pe[M] = cplx(pe[M].real(), 0.0);
    return pe;
}

template <class T>
DeviceTables<T> build_tables(const Geometry& g, double fs, const std::vector<Channel>& chans, double fcut, double scale) {
    DeviceTables<T> dt;
    const int N = g.N, M = N / 2, NT = g.NT, MS = g.MS;
    const double df = fs / N;
    dt.scale = scale;
    dt.nlow = count_low_bins(N, fs, fcut);
    if (dt.nlow > 2 * NT || dt.nlow > DP_NLOW_MAX)
        throw std::invalid_argument("lowchi2_fcutoff too high for the fused kernel (needs <= " +
                                    std::to_string(std::min(2 * NT, DP_NLOW_MAX)) + " bins)");
    auto cxT = [](cplx z) { return cx<T>{(T)z.real(), (T)z.imag()}; };
    auto root = [](long long num, long long den) { return unit_root(num, den); };
    dt.tw1.resize(512);
    for (int m = 0; m < 512; ++m) dt.tw1[m] = cxT(root(m, MS));
    dt.tw2.resize(16);
    for (int m = 0; m < 16; ++m) dt.tw2[m] = cxT(root(m, 512));
    dt.twn.resize(NT);
    dt.twp.resize(NT);
    for (int t = 0; t < NT; ++t) {
        const int K12 = (t >> 4) + g.R1 * (t & 15);
        dt.twn[t] = cxT(root(K12, N));
        dt.twp[t] = cxT(root(K12, M));
    }
    const int NE = 32 * g.P;
    for (const auto& ch : chans) {
        typename DeviceTables<T>::Chan dc;
        if ((int)ch.J.size() != N) throw std::invalid_argument("psd not set for a channel");
        const std::vector<double> wJ = chi0_weights(ch.J, fs, scale);
        dc.wj.resize((size_t)NE * NT);
        for (int e = 0; e < NE; ++e)
            for (int t = 0; t < NT; ++t) dc.wj[(size_t)e * NT + t] = (T)wJ[k_of(g, t, e)];
        dc.wj_self.resize(17 * 2 * g.P);
        for (int l = 0; l < 17; ++l) {
            int bins[4];
            bool dup[4];
            self_bins(g, l, bins, dup);
            for (int j = 0; j < 2 * g.P; ++j) dc.wj_self[l * 2 * g.P + j] = dup[j] ? (T)0 : (T)wJ[bins[j]];
        }
        dc.wj_low.resize(dt.nlow);
        for (int k = 0; k < dt.nlow; ++k) dc.wj_low[k] = (T)wJ[k];
        for (const auto& tp : ch.templ) {
            typename DeviceTables<T>::Templ d;
            const std::vector<cplx> pe = filter_onesided(tp, scale);
            d.phi.resize((size_t)NE * NT);
            for (int e = 0; e < NE; ++e)
                for (int t = 0; t < NT; ++t) d.phi[(size_t)e * NT + t] = cxT(pe[k_of(g, t, e)]);
            d.phi_self.resize(17 * 2 * g.P);
            for (int l = 0; l < 17; ++l) {
                int bins[4];
                bool dup[4];
                self_bins(g, l, bins, dup);
                for (int j = 0; j < 2 * g.P; ++j) d.phi_self[l * 2 * g.P + j] = cxT(pe[bins[j]]);
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

}  // namespace dpplan
