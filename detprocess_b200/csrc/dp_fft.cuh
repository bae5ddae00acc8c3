// In-register complex arithmetic and small DFTs (R = 2..32) used by the fused
// optimal-filter kernel.  Everything here is fully unrolled at compile time so the
// arrays live in registers; no shared memory, no shuffles.
#pragma once
#include "dp_platform.cuh"

// ---------------------------------------------------------------- complex helpers
template <class T> struct alignas(2 * sizeof(T)) cx {
    T re, im;
};
template <class T> DP_HD cx<T> cmk(T a, T b) { return cx<T>{a, b}; }
template <class T> DP_HD cx<T> cadd(cx<T> a, cx<T> b) { return cx<T>{a.re + b.re, a.im + b.im}; }
template <class T> DP_HD cx<T> csub(cx<T> a, cx<T> b) { return cx<T>{a.re - b.re, a.im - b.im}; }
template <class T> DP_HD cx<T> cconj(cx<T> a) { return cx<T>{a.re, -a.im}; }
// a*b : 2 mul + 2 fma
template <class T> DP_HD cx<T> cmul(cx<T> a, cx<T> b) {
    return cx<T>{dp_fma(a.re, b.re, -(a.im * b.im)), dp_fma(a.re, b.im, a.im * b.re)};
}
// a*conj(b)
template <class T> DP_HD cx<T> cmulc(cx<T> a, cx<T> b) {
    return cx<T>{dp_fma(a.re, b.re, a.im * b.im), dp_fma(a.im, b.re, -(a.re * b.im))};
}
template <class T> DP_HD cx<T> cmuli(cx<T> a) { return cx<T>{-a.im, a.re}; }    //  i*a
template <class T> DP_HD cx<T> cmulni(cx<T> a) { return cx<T>{a.im, -a.re}; }   // -i*a
template <class T> DP_HD T cnorm2(cx<T> a) { return dp_fma(a.re, a.re, a.im * a.im); }

// read-only-path loads
template <class T> DP_DEV cx<T> dp_ldg(const cx<T>* p) {
#ifdef DP_HOST_EMU
    return *p;
#else
    const typename dp_vec2<T>::type v = __ldg(reinterpret_cast<const typename dp_vec2<T>::type*>(p));
    return cx<T>{v.x, v.y};
#endif
}
DP_DEV float dp_ldg(const float* p) { return __ldg(p); }
DP_DEV double dp_ldg(const double* p) { return __ldg(p); }

// --------------------------------------------------------- compile-time twiddles
// cos(2*pi*j/64), j = 0..16
DP_HD constexpr double dp_cos64_q(int j) {
    constexpr double t[17] = {1.0,
                              0.995184726672196886245,
                              0.980785280403230449126,
                              0.956940335732208864936,
                              0.923879532511286756128,
                              0.881921264348355029713,
                              0.831469612302545237079,
                              0.773010453362736960811,
                              0.707106781186547524401,
                              0.634393284163645498215,
                              0.555570233019602224743,
                              0.471396736825997648556,
                              0.382683432365089771728,
                              0.290284677254462367636,
                              0.195090322016128267848,
                              0.0980171403295606019942,
                              0.0};
    return t[j];
}
// cos / sin of 2*pi*j/64 for any integer j (symmetry folding)
DP_HD constexpr double dp_cos64(int j) {
    j = ((j % 64) + 64) % 64;
    if (j > 32) j = 64 - j;                 // cos even about pi
    return (j <= 16) ? dp_cos64_q(j) : -dp_cos64_q(32 - j);
}
DP_HD constexpr double dp_sin64(int j) { return dp_cos64(j - 16); }

// exp(sign * 2*pi*i * j/64) as compile-time constants of type T
template <class T, int J64, int SIGN> DP_HD cx<T> dp_w64() {
    constexpr double c = dp_cos64(J64);
    constexpr double s = dp_sin64(J64);
    return cx<T>{(T)c, (T)(SIGN * s)};
}

// exp(-2*pi*i*j/64) for a loop index that the compiler unrolls (j known at compile time
// after unrolling; falls back to a 64-entry select chain otherwise)
template <class T> DP_HD cx<T> dp_w64_rt(int j) {
    return cx<T>{(T)dp_cos64(j), (T)(-dp_sin64(j))};
}

// ----------------------------------------------------- radix-2 DIT, natural order
// out[k] = sum_n in[n] * exp(SIGN*2*pi*i*n*k/R); in and out natural order.
// Butterfly (e + w*o, e - w*o) in the 6-FMA form: x = e + w*o; x' = 2e - x.
template <int R, int SIGN, class T> struct dp_dft {
    static DP_HD void run(cx<T> (&x)[R]) {
        cx<T> e[R / 2], o[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; ++i) {
            e[i] = x[2 * i];
            o[i] = x[2 * i + 1];
        }
        dp_dft<R / 2, SIGN, T>::run(e);
        dp_dft<R / 2, SIGN, T>::run(o);
        bfly<0>(x, e, o);
    }
    template <int K> static DP_HD void bfly(cx<T> (&x)[R], const cx<T> (&e)[R / 2], const cx<T> (&o)[R / 2]) {
        if constexpr (K < R / 2) {
            if constexpr (K == 0) {
                x[K] = cadd(e[K], o[K]);
                x[K + R / 2] = csub(e[K], o[K]);
            } else if constexpr (4 * K == R) {
                // w = exp(SIGN*i*pi/2) = SIGN*i
                cx<T> t = (SIGN > 0) ? cmuli(o[K]) : cmulni(o[K]);
                x[K] = cadd(e[K], t);
                x[K + R / 2] = csub(e[K], t);
            } else {
                const cx<T> w = dp_w64<T, K*(64 / R), SIGN>();
                cx<T> a;
                a.re = dp_fma(w.re, o[K].re, dp_fma(-w.im, o[K].im, e[K].re));
                a.im = dp_fma(w.re, o[K].im, dp_fma(w.im, o[K].re, e[K].im));
                x[K] = a;
                x[K + R / 2] = cx<T>{dp_fma((T)2, e[K].re, -a.re), dp_fma((T)2, e[K].im, -a.im)};
            }
            bfly<K + 1>(x, e, o);
        }
    }
};
template <int SIGN, class T> struct dp_dft<1, SIGN, T> {
    static DP_HD void run(cx<T> (&)[1]) {}
};
template <int SIGN, class T> struct dp_dft<2, SIGN, T> {
    static DP_HD void run(cx<T> (&x)[2]) {
        cx<T> a = x[0], b = x[1];
        x[0] = cadd(a, b);
        x[1] = csub(a, b);
    }
};

// ------------------------------------------------------------- twiddle powers
// x[k] *= w^k, k = 1..R-1 (CONJ: conj(w)^k).  Powers are built level by level,
// p[k] = p[k/2] * p[k - k/2] (log-depth => ~log2(R) ulp), and applied as soon as they
// exist; only the previous level (<= R/4 + 1 values) stays live, the last level is
// never stored -- this keeps the radix-32 passes inside the register budget.
template <int R, int LO, class T> struct dp_tw_level {
    // prev holds p[LO/2 .. LO] (LO/2 + 1 values); this level produces p[LO .. 2*LO-1]
    static DP_HD void run(cx<T> (&x)[R], const cx<T> (&prev)[LO / 2 + 1]) {
        if constexpr (2 * LO >= R) {
            // last level: compute, apply, forget
#pragma unroll
            for (int k = LO; k < R; ++k) {
                const int h = k / 2, g = k - h;
                const cx<T> a = prev[h - LO / 2], b = prev[g - LO / 2];
                const cx<T> p = (h == g) ? cx<T>{dp_fma(a.re, a.re, -(a.im * a.im)), (T)2 * a.re * a.im} : cmul(a, b);
                x[k] = cmul(x[k], p);
            }
        } else {
            cx<T> cur[LO + 1];  // p[LO .. 2*LO]
#pragma unroll
            for (int k = LO; k <= 2 * LO; ++k) {
                const int h = k / 2, g = k - h;
                const cx<T> a = prev[h - LO / 2], b = prev[g - LO / 2];
                cur[k - LO] = (h == g) ? cx<T>{dp_fma(a.re, a.re, -(a.im * a.im)), (T)2 * a.re * a.im} : cmul(a, b);
            }
#pragma unroll
            for (int k = LO; k < 2 * LO; ++k) x[k] = cmul(x[k], cur[k - LO]);
            dp_tw_level<R, 2 * LO, T>::run(x, cur);
        }
    }
};

template <int R, bool CONJ, class T> DP_HD void dp_twiddle(cx<T> (&x)[R], cx<T> w) {
    if constexpr (R > 1) {
        if (CONJ) w.im = -w.im;
        x[1] = cmul(x[1], w);
        if constexpr (R > 2) {
            // level LO = 2 needs prev = p[1..2]
            cx<T> prev[2];
            prev[0] = w;
            prev[1] = cx<T>{dp_fma(w.re, w.re, -(w.im * w.im)), (T)2 * w.re * w.im};
            dp_tw_level<R, 2, T>::run(x, prev);
        }
    }
}
