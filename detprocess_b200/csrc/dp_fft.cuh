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
// p[k] = w^k for k = 0..R-1 with a log-depth product tree (error ~ log2(R) ulp).
template <int R, class T> struct dp_powers {
    template <int K> static DP_HD void fill(cx<T> (&p)[R]) {
        if constexpr (K < R) {
            if constexpr (K % 2 == 0) {
                const cx<T> h = p[K / 2];
                p[K] = cx<T>{dp_fma(h.re, h.re, -(h.im * h.im)), (T)2 * h.re * h.im};
            } else {
                p[K] = cmul(p[K / 2], p[K - K / 2]);
            }
            fill<K + 1>(p);
        }
    }
    static DP_HD void run(cx<T> w, cx<T> (&p)[R]) {
        p[0] = cx<T>{(T)1, (T)0};
        if constexpr (R > 1) p[1] = w;
        fill<2>(p);
    }
};

// x[k] *= w^k (k = 1..R-1);  CONJ: x[k] *= conj(w)^k
template <int R, bool CONJ, class T> DP_HD void dp_twiddle(cx<T> (&x)[R], cx<T> w) {
    if constexpr (R > 1) {
        if (CONJ) w.im = -w.im;
        cx<T> p[R];
        dp_powers<R, T>::run(w, p);
#pragma unroll
        for (int k = 1; k < R; ++k) x[k] = cmul(x[k], p[k]);
    }
}
