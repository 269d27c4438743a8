// Statically unrolled dual arithmetic kept in registers.
//
// Replaces the reference's generic ∂ℝ{P,N,R} (src/Adiff.jl:13-304) on the GPU path by two fixed-shape number types:
//   Dual<W>  : value + W directional partials (the lane's share of the solver's seed directions, SweepX.jl:46-63)
//   Jet<S>   : second-order Taylor number in time (q, q̇, q̈) over S, standing for the reference's
//              motion{P} packing ∂ℝ{P+1,1,∂ℝ{P,1,·}} (src/Taylor.jl:24-29) without its duplicated q̇ slot.
// Differentiation rules are the reference's (Adiff.jl:223-266) with FMA contraction allowed.
//
// The header compiles for the device (nvcc) and for the host (g++, used by tests/ to validate the kernel's
// formulation against the oracle without a GPU).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define MB_HD __host__ __device__ __forceinline__
#else
#define MB_HD inline
#endif
// MB_FN: larger building blocks (rotations, adjoints) — optionally kept out of line on the device to shrink the instruction footprint
#if defined(__CUDACC__) && defined(MB_NOINLINE)
#define MB_FN __host__ __device__ __noinline__
#else
#define MB_FN MB_HD
#endif

// MB_ALIGN(): block-wide rendezvous between the big phases of a kernel whose straight-line code is far larger than the
// instruction caches: it keeps the warps of a CTA on the same code region so that instruction lines are fetched once per CTA.
#if defined(__CUDA_ARCH__) && defined(MB_PHASE_SYNC)
#define MB_ALIGN() __syncthreads()
#else
#define MB_ALIGN() ((void)0)
#endif

namespace mb {

// ------------------------------------------------------------------------------------------------ double helpers
MB_HD double mb_sinpi(double x) {
#ifdef __CUDA_ARCH__
    return sinpi(x);
#else
    double n = nearbyint(x), r = x - n;
    double s = sin(3.141592653589793 * r);
    return (((long long)n) & 1) ? -s : s;
#endif
}
MB_HD void mb_sincos(double x, double* s, double* c) {
#ifdef __CUDA_ARCH__
    sincos(x, s, c);
#else
    *s = sin(x); *c = cos(x);
#endif
}

// sinc1 and its derivatives on Float64, with the reference's branch thresholds (toolbox/Rotations.jl:13-44)
// and Julia Base.sinc's small-argument polynomial (|x/π| < 1e-3).
template <int K> MB_HD double sinc1k(double x);
template <> MB_HD double sinc1k<0>(double x) {
    const double PI = 3.141592653589793;
    double y = x / PI;
    if (fabs(y) < 0.001) {
        double y2 = y * y;
        return fma(y2, fma(y2, (PI * PI) * (PI * PI) / 120, -(PI * PI) / 6), 1.0);
    }
    return mb_sinpi(y) / (PI * y);
}
template <> MB_HD double sinc1k<1>(double x) {
    if (fabs(x) > 1e-3) { double s, c; mb_sincos(x, &s, &c); return c / x - s / (x * x); }
    double x2 = x * x;
    return x * (-1. / 3 + x2 / 30);
}
template <> MB_HD double sinc1k<2>(double x) {
    if (fabs(x) > 1e-1) { double s, c; mb_sincos(x, &s, &c); return -s / x - 2 * c / (x * x) + 2 * s / (x * x * x); }
    double x2 = x * x;
    return -1. / 3 + x2 * (1. / 10 + x2 * (-1. / 168 + x2 * (1. / 6480)));
}
template <> MB_HD double sinc1k<3>(double x) {
    if (fabs(x) > 0.4) {
        double s, c; mb_sincos(x, &s, &c);
        double x2 = x * x;
        return -c / x + 3 * s / x2 + 6 * c / (x * x2) - 6 * s / (x2 * x2);
    }
    double x2 = x * x;
    return x * (1. / 5 + x2 * (-1. / 42 + x2 * (1. / 1080 + x2 * (-1. / 55440 + x2 * (1. / 4717440)))));
}
template <> MB_HD double sinc1k<4>(double x) {
    double x2 = x * x;
    return 1. / 5 + x2 * (-1. / 14 + x2 * (1. / 216 + x2 * (-1. / 7920 + x2 * (1. / 524160 + x2 * (-1. / 54432000 + x2 * (1. / 54432000 + x2 * (-1. / 8143027200. + x2 * (1. / 1656387532800.))))))));
}
template <> MB_HD double sinc1k<5>(double x) { return x * NAN; }
template <> MB_HD double sinc1k<6>(double x) { return x * NAN; }

MB_HD double value(double a) { return a; }
MB_HD double mb_sqrt(double a) { return sqrt(a); }
MB_HD double mb_acos(double a) { return acos(a); }
MB_HD double mb_rcp(double a) { return 1.0 / a; }
template <class T> struct Make { static MB_HD T c(double v); };
template <> struct Make<double> { static MB_HD double c(double v) { return v; } };

// ------------------------------------------------------------------------------------------------ Dual<W>
template <int W> struct Dual {
    double v;
    double d[W];
};
template <int W> struct Make<Dual<W>> {
    static MB_HD Dual<W> c(double v) { Dual<W> r; r.v = v;
#pragma unroll
        for (int i = 0; i < W; ++i) r.d[i] = 0.; return r; }
};
template <int W> MB_HD double value(const Dual<W>& a) { return a.v; }

#define MB_FORW _Pragma("unroll") for (int i = 0; i < W; ++i)

template <int W> MB_HD Dual<W> operator+(const Dual<W>& a, const Dual<W>& b) { Dual<W> r; r.v = a.v + b.v; MB_FORW r.d[i] = a.d[i] + b.d[i]; return r; }
template <int W> MB_HD Dual<W> operator+(const Dual<W>& a, double b) { Dual<W> r = a; r.v = a.v + b; return r; }
template <int W> MB_HD Dual<W> operator+(double a, const Dual<W>& b) { Dual<W> r = b; r.v = a + b.v; return r; }
template <int W> MB_HD Dual<W> operator-(const Dual<W>& a, const Dual<W>& b) { Dual<W> r; r.v = a.v - b.v; MB_FORW r.d[i] = a.d[i] - b.d[i]; return r; }
template <int W> MB_HD Dual<W> operator-(const Dual<W>& a, double b) { Dual<W> r = a; r.v = a.v - b; return r; }
template <int W> MB_HD Dual<W> operator-(double a, const Dual<W>& b) { Dual<W> r; r.v = a - b.v; MB_FORW r.d[i] = -b.d[i]; return r; }
template <int W> MB_HD Dual<W> operator-(const Dual<W>& a) { Dual<W> r; r.v = -a.v; MB_FORW r.d[i] = -a.d[i]; return r; }
template <int W> MB_HD Dual<W> operator*(const Dual<W>& a, const Dual<W>& b) {
    Dual<W> r; r.v = a.v * b.v; MB_FORW r.d[i] = fma(a.v, b.d[i], a.d[i] * b.v); return r;
}
template <int W> MB_HD Dual<W> operator*(const Dual<W>& a, double b) { Dual<W> r; r.v = a.v * b; MB_FORW r.d[i] = a.d[i] * b; return r; }
template <int W> MB_HD Dual<W> operator*(double a, const Dual<W>& b) { Dual<W> r; r.v = a * b.v; MB_FORW r.d[i] = a * b.d[i]; return r; }
template <int W> MB_HD Dual<W> mb_rcp(const Dual<W>& a) { Dual<W> r; r.v = 1.0 / a.v; double m = -r.v * r.v; MB_FORW r.d[i] = m * a.d[i]; return r; }
template <int W> MB_HD Dual<W> operator/(const Dual<W>& a, const Dual<W>& b) {
    Dual<W> r; double inv = 1.0 / b.v; r.v = a.v * inv; MB_FORW r.d[i] = (a.d[i] - r.v * b.d[i]) * inv; return r;
}
template <int W> MB_HD Dual<W> operator/(const Dual<W>& a, double b) { return a * (1.0 / b); }
template <int W> MB_HD Dual<W> operator/(double a, const Dual<W>& b) { Dual<W> r; double inv = 1.0 / b.v; r.v = a * inv; double m = -r.v * inv; MB_FORW r.d[i] = m * b.d[i]; return r; }
template <int W> MB_HD Dual<W> mb_sqrt(const Dual<W>& a) { Dual<W> r; r.v = sqrt(a.v); double m = 0.5 / r.v; MB_FORW r.d[i] = m * a.d[i]; return r; }
template <int W> MB_HD Dual<W> mb_acos(const Dual<W>& a) { Dual<W> r; r.v = acos(a.v); double m = -1.0 / sqrt(1.0 - a.v * a.v); MB_FORW r.d[i] = m * a.d[i]; return r; }
// chained rule sinc1⁽ᵏ⁾(a) = (sinc1⁽ᵏ⁾(a.v), sinc1⁽ᵏ⁺¹⁾(a.v)·a.d)   (Rotations.jl:47-51)
template <int K, int W> MB_HD Dual<W> sinc1k(const Dual<W>& a) { Dual<W> r; r.v = sinc1k<K>(a.v); double m = sinc1k<K + 1>(a.v); MB_FORW r.d[i] = m * a.d[i]; return r; }

// ------------------------------------------------------------------------------------------------ Jet<S>  (q, q̇, q̈)
// Jet<S0,S1,S2>: the three Taylor coefficients may have different number types.  Jet<S> = Jet<S,S,S> is the general case; Jet<V,S,S> with V a plain value serves lanes whose
// seeds sit on the VELOCITIES (DirectXUA's ∂/∂X′ lanes): every order-0 coefficient is then free of partials by construction and no arithmetic is emitted for them.
template <class S0, class S1 = S0, class S2 = S1> struct Jet {
    S0 c0; S1 c1; S2 c2;
};
template <class S0, class S1, class S2> struct Make<Jet<S0, S1, S2>> {
    static MB_HD Jet<S0, S1, S2> c(double v) { Jet<S0, S1, S2> r; r.c0 = Make<S0>::c(v); r.c1 = Make<S1>::c(0.); r.c2 = Make<S2>::c(0.); return r; }
};
#define MB_J3(X) class X##0, class X##1, class X##2
#define MB_JT(X) Jet<X##0, X##1, X##2>
template <MB_J3(S)> MB_HD double value(const MB_JT(S)& a) { return value(a.c0); }
template <MB_J3(A), MB_J3(B)> MB_HD auto operator+(const MB_JT(A)& a, const MB_JT(B)& b) -> Jet<decltype(a.c0 + b.c0), decltype(a.c1 + b.c1), decltype(a.c2 + b.c2)> {
    Jet<decltype(a.c0 + b.c0), decltype(a.c1 + b.c1), decltype(a.c2 + b.c2)> r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; r.c2 = a.c2 + b.c2; return r;
}
template <MB_J3(S)> MB_HD MB_JT(S) operator+(const MB_JT(S)& a, double b) { MB_JT(S) r = a; r.c0 = a.c0 + b; return r; }
template <MB_J3(S)> MB_HD MB_JT(S) operator+(double a, const MB_JT(S)& b) { MB_JT(S) r = b; r.c0 = a + b.c0; return r; }
template <MB_J3(A), MB_J3(B)> MB_HD auto operator-(const MB_JT(A)& a, const MB_JT(B)& b) -> Jet<decltype(a.c0 - b.c0), decltype(a.c1 - b.c1), decltype(a.c2 - b.c2)> {
    Jet<decltype(a.c0 - b.c0), decltype(a.c1 - b.c1), decltype(a.c2 - b.c2)> r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; r.c2 = a.c2 - b.c2; return r;
}
template <MB_J3(S)> MB_HD MB_JT(S) operator-(const MB_JT(S)& a, double b) { MB_JT(S) r = a; r.c0 = a.c0 - b; return r; }
template <MB_J3(S)> MB_HD MB_JT(S) operator-(double a, const MB_JT(S)& b) { MB_JT(S) r; r.c0 = a - b.c0; r.c1 = -b.c1; r.c2 = -b.c2; return r; }
template <MB_J3(S)> MB_HD MB_JT(S) operator-(const MB_JT(S)& a) { MB_JT(S) r; r.c0 = -a.c0; r.c1 = -a.c1; r.c2 = -a.c2; return r; }
template <MB_J3(A), MB_J3(B)> MB_HD auto operator*(const MB_JT(A)& a, const MB_JT(B)& b)
    -> Jet<decltype(a.c0 * b.c0), decltype(a.c0 * b.c1 + a.c1 * b.c0), decltype((a.c0 * b.c2 + a.c2 * b.c0) + 2.0 * (a.c1 * b.c1))> {
    Jet<decltype(a.c0 * b.c0), decltype(a.c0 * b.c1 + a.c1 * b.c0), decltype((a.c0 * b.c2 + a.c2 * b.c0) + 2.0 * (a.c1 * b.c1))> r;
    r.c0 = a.c0 * b.c0;
    r.c1 = a.c0 * b.c1 + a.c1 * b.c0;
    r.c2 = (a.c0 * b.c2 + a.c2 * b.c0) + 2.0 * (a.c1 * b.c1);
    return r;
}
template <MB_J3(S)> MB_HD MB_JT(S) operator*(const MB_JT(S)& a, double b) { MB_JT(S) r; r.c0 = a.c0 * b; r.c1 = a.c1 * b; r.c2 = a.c2 * b; return r; }
template <MB_J3(S)> MB_HD MB_JT(S) operator*(double a, const MB_JT(S)& b) { MB_JT(S) r; r.c0 = a * b.c0; r.c1 = a * b.c1; r.c2 = a * b.c2; return r; }
// compose a univariate function with derivatives f, f′, f″ (numbers of the type of x.c0) taken at x.c0
template <MB_J3(S)> MB_HD MB_JT(S) jet_compose(const MB_JT(S)& x, const S0& f0, const S0& f1, const S0& f2) {
    MB_JT(S) r; r.c0 = f0; r.c1 = f1 * x.c1; r.c2 = f2 * (x.c1 * x.c1) + f1 * x.c2; return r;
}
template <MB_J3(S)> MB_HD MB_JT(S) mb_rcp(const MB_JT(S)& a) { S0 f0 = mb_rcp(a.c0); S0 f1 = -(f0 * f0); S0 f2 = -2.0 * (f1 * f0); return jet_compose(a, f0, f1, f2); }
template <MB_J3(A), MB_J3(B)> MB_HD auto operator/(const MB_JT(A)& a, const MB_JT(B)& b) -> decltype(a * mb_rcp(b)) { return a * mb_rcp(b); }
template <MB_J3(S)> MB_HD MB_JT(S) operator/(const MB_JT(S)& a, double b) { return a * (1.0 / b); }
template <MB_J3(S)> MB_HD MB_JT(S) operator/(double a, const MB_JT(S)& b) { return a * mb_rcp(b); }
template <MB_J3(S)> MB_HD MB_JT(S) mb_sqrt(const MB_JT(S)& a) { S0 f0 = mb_sqrt(a.c0); S0 f1 = 0.5 / f0; S0 f2 = -0.5 * (f1 / a.c0); return jet_compose(a, f0, f1, f2); }
template <MB_J3(S)> MB_HD MB_JT(S) mb_acos(const MB_JT(S)& a) {
    S0 f0 = mb_acos(a.c0);
    S0 w = 1.0 / (1.0 - a.c0 * a.c0);         // 1/(1-x²)
    S0 f1 = -mb_sqrt(w);                      // -1/√(1-x²)
    S0 f2 = f1 * a.c0 * w;                    // -x/(1-x²)^{3/2}
    return jet_compose(a, f0, f1, f2);
}
template <int K, MB_J3(S)> MB_HD MB_JT(S) sinc1k(const MB_JT(S)& a) { return jet_compose(a, sinc1k<K>(a.c0), sinc1k<K + 1>(a.c0), sinc1k<K + 2>(a.c0)); }

template <class T> MB_HD T sqr(const T& a) { return a * a; }

// ------------------------------------------------------------------------------------------------ derivative packs
// sinc1 and its first three derivatives at a plain argument, with ONE sincos shared by all orders; each order keeps the
// reference's own branch (closed form vs series, toolbox/Rotations.jl:13-40).  S[k] = sinc1⁽ᵏ⁾(x), k = 0..3.
MB_HD void sinc_pack_sc(double x, double s, double c, double* S) {
    const double PI = 3.141592653589793;
    const double x2 = x * x, ix = 1.0 / x, ix2 = ix * ix;
    const double y = x / PI;
    if (fabs(y) < 0.001) { const double y2 = y * y; S[0] = fma(y2, fma(y2, (PI * PI) * (PI * PI) / 120, -(PI * PI) / 6), 1.0); }
    else S[0] = s * ix;
    S[1] = (fabs(x) > 1e-3) ? (c * ix - s * ix2) : x * (-1. / 3 + x2 / 30);
    S[2] = (fabs(x) > 1e-1) ? (-s * ix - 2 * c * ix2 + 2 * s * (ix2 * ix))
                            : (-1. / 3 + x2 * (1. / 10 + x2 * (-1. / 168 + x2 * (1. / 6480))));
    S[3] = (fabs(x) > 0.4) ? (-c * ix + 3 * s * ix2 + 6 * c * (ix2 * ix) - 6 * s * (ix2 * ix2))
                           : x * (1. / 5 + x2 * (-1. / 42 + x2 * (1. / 1080 + x2 * (-1. / 55440 + x2 * (1. / 4717440)))));
}
MB_HD void sinc_pack(double x, double* S) {
    double s = 0., c = 1.;
    if (fabs(x) > 1e-3) mb_sincos(x, &s, &c);       // below 1e-3 every order is on its series branch
    if (x == 0.) { S[0] = 1.; S[1] = 0.; S[2] = -1. / 3; S[3] = 0.; return; }
    sinc_pack_sc(x, s, c, S);
}
// scac = sinc1∘acos and its first three derivatives (toolbox/Rotations.jl:61-68) — what the reference obtains by chasing
// @DiffRule1(acos), @DiffRule1(sinc1…) through nested duals. sin(acos x) = √(1−x²), cos(acos x) = x: no trigonometric call.
MB_HD void scac_pack(double x, double* C) {
    const double dx = x - 1.0;
    if (fabs(dx) > 1e-3) {
        const double w = 1.0 - x * x, sq = sqrt(w), q = 1.0 / sq, q2 = q * q, q3 = q2 * q;
        const double phi = acos(x);
        double G[4];
        sinc_pack_sc(phi, sq, x, G);
        const double p1 = -q, p2 = -x * q3, p3 = -q3 - 3 * (x * x) * (q3 * q2);
        C[0] = G[0];
        C[1] = G[1] * p1;
        C[2] = G[2] * (p1 * p1) + G[1] * p2;
        C[3] = G[3] * (p1 * p1 * p1) + 3 * G[2] * (p1 * p2) + G[1] * p3;
    } else {
        const double c1 = 1. / 3, c2 = -2. / 90, c3 = 0.0052911879917544626, c4 = -0.0016229317117234072, c5 = 0.0005625;
        C[0] = 1.0 + dx * (c1 + dx * (c2 + dx * (c3 + dx * (c4 + dx * c5))));
        C[1] = c1 + dx * (2 * c2 + dx * (3 * c3 + dx * (4 * c4 + dx * (5 * c5))));
        C[2] = 2 * c2 + dx * (6 * c3 + dx * (12 * c4 + dx * (20 * c5)));
        C[3] = 6 * c3 + dx * (24 * c4 + dx * (60 * c5));
    }
}
// f(x) for a number x of any of our types, given F[k] = f⁽ᵏ⁾(value(x)) for k up to the nesting depth of x
MB_HD double apply_fn(double, const double* F) { return F[0]; }
template <int W> MB_HD Dual<W> apply_fn(const Dual<W>& x, const double* F) { Dual<W> r; r.v = F[0]; MB_FORW r.d[i] = F[1] * x.d[i]; return r; }
template <MB_J3(S)> MB_HD MB_JT(S) apply_fn(const MB_JT(S)& x, const double* F) { return jet_compose(x, apply_fn(x.c0, F), apply_fn(x.c0, F + 1), apply_fn(x.c0, F + 2)); }

// widen<To>(x): embed a number into a type with more partial slots (missing slots are zero)
template <class To, class From> struct Widen;
template <class T> struct Widen<T, T> { static MB_HD T w(const T& x) { return x; } };
template <class To> struct Widen<To, double> { static MB_HD To w(double x) { return Make<To>::c(x); } };
template <> struct Widen<double, double> { static MB_HD double w(double x) { return x; } };
template <class To, class From> MB_HD To widen(const From& x) { return Widen<To, From>::w(x); }
template <MB_J3(A), MB_J3(B)> struct Widen<MB_JT(A), MB_JT(B)> {
    static MB_HD MB_JT(A) w(const MB_JT(B)& x) { MB_JT(A) r; r.c0 = widen<A0>(x.c0); r.c1 = widen<A1>(x.c1); r.c2 = widen<A2>(x.c2); return r; }
};
template <MB_J3(S)> struct Widen<MB_JT(S), MB_JT(S)> { static MB_HD MB_JT(S) w(const MB_JT(S)& x) { return x; } };

// a^2 exactly as the reference evaluates it (src/Adiff.jl:230):  ∂ℝ(a.x^b, a.dx*b*a.x^(b-1))  with  b==0 ? zero(a).
// On numbers nested three deep (time-packed by motion{P}, Taylor.jl:24-29, over the solver's ∂ℝ{1,Np}) the inner a.x^1 hits
// "x^0 → zero" one level down and loses its partial, so the second time derivative comes out as 2·a·ä instead of
// 2·a·ä + 2·ȧ².  Only toolbox/Rotations.jl:134 (sinc1(θ/2)^2) feeds such a square into results that are used.
// Reproduced here because parity with the reference is the contract; plain numbers and Dual<W> square normally.
MB_HD double sqr_ref(double a) { return a * a; }
template <int W> MB_HD Dual<W> sqr_ref(const Dual<W>& a) { return a * a; }
template <MB_J3(S)> MB_HD MB_JT(S) sqr_ref(const MB_JT(S)& a) {
    MB_JT(S) r; r.c0 = a.c0 * a.c0; r.c1 = 2.0 * (a.c0 * a.c1); r.c2 = 2.0 * (a.c0 * a.c2); return r;
}

}  // namespace mb
