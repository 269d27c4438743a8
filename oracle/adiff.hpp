// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Literal CPU restatement of Muscade's forward-mode automatic differentiation type
//     struct ∂ℝ{P,N,R} <: Real       (reference: src/Adiff.jl:13-16)
// and of the differentiation rules the toolbox elements use
//     @DiffRule2 + - * /             (src/Adiff.jl:204-229)
//     ^ with integer exponent        (src/Adiff.jl:230)
//     @DiffRule1 + - sqrt sin cos acos (src/Adiff.jl:233-266)
//     comparisons act on VALUE       (src/Adiff.jl:198-202)
// The order of floating point operations inside each rule follows the Julia expressions
// (left-to-right evaluation, same association), so that the oracle differs from the reference
// only through libm (sin/cos/acos/sqrt) — see DESIGN.md "oracle pinning".
//
// Two dual types are provided:
//   D<P,N,R>  : static number of partials N, nestable (R may itself be a dual)   — ∂ℝ{P,N,R}
//   DV        : precedence-1 dual over double with a *runtime* number of partials (≤ DV_MAX),
//               standing for ∂ℝ{1,Np,Float64} where Np depends on the solver's seeding
//               (12 SweepX :iter, 13 SweepX :step, 12(OX+1)+3(OU+1) DirectXUA).
#pragma once
#include <cmath>
#include <cstring>

namespace orc {

// ---------------------------------------------------------------- DV : ∂ℝ{1,Np,Float64}, Np runtime
// ORC_NP_FIXED (build flag, oracle/Makefile → liboracle_np12.so): the number of partials is a compile-time constant, as it is in the reference
// (∂ℝ{1,12,Float64} is an isbits struct of 13 doubles whose loops StaticArrays unrolls) — the build the CPU baseline is timed with.  The default build keeps
// it a run-time value so that one library serves every seeding (12, 13, 39 … partials); both give bit-identical numbers (tests/test_oracle_goldens.py).
#ifdef ORC_NP_FIXED
constexpr int DV_MAX = ORC_NP_FIXED;
struct DVctx { static constexpr int np = ORC_NP_FIXED; static bool set(int n) { return n <= ORC_NP_FIXED; } };     // fewer seeds: the spare partials stay zero
#else
constexpr int DV_MAX = 48;
struct DVctx { static inline thread_local int np = 0; static bool set(int n) { np = n; return n <= DV_MAX; } };
#endif
struct DV {
    double x;
    double dx[DV_MAX];
    DV() : x(0.) { for (int i = 0; i < DVctx::np; ++i) dx[i] = 0.; }
    DV(double v) : x(v) { for (int i = 0; i < DVctx::np; ++i) dx[i] = 0.; }   // convert(∂ℝ,b::ℝ) Adiff.jl:56
};

// ---------------------------------------------------------------- D<P,N,R> : ∂ℝ{P,N,R}
template <int P, int N, class R> struct D {
    R x;
    R dx[N];
    D() : x(R(0.)) { for (int i = 0; i < N; ++i) dx[i] = R(0.); }
    D(double v) : x(R(v)) { for (int i = 0; i < N; ++i) dx[i] = R(0.); }       // Adiff.jl:23,56
    D(const R& v, const R* d) : x(v) { for (int i = 0; i < N; ++i) dx[i] = d[i]; }
};

// number of partials of a dual (runtime for DV)
template <class T> struct npart;
template <> struct npart<DV> { static int n() { return DVctx::np; } };
template <int P, int N, class R> struct npart<D<P, N, R>> { static int n() { return N; } };

// VALUE: strip all partials (Adiff.jl:136-139)
inline double VALUE(double a) { return a; }
inline double VALUE(const DV& a) { return a.x; }
template <int P, int N, class R> inline double VALUE(const D<P, N, R>& a) { return VALUE(a.x); }

// ---------------------------------------------------------------- generic rule engine
// All rules are written once, for any dual type T with members x, dx[] and inner type inner<T>.
template <class T> struct inner;
template <> struct inner<DV> { using type = double; };
template <int P, int N, class R> struct inner<D<P, N, R>> { using type = R; };
template <class T> struct is_dual { static constexpr bool value = false; };
template <> struct is_dual<DV> { static constexpr bool value = true; };
template <int P, int N, class R> struct is_dual<D<P, N, R>> { static constexpr bool value = true; };

#define ORC_DUAL template <class T, class = typename std::enable_if<is_dual<T>::value>::type>

// integer power on plain doubles: Julia's x^2 == x*x, x^1 == x, x^0 == 1 (Base.literal_pow / pow_body)
inline double ipow(double a, int b) {
    if (b == 0) return 1.;
    if (b == 1) return a;
    if (b == 2) return a * a;
    if (b == 3) return a * a * a;
    return std::pow(a, b);
}

}  // namespace orc

#include <type_traits>

namespace orc {

// a + b    (Adiff.jl:223  a.dx+b.dx | a.dx | b.dx)
ORC_DUAL inline T operator+(const T& a, const T& b) {
    T r; r.x = a.x + b.x;
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = a.dx[i] + b.dx[i];
    return r;
}
ORC_DUAL inline T operator+(const T& a, double b) {
    T r; r.x = a.x + b;
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = a.dx[i];
    return r;
}
ORC_DUAL inline T operator+(double a, const T& b) {
    T r; r.x = a + b.x;
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = b.dx[i];
    return r;
}
// a - b    (Adiff.jl:224  a.dx-b.dx | a.dx | -b.dx)
ORC_DUAL inline T operator-(const T& a, const T& b) {
    T r; r.x = a.x - b.x;
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = a.dx[i] - b.dx[i];
    return r;
}
ORC_DUAL inline T operator-(const T& a, double b) {
    T r; r.x = a.x - b;
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = a.dx[i];
    return r;
}
ORC_DUAL inline T operator-(double a, const T& b) {
    T r; r.x = a - b.x;
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = -b.dx[i];
    return r;
}
// unary -  (Adiff.jl:237)
ORC_DUAL inline T operator-(const T& a) {
    T r; r.x = -a.x;
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = -a.dx[i];
    return r;
}
// a * b    (Adiff.jl:225  a.dx*b.x+a.x*b.dx | a.dx*b | a*b.dx)
ORC_DUAL inline T operator*(const T& a, const T& b) {
    T r; r.x = a.x * b.x;
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = a.dx[i] * b.x + a.x * b.dx[i];
    return r;
}
ORC_DUAL inline T operator*(const T& a, double b) {
    T r; r.x = a.x * b;
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = a.dx[i] * b;
    return r;
}
ORC_DUAL inline T operator*(double a, const T& b) {
    T r; r.x = a * b.x;
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = a * b.dx[i];
    return r;
}
// a ^ b::Int   (Adiff.jl:230   ∂ℝ(a.x^b , a.dx*b*a.x^(b-1)) ; b==0 → zero(a))
ORC_DUAL inline T ipow(const T& a, int b) {
    if (b == 0) return T(0.);
    T r; r.x = ipow(a.x, b);
    auto p = ipow(a.x, b - 1);
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = (a.dx[i] * double(b)) * p;
    return r;
}
// a / b    (Adiff.jl:226  a.dx/b.x-a.x/b.x^2*b.dx | a.dx/b | -a/b.x^2*b.dx)
ORC_DUAL inline T operator/(const T& a, const T& b) {
    T r; r.x = a.x / b.x;
    auto q = a.x / ipow(b.x, 2);
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = a.dx[i] / b.x - q * b.dx[i];
    return r;
}
ORC_DUAL inline T operator/(const T& a, double b) {
    T r; r.x = a.x / b;
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = a.dx[i] / b;
    return r;
}
ORC_DUAL inline T operator/(double a, const T& b) {
    T r; r.x = a / b.x;
    auto q = (-a) / ipow(b.x, 2);
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = q * b.dx[i];
    return r;
}

// ---------------------------------------------------------------- unary functions
inline double dsqrt(double a) { return std::sqrt(a); }
inline double dsin(double a) { return std::sin(a); }
inline double dcos(double a) { return std::cos(a); }
inline double dacos(double a) { return std::acos(a); }

// sqrt   (Adiff.jl:240  a.dx / 2. / sqrt(a.x))
ORC_DUAL inline T dsqrt(const T& a) {
    T r; r.x = dsqrt(a.x);
    auto s = dsqrt(a.x);
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = (a.dx[i] / 2.) / s;
    return r;
}
ORC_DUAL inline T dsin(const T& a);
// cos    (Adiff.jl:253  -sin(a.x) * a.dx)
ORC_DUAL inline T dcos(const T& a) {
    T r; r.x = dcos(a.x);
    auto s = -dsin(a.x);
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = s * a.dx[i];
    return r;
}
// sin    (Adiff.jl:252  cos(a.x) * a.dx)
template <class T, class> inline T dsin(const T& a) {
    T r; r.x = dsin(a.x);
    auto c = dcos(a.x);
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = c * a.dx[i];
    return r;
}
// acos   (Adiff.jl:266  -a.dx / sqrt(1. - a.x^2))
ORC_DUAL inline T dacos(const T& a) {
    T r; r.x = dacos(a.x);
    auto s = dsqrt(1. - ipow(a.x, 2));
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = (-a.dx[i]) / s;
    return r;
}

// hasnan (Adiff.jl:309-311)
inline bool hasnan(double a) { return std::isnan(a); }
ORC_DUAL inline bool hasnan(const T& a) {
    if (hasnan(a.x)) return true;
    for (int i = 0; i < npart<T>::n(); ++i) if (hasnan(a.dx[i])) return true;
    return false;
}

}  // namespace orc
