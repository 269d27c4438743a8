// ORACLE — TEST INFRASTRUCTURE ONLY (see adiff.hpp).
//
// Literal restatement of toolbox/Rotations.jl:
//   sinc1, sinc1′, sinc1″, sinc1‴, sinc1⁗ and their chained @DiffRule1   (Rotations.jl:13-51)
//   scac                                                                  (Rotations.jl:61-68)
//   spin, spin⁻¹, trace, Rodrigues⁻¹, norm3, spin², Rodrigues             (Rotations.jl:81-135)
//   intrinsicrotationrates                                                (Rotations.jl:177-182)
// and of the StaticArrays products they rely on (src/Dots.jl:73-85 → SMatrix*SMatrix, left-to-right sums).
//
// Julia Base.sinc (un-vendored dependency, Base/special/trig.jl, Julia 1.11): for Float64
//   |x| < 0.001 ? evalpoly(x^2,(1,-π²/6,π⁴/120)) : sinpi(x)/(π*x)
#pragma once
#include "adiff.hpp"

namespace orc {

// ---------------------------------------------------------------- small static arrays (column-major like SMatrix)
template <class T> struct V3 { T v[3]; T& operator[](int i) { return v[i]; } const T& operator[](int i) const { return v[i]; } };
template <class T> struct M33 {
    T m[9];                                                   // column-major: m[i+3j] = M[i,j]
    T& operator()(int i, int j) { return m[i + 3 * j]; }
    const T& operator()(int i, int j) const { return m[i + 3 * j]; }
};
// SMatrix{3,3}*SMatrix{3,3}: c_ij = (a_i1*b_1j + a_i2*b_2j) + a_i3*b_3j
template <class A, class B> inline auto matmul(const M33<A>& a, const M33<B>& b) -> M33<decltype(a.m[0] * b.m[0])> {
    M33<decltype(a.m[0] * b.m[0])> c;
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) c(i, j) = (a(i, 0) * b(0, j) + a(i, 1) * b(1, j)) + a(i, 2) * b(2, j);
    return c;
}
template <class T> inline M33<T> transpose(const M33<T>& a) {
    M33<T> c;
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) c(i, j) = a(j, i);
    return c;
}
// SMatrix{3,3}*SVector{3}
template <class A, class B> inline auto matvec(const M33<A>& a, const V3<B>& b) -> V3<decltype(a.m[0] * b.v[0])> {
    V3<decltype(a.m[0] * b.v[0])> c;
    for (int i = 0; i < 3; ++i) c[i] = (a(i, 0) * b[0] + a(i, 1) * b[1]) + a(i, 2) * b[2];
    return c;
}
// SVector{3} ∘₁ SMatrix{3,3}  (row vector times matrix, Dots.jl:83)
template <class A, class B> inline auto vecmat(const V3<A>& a, const M33<B>& b) -> V3<decltype(a.v[0] * b.m[0])> {
    V3<decltype(a.v[0] * b.m[0])> c;
    for (int j = 0; j < 3; ++j) c[j] = (a[0] * b(0, j) + a[1] * b(1, j)) + a[2] * b(2, j);
    return c;
}
template <class T> inline V3<T> operator+(const V3<T>& a, const V3<T>& b) { V3<T> c; for (int i = 0; i < 3; ++i) c[i] = a[i] + b[i]; return c; }
template <class T> inline V3<T> operator-(const V3<T>& a, const V3<T>& b) { V3<T> c; for (int i = 0; i < 3; ++i) c[i] = a[i] - b[i]; return c; }
template <class T> inline M33<T> operator+(const M33<T>& a, const M33<T>& b) { M33<T> c; for (int i = 0; i < 9; ++i) c.m[i] = a.m[i] + b.m[i]; return c; }

// ---------------------------------------------------------------- sinc1 family on Float64
inline double sinpi_(double x) {   // Julia sinpi: exact zeros at integers; argument reduction is exact
    double n = std::nearbyint(x);
    double r = x - n;
    double s = std::sin(M_PI * r);
    long long k = (long long)n;
    return (k & 1) ? -s : s;
}
inline double jl_sinc(double x) {  // Base.sinc (Julia ≥1.6)
    if (std::fabs(x) < 0.001) {
        const double a1 = -(M_PI * M_PI) / 6, a2 = (M_PI * M_PI) * (M_PI * M_PI) / 120;  // T(pi)^4 == (π²)² by power_by_squaring
        double x2 = x * x;
        return std::fma(x2, std::fma(x2, a2, a1), 1.0);
    }
    if (std::isinf(x)) return 0.;
    return sinpi_(x) / (M_PI * x);
}
template <int K> inline double sinc1k(double x);
template <> inline double sinc1k<0>(double x) { return jl_sinc(x / M_PI); }                      // Rotations.jl:13
template <> inline double sinc1k<1>(double x) {                                                  // Rotations.jl:14-22
    if (std::fabs(x) > 1e-3) { double s = std::sin(x), c = std::cos(x); return c / x - s / (x * x); }
    double x2 = x * x; return x * (-1. / 3 + x2 / 30);
}
template <> inline double sinc1k<2>(double x) {                                                  // Rotations.jl:23-31
    if (std::fabs(x) > 1e-1) { double s = std::sin(x), c = std::cos(x); return -s / x - 2 * c / (x * x) + 2 * s / (x * x * x); }
    double x2 = x * x; return -1. / 3 + x2 * (1. / 10 + x2 * (-1. / 168 + x2 * (1. / 6480)));
}
template <> inline double sinc1k<3>(double x) {                                                  // Rotations.jl:32-40
    if (std::fabs(x) > 0.4) {
        double s = std::sin(x), c = std::cos(x);
        double x2 = x * x;
        return -c / x + 3 * s / x2 + 6 * c / (x * x * x) - 6 * s / (x2 * x2);
    }
    double x2 = x * x;
    return x * (1. / 5 + x2 * (-1. / 42 + x2 * (1. / 1080 + x2 * (-1. / 55440 + x2 * (1. / 4717440)))));
}
template <> inline double sinc1k<4>(double x) {                                                  // Rotations.jl:41-44
    double x2 = x * x;
    return 1. / 5 + x2 * (-1. / 14 + x2 * (1. / 216 + x2 * (-1. / 7920 + x2 * (1. / 524160 + x2 * (-1. / 54432000 + x2 * (1. / 54432000 + x2 * (-1. / 8143027200. + x2 * (1. / 1656387532800.))))))));
}
template <> inline double sinc1k<5>(double x) { return x * NAN; }                                // Rotations.jl:45
template <> inline double sinc1k<6>(double x) { return x * NAN; }
template <> inline double sinc1k<7>(double x) { return x * NAN; }
// chained @DiffRule1 (Rotations.jl:47-51):  sinc1⁽ᵏ⁾(a) = ∂ℝ(sinc1⁽ᵏ⁾(a.x), sinc1⁽ᵏ⁺¹⁾(a.x)*a.dx)
template <int K, class T, class = typename std::enable_if<is_dual<T>::value>::type> inline T sinc1k(const T& a) {
    T r; r.x = sinc1k<K>(a.x);
    auto d = sinc1k<K + 1>(a.x);
    for (int i = 0; i < npart<T>::n(); ++i) r.dx[i] = d * a.dx[i];
    return r;
}
template <class T> inline T sinc1(const T& a) { return sinc1k<0>(a); }

// scac(x) = sinc1(acos(x)), with a Taylor series about x=1 (Rotations.jl:61-68). Generic code, differentiated by rule-chasing.
template <class T> inline T scac(const T& x) {
    T dx = x - 1.;
    if (std::fabs(VALUE(dx)) > 1e-3) return sinc1(dacos(x));
    return 1. + dx * (1. / 3 + dx * (-2. / 90 + dx * (0.0052911879917544626 + dx * (-0.0016229317117234072 + dx * (0.0005625)))));
}

// ---------------------------------------------------------------- rotation algebra
template <class T> inline M33<T> spin(const V3<T>& v) {                                          // Rotations.jl:81
    M33<T> s;
    s.m[0] = T(0.); s.m[1] = v[2];  s.m[2] = -v[1];
    s.m[3] = -v[2]; s.m[4] = T(0.); s.m[5] = v[0];
    s.m[6] = v[1];  s.m[7] = -v[0]; s.m[8] = T(0.);
    return s;
}
template <class T> inline V3<T> spin_inv(const M33<T>& m) {                                      // Rotations.jl:90
    V3<T> v;
    v[0] = (m(2, 1) - m(1, 2)) / 2.;
    v[1] = (m(0, 2) - m(2, 0)) / 2.;
    v[2] = (m(1, 0) - m(0, 1)) / 2.;
    return v;
}
template <class T> inline T trace(const M33<T>& m) { return (m(0, 0) + m(1, 1)) + m(2, 2); }     // Rotations.jl:96
template <class T> inline V3<T> rodrigues_inv(const M33<T>& m) {                                 // Rotations.jl:105
    V3<T> s = spin_inv(m);
    T d = scac((trace(m) - 1.) / 2.);
    V3<T> v; for (int i = 0; i < 3; ++i) v[i] = s[i] / d;
    return v;
}
template <class T> inline T norm3(const V3<T>& v) {                                              // Rotations.jl:106-112
    T n = dsqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
    if (VALUE(n) < 1e-14) n = T(0.);
    return n;
}
template <class T> inline M33<T> spin2(const M33<T>& S) {                                        // Rotations.jl:114-122
    T ab = S(1, 2) * S(2, 0);
    T ca = S(0, 1) * S(1, 2);
    T bc = S(2, 0) * S(0, 1);
    T ma2 = S(2, 1) * S(1, 2);
    T mb2 = S(2, 0) * S(0, 2);
    T mc2 = S(1, 0) * S(0, 1);
    M33<T> r;
    r.m[0] = mc2 + mb2; r.m[1] = ab;        r.m[2] = ca;
    r.m[3] = ab;        r.m[4] = ma2 + mc2; r.m[5] = bc;
    r.m[6] = ca;        r.m[7] = bc;        r.m[8] = ma2 + mb2;
    return r;
}
template <class T> inline M33<T> rodrigues(const V3<T>& v) {                                     // Rotations.jl:131-135
    M33<T> S = spin(v);
    T th = norm3(v);
    T a = sinc1(th);
    T b = ipow(sinc1(th / 2.), 2) / 2.;
    M33<T> S2 = spin2(S);
    M33<T> r;
    // (I + a.*S) + b.*S²  : UniformScaling + SMatrix adds 1 on the diagonal
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) {
            T t = a * S(i, j);
            if (i == j) t = 1. + t;
            r(i, j) = t + b * S2(i, j);
        }
    return r;
}

}  // namespace orc
