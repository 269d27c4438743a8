// Statically sparse two-slot dual numbers.
//
// SD<A,B> = value + partial d0 (present iff A) + partial d1 (present iff B).  Absent partials are compile-time zeros:
// operators return SD<A1|A2, B1|B2> and emit no arithmetic for absent slots.  The beam kernel gives every lane two seed
// directions — slot 0 a rotation dof, slot 1 a translation dof (or U-dof) — so the whole corotational rotation algebra
// (toolbox/BeamElement.jl:192-208), which does not depend on translations, runs as SD<1,0> and costs one direction, while
// everything downstream carries both.  Same differentiation rules as Dual<W> (src/Adiff.jl:223-266).
#pragma once
#include "dual.cuh"

namespace mb {

template <bool A, bool B> struct SD;
template <> struct SD<false, false> { double v; };
template <> struct SD<true, false> { double v, d0; };
template <> struct SD<false, true> { double v, d1; };
template <> struct SD<true, true> { double v, d0, d1; };

template <bool A, bool B> MB_HD double sd0(const SD<A, B>& x) { if constexpr (A) return x.d0; else return 0.; }
template <bool A, bool B> MB_HD double sd1(const SD<A, B>& x) { if constexpr (B) return x.d1; else return 0.; }
template <bool A, bool B> MB_HD void set0(SD<A, B>& x, double v) { if constexpr (A) x.d0 = v; }
template <bool A, bool B> MB_HD void set1(SD<A, B>& x, double v) { if constexpr (B) x.d1 = v; }
template <bool A, bool B> MB_HD double value(const SD<A, B>& a) { return a.v; }
template <bool A, bool B> struct Make<SD<A, B>> {
    static MB_HD SD<A, B> c(double v) { SD<A, B> r; r.v = v; set0(r, 0.); set1(r, 0.); return r; }
};
template <bool A1, bool B1, bool A2, bool B2> struct Widen<SD<A1, B1>, SD<A2, B2>> {
    static_assert((A1 || !A2) && (B1 || !B2), "widen cannot drop a partial");
    static MB_HD SD<A1, B1> w(const SD<A2, B2>& x) { SD<A1, B1> r; r.v = x.v; set0(r, sd0(x)); set1(r, sd1(x)); return r; }
};
template <bool A, bool B> struct Widen<SD<A, B>, SD<A, B>> { static MB_HD SD<A, B> w(const SD<A, B>& x) { return x; } };

// per-slot combination rules
#define MB_SLOT(P1, P2, both, only1, only2) ((P1 && P2) ? (both) : (P1 ? (only1) : (only2)))

template <bool A1, bool B1, bool A2, bool B2> MB_HD SD<A1 || A2, B1 || B2> operator+(const SD<A1, B1>& a, const SD<A2, B2>& b) {
    SD<A1 || A2, B1 || B2> r; r.v = a.v + b.v;
    if constexpr (A1 && A2) r.d0 = a.d0 + b.d0; else if constexpr (A1) r.d0 = a.d0; else if constexpr (A2) r.d0 = b.d0;
    if constexpr (B1 && B2) r.d1 = a.d1 + b.d1; else if constexpr (B1) r.d1 = a.d1; else if constexpr (B2) r.d1 = b.d1;
    return r;
}
template <bool A, bool B> MB_HD SD<A, B> operator+(const SD<A, B>& a, double b) { SD<A, B> r = a; r.v = a.v + b; return r; }
template <bool A, bool B> MB_HD SD<A, B> operator+(double a, const SD<A, B>& b) { SD<A, B> r = b; r.v = a + b.v; return r; }
template <bool A1, bool B1, bool A2, bool B2> MB_HD SD<A1 || A2, B1 || B2> operator-(const SD<A1, B1>& a, const SD<A2, B2>& b) {
    SD<A1 || A2, B1 || B2> r; r.v = a.v - b.v;
    if constexpr (A1 && A2) r.d0 = a.d0 - b.d0; else if constexpr (A1) r.d0 = a.d0; else if constexpr (A2) r.d0 = -b.d0;
    if constexpr (B1 && B2) r.d1 = a.d1 - b.d1; else if constexpr (B1) r.d1 = a.d1; else if constexpr (B2) r.d1 = -b.d1;
    return r;
}
template <bool A, bool B> MB_HD SD<A, B> operator-(const SD<A, B>& a, double b) { SD<A, B> r = a; r.v = a.v - b; return r; }
template <bool A, bool B> MB_HD SD<A, B> operator-(double a, const SD<A, B>& b) { SD<A, B> r; r.v = a - b.v; set0(r, -sd0(b)); set1(r, -sd1(b)); return r; }
template <bool A, bool B> MB_HD SD<A, B> operator-(const SD<A, B>& a) { SD<A, B> r; r.v = -a.v; set0(r, -sd0(a)); set1(r, -sd1(a)); return r; }
template <bool A1, bool B1, bool A2, bool B2> MB_HD SD<A1 || A2, B1 || B2> operator*(const SD<A1, B1>& a, const SD<A2, B2>& b) {
    SD<A1 || A2, B1 || B2> r; r.v = a.v * b.v;
    if constexpr (A1 && A2) r.d0 = fma(a.v, b.d0, a.d0 * b.v); else if constexpr (A1) r.d0 = a.d0 * b.v; else if constexpr (A2) r.d0 = a.v * b.d0;
    if constexpr (B1 && B2) r.d1 = fma(a.v, b.d1, a.d1 * b.v); else if constexpr (B1) r.d1 = a.d1 * b.v; else if constexpr (B2) r.d1 = a.v * b.d1;
    return r;
}
template <bool A, bool B> MB_HD SD<A, B> operator*(const SD<A, B>& a, double b) { SD<A, B> r; r.v = a.v * b; set0(r, sd0(a) * b); set1(r, sd1(a) * b); return r; }
template <bool A, bool B> MB_HD SD<A, B> operator*(double a, const SD<A, B>& b) { SD<A, B> r; r.v = a * b.v; set0(r, a * sd0(b)); set1(r, a * sd1(b)); return r; }
template <bool A, bool B> MB_HD SD<A, B> mb_rcp(const SD<A, B>& a) { SD<A, B> r; r.v = 1.0 / a.v; double m = -r.v * r.v; set0(r, m * sd0(a)); set1(r, m * sd1(a)); return r; }
template <bool A1, bool B1, bool A2, bool B2> MB_HD SD<A1 || A2, B1 || B2> operator/(const SD<A1, B1>& a, const SD<A2, B2>& b) {
    SD<A1 || A2, B1 || B2> r; const double inv = 1.0 / b.v; r.v = a.v * inv;
    if constexpr (A1 && A2) r.d0 = (a.d0 - r.v * b.d0) * inv; else if constexpr (A1) r.d0 = a.d0 * inv; else if constexpr (A2) r.d0 = -(r.v * b.d0) * inv;
    if constexpr (B1 && B2) r.d1 = (a.d1 - r.v * b.d1) * inv; else if constexpr (B1) r.d1 = a.d1 * inv; else if constexpr (B2) r.d1 = -(r.v * b.d1) * inv;
    return r;
}
template <bool A, bool B> MB_HD SD<A, B> operator/(const SD<A, B>& a, double b) { return a * (1.0 / b); }
template <bool A, bool B> MB_HD SD<A, B> operator/(double a, const SD<A, B>& b) { SD<A, B> r; const double inv = 1.0 / b.v; r.v = a * inv; const double m = -r.v * inv; set0(r, m * sd0(b)); set1(r, m * sd1(b)); return r; }
template <bool A, bool B> MB_HD SD<A, B> mb_sqrt(const SD<A, B>& a) { SD<A, B> r; r.v = sqrt(a.v); const double m = 0.5 / r.v; set0(r, m * sd0(a)); set1(r, m * sd1(a)); return r; }
template <bool A, bool B> MB_HD SD<A, B> mb_acos(const SD<A, B>& a) { SD<A, B> r; r.v = acos(a.v); const double m = -1.0 / sqrt(1.0 - a.v * a.v); set0(r, m * sd0(a)); set1(r, m * sd1(a)); return r; }
template <int K, bool A, bool B> MB_HD SD<A, B> sinc1k(const SD<A, B>& a) { SD<A, B> r; r.v = sinc1k<K>(a.v); const double m = sinc1k<K + 1>(a.v); set0(r, m * sd0(a)); set1(r, m * sd1(a)); return r; }
template <bool A, bool B> MB_HD SD<A, B> sqr_ref(const SD<A, B>& a) { return a * a; }
template <bool A, bool B> MB_HD SD<A, B> apply_fn(const SD<A, B>& x, const double* F) { SD<A, B> r; r.v = F[0]; set0(r, F[1] * sd0(x)); set1(r, F[1] * sd1(x)); return r; }

}  // namespace mb
