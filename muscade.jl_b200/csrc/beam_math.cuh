// EulerBeam3D residual + tangent, B200 formulation.
//
// What the reference computes (toolbox/BeamElement.jl:151-174):
//     R_j = Σ_gp dL_gp [ fᵢ·∂ε/∂X₀ⱼ + mᵢ·∂κ_gp/∂X₀ⱼ + fₑ·∂x_gp/∂X₀ⱼ + mₑ·∂vₛₘ/∂X₀ⱼ ]          (virtual work)
// as a ∂ℝ{1,Np} whose partials are the tangent.  It obtains the Jacobians AND their partials (Hessian·seed) from a
// second-order forward pass in 12 variables with a 3→6→12 chain rule (Taylor.jl:225-323) — 6.5e5 flops/element.
//
// Same mathematics here, evaluated as forward-over-reverse:
//   1. forward sweep of the corotated kinematics (BeamElement.jl:176-208) in the lane's number type
//        S = Dual<W>            (static)        or   Jet<Dual<W>>  (Newmark / DirectXUA: value, velocity, acceleration)
//      giving the resultants fᵢ,mᵢ,fₑ,mₑ (BeamElement.jl:28-64) with their directional partials;
//   2. reverse (adjoint) sweep of the order-0 kinematics with cotangents w = dL·(fᵢ,mᵢ,fₑ,mₑ), in Dual<W>:
//      the value part is R = Jᵀw, the dual part is ∂R/∂seed = Jᵀ∂w + (∂Jᵀ)w, i.e. material + geometric tangent.
// Each lane carries a few of the Np seed directions; no second-order dual, no McLaurin expansion, nothing leaves registers.
// The code is generic over three number types (struct Num): TR for what depends on the rotation dofs only, TU for what
// depends on translation dofs / U only, TS for everything else.  With Num = NumSD they are SD<1,0>, SD<0,1>, SD<1,1>
// (sdual.cuh): one lane carries (rotation dof l, translation dof l) and the rotation algebra costs one direction, not two.
// With Num = NumDual<W> all three are the dense Dual<W>.
// Branches of the reference's special functions (sinc1 family thresholds, scac series, norm3 cutoff, drag sign)
// are reproduced on VALUE, as the reference does (Adiff.jl:198-202).
#pragma once
#include <type_traits>
#include "sdual.cuh"

namespace mb {

template <class T> struct Vec3 { T a[3]; MB_HD T& operator[](int i) { return a[i]; } MB_HD const T& operator[](int i) const { return a[i]; } };
template <class T> struct Mat3 { T a[9]; MB_HD T& operator()(int i, int j) { return a[i + 3 * j]; } MB_HD const T& operator()(int i, int j) const { return a[i + 3 * j]; } };

template <class A, class B> MB_HD auto mul(const Mat3<A>& a, const Mat3<B>& b) -> Mat3<decltype(a.a[0] * b.a[0])> {
    Mat3<decltype(a.a[0] * b.a[0])> c;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) c(i, j) = (a(i, 0) * b(0, j) + a(i, 1) * b(1, j)) + a(i, 2) * b(2, j);
    return c;
}
// a * bᵀ
template <class A, class B> MB_HD auto mul_nt(const Mat3<A>& a, const Mat3<B>& b) -> Mat3<decltype(a.a[0] * b.a[0])> {
    Mat3<decltype(a.a[0] * b.a[0])> c;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) c(i, j) = (a(i, 0) * b(j, 0) + a(i, 1) * b(j, 1)) + a(i, 2) * b(j, 2);
    return c;
}
// aᵀ * b
template <class A, class B> MB_HD auto mul_tn(const Mat3<A>& a, const Mat3<B>& b) -> Mat3<decltype(a.a[0] * b.a[0])> {
    Mat3<decltype(a.a[0] * b.a[0])> c;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) c(i, j) = (a(0, i) * b(0, j) + a(1, i) * b(1, j)) + a(2, i) * b(2, j);
    return c;
}
template <class A, class B> MB_HD auto mulv(const Mat3<A>& a, const Vec3<B>& b) -> Vec3<decltype(a.a[0] * b.a[0])> {
    Vec3<decltype(a.a[0] * b.a[0])> c;
#pragma unroll
    for (int i = 0; i < 3; ++i) c[i] = (a(i, 0) * b[0] + a(i, 1) * b[1]) + a(i, 2) * b[2];
    return c;
}
// aᵀ * v
template <class A, class B> MB_HD auto mulv_t(const Mat3<A>& a, const Vec3<B>& b) -> Vec3<decltype(a.a[0] * b.a[0])> {
    Vec3<decltype(a.a[0] * b.a[0])> c;
#pragma unroll
    for (int i = 0; i < 3; ++i) c[i] = (a(0, i) * b[0] + a(1, i) * b[1]) + a(2, i) * b[2];
    return c;
}

// ------------------------------------------------------------------------------------------------ rotations (forward)
// Rodrigues(v) = I + sinc1(θ)·S + ½sinc1(θ/2)²·S²   (toolbox/Rotations.jl:131-135, spin² :114-122, norm3 :106-112)
template <class T> struct RodAux { T a, b, th; double Sth[4], Sh[4]; bool small; };     // what the adjoint needs again: coefficients, θ, sinc1 packs at θ and θ/2
template <class T> MB_FN Mat3<T> rodrigues(const Vec3<T>& v, RodAux<T>& aux) {
    T t2 = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2];
    T a, b;
    T th = mb_sqrt(t2);
    aux.small = value(th) < 1e-14;
    if (aux.small) { a = Make<T>::c(1.0); b = Make<T>::c(0.5); }                 // θ := 0 without partials
    else {
        sinc_pack(value(th), aux.Sth); sinc_pack(0.5 * value(th), aux.Sh);
        a = apply_fn(th, aux.Sth); T c = apply_fn(th * 0.5, aux.Sh); b = sqr_ref(c) * 0.5;
    }
    aux.a = a; aux.b = b; aux.th = th;
    T v00 = v[0] * v[0], v11 = v[1] * v[1], v22 = v[2] * v[2];
    T b01 = b * (v[0] * v[1]), b02 = b * (v[0] * v[2]), b12 = b * (v[1] * v[2]);
    T a0 = a * v[0], a1 = a * v[1], a2 = a * v[2];
    Mat3<T> r;
    r(0, 0) = 1.0 - b * (v11 + v22); r(1, 1) = 1.0 - b * (v00 + v22); r(2, 2) = 1.0 - b * (v00 + v11);
    r(1, 0) = b01 + a2; r(0, 1) = b01 - a2;
    r(2, 0) = b02 - a1; r(0, 2) = b02 + a1;
    r(2, 1) = b12 + a0; r(1, 2) = b12 - a0;
    return r;
}
// Rodrigues⁻¹(m) = spin⁻¹(m)/scac((tr m − 1)/2)   (Rotations.jl:90,105)
template <class T> struct RinvAux { T x, s; double C[4]; };     // scac pack at value(x)
template <class T> MB_FN Vec3<T> rodrigues_inv(const Mat3<T>& m, RinvAux<T>& aux) {
    T x = (((m(0, 0) + m(1, 1)) + m(2, 2)) - 1.0) * 0.5;
    scac_pack(value(x), aux.C);
    T s = apply_fn(x, aux.C);
    aux.x = x; aux.s = s;
    T is = mb_rcp(s);
    Vec3<T> v;
    v[0] = ((m(2, 1) - m(1, 2)) * 0.5) * is;
    v[1] = ((m(0, 2) - m(2, 0)) * 0.5) * is;
    v[2] = ((m(1, 0) - m(0, 1)) * 0.5) * is;
    return v;
}

// ------------------------------------------------------------------------------------------------ number-type bundles
template <class TR_, class TU_, class TS_> struct Num { using TR = TR_; using TU = TU_; using TS = TS_; };
template <int W> using NumDual = Num<Dual<W>, Dual<W>, Dual<W>>;
using NumSD = Num<SD<true, false>, SD<false, true>, SD<true, true>>;
using NumVal = Num<SD<false, false>, SD<false, false>, SD<false, false>>;     // plain values behind the SD interface
template <class N> using NumJet = Num<Jet<typename N::TR>, Jet<typename N::TU>, Jet<typename N::TS>>;
// time-jets of lanes whose seeds sit on the VELOCITIES (DirectXUA's ∂/∂X′ lanes): every order-0 coefficient is free of partials by construction, so it is a plain value
// and products with it cost a direction less (the velocity and acceleration coefficients carry the partials)
using PlainSD = SD<false, false>;
template <class N> using NumJet1 = Num<Jet<PlainSD, typename N::TR, typename N::TR>, Jet<PlainSD, typename N::TU, typename N::TU>, Jet<PlainSD, typename N::TS, typename N::TS>>;

// ------------------------------------------------------------------------------------------------ scratch (explicit spill space)
// The reverse sweep needs rₛ₂ and Rodrigues(Δvᵧ) only at its very end; parking them in per-thread scratch (shared memory on the
// device) instead of registers keeps ptxas from spilling to local memory, whose stores all reach DRAM (ncu: 2 KB/element).
template <class T> struct Comp;
template <> struct Comp<double> { static constexpr int n = 1; static MB_HD double get(const double& x, int) { return x; } static MB_HD void set(double& x, int, double v) { x = v; } };
template <int W> struct Comp<Dual<W>> {
    static constexpr int n = W + 1;
    static MB_HD double get(const Dual<W>& x, int i) { return i == 0 ? x.v : x.d[i - 1]; }
    static MB_HD void set(Dual<W>& x, int i, double v) { if (i == 0) x.v = v; else x.d[i - 1] = v; }
};
template <bool A, bool B> struct Comp<SD<A, B>> {
    static constexpr int n = 1 + (A ? 1 : 0) + (B ? 1 : 0);
    static MB_HD double get(const SD<A, B>& x, int i) { return i == 0 ? x.v : ((A && i == 1) ? sd0(x) : sd1(x)); }
    static MB_HD void set(SD<A, B>& x, int i, double v) { if (i == 0) x.v = v; else if (A && i == 1) set0(x, v); else set1(x, v); }
};
struct HostScratch {                       // host build / kernels that do not stash
    static constexpr bool enabled = false;
    MB_HD void put(int, double) {}
    MB_HD double get(int) const { return 0.; }
};
template <class SC, class T> MB_HD void stash(SC& sc, int slot0, const Mat3<T>& m) {
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int c = 0; c < Comp<T>::n; ++c) sc.put(slot0 + k * Comp<T>::n + c, Comp<T>::get(m.a[k], c));
}
template <class SC, class T> MB_HD void unstash(const SC& sc, int slot0, Mat3<T>& m) {
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int c = 0; c < Comp<T>::n; ++c) Comp<T>::set(m.a[k], c, sc.get(slot0 + k * Comp<T>::n + c));
}

// ------------------------------------------------------------------------------------------------ rotations (adjoint)
// v̄ += ∂Rodrigues(v)ᵀ·R̄          (v, aux in TR;  R̄, v̄ in TS)
template <class TR, class TS> MB_FN void rodrigues_adj(const Vec3<TR>& v, const RodAux<TR>& aux, const Mat3<TS>& Rb, Vec3<TS>& vb) {
    const TR &a = aux.a, &b = aux.b;
    TS k0 = Rb(2, 1) - Rb(1, 2), k1 = Rb(0, 2) - Rb(2, 0), k2 = Rb(1, 0) - Rb(0, 1);          // skew part
    TS s01 = Rb(0, 1) + Rb(1, 0), s02 = Rb(0, 2) + Rb(2, 0), s12 = Rb(1, 2) + Rb(2, 1);        // symmetric part
    TS d0 = Rb(1, 1) + Rb(2, 2), d1 = Rb(0, 0) + Rb(2, 2), d2 = Rb(0, 0) + Rb(1, 1);
    TS g0 = (s01 * v[1] + s02 * v[2]) - 2.0 * (v[0] * d0);                                     // ∂(S²:R̄)/∂v
    TS g1 = (s01 * v[0] + s12 * v[2]) - 2.0 * (v[1] * d1);
    TS g2 = (s02 * v[0] + s12 * v[1]) - 2.0 * (v[2] * d2);
    vb[0] = vb[0] + (a * k0 + b * g0);
    vb[1] = vb[1] + (a * k1 + b * g1);
    vb[2] = vb[2] + (a * k2 + b * g2);
    if (!aux.small) {
        TS ab = (v[0] * k0 + v[1] * k1) + v[2] * k2;                                           // ā = S:R̄
        TS bb = 0.5 * ((v[0] * g0 + v[1] * g1) + v[2] * g2);                                   // b̄ = S²:R̄ (Euler: homogeneous degree 2)
        const TR& th = aux.th;
        TR hth = th * 0.5;
        TR da = apply_fn(th, aux.Sth + 1);                                                       // sinc1′(θ)
        TR db = (apply_fn(hth, aux.Sh) * apply_fn(hth, aux.Sh + 1)) * 0.5;                       // d/dθ ½sinc1(θ/2)²
        TS tb = (ab * da + bb * db) / th;                                                      // θ̄/θ
        vb[0] = vb[0] + tb * v[0]; vb[1] = vb[1] + tb * v[1]; vb[2] = vb[2] + tb * v[2];
    }
}
// M̄ += ∂Rodrigues⁻¹(M)ᵀ·v̄ , v = Rodrigues⁻¹(M)
template <class TR, class TS> MB_FN void rodrigues_inv_adj(const Vec3<TR>& v, const RinvAux<TR>& aux, const Vec3<TS>& vb, Mat3<TS>& Mb) {
    TR is = mb_rcp(aux.s);
    TS sb = -(((vb[0] * v[0] + vb[1] * v[1]) + vb[2] * v[2]) * is);
    TS xb = (sb * apply_fn(aux.x, aux.C + 1)) * 0.5;
    TS w0 = (vb[0] * is) * 0.5, w1 = (vb[1] * is) * 0.5, w2 = (vb[2] * is) * 0.5;
    Mb(0, 0) = Mb(0, 0) + xb; Mb(1, 1) = Mb(1, 1) + xb; Mb(2, 2) = Mb(2, 2) + xb;
    Mb(2, 1) = Mb(2, 1) + w0; Mb(1, 2) = Mb(1, 2) - w0;
    Mb(0, 2) = Mb(0, 2) + w1; Mb(2, 0) = Mb(2, 0) - w1;
    Mb(1, 0) = Mb(1, 0) + w2; Mb(0, 1) = Mb(0, 1) - w2;
}

// ------------------------------------------------------------------------------------------------ element data
constexpr int NGP = 4;
struct BeamMat { double EA, EI2, EI3, GJ, mu, iota1, w, Ca1, Cl1, Cq1, Ca2, Cl2, Cq2, Ca3, Cl3, Cq3; };   // BeamElement.jl:6-23
struct BeamGeo { double cm[3]; Mat3<double> rm; double tgm[3]; double L; };                                // the non-constant part of BeamElement.jl:87-103

// Gauss abscissae, weights/L and shape values that the reference stores per element (BeamElement.jl:141-146) are compile-time
// constants times powers of L; computed from the Gauss point index so that the Gauss loop can stay rolled.
struct GpConst { double z, w, ya, yu, yv, ku; };
MB_HD GpConst gp_const(int gp) {
    const double s65 = 1.0954451150103321;    // sqrt(6/5)
    const double s30 = 5.477225575051661;     // sqrt(30)
    const bool outer = (gp == 0) || (gp == 3);
    const double zm = 0.5 * sqrt(3. / 7 + (outer ? 2. / 7 : -2. / 7) * s65);
    GpConst c;
    c.z = (gp < 2) ? -zm : zm;
    c.w = 0.5 * (outer ? (18 - s30) : (18 + s30)) / 36;
    c.ya = 2 * c.z; c.yu = -4 * (c.z * c.z * c.z) + 3 * c.z; c.yv = c.z * c.z - 0.25; c.ku = -24 * c.z;
    return c;
}
// Gauss loop unrolling, measured on B200 (2e6 elements): static 10.3 ms rolled vs 11.4 ms unrolled; Newmark 26.7 ms rolled vs 24.6 ms unrolled
#ifndef MB_GP_UNROLL_STATIC
#define MB_GP_UNROLL_STATIC 1
#endif
#ifndef MB_GP_UNROLL_DYN
#define MB_GP_UNROLL_DYN 1
#endif
#define MB_PRAGMA_(x) _Pragma(#x)
#define MB_PRAGMA(x) MB_PRAGMA_(x)

// ------------------------------------------------------------------------------------------------ forward kinematics
template <class N> struct BeamFwd {
    using TR = typename N::TR; using TU = typename N::TU; using TS = typename N::TS;
    Vec3<TR> v1, v2, dv, vsm, vl;             // rotation-only quantities; vl = rₛₘᵀΔvᵧ
    Mat3<TR> r1, r2, rd, r;                   // rₛ₁, rₛ₂, Rodrigues(Δvᵧ), rₛₘ
    RodAux<TR> a1, a2, ad;
    RinvAux<TR> im, ir;
    Vec3<TU> dp, cs;                          // dp = uᵧ₂ + tgₘ/2 − cₛ ; cs = cₛ + cₘ
    Vec3<TS> ul, q; TS eps, qn;               // q = uₗ₂ + (L/2,0,0), qn = |q|
};
// corotated{:direct} + ε  (BeamElement.jl:181,192-208).  u1,u2: translations (TU); v1,v2: rotation vectors (TR)
// VSM = false skips vₛₘ = Rodrigues⁻¹(rₛₘ): it only feeds the roll-inertia cotangent v̄ₛₘ, which is zero without accelerations.
template <class N, bool VSM = true> MB_HD void beam_forward(const BeamGeo& g, const Vec3<typename N::TU>& u1, const Vec3<typename N::TR>& v1,
                                                            const Vec3<typename N::TU>& u2, const Vec3<typename N::TR>& v2, BeamFwd<N>& f) {
    using TR = typename N::TR; using TU = typename N::TU;
    f.v1 = v1; f.v2 = v2;
    f.r1 = rodrigues(f.v1, f.a1);
    f.r2 = rodrigues(f.v2, f.a2);
    Mat3<TR> M = mul_nt(f.r2, f.r1);
    Vec3<TR> h = rodrigues_inv(M, f.im);
#pragma unroll
    for (int i = 0; i < 3; ++i) f.dv[i] = 0.5 * h[i];
    f.rd = rodrigues(f.dv, f.ad);
    f.r = mul(mul(f.rd, f.r1), g.rm);
    if constexpr (VSM) f.vsm = rodrigues_inv(f.r, f.ir);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        TU cs = 0.5 * (u1[i] + u2[i]);
        f.dp[i] = (u2[i] + g.tgm[i] * 0.5) - cs;
        f.cs[i] = cs + g.cm[i];
    }
    f.q = mulv_t(f.r, f.dp);
    f.ul = f.q; f.ul[0] = f.ul[0] - g.L * 0.5;
    f.vl = mulv_t(f.r, f.dv);
    f.qn = mb_sqrt((f.q[0] * f.q[0] + f.q[1] * f.q[1]) + f.q[2] * f.q[2]);
    f.eps = f.qn * (2.0 / g.L) - 1.0;
}
// Gauss point position in the corotated frame  p = tgₑζ + y  (BeamElement.jl:183-187)
template <class A, class B> MB_HD auto beam_gp_local(const GpConst& c, double L, const Vec3<A>& ul, const Vec3<B>& vl) -> Vec3<decltype(ul[0] + vl[0])> {
    Vec3<decltype(ul[0] + vl[0])> p;
    double yv = c.yv * L;
    p[0] = widen<decltype(ul[0] + vl[0])>(c.ya * ul[0] + L * c.z);
    p[1] = c.yu * ul[1] + yv * vl[2];
    p[2] = c.yu * ul[2] - yv * vl[1];
    return p;
}

// ------------------------------------------------------------------------------------------------ resultants → cotangents
// external force at a Gauss point (BeamElement.jl:28-58) from velocity/acceleration of the point and the frame rₛₘ
template <class TR, class S> MB_HD Vec3<S> beam_fe(const BeamMat& m, const Mat3<TR>& r0, const Vec3<S>& x1, const Vec3<S>& x2) {
    Vec3<S> xl1 = mulv_t(r0, x1), xl2 = mulv_t(r0, x2);
    const double Ca[3] = {m.Ca1, m.Ca2, m.Ca3}, Cl[3] = {m.Cl1, m.Cl2, m.Cl3}, Cq[3] = {m.Cq1, m.Cq2, m.Cq3};
    Vec3<S> fl;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        S fq = Cq[i] * (xl1[i] * xl1[i]);
        if (value(xl1[i]) < 0) fq = -fq;
        fl[i] = (Ca[i] * xl2[i] + Cl[i] * xl1[i]) + fq;
    }
    Vec3<S> fg = mulv(r0, fl);
    Vec3<S> fe;
#pragma unroll
    for (int i = 0; i < 3; ++i) fe[i] = m.mu * x2[i] + fg[i];
    fe[2] = fe[2] + m.w;
    return fe;
}

// ------------------------------------------------------------------------------------------------ reverse sweep
// Adjoint accumulators of the quantities the Gauss loop feeds: r̄ₛₘ, ūₗ, v̄ₗ, c̄ₛₘ
template <class S> struct BeamAcc { Mat3<S> rb; Vec3<S> ulb, vlb, cb; };
// Internal moments mᵢ = (GJκ₁, EI₃κ₂, EI₂κ₃) (BeamElement.jl:62) with κ = (κₐvₗ₁, κᵤuₗ₂+κᵥvₗ₃, κᵤuₗ₃−κᵥvₗ₂) (:184), κₐ = κᵥ = 2/L, κᵤ = −24ζ/L²: the
// Gauss sum Σ dL·κ̄·∂κ is a quadratic form in the abscissae and is taken in closed form — Σw = 1, Σwζ = 0, Σwζ² = 1/12 (4-point rule,
// exact to degree 7) — which gives the textbook coefficients 48EI/L³ on uₗ and 4EI/L, 4GJ/L on vₗ.
template <class N, class S = typename N::TS> MB_HD void beam_internal_reverse(double L, const BeamMat& m, const BeamFwd<N>& f, BeamAcc<S>& a) {
    const double c48 = 48.0 / (L * L * L), c4 = 4.0 / L;
    a.ulb[1] = a.ulb[1] + (m.EI3 * c48) * f.ul[1];
    a.ulb[2] = a.ulb[2] + (m.EI2 * c48) * f.ul[2];
    a.vlb[0] = a.vlb[0] + (m.GJ * c4) * f.vl[0];
    a.vlb[1] = a.vlb[1] + (m.EI2 * c4) * f.vl[1];
    a.vlb[2] = a.vlb[2] + (m.EI3 * c4) * f.vl[2];
}
// One Gauss point, external loads: x̄ = dL·fₑ given; x = rₛₘ p + cₛₘ, p = (yₐuₗ₁+Lζ, yᵤuₗ₂+yᵥvₗ₃, yᵤuₗ₃−yᵥvₗ₂)  (BeamElement.jl:183-187)
template <class N, class S = typename N::TS> MB_HD void beam_gp_reverse(const GpConst& c, double L, const BeamFwd<N>& f, const Vec3<S>& xb, BeamAcc<S>& a) {
    const double yv = c.yv * L;
    auto p = beam_gp_local(c, L, f.ul, f.vl);
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) a.rb(i, j) = a.rb(i, j) + xb[i] * p[j];
    Vec3<S> pb = mulv_t(f.r, xb);
#pragma unroll
    for (int i = 0; i < 3; ++i) a.cb[i] = a.cb[i] + xb[i];
    a.ulb[0] = a.ulb[0] + c.ya * pb[0];
    a.ulb[1] = a.ulb[1] + c.yu * pb[1];
    a.ulb[2] = a.ulb[2] + c.yu * pb[2];
    a.vlb[1] = a.vlb[1] - yv * pb[2];
    a.vlb[2] = a.vlb[2] + yv * pb[1];
}
// The same summed over the Gauss points when fₑ does not depend on the point (statics: weight and −U): Σ dL·p = (0, −L²vₗ₃/6, L²vₗ₂/6)
// (Σw·yₐ = Σw·yᵤ = Σwζ = 0, Σw·yᵥ = Σwζ² − ¼ = −1/6), Σ dL = L.
template <class N, class S = typename N::TS> MB_HD void beam_uniform_reverse(double L, const BeamFwd<N>& f, const Vec3<S>& fe, BeamAcc<S>& a) {
    const double k6 = L * L / 6.0;
    auto P1 = -k6 * f.vl[2]; auto P2 = k6 * f.vl[1];
#pragma unroll
    for (int i = 0; i < 3; ++i) { a.rb(i, 1) = a.rb(i, 1) + fe[i] * P1; a.rb(i, 2) = a.rb(i, 2) + fe[i] * P2; a.cb[i] = a.cb[i] + L * fe[i]; }
    Vec3<S> gl = mulv_t(f.r, fe);
    a.vlb[1] = a.vlb[1] + k6 * gl[2];
    a.vlb[2] = a.vlb[2] - k6 * gl[1];
}
// Everything upstream of the Gauss loop: ε, uₗ/vₗ, vₛₘ, the corotated frame and the three Rodrigues maps → X̄[12]
template <class N, class SC, class S = typename N::TS, bool VSM = true> MB_HD void beam_reverse_rot(const BeamGeo& g, const BeamFwd<N>& f, const S& epsb, const Vec3<S>& vsmb,
                                                                                                   BeamAcc<S>& a, S* Xb, const SC& sc) {
    const double L = g.L;
    S z = Make<S>::c(0.);
    Mat3<S>& rb = a.rb; Vec3<S>&ulb = a.ulb, &vlb = a.vlb, &cb = a.cb;
    // ε = 2|q|/L − 1
    {
        S k = (epsb * (2.0 / L)) / f.qn;
#pragma unroll
        for (int i = 0; i < 3; ++i) ulb[i] = ulb[i] + k * f.q[i];
    }
    // uₗ = rᵀ d' − tgₑ/2 ; vₗ = rᵀ Δv
    Vec3<S> dvb = mulv(f.r, vlb);
    Vec3<S> dpb = mulv(f.r, ulb);
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) rb(i, j) = rb(i, j) + (f.dp[i] * ulb[j] + f.dv[i] * vlb[j]);
    // vₛₘ = Rodrigues⁻¹(r)
    if constexpr (VSM) rodrigues_inv_adj(f.vsm, f.ir, vsmb, rb);
    // r = rd · r1 · rm
    Mat3<S> t = mul_nt(rb, g.rm);                  // r̄ rₘᵀ
    Mat3<S> rdb = mul_nt(t, f.r1);                 // (r̄ rₘᵀ) r₁ᵀ
    Mat3<S> r1b;                                   // r_dᵀ (r̄ rₘᵀ)
    if constexpr (SC::enabled) { Mat3<typename N::TR> rd_; unstash(sc, 9 * Comp<typename N::TR>::n, rd_); r1b = mul_tn(rd_, t); }
    else r1b = mul_tn(f.rd, t);
    // rd = Rodrigues(Δv)
    rodrigues_adj(f.dv, f.ad, rdb, dvb);
    // Δv = ½ Rodrigues⁻¹(M), M = r2 r1ᵀ
    Mat3<S> Mb; for (int i = 0; i < 9; ++i) Mb.a[i] = z;
    Vec3<S> hb{0.5 * dvb[0], 0.5 * dvb[1], 0.5 * dvb[2]};
    Vec3<typename N::TR> h{2.0 * f.dv[0], 2.0 * f.dv[1], 2.0 * f.dv[2]};
    rodrigues_inv_adj(h, f.im, hb, Mb);
    Mat3<S> r2b = mul(Mb, f.r1);                   // M̄ r₁
    Mat3<S> r1b2;                                  // M̄ᵀ r₂
    if constexpr (SC::enabled) { Mat3<typename N::TR> r2_; unstash(sc, 0, r2_); r1b2 = mul_tn(Mb, r2_); }
    else r1b2 = mul_tn(Mb, f.r2);
    for (int i = 0; i < 9; ++i) r1b.a[i] = r1b.a[i] + r1b2.a[i];
    Vec3<S> v1b{z, z, z}, v2b{z, z, z};
    rodrigues_adj(f.v1, f.a1, r1b, v1b);
    rodrigues_adj(f.v2, f.a2, r2b, v2b);
    // d' = (u₂ + tgₘ/2) − ½(u₁+u₂) ; cₛₘ = ½(u₁+u₂) + cₘ
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        S hcs = 0.5 * (cb[i] - dpb[i]);
        Xb[i] = hcs; Xb[3 + i] = v1b[i]; Xb[6 + i] = hcs + dpb[i]; Xb[9 + i] = v2b[i];
    }
}

// ------------------------------------------------------------------------------------------------ whole element
// ND = 1 (static), 2 (first order), 3 (Newmark / DirectXUA second order).
// Xu[ider][6] = translations of node 1,2 (TU), Xv[ider][6] = rotation vectors of node 1,2 (TR), U0[3] (TU) carry the lane's
// seeds.  Returns R[12] in TS (element dof order t1..r3 of node 1, then node 2): value = residual, partials = ∂R/∂(lane's directions).
template <int ND, class N, class SC> MB_HD void beam_residual_n(const BeamGeo& g, const BeamMat& m, const typename N::TU (*Xu)[6], const typename N::TR (*Xv)[6],
                                                     bool udof, const typename N::TU* U0, typename N::TS* R, SC& sc) {
    using TR = typename N::TR; using TU = typename N::TU; using S = typename N::TS;
    const double L = g.L;
    BeamFwd<N> f;
    BeamAcc<S> acc;
    {
        S z = Make<S>::c(0.);
        for (int i = 0; i < 9; ++i) acc.rb.a[i] = z;
        for (int i = 0; i < 3; ++i) { acc.ulb[i] = z; acc.vlb[i] = z; acc.cb[i] = z; }
    }
    Vec3<S> vsmb{Make<S>::c(0.), Make<S>::c(0.), Make<S>::c(0.)};
    if (ND == 1) {
        beam_forward<N, false>(g, Vec3<TU>{Xu[0][0], Xu[0][1], Xu[0][2]}, Vec3<TR>{Xv[0][0], Xv[0][1], Xv[0][2]},
                               Vec3<TU>{Xu[0][3], Xu[0][4], Xu[0][5]}, Vec3<TR>{Xv[0][3], Xv[0][4], Xv[0][5]}, f);
        if constexpr (SC::enabled) { stash(sc, 0, f.r2); stash(sc, 9 * Comp<TR>::n, f.rd); }
        Vec3<S> fe{Make<S>::c(0.), Make<S>::c(0.), Make<S>::c(m.w)};               // weight (BeamElement.jl:37); − U (:169)
        if (udof) for (int i = 0; i < 3; ++i) fe[i] = fe[i] - U0[i];
        beam_uniform_reverse<N>(L, f, fe, acc);
        beam_internal_reverse<N>(L, m, f, acc);
        S epsb = (m.EA * L) * f.eps;                                                // internal axial force (:59): Σ_gp dL·fᵢ = EA·L·ε
        beam_reverse_rot<N, SC, S, false>(g, f, epsb, vsmb, acc, R, sc);
        return;
    } else {
        using NJ = NumJet<N>;
        using JR = typename NJ::TR; using JU = typename NJ::TU; using JS = typename NJ::TS;
        JU XuJ[6]; JR XvJ[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            XuJ[i].c0 = Xu[0][i]; XuJ[i].c1 = Xu[1][i]; XuJ[i].c2 = (ND >= 3) ? Xu[2][i] : Make<TU>::c(0.);
            XvJ[i].c0 = Xv[0][i]; XvJ[i].c1 = Xv[1][i]; XvJ[i].c2 = (ND >= 3) ? Xv[2][i] : Make<TR>::c(0.);
        }
        BeamFwd<NJ> fj;
        beam_forward<NJ>(g, Vec3<JU>{XuJ[0], XuJ[1], XuJ[2]}, Vec3<JR>{XvJ[0], XvJ[1], XvJ[2]}, Vec3<JU>{XuJ[3], XuJ[4], XuJ[5]},
                         Vec3<JR>{XvJ[3], XvJ[4], XvJ[5]}, fj);
        // order-0 state for the reverse sweep
        for (int i = 0; i < 3; ++i) {
            f.v1[i] = fj.v1[i].c0; f.v2[i] = fj.v2[i].c0; f.dv[i] = fj.dv[i].c0; f.vsm[i] = fj.vsm[i].c0;
            f.ul[i] = fj.ul[i].c0; f.vl[i] = fj.vl[i].c0; f.dp[i] = fj.dp[i].c0; f.q[i] = fj.q[i].c0;
        }
        for (int i = 0; i < 9; ++i) { f.r1.a[i] = fj.r1.a[i].c0; f.r2.a[i] = fj.r2.a[i].c0; f.rd.a[i] = fj.rd.a[i].c0; f.r.a[i] = fj.r.a[i].c0; }
        f.a1.a = fj.a1.a.c0; f.a1.b = fj.a1.b.c0; f.a1.th = fj.a1.th.c0; f.a1.small = fj.a1.small;
        f.a2.a = fj.a2.a.c0; f.a2.b = fj.a2.b.c0; f.a2.th = fj.a2.th.c0; f.a2.small = fj.a2.small;
        f.ad.a = fj.ad.a.c0; f.ad.b = fj.ad.b.c0; f.ad.th = fj.ad.th.c0; f.ad.small = fj.ad.small;
        for (int k = 0; k < 4; ++k) { f.a1.Sth[k] = fj.a1.Sth[k]; f.a1.Sh[k] = fj.a1.Sh[k]; f.a2.Sth[k] = fj.a2.Sth[k]; f.a2.Sh[k] = fj.a2.Sh[k];
                                      f.ad.Sth[k] = fj.ad.Sth[k]; f.ad.Sh[k] = fj.ad.Sh[k]; f.im.C[k] = fj.im.C[k]; f.ir.C[k] = fj.ir.C[k]; }
        f.im.x = fj.im.x.c0; f.im.s = fj.im.s.c0; f.ir.x = fj.ir.x.c0; f.ir.s = fj.ir.s.c0;
        f.eps = fj.eps.c0; f.qn = fj.qn.c0;
        if constexpr (SC::enabled) { stash(sc, 0, f.r2); stash(sc, 9 * Comp<TR>::n, f.rd); }
        // external loads at the Gauss points from (x, ẋ, ẍ) and rₛₘ (BeamElement.jl:28-58)
        MB_PRAGMA(unroll MB_GP_UNROLL_DYN)
        for (int gp = 0; gp < NGP; ++gp) {
            const GpConst c = gp_const(gp);
            Vec3<JS> p = beam_gp_local(c, L, fj.ul, fj.vl);
            Vec3<S> x1, x2;
#pragma unroll
            for (int i = 0; i < 3; ++i) {     // velocity / acceleration of x = rₛₘ p + cₛₘ (only these Taylor coefficients are needed)
                x1[i] = ((fj.r(i, 0).c0 * p[0].c1 + fj.r(i, 0).c1 * p[0].c0) + (fj.r(i, 1).c0 * p[1].c1 + fj.r(i, 1).c1 * p[1].c0))
                      + ((fj.r(i, 2).c0 * p[2].c1 + fj.r(i, 2).c1 * p[2].c0) + fj.cs[i].c1);
                if (ND >= 3) {
                    x2[i] = (((fj.r(i, 0).c0 * p[0].c2 + fj.r(i, 0).c2 * p[0].c0) + 2.0 * (fj.r(i, 0).c1 * p[0].c1))
                           + ((fj.r(i, 1).c0 * p[1].c2 + fj.r(i, 1).c2 * p[1].c0) + 2.0 * (fj.r(i, 1).c1 * p[1].c1)))
                          + (((fj.r(i, 2).c0 * p[2].c2 + fj.r(i, 2).c2 * p[2].c0) + 2.0 * (fj.r(i, 2).c1 * p[2].c1)) + fj.cs[i].c2);
                } else x2[i] = Make<S>::c(0.);   // ∂2(x) is zero when the solver gives no acceleration (ElementAPI.jl:48)
            }
            Vec3<S> fe = beam_fe(m, f.r, x1, x2);
            const double dL = c.w * L;
            Vec3<S> xb;
            for (int i = 0; i < 3; ++i) { if (udof) fe[i] = fe[i] - U0[i]; xb[i] = dL * fe[i]; }
            beam_gp_reverse<N>(c, L, f, xb, acc);
        }
        beam_internal_reverse<N>(L, m, f, acc);
        // roll inertia: mₑ = rₛₘ[:,1]·ι₁·vᵢ₂[1], vᵢ₂ = spin⁻¹(ṙᵀṙ + rᵀr̈)  (Rotations.jl:177-182; the symmetric ṙᵀṙ drops out of spin⁻¹)
        if (ND >= 3) {
            TR m21 = (fj.r(0, 2).c0 * fj.r(0, 1).c2 + fj.r(1, 2).c0 * fj.r(1, 1).c2) + fj.r(2, 2).c0 * fj.r(2, 1).c2;   // (rᵀr̈)[3,2]
            TR m12 = (fj.r(0, 1).c0 * fj.r(0, 2).c2 + fj.r(1, 1).c0 * fj.r(1, 2).c2) + fj.r(2, 1).c0 * fj.r(2, 2).c2;   // (rᵀr̈)[2,3]
            TR m1l = (m.iota1 * L) * ((m21 - m12) * 0.5);                     // Σ_gp dL = L
            for (int i = 0; i < 3; ++i) vsmb[i] = widen<S>(f.r(i, 0) * m1l);
        }
    }
    // internal axial force (BeamElement.jl:59): Σ_gp dL·fᵢ = EA·L·ε
    S epsb = (m.EA * L) * f.eps;
    beam_reverse_rot<N, SC>(g, f, epsb, vsmb, acc, R, sc);
}

// ------------------------------------------------------------------------------------------------ active / passive formulation, any number type
// (see "statics: active / passive node" below for the identities).  A lane whose seed directions all belong to ONE node — its ACTIVE node a — carries
// the other node's rotation as plain numbers of type TP (no partial slots), so every product with them costs a direction less; the same holds one level
// up for time-jets.  TA: number type of the active rotation (SD<1,0>, Jet<SD<1,0>>, or plain), TP: of the passive one, TU: of the translations.
struct RodPacks { double Sth[4], Sh[4]; };
// I + a·S + b·S² from the rotation vector and the two coefficients (the tail of rodrigues)
template <class T> MB_HD Mat3<T> rod_matrix(const Vec3<T>& v, const T& a, const T& b) {
    T v00 = v[0] * v[0], v11 = v[1] * v[1], v22 = v[2] * v[2];
    T b01 = b * (v[0] * v[1]), b02 = b * (v[0] * v[2]), b12 = b * (v[1] * v[2]);
    T a0 = a * v[0], a1 = a * v[1], a2 = a * v[2];
    Mat3<T> r;
    r(0, 0) = 1.0 - b * (v11 + v22); r(1, 1) = 1.0 - b * (v00 + v22); r(2, 2) = 1.0 - b * (v00 + v11);
    r(1, 0) = b01 + a2; r(0, 1) = b01 - a2;
    r(2, 0) = b02 - a1; r(0, 2) = b02 + a1;
    r(2, 1) = b12 + a0; r(1, 2) = b12 - a0;
    return r;
}
// Rematerialisation fence: the value is unchanged, but the compiler may no longer assume so — what is recomputed from it is not merged with the first
// evaluation and kept in registers from the forward sweep to the end of the reverse sweep (the kernel is bound by its register live set, not by flops).
MB_HD void opaque(double& x) {
#ifdef __CUDA_ARCH__
    asm volatile("" : "+d"(x));
#else
    (void)x;
#endif
}
template <bool A, bool B> MB_HD void opaque(SD<A, B>& x) { opaque(x.v); if constexpr (A) opaque(x.d0); if constexpr (B) opaque(x.d1); }
#ifndef MB_AP_REMAT
#define MB_AP_REMAT 1
#endif
template <class T> MB_FN Mat3<T> rodrigues_pk(const Vec3<T>& v, RodAux<T>& aux, const T& th, bool small, const RodPacks& pk) {
    T a, b;
    aux.small = small;
#pragma unroll
    for (int k = 0; k < 4; ++k) { aux.Sth[k] = pk.Sth[k]; aux.Sh[k] = pk.Sh[k]; }
    if (small) { a = Make<T>::c(1.0); b = Make<T>::c(0.5); }
    else { a = apply_fn(th, aux.Sth); T c = apply_fn(th * 0.5, aux.Sh); b = sqr_ref(c) * 0.5; }
    aux.a = a; aux.b = b; aux.th = th;
    return rod_matrix(v, a, b);
}
template <class T> MB_HD T norm_of(const Vec3<T>& v) { return mb_sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }
MB_HD void rod_packs(double th, RodPacks& pk) { sinc_pack(th, pk.Sth); sinc_pack(0.5 * th, pk.Sh); }

struct PacksLocal { MB_HD void operator()(int, double th, RodPacks& pk) const { rod_packs(th, pk); } };
template <class TA, class TP, class TU> struct BeamFwdAP {
    using TS = decltype(TA() * TU());
    Vec3<TA> va, h, dvp, dv, vl, vsm;        // dvp = Δv′ = ½Rodrigues⁻¹(rₐrₚᵀ), dv = Δvᵧ = σΔv′, vl = rₛₘᵀΔvᵧ
    Vec3<TP> vp;
    RodAux<TA> aa, ad; RodAux<TP> ap; RinvAux<TA> ai, ir;
    Mat3<TA> ra, rd, r; Mat3<TP> rp, B;      // B = rₚ·rₘ, r = rₛₘ = Rodrigues(Δv′)·B
    Vec3<TU> dp, cs;
    Vec3<TS> ul, q; TS eps, qn;
};
template <bool VSM, class TA, class TP, class TU, class EX>
MB_HD void beam_forward_ap(const BeamGeo& g, const Vec3<TU>& u1, const Vec3<TU>& u2, const Vec3<TA>& va, const Vec3<TP>& vp, double sigma,
                           BeamFwdAP<TA, TP, TU>& f, const EX& ex) {
    f.va = va; f.vp = vp;
    RodPacks pk;
    TA tha = norm_of(va); TP thp = norm_of(vp);
    ex(0, value(tha), pk);
    f.ra = rodrigues_pk(f.va, f.aa, tha, value(tha) < 1e-14, pk);
    ex(1, value(thp), pk);
    f.rp = rodrigues_pk(f.vp, f.ap, thp, value(thp) < 1e-14, pk);
    Mat3<TA> Nm = mul_nt(f.ra, f.rp);
    f.h = rodrigues_inv(Nm, f.ai);
#pragma unroll
    for (int i = 0; i < 3; ++i) { f.dvp[i] = 0.5 * f.h[i]; f.dv[i] = sigma * f.dvp[i]; }
    TA thd = norm_of(f.dvp);
    ex(2, value(thd), pk);
    f.rd = rodrigues_pk(f.dvp, f.ad, thd, value(thd) < 1e-14, pk);
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) f.B(i, j) = (f.rp(i, 0) * g.rm(0, j) + f.rp(i, 1) * g.rm(1, j)) + f.rp(i, 2) * g.rm(2, j);
    f.r = mul(f.rd, f.B);
    if constexpr (VSM) f.vsm = rodrigues_inv(f.r, f.ir);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        TU c2 = 0.5 * (u1[i] + u2[i]);
        f.dp[i] = (u2[i] + g.tgm[i] * 0.5) - c2;
        f.cs[i] = c2 + g.cm[i];
    }
    f.q = mulv_t(f.r, f.dp);
    f.ul = f.q; f.ul[0] = f.ul[0] - g.L * 0.5;
    f.vl = mulv_t(f.r, f.dv);
    f.qn = mb_sqrt((f.q[0] * f.q[0] + f.q[1] * f.q[1]) + f.q[2] * f.q[2]);
    f.eps = f.qn * (2.0 / g.L) - 1.0;
}
// Everything upstream of the Gauss loop, reverse: ε, uₗ/vₗ, vₛₘ, r = rd·B, Δv′, N = rₐrₚᵀ and the three Rodrigues maps → rows of u₁, u₂, vₐ, vₚ
template <bool VSM, class TA, class TP, class TU, class S>
MB_HD void beam_reverse_ap(const BeamGeo& g, const BeamFwdAP<TA, TP, TU>& f, const S& epsb, const Vec3<S>& vsmb, BeamAcc<S>& a, double sigma,
                           S* Ru1, S* Ru2, S* Rva, S* Rvp) {
    const double L = g.L;
    Mat3<S>& rb = a.rb; Vec3<S>&ulb = a.ulb, &vlb = a.vlb, &cb = a.cb;
    {
        S k = (epsb * (2.0 / L)) / f.qn;
#pragma unroll
        for (int i = 0; i < 3; ++i) ulb[i] = ulb[i] + k * f.q[i];
    }
    Vec3<S> dvb = mulv(f.r, vlb);
    {
        Vec3<S> dpb = mulv(f.r, ulb);
#pragma unroll
        for (int i = 0; i < 3; ++i) { S hcs = 0.5 * (cb[i] - dpb[i]); Ru1[i] = hcs; Ru2[i] = hcs + dpb[i]; }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) rb(i, j) = rb(i, j) + (f.dp[i] * ulb[j] + f.dv[i] * vlb[j]);
    if constexpr (VSM) rodrigues_inv_adj(f.vsm, f.ir, vsmb, rb);
    Mat3<S> rdb = mul_nt(rb, f.B);
    Mat3<S> Bb = mul_tn(f.rd, rb);
    Mat3<S> rpb;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) rpb(i, j) = (Bb(i, 0) * g.rm(j, 0) + Bb(i, 1) * g.rm(j, 1)) + Bb(i, 2) * g.rm(j, 2);
    Vec3<S> dvpb{sigma * dvb[0], sigma * dvb[1], sigma * dvb[2]};
    rodrigues_adj(f.dvp, f.ad, rdb, dvpb);
    Mat3<S> Nb; { S z = Make<S>::c(0.); for (int i = 0; i < 9; ++i) Nb.a[i] = z; }
    Vec3<S> hb{0.5 * dvpb[0], 0.5 * dvpb[1], 0.5 * dvpb[2]};
    rodrigues_inv_adj(f.h, f.ai, hb, Nb);
    Mat3<S> rab = mul(Nb, f.rp);
    Mat3<S> rpb2 = mul_tn(Nb, f.ra);
    for (int i = 0; i < 9; ++i) rpb.a[i] = rpb.a[i] + rpb2.a[i];
    S z = Make<S>::c(0.);
    Vec3<S> vab{z, z, z}, vpb{z, z, z};
    rodrigues_adj(f.va, f.aa, rab, vab);
    rodrigues_adj(f.vp, f.ap, rpb, vpb);
#pragma unroll
    for (int i = 0; i < 3; ++i) { Rva[i] = vab[i]; Rvp[i] = vpb[i]; }
}
// Gauss-loop accumulators for the AP forward state (same arithmetic as beam_internal_reverse / beam_gp_reverse on BeamFwd)
template <class F, class S> MB_HD void beam_internal_reverse_f(double L, const BeamMat& m, const F& f, BeamAcc<S>& a) {
    const double c48 = 48.0 / (L * L * L), c4 = 4.0 / L;
    a.ulb[1] = a.ulb[1] + (m.EI3 * c48) * f.ul[1];
    a.ulb[2] = a.ulb[2] + (m.EI2 * c48) * f.ul[2];
    a.vlb[0] = a.vlb[0] + (m.GJ * c4) * f.vl[0];
    a.vlb[1] = a.vlb[1] + (m.EI2 * c4) * f.vl[1];
    a.vlb[2] = a.vlb[2] + (m.EI3 * c4) * f.vl[2];
}
template <class F, class S> MB_HD void beam_gp_reverse_f(const GpConst& c, double L, const F& f, const Vec3<S>& xb, BeamAcc<S>& a) {
    const double yv = c.yv * L;
    auto p = beam_gp_local(c, L, f.ul, f.vl);
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) a.rb(i, j) = a.rb(i, j) + xb[i] * p[j];
    Vec3<S> pb = mulv_t(f.r, xb);
#pragma unroll
    for (int i = 0; i < 3; ++i) a.cb[i] = a.cb[i] + xb[i];
    a.ulb[0] = a.ulb[0] + c.ya * pb[0];
    a.ulb[1] = a.ulb[1] + c.yu * pb[1];
    a.ulb[2] = a.ulb[2] + c.yu * pb[2];
    a.vlb[1] = a.vlb[1] - yv * pb[2];
    a.vlb[2] = a.vlb[2] + yv * pb[1];
}
// strips the partial slots: the passive node's numbers
MB_HD SD<false, false> strip(const SD<false, false>& x) { return x; }
template <bool A, bool B> MB_HD SD<false, false> strip(const SD<A, B>& x) { SD<false, false> r; r.v = x.v; return r; }
template <MB_J3(T)> MB_HD auto strip(const MB_JT(T)& x) -> Jet<decltype(strip(x.c0))> { Jet<decltype(strip(x.c0))> r; r.c0 = strip(x.c0); r.c1 = strip(x.c1); r.c2 = strip(x.c2); return r; }
// MB_DYN_AP: the Newmark / DirectXUA lanes (NumSD: rotation dof l and translation dof l, both of node ⌊l/3⌋+1) run the active/passive formulation
#ifndef MB_DYN_AP
#define MB_DYN_AP 1
#endif
template <class N> struct ApOk { static constexpr bool value = false; };
template <> struct ApOk<NumSD> { static constexpr bool value = MB_DYN_AP != 0; };

// ------------------------------------------------------------------------------------------------ two-phase variant (ND ≥ 2)
// The fused Newmark kernel is ~210 KB of straight-line code and runs out of the instruction cache; split in two, each half fits.
// Phase A: time-jets forward only → external-load cotangents x̄_gp = dL·fₑ (BeamElement.jl:28-58,169) and v̄ₛₘ = Σ dL·mₑ (:56-57).
// DD (SD number types, ND = 3): also the partials of the cotangents with respect to the ACCELERATIONS of the lane's two dofs, xb2/vsmb2 (value
// parts unused).  Along q(t) = X₀+X′t+X″t²/2, ∂/∂X″ⱼ f(q(t)) = (t²/2)·(∂ⱼf)(q(t)), whose second time derivative at 0 is ∂ⱼf(X₀): ∂ẍ/∂X″ⱼ = ∂x/∂X₀ⱼ,
// ∂r̈/∂X″ⱼ = ∂r/∂X₀ⱼ, ∂ẋ/∂X″ⱼ = 0 — all of it already in the order-0 partials of this lane, so DirectXUA needs no time-jet lanes seeded at X″.
// (Exact also under the reference's x^0 rule, whose missing term depends on X₀ and X′ only.)
// the Gauss loop of phase A on a forward state F (BeamFwd<NumJet<N>> or BeamFwdAP of jets): fj.r, fj.ul, fj.vl, fj.cs
template <int ND, class N, bool DD, class F> MB_HD void beam_dyn_cot_tail(const BeamGeo& g, const BeamMat& m, const F& fj, bool udof, const typename N::TU* U0,
                                                                          Vec3<typename N::TS>* xb, Vec3<typename N::TS>& vsmb,
                                                                          Vec3<typename N::TS>* xb2, Vec3<typename N::TS>* vsmb2) {
    using TR = typename N::TR; using S = typename N::TS;
    using R0 = decltype(fj.r.a[0].c0);                    // TR, or a plain value for the velocity-seeded lanes (NumJet1)
    const double L = g.L;
    Mat3<R0> r0; for (int i = 0; i < 9; ++i) r0.a[i] = fj.r.a[i].c0;
    MB_PRAGMA(unroll MB_GP_UNROLL_DYN)
    for (int gp = 0; gp < NGP; ++gp) {
        const GpConst c = gp_const(gp);
        auto p = beam_gp_local(c, L, fj.ul, fj.vl);
        Vec3<S> x1, x2;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            x1[i] = ((fj.r(i, 0).c0 * p[0].c1 + fj.r(i, 0).c1 * p[0].c0) + (fj.r(i, 1).c0 * p[1].c1 + fj.r(i, 1).c1 * p[1].c0))
                  + ((fj.r(i, 2).c0 * p[2].c1 + fj.r(i, 2).c1 * p[2].c0) + fj.cs[i].c1);
            if (ND >= 3) {
                x2[i] = (((fj.r(i, 0).c0 * p[0].c2 + fj.r(i, 0).c2 * p[0].c0) + 2.0 * (fj.r(i, 0).c1 * p[0].c1))
                       + ((fj.r(i, 1).c0 * p[1].c2 + fj.r(i, 1).c2 * p[1].c0) + 2.0 * (fj.r(i, 1).c1 * p[1].c1)))
                      + (((fj.r(i, 2).c0 * p[2].c2 + fj.r(i, 2).c2 * p[2].c0) + 2.0 * (fj.r(i, 2).c1 * p[2].c1)) + fj.cs[i].c2);
            } else x2[i] = Make<S>::c(0.);
        }
        Vec3<S> fe = beam_fe(m, r0, x1, x2);
        const double dL = c.w * L;
        for (int i = 0; i < 3; ++i) { if (udof) fe[i] = fe[i] - U0[i]; xb[gp][i] = dL * fe[i]; }
        if constexpr (DD) {
            // fₑ is linear in ẍ: ∂fₑ/∂X″ = μ·∂x + r₀·(Cₐ∘(r₀ᵀ∂x)) with ∂x = ∂x/∂X₀ (BeamElement.jl:37-45), r₀ taken as plain values
            using V = SD<false, false>;
            Vec3<S> dx;
            for (int i = 0; i < 3; ++i) {
                dx[i] = ((fj.r(i, 0).c0 * p[0].c0 + fj.r(i, 1).c0 * p[1].c0) + fj.r(i, 2).c0 * p[2].c0) + fj.cs[i].c0;
                dx[i].v = 0.;
            }
            Mat3<V> rv; for (int i = 0; i < 9; ++i) rv.a[i].v = r0.a[i].v;
            Vec3<S> dl = mulv_t(rv, dx);
            const double Ca[3] = {m.Ca1, m.Ca2, m.Ca3};
            for (int i = 0; i < 3; ++i) dl[i] = Ca[i] * dl[i];
            Vec3<S> dg = mulv(rv, dl);
            for (int i = 0; i < 3; ++i) xb2[gp][i] = dL * (m.mu * dx[i] + dg[i]);
        }
    }
    vsmb = Vec3<S>{Make<S>::c(0.), Make<S>::c(0.), Make<S>::c(0.)};
    if (ND >= 3) {
        TR m21 = (fj.r(0, 2).c0 * fj.r(0, 1).c2 + fj.r(1, 2).c0 * fj.r(1, 1).c2) + fj.r(2, 2).c0 * fj.r(2, 1).c2;
        TR m12 = (fj.r(0, 1).c0 * fj.r(0, 2).c2 + fj.r(1, 1).c0 * fj.r(1, 2).c2) + fj.r(2, 1).c0 * fj.r(2, 2).c2;
        TR m1l = (m.iota1 * L) * ((m21 - m12) * 0.5);
        for (int i = 0; i < 3; ++i) vsmb[i] = widen<S>(r0(i, 0) * m1l);
        if constexpr (DD) {
            using V = SD<false, false>;
            auto val = [](const TR& x) { V r; r.v = x.v; return r; };
            TR a21 = (val(fj.r(0, 2).c0) * fj.r(0, 1).c0 + val(fj.r(1, 2).c0) * fj.r(1, 1).c0) + val(fj.r(2, 2).c0) * fj.r(2, 1).c0;
            TR a12 = (val(fj.r(0, 1).c0) * fj.r(0, 2).c0 + val(fj.r(1, 1).c0) * fj.r(1, 2).c0) + val(fj.r(2, 1).c0) * fj.r(2, 2).c0;
            TR d1l = (m.iota1 * L) * ((a21 - a12) * 0.5);
            for (int i = 0; i < 3; ++i) { (*vsmb2)[i] = widen<S>(val(r0(i, 0)) * d1l); (*vsmb2)[i].v = 0.; }
        }
    }
}
// n1: the lane's seeds belong to node 1 (lanes 0-2) or node 2 (lanes 3-5) — only read by the active/passive formulation (ApOk<N>)
// V1: the lane's seeds sit on X′ only (Xu[0], Xv[0], Xu[2], Xv[2] carry zero partials): the jets run in NumJet1<N>
template <int ND, class N, bool DD = false, bool V1 = false> MB_HD void beam_dyn_cotangents(const BeamGeo& g, const BeamMat& m, const typename N::TU (*Xu)[6], const typename N::TR (*Xv)[6],
                                                         bool udof, const typename N::TU* U0, Vec3<typename N::TS>* xb, Vec3<typename N::TS>& vsmb,
                                                         Vec3<typename N::TS>* xb2 = nullptr, Vec3<typename N::TS>* vsmb2 = nullptr, bool n1 = true) {
    using TR = typename N::TR; using TU = typename N::TU;
    static_assert(!V1 || (ND >= 3 && !DD), "velocity-seeded jets: DirectXUA{2,…} lanes of ∂/∂X′");
    using NJ = std::conditional_t<V1, NumJet1<N>, NumJet<N>>;
    using JR = typename NJ::TR; using JU = typename NJ::TU;
    JU XuJ[6]; JR XvJ[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        if constexpr (V1) { XuJ[i].c0 = strip(Xu[0][i]); XvJ[i].c0 = strip(Xv[0][i]); }
        else { XuJ[i].c0 = Xu[0][i]; XvJ[i].c0 = Xv[0][i]; }
        XuJ[i].c1 = Xu[1][i]; XuJ[i].c2 = (ND >= 3) ? Xu[2][i] : Make<TU>::c(0.);
        XvJ[i].c1 = Xv[1][i]; XvJ[i].c2 = (ND >= 3) ? Xv[2][i] : Make<TR>::c(0.);
    }
    // ND = 3 stays on the symmetric form: under the reference's x^0 rule (sqr_ref, dual.cuh) the SECOND time derivative of Rodrigues is not the derivative of a
    // rotation, so Rodrigues(Δvᵧ)·rₛ₁ = Rodrigues(Δvᵧ)ᵀ·rₛ₂ — which the active/passive form rests on for the lanes of node 1 — holds for values and velocities
    // only; the accelerations must be formed exactly as the reference forms them (BeamElement.jl:195-199)
    if constexpr (ApOk<N>::value && ND < 3) {
        using JP = decltype(strip(XvJ[0]));
        Vec3<JR> va; Vec3<JP> vp;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const JP p1 = strip(XvJ[i]), p2 = strip(XvJ[3 + i]);
            va[i].c0 = n1 ? XvJ[i].c0 : XvJ[3 + i].c0; va[i].c1 = n1 ? XvJ[i].c1 : XvJ[3 + i].c1; va[i].c2 = n1 ? XvJ[i].c2 : XvJ[3 + i].c2;
            vp[i].c0 = n1 ? p2.c0 : p1.c0; vp[i].c1 = n1 ? p2.c1 : p1.c1; vp[i].c2 = n1 ? p2.c2 : p1.c2;
        }
        BeamFwdAP<JR, JP, JU> fj;
        beam_forward_ap<false>(g, Vec3<JU>{XuJ[0], XuJ[1], XuJ[2]}, Vec3<JU>{XuJ[3], XuJ[4], XuJ[5]}, va, vp, n1 ? -1.0 : 1.0, fj, PacksLocal());
        beam_dyn_cot_tail<ND, N, DD>(g, m, fj, udof, U0, xb, vsmb, xb2, vsmb2);
    } else {
        BeamFwd<NJ> fj;
        beam_forward<NJ, false>(g, Vec3<JU>{XuJ[0], XuJ[1], XuJ[2]}, Vec3<JR>{XvJ[0], XvJ[1], XvJ[2]}, Vec3<JU>{XuJ[3], XuJ[4], XuJ[5]},
                                Vec3<JR>{XvJ[3], XvJ[4], XvJ[5]}, fj);
        beam_dyn_cot_tail<ND, N, DD>(g, m, fj, udof, U0, xb, vsmb, xb2, vsmb2);
    }
}
// getresult (src/Output.jl:131-181) for EulerBeam3D, values only: the ☼/♢ requestables of residual (BeamElement.jl:151-174) and resultants (:28-64).
// out[77]: ε, rₛₘ (column-major 9), ♢κ (3), then per Gauss point x(3), κgp(3), fᵢ, mᵢ(3), fₑ(3) (before the −U of :169), mₑ(3).
constexpr int MB_NRES = 13 + 16 * NGP;
template <int ND> MB_HD void beam_results(const BeamGeo& g, const BeamMat& m, const double (*Xu)[6], const double (*Xv)[6], double* out) {
    using V = SD<false, false>; using NJ = NumJet<NumVal>; using J = Jet<V>;
    const double L = g.L;
    J XuJ[6], XvJ[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        XuJ[i].c0.v = Xu[0][i]; XuJ[i].c1.v = (ND >= 2) ? Xu[1][i] : 0.; XuJ[i].c2.v = (ND >= 3) ? Xu[2][i] : 0.;
        XvJ[i].c0.v = Xv[0][i]; XvJ[i].c1.v = (ND >= 2) ? Xv[1][i] : 0.; XvJ[i].c2.v = (ND >= 3) ? Xv[2][i] : 0.;
    }
    BeamFwd<NJ> fj;
    beam_forward<NJ, false>(g, Vec3<J>{XuJ[0], XuJ[1], XuJ[2]}, Vec3<J>{XvJ[0], XvJ[1], XvJ[2]}, Vec3<J>{XuJ[3], XuJ[4], XuJ[5]}, Vec3<J>{XvJ[3], XvJ[4], XvJ[5]}, fj);
    Mat3<V> r0; for (int i = 0; i < 9; ++i) r0.a[i] = fj.r.a[i].c0;
    out[0] = fj.eps.c0.v;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out[1 + i + 3 * j] = r0(i, j).v;
    out[10] = fj.vl[0].c0.v * (2 / L); out[11] = fj.vl[2].c0.v * (2 / L); out[12] = -fj.vl[1].c0.v * (2 / L);
    // roll inertia moment (Rotations.jl:177-182 → BeamElement.jl:55-57)
    V m1l = Make<V>::c(0.);
    if (ND >= 3) {
        V m21 = (fj.r(0, 2).c0 * fj.r(0, 1).c2 + fj.r(1, 2).c0 * fj.r(1, 1).c2) + fj.r(2, 2).c0 * fj.r(2, 1).c2;
        V m12 = (fj.r(0, 1).c0 * fj.r(0, 2).c2 + fj.r(1, 1).c0 * fj.r(1, 2).c2) + fj.r(2, 1).c0 * fj.r(2, 2).c2;
        m1l = m.iota1 * ((m21 - m12) * 0.5);
    }
    for (int gp = 0; gp < NGP; ++gp) {
        const GpConst c = gp_const(gp);
        double* q = out + 13 + 16 * gp;
        Vec3<J> p = beam_gp_local(c, L, fj.ul, fj.vl);
        const double ka = 2.0 / L, ku = c.ku / (L * L), kv = 2.0 / L;
        const double kap[3] = {ka * fj.vl[0].c0.v, ku * fj.ul[1].c0.v + kv * fj.vl[2].c0.v, ku * fj.ul[2].c0.v - kv * fj.vl[1].c0.v};
        Vec3<V> x1, x2;
        for (int i = 0; i < 3; ++i) {
            V x0 = ((fj.r(i, 0).c0 * p[0].c0 + fj.r(i, 1).c0 * p[1].c0) + fj.r(i, 2).c0 * p[2].c0) + fj.cs[i].c0;
            q[i] = x0.v;
            x1[i] = ((fj.r(i, 0).c0 * p[0].c1 + fj.r(i, 0).c1 * p[0].c0) + (fj.r(i, 1).c0 * p[1].c1 + fj.r(i, 1).c1 * p[1].c0))
                  + ((fj.r(i, 2).c0 * p[2].c1 + fj.r(i, 2).c1 * p[2].c0) + fj.cs[i].c1);
            if (ND >= 3) {
                x2[i] = (((fj.r(i, 0).c0 * p[0].c2 + fj.r(i, 0).c2 * p[0].c0) + 2.0 * (fj.r(i, 0).c1 * p[0].c1))
                       + ((fj.r(i, 1).c0 * p[1].c2 + fj.r(i, 1).c2 * p[1].c0) + 2.0 * (fj.r(i, 1).c1 * p[1].c1)))
                      + (((fj.r(i, 2).c0 * p[2].c2 + fj.r(i, 2).c2 * p[2].c0) + 2.0 * (fj.r(i, 2).c1 * p[2].c1)) + fj.cs[i].c2);
            } else x2[i] = Make<V>::c(0.);
            q[3 + i] = kap[i];
        }
        Vec3<V> fe = beam_fe(m, r0, x1, x2);
        q[6] = m.EA * fj.eps.c0.v;
        q[7] = m.GJ * kap[0]; q[8] = m.EI3 * kap[1]; q[9] = m.EI2 * kap[2];
        for (int i = 0; i < 3; ++i) { q[10 + i] = fe[i].v; q[13 + i] = (r0(i, 0) * m1l).v; }
    }
}
// Phase B: order-0 forward + reverse sweep with the cotangents of phase A
// (S ≠ N::TS: forward in plain values, cotangents carrying partials — the linear lanes ∂R/∂X′, ∂R/∂X″, ∂R/∂U of DirectXUA)
template <class N, class S = typename N::TS, bool VSM = true> MB_HD void beam_residual_cot(const BeamGeo& g, const BeamMat& m, const typename N::TU* Xu0,
                                                                                           const typename N::TR* Xv0, const Vec3<S>* xb, const Vec3<S>& vsmb, S* R,
                                                                                           bool n1 = true) {
    using TR = typename N::TR; using TU = typename N::TU;
    const double L = g.L;
    BeamAcc<S> acc;
    {
        S z = Make<S>::c(0.);
        for (int i = 0; i < 9; ++i) acc.rb.a[i] = z;
        for (int i = 0; i < 3; ++i) { acc.ulb[i] = z; acc.vlb[i] = z; acc.cb[i] = z; }
    }
    // measured on B200 (4 M elements): with accelerations (VSM: the v̄ₛₘ branch adds Rodrigues⁻¹(rₛₘ) to the live set) the active/passive phase B spills more than it
    // saves (SweepX{2} 26.2 vs 25.3 ms), without it wins (SweepX{1} 16.3 vs 17.2 ms) — so only the first-order kernels use it
    if constexpr (ApOk<N>::value && !VSM) {
        using TP = decltype(strip(Xv0[0]));
        Vec3<TR> va; Vec3<TP> vp;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const TP p1 = strip(Xv0[i]), p2 = strip(Xv0[3 + i]);
            va[i] = n1 ? Xv0[i] : Xv0[3 + i];
            vp[i] = n1 ? p2 : p1;
        }
        const double sigma = n1 ? -1.0 : 1.0;
        BeamFwdAP<TR, TP, TU> f;
        beam_forward_ap<VSM>(g, Vec3<TU>{Xu0[0], Xu0[1], Xu0[2]}, Vec3<TU>{Xu0[3], Xu0[4], Xu0[5]}, va, vp, sigma, f, PacksLocal());
        MB_PRAGMA(unroll MB_GP_UNROLL_STATIC)
        for (int gp = 0; gp < NGP; ++gp) beam_gp_reverse_f(gp_const(gp), L, f, xb[gp], acc);
        beam_internal_reverse_f(L, m, f, acc);
        S epsb = widen<S>((m.EA * L) * f.eps);
        S Ru1[3], Ru2[3], Rva[3], Rvp[3];
        beam_reverse_ap<VSM>(g, f, epsb, vsmb, acc, sigma, Ru1, Ru2, Rva, Rvp);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            R[i] = Ru1[i]; R[6 + i] = Ru2[i];
            R[3 + i] = n1 ? Rva[i] : Rvp[i];
            R[9 + i] = n1 ? Rvp[i] : Rva[i];
        }
    } else {
        BeamFwd<N> f;
        beam_forward<N, VSM>(g, Vec3<TU>{Xu0[0], Xu0[1], Xu0[2]}, Vec3<TR>{Xv0[0], Xv0[1], Xv0[2]}, Vec3<TU>{Xu0[3], Xu0[4], Xu0[5]}, Vec3<TR>{Xv0[3], Xv0[4], Xv0[5]}, f);
        MB_PRAGMA(unroll MB_GP_UNROLL_STATIC)
        for (int gp = 0; gp < NGP; ++gp) beam_gp_reverse<N, S>(gp_const(gp), L, f, xb[gp], acc);
        beam_internal_reverse<N, S>(L, m, f, acc);
        S epsb = widen<S>((m.EA * L) * f.eps);
        HostScratch sc;
        beam_reverse_rot<N, HostScratch, S, VSM>(g, f, epsb, vsmb, acc, R, sc);
    }
}

template <int ND, class N> MB_HD void beam_residual_n(const BeamGeo& g, const BeamMat& m, const typename N::TU (*Xu)[6], const typename N::TR (*Xv)[6],
                                                     bool udof, const typename N::TU* U0, typename N::TS* R) {
    HostScratch sc;
    beam_residual_n<ND, N, HostScratch>(g, m, Xu, Xv, udof, U0, R, sc);
}

// ------------------------------------------------------------------------------------------------ statics: symmetric tangent
// In statics the residual is the gradient of a potential (strain energy − weight·x + U·x: fₑ does not depend on X, BeamElement.jl:37,169),
// so the tangent ∂R/∂X₀ is symmetric and only the six ROTATION directions need a forward-over-reverse sweep:
//   columns of the rotation dofs      from the sweep seeded at rotation dof l (all in SD<1,0>: one direction instead of two),
//   rows of the rotation dofs         by symmetry (the translation rows of those columns, transposed),
//   translation × translation block   in closed form: translations enter only through d′ = (u₂−u₁)/2 + tgₘ/2 and only through
//       d̄′ = rₛₘ·ūₗ,  ūₗ = 2EA(2/L − 1/|q|)·q + diag(0, 48EI₃/L³, 48EI₂/L³)·uₗ,  q = rₛₘᵀd′  (ε = 2|q|/L − 1, beam_reverse_rot / beam_internal_reverse)
//       ⇒ G = ∂d̄′/∂d′ = rₛₘ·[2EA(2/L − 1/|q|)·I + 2EA·qqᵀ/|q|³ + diag(0,48EI₃/L³,48EI₂/L³)]·rₛₘᵀ and ∂R_{uₐ}/∂u_b = ±G/4.
// c ∈ {0,1,2}: the column of G this lane returns in Gc[3].
using NumRot = Num<SD<true, false>, SD<false, false>, SD<true, false>>;
MB_HD void beam_static_sym(const BeamGeo& g, const BeamMat& m, const SD<false, false>* Xu0, const SD<true, false>* Xv0, bool udof,
                           const SD<false, false>* U0, int c, SD<true, false>* R, double* Gc) {
    using N = NumRot; using TR = N::TR; using TU = N::TU; using S = N::TS;
    const double L = g.L;
    BeamFwd<N> f;
    BeamAcc<S> acc;
    {
        S z = Make<S>::c(0.);
        for (int i = 0; i < 9; ++i) acc.rb.a[i] = z;
        for (int i = 0; i < 3; ++i) { acc.ulb[i] = z; acc.vlb[i] = z; acc.cb[i] = z; }
    }
    Vec3<S> vsmb{Make<S>::c(0.), Make<S>::c(0.), Make<S>::c(0.)};
    beam_forward<N, false>(g, Vec3<TU>{Xu0[0], Xu0[1], Xu0[2]}, Vec3<TR>{Xv0[0], Xv0[1], Xv0[2]}, Vec3<TU>{Xu0[3], Xu0[4], Xu0[5]},
                           Vec3<TR>{Xv0[3], Xv0[4], Xv0[5]}, f);
    {
        const double iq = 1.0 / f.qn.v, k = (2.0 * m.EA) * (2.0 / L - iq), h = (2.0 * m.EA) * (iq * iq * iq);
        const double c48 = 48.0 / (L * L * L);
        const double D[3] = {0., m.EI3 * c48, m.EI2 * c48};
        double w[3], Aw[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) w[j] = (c == 0) ? f.r(0, j).v : ((c == 1) ? f.r(1, j).v : f.r(2, j).v);     // rₛₘᵀe_c
        const double qw = (f.q[0].v * w[0] + f.q[1].v * w[1]) + f.q[2].v * w[2];
#pragma unroll
        for (int j = 0; j < 3; ++j) Aw[j] = (k + D[j]) * w[j] + (h * qw) * f.q[j].v;
#pragma unroll
        for (int i = 0; i < 3; ++i) Gc[i] = (f.r(i, 0).v * Aw[0] + f.r(i, 1).v * Aw[1]) + f.r(i, 2).v * Aw[2];
    }
    Vec3<S> fe{Make<S>::c(0.), Make<S>::c(0.), Make<S>::c(m.w)};
    if (udof) for (int i = 0; i < 3; ++i) fe[i] = fe[i] - U0[i];
    beam_uniform_reverse<N>(L, f, fe, acc);
    beam_internal_reverse<N>(L, m, f, acc);
    S epsb = (m.EA * L) * f.eps;
    HostScratch sc;
    beam_reverse_rot<N, HostScratch, S, false>(g, f, epsb, vsmb, acc, R, sc);
}

// Where lane l (rotation dof l; element dof cv = l+3 or l+6) of beam_static_sym puts its share of the scaled element tangent
// Ke[i + 12j] = scale_i·∂R_i/∂X_j·scale_j (src/SweepX.jl:55,63): the rotation column cv, the same six translation-row entries transposed into
// row cv of the translation columns, and the translation rows of translation column cu = cv − 3 (±G/4).  R must have been seeded with scale_cv.
// put(k, value) stores entry k of the element's 144, put2(k, v0, v1) entries k (even) and k+1; returns true if any value is NaN.
template <class Put, class Put2> MB_HD bool beam_static_sym_store(int l, const SD<true, false>* R, const double* Gc, const double* scaleX, Put put, Put2 put2) {
    const int cu = (l < 3) ? l : l + 3, cv = cu + 3;
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 12; i += 2) {
        const double a = R[i].d0 * scaleX[i], b = R[i + 1].d0 * scaleX[i + 1];
        bad |= (a != a) | (b != b);
        put2(12 * cv + i, a, b);                              // even offsets: 16-byte stores
        if (i < 3 || (i >= 6 && i < 9)) put(12 * i + cv, a);
        if (i + 1 < 3 || (i + 1 >= 6 && i + 1 < 9)) put(12 * (i + 1) + cv, b);
    }
    const double sc = 0.25 * scaleX[cu];
    double v1[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double v = Gc[a] * sc;
        bad |= (v != v);
        v1[a] = (l < 3) ? v : -v;                             // rows of node 1: + when the column belongs to node 1
    }
    put2(12 * cu, v1[0] * scaleX[0], v1[1] * scaleX[1]); put(12 * cu + 2, v1[2] * scaleX[2]);
    put2(12 * cu + 6, -v1[0] * scaleX[6], -v1[1] * scaleX[7]); put(12 * cu + 8, -v1[2] * scaleX[8]);
    return bad;
}

// ------------------------------------------------------------------------------------------------ statics: active / passive node
// Lane l of beam_static_sym seeds ONE rotation dof, which belongs to one of the two nodes — the lane's ACTIVE node a; the rotation vector of the
// other, PASSIVE node p carries no partial, yet the SD<1,0> sweep above drags a (zero) partial through everything that depends on it.  The corotated
// frame is symmetric in the two nodes: with N = rₐ·rₚᵀ, Δv′ = ½Rodrigues⁻¹(N),
//     rₛₘ = Rodrigues(Δvᵧ)·rₛ₁·rₘ = Rodrigues(Δv′)·rₚ·rₘ                     (BeamElement.jl:195-199; rₛ₂ = Rodrigues(Δvᵧ)²·rₛ₁, Rodrigues(−v) = Rodrigues(v)ᵀ)
//     Δvᵧ = σ·Δv′, σ = +1 if a is node 2, −1 if a is node 1               (Rodrigues⁻¹(Mᵀ) = −Rodrigues⁻¹(M))
//     vₗ  = rₛₘᵀ·Δvᵧ = σ·(rₚ·rₘ)ᵀ·Δv′                                     (a rotation leaves its own axis alone)
// so all six lanes run the same code with their own (a,p) and everything that depends on the passive node only — Rodrigues(vₚ), B = rₚ·rₘ — is plain
// values (SD<0,0>): products with them cost one direction less, and the passive adjoints need the tangent of the cotangent only.
// Xv_a: rotation vector of the active node with the lane's seed, Xv_p: of the passive node, sigma as above.  Returns the residual rows of
// u₁, u₂ (Ru1, Ru2), of the active node's rotation (Rva) and of the passive node's (Rvp): value = R, d0 = the lane's tangent column.
// Special-function packs: PK = nullptr evaluates them here; otherwise PK[0..7] = sinc packs at θₐ, θₐ/2 and PK[8..15] at θₚ, θₚ/2 come from the caller
// (one lane of the element evaluates each of them and they are exchanged by warp shuffles, see beam_static_ap_kernel).
// EX: how the packs of the three Rodrigues maps are obtained — a functor  ex(which, θ, pk)  with which = 0 (active), 1 (passive), 2 (Δv′)
// OUT: where results go as soon as they exist (holding them to the end of the sweep costs registers the kernel does not have):
//   out.tt(Gc)          column c of G (closed-form translation × translation block), after the forward sweep
//   out.trans(Ru1,Ru2)  residual rows of u₁, u₂ (value = R, d0 = the lane's tangent column), early in the reverse sweep
//   out.rot(Rva,Rvp)    rows of the active / passive node's rotation, at the end
// S = SD<1,0> (the kernel's lanes) or SD<0,0> (values only: what tools/opcount counts as the primal share of a lane)
template <class S, class EX, class OUT>
MB_HD void beam_static_ap(const BeamGeo& g, const BeamMat& m, const SD<false, false>* Xu0, const S* Xv_a, const SD<false, false>* Xv_p,
                          double sigma, bool udof, const SD<false, false>* U0, int c, const EX& ex, OUT& out) {
    using V = SD<false, false>;
    const double L = g.L;
    // ---- forward
    Vec3<S> va{Xv_a[0], Xv_a[1], Xv_a[2]}; Vec3<V> vp{Xv_p[0], Xv_p[1], Xv_p[2]};
    RodAux<S> aa; RodAux<V> ap; RodAux<S> ad; RinvAux<S> ai;
    RodPacks pk;
    S tha = norm_of(va); V thp = norm_of(vp);
    ex(0, tha.v, pk);
    Mat3<S> ra = rodrigues_pk(va, aa, tha, tha.v < 1e-14, pk);
    ex(1, thp.v, pk);
    Mat3<V> rp = rodrigues_pk(vp, ap, thp, thp.v < 1e-14, pk);
    Mat3<S> Nm = mul_nt(ra, rp);
    Vec3<S> h = rodrigues_inv(Nm, ai);
    Vec3<S> dvp{0.5 * h[0], 0.5 * h[1], 0.5 * h[2]};
    S thd = norm_of(dvp);
    ex(2, thd.v, pk);
    Mat3<S> rd = rodrigues_pk(dvp, ad, thd, thd.v < 1e-14, pk);
    Mat3<V> gm; for (int i = 0; i < 9; ++i) gm.a[i].v = g.rm.a[i];
    Mat3<V> B = mul(rp, gm);
    Mat3<S> r = mul(rd, B);
    Vec3<S> dv{sigma * dvp[0], sigma * dvp[1], sigma * dvp[2]};
    Vec3<V> dp, cs;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        V c2 = 0.5 * (Xu0[i] + Xu0[3 + i]);
        dp[i] = (Xu0[3 + i] + g.tgm[i] * 0.5) - c2;
        cs[i] = c2 + g.cm[i];
    }
    Vec3<S> q = mulv_t(r, dp);
    Vec3<S> ul = q; ul[0] = ul[0] - L * 0.5;
    Vec3<S> vl = mulv_t(r, dv);
    S qn = mb_sqrt((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]);
    S eps = qn * (2.0 / L) - 1.0;
    // ---- translation × translation block, closed form (see beam_static_sym)
    {
        const double iq = 1.0 / qn.v, k = (2.0 * m.EA) * (2.0 / L - iq), hh = (2.0 * m.EA) * (iq * iq * iq);
        const double c48 = 48.0 / (L * L * L);
        const double D[3] = {0., m.EI3 * c48, m.EI2 * c48};
        double w[3], Aw[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) w[j] = (c == 0) ? r(0, j).v : ((c == 1) ? r(1, j).v : r(2, j).v);
        const double qw = (q[0].v * w[0] + q[1].v * w[1]) + q[2].v * w[2];
#pragma unroll
        for (int j = 0; j < 3; ++j) Aw[j] = (k + D[j]) * w[j] + (hh * qw) * q[j].v;
#pragma unroll
        double Gc[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) Gc[i] = (r(i, 0).v * Aw[0] + r(i, 1).v * Aw[1]) + r(i, 2).v * Aw[2];
        out.tt(Gc);
    }
    // ---- reverse: Gauss sums in closed form (beam_uniform_reverse, beam_internal_reverse), then the frame
    Mat3<S> rb; Vec3<S> ulb, vlb, cb;
    {
        V fe[3]; fe[0].v = 0.; fe[1].v = 0.; fe[2].v = m.w;
        if (udof) for (int i = 0; i < 3; ++i) fe[i] = fe[i] - U0[i];
        const double k6 = L * L / 6.0;
        S P1 = -k6 * vl[2], P2 = k6 * vl[1];
        S zero = Make<S>::c(0.);
#pragma unroll
        for (int i = 0; i < 3; ++i) { rb(i, 0) = zero; rb(i, 1) = fe[i] * P1; rb(i, 2) = fe[i] * P2; cb[i] = widen<S>(L * fe[i]); }
        Vec3<V> fev{fe[0], fe[1], fe[2]};
        Vec3<S> gl = mulv_t(r, fev);
        const double c48 = 48.0 / (L * L * L), c4 = 4.0 / L;
        S k = (((m.EA * L) * eps) * (2.0 / L)) / qn;
        ulb[0] = k * q[0];
        ulb[1] = (m.EI3 * c48) * ul[1] + k * q[1];
        ulb[2] = (m.EI2 * c48) * ul[2] + k * q[2];
        vlb[0] = (m.GJ * c4) * vl[0];
        vlb[1] = (m.EI2 * c4) * vl[1] + k6 * gl[2];
        vlb[2] = (m.EI3 * c4) * vl[2] - k6 * gl[1];
    }
    Vec3<S> dvb = mulv(r, vlb);
    {
        Vec3<S> dpb = mulv(r, ulb);
        S Ru1[3], Ru2[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { S hcs = 0.5 * (cb[i] - dpb[i]); Ru1[i] = hcs; Ru2[i] = hcs + dpb[i]; }      // d′ = (u₂ + tgₘ/2) − ½(u₁+u₂) ; cₛₘ = ½(u₁+u₂) + cₘ
        out.trans(Ru1, Ru2);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) rb(i, j) = rb(i, j) + (dp[i] * ulb[j] + dv[i] * vlb[j]);
    // r = rd·B, B = rₚ·rₘ
    Mat3<S> rdb = mul_nt(rb, B);
    Mat3<S> Bb = mul_tn(rd, rb);
    Mat3<S> rpb = mul_nt(Bb, gm);
    // rd = Rodrigues(Δv′), Δvᵧ = σΔv′
    Vec3<S> dvpb{sigma * dvb[0], sigma * dvb[1], sigma * dvb[2]};
    rodrigues_adj(dvp, ad, rdb, dvpb);
    // Δv′ = ½Rodrigues⁻¹(N), N = rₐ·rₚᵀ
    Mat3<S> Nb; { S z = Make<S>::c(0.); for (int i = 0; i < 9; ++i) Nb.a[i] = z; }
    Vec3<S> hb{0.5 * dvpb[0], 0.5 * dvpb[1], 0.5 * dvpb[2]};
    rodrigues_inv_adj(h, ai, hb, Nb);
#if MB_AP_REMAT
    {   // rₐ, rₚ again from (v, a, b) rather than held in registers since the forward sweep
        S a2 = aa.a, b2 = aa.b; opaque(a2); opaque(b2);
        ra = rod_matrix(va, a2, b2);
        V a3 = ap.a, b3 = ap.b; opaque(a3); opaque(b3);
        rp = rod_matrix(vp, a3, b3);
    }
#endif
    Mat3<S> rab = mul(Nb, rp);
    Mat3<S> rpb2 = mul_tn(Nb, ra);
    for (int i = 0; i < 9; ++i) rpb.a[i] = rpb.a[i] + rpb2.a[i];
    S z = Make<S>::c(0.);
    Vec3<S> vab{z, z, z}, vpb{z, z, z};
    rodrigues_adj(va, aa, rab, vab);
    rodrigues_adj(vp, ap, rpb, vpb);
    out.rot(vab.a, vpb.a);
}
// collects the outputs in element dof order, as beam_static_sym returns them (host tests; feeds beam_static_sym_store)
struct ApCollect {
    bool n1; SD<true, false>* R; double* G;
    MB_HD void tt(const double* Gc) { for (int i = 0; i < 3; ++i) G[i] = Gc[i]; }
    MB_HD void trans(const SD<true, false>* Ru1, const SD<true, false>* Ru2) { for (int i = 0; i < 3; ++i) { R[i] = Ru1[i]; R[6 + i] = Ru2[i]; } }
    MB_HD void rot(const SD<true, false>* Rva, const SD<true, false>* Rvp) { for (int i = 0; i < 3; ++i) { R[(n1 ? 3 : 9) + i] = Rva[i]; R[(n1 ? 9 : 3) + i] = Rvp[i]; } }
};

// Lane l (rotation dof l of the element: node 1 for l < 3, node 2 otherwise) of the active/passive sweep: xu = (u₁,u₂), xv = (v₁,v₂) plain values,
// seed = scale.X of the lane's dof.  Results go to `out` (see beam_static_ap).
template <class EX, class OUT>
MB_HD void beam_static_ap_lane(const BeamGeo& g, const BeamMat& m, const double* xu, const double* xv, double seed, int l, bool udof,
                               const SD<false, false>* U0, const EX& ex, OUT& out) {
    using S = SD<true, false>; using V = SD<false, false>;
    const bool n1 = l < 3;
    const int c = n1 ? l : l - 3;
    V Xu0[6]; S va[3]; V vp[3];
#pragma unroll
    for (int i = 0; i < 6; ++i) Xu0[i].v = xu[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) { va[i].v = n1 ? xv[i] : xv[3 + i]; va[i].d0 = (i == c) ? seed : 0.; vp[i].v = n1 ? xv[3 + i] : xv[i]; }
    beam_static_ap(g, m, Xu0, va, vp, n1 ? -1.0 : 1.0, udof, U0, c, ex, out);
}

// dense Dual<W> front-end (element dof order X[ider][12]); used by the δr lane of the :step mission and by the host tests
template <int ND, int W> MB_HD void beam_residual(const BeamGeo& g, const BeamMat& m, const Dual<W> (*X)[12], bool udof, const Dual<W>* U0, Dual<W>* R) {
    Dual<W> Xu[3][6], Xv[3][6];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int i = 0; i < 3; ++i) { Xu[d][i] = X[d][i]; Xv[d][i] = X[d][3 + i]; Xu[d][3 + i] = X[d][6 + i]; Xv[d][3 + i] = X[d][9 + i]; }
    beam_residual_n<ND, NumDual<W>>(g, m, Xu, Xv, udof, U0, R);
}

}  // namespace mb
