// ORACLE — TEST INFRASTRUCTURE ONLY (see adiff.hpp).
//
// Literal restatement of
//   toolbox/BeamElement.jl : BeamCrossSection (:6-25), resultants (:28-64), EulerBeam3D ctor (:121-148),
//                            residual (:151-174), kinematics{Mode} (:176-189), corotated{Mode} (:192-208)
//   src/Taylor.jl          : motion{P} (:24-29), motion⁻¹{P,ND} (:46-65), revariate{P} (:143-151),
//                            McLaurin / McLaurin_right (:225-243), chainrule (:271-274),
//                            apply{:chainrule|:direct} (:293-295), composeJacobian{P} (:319-323)
//   toolbox/Rotations.jl   : intrinsicrotationrates (:177-182)
// The nested-dual algorithm is reproduced as is, including the three-level chain rule 3→6→12 variables.
#pragma once
#include "rotations.hpp"

namespace orc {

struct BeamCrossSection {   // BeamElement.jl:6-23, same field order
    double EA, EI2, EI3, GJ, mu, iota1, w, Ca1, Cl1, Cq1, Ca2, Cl2, Cq2, Ca3, Cl3, Cq3;
};
constexpr int ngp = 4;
struct EulerBeam3D {        // BeamElement.jl:87-103
    double cm[3];
    M33<double> rm;
    double zgp[ngp], znod[2], tgm[3], tge[3];
    double ya[ngp], yu[ngp], yv[ngp], ka[ngp], ku[ngp], kv[ngp];
    double L, dL[ngp];
    BeamCrossSection mat;
};

// EulerBeam3D{Udof}(nod; mat, orient2)   BeamElement.jl:121-148.  Returns 0, or 1 for the "nearly parallel" muscadeerror.
inline int beam_ctor(EulerBeam3D& o, const double* c1, const double* c2, const double* orient2_in, const BeamCrossSection& mat) {
    double tgm[3], t[3], o2[3], n[3], b[3];
    for (int i = 0; i < 3; ++i) o.cm[i] = (c1[i] + c2[i]) / 2;
    for (int i = 0; i < 3; ++i) tgm[i] = c2[i] - c1[i];
    double L = std::sqrt(tgm[0] * tgm[0] + tgm[1] * tgm[1] + tgm[2] * tgm[2]);     // norm
    for (int i = 0; i < 3; ++i) t[i] = tgm[i] / L;
    double no = std::sqrt(orient2_in[0] * orient2_in[0] + orient2_in[1] * orient2_in[1] + orient2_in[2] * orient2_in[2]);
    for (int i = 0; i < 3; ++i) o2[i] = orient2_in[i] / no;
    double d = o2[0] * t[0] + o2[1] * t[1] + o2[2] * t[2];
    for (int i = 0; i < 3; ++i) n[i] = o2[i] - t[i] * d;
    double nn = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    if (!(nn > 1e-3)) return 1;
    for (int i = 0; i < 3; ++i) n[i] /= nn;
    b[0] = t[1] * n[2] - t[2] * n[1];
    b[1] = t[2] * n[0] - t[0] * n[2];
    b[2] = t[0] * n[1] - t[1] * n[0];
    for (int i = 0; i < 3; ++i) { o.rm.m[i] = t[i]; o.rm.m[3 + i] = n[i]; o.rm.m[6 + i] = b[i]; }   // SMatrix(t...,n...,b...): columns
    for (int i = 0; i < 3; ++i) { o.tgm[i] = tgm[i]; o.tge[i] = 0.; }
    o.tge[0] = L;
    const double s30 = std::sqrt(30.);
    o.dL[0] = L / 2 * (18 - s30) / 36; o.dL[1] = L / 2 * (18 + s30) / 36; o.dL[2] = L / 2 * (18 + s30) / 36; o.dL[3] = L / 2 * (18 - s30) / 36;
    const double s65 = std::sqrt(6. / 5);
    o.zgp[0] = -1. / 2 * std::sqrt(3. / 7 + 2. / 7 * s65); o.zgp[1] = -1. / 2 * std::sqrt(3. / 7 - 2. / 7 * s65);
    o.zgp[2] = +1. / 2 * std::sqrt(3. / 7 - 2. / 7 * s65); o.zgp[3] = +1. / 2 * std::sqrt(3. / 7 + 2. / 7 * s65);
    o.znod[0] = -1. / 2; o.znod[1] = 1. / 2;
    for (int g = 0; g < ngp; ++g) {
        double z = o.zgp[g];
        o.ya[g] = 2 * z;                                  // yₐ(ζ)
        o.yu[g] = -4 * (z * z * z) + 3 * z;               // yᵤ(ζ)
        o.yv[g] = ((z * z) - 1. / 4) * L;                 // yᵥ(ζ)*L
        o.ka[g] = 2. / L;                                 // κₐ(ζ)/L
        o.ku[g] = (-24 * z) / (L * L);                    // κᵤ(ζ)/L²
        o.kv[g] = 2. / L;                                 // κᵥ(ζ)/L
    }
    o.L = L;
    o.mat = mat;
    return 0;
}

// ---------------------------------------------------------------- Taylor.jl: McLaurin
// McLaurin_right(y::∂ℝ{1,N,𝕣},Δx)  (Taylor.jl:230-242 with P==1; "/1" is exact)
template <int N, class T> inline T mclaurin_right1(const D<1, N, double>& y, const T* dx) {
    T s = y.dx[0] * dx[0];
    for (int i = 1; i < N; ++i) s = s + y.dx[i] * dx[i];
    return s;
}
// McLaurin(y::∂ℝ{1,N,𝕣},Δx) = y.x .+ McLaurin_right(y,Δx)   (Taylor.jl:228-229)
template <int N, class T> inline T mclaurin(const D<1, N, double>& y, const T* dx) { return y.x + mclaurin_right1(y, dx); }
// McLaurin(y::∂ℝ{2,N,∂ℝ{1,N,𝕣}},Δx) = McLaurin(y.x,Δx) .+ McLaurin_right(y,Δx), the latter divided by P=2
template <int N, class T> inline T mclaurin(const D<2, N, D<1, N, double>>& y, const T* dx) {
    T a = mclaurin(y.x, dx);
    T s = mclaurin_right1(y.dx[0], dx) * dx[0];
    for (int i = 1; i < N; ++i) s = s + mclaurin_right1(y.dx[i], dx) * dx[i];
    s = s / 2.;
    return a + s;
}
template <class TY, class T> inline V3<T> mclaurin(const V3<TY>& y, const T* dx) { V3<T> r; for (int i = 0; i < 3; ++i) r[i] = mclaurin(y[i], dx); return r; }
template <class TY, class T> inline M33<T> mclaurin(const M33<TY>& y, const T* dx) { M33<T> r; for (int i = 0; i < 9; ++i) r.m[i] = mclaurin(y.m[i], dx); return r; }

// second-order multivariate number  type_multivariate_𝕣{2,N} (Taylor.jl:96-105)
template <int N> using T2 = D<2, N, D<1, N, double>>;
// revariate{2,N,:variate}(a,i) = multivariate_𝕣{2,N}(VALUE(a),i)   (Taylor.jl:151,103)
template <int N> inline T2<N> revariate2(double a, int i) {
    T2<N> r;
    r.x.x = a; r.x.dx[i] = 1.;
    r.dx[i].x = 1.;
    return r;
}

// ---------------------------------------------------------------- corotated{Mode}  BeamElement.jl:192-208
// the closure handed to apply{Mode} (BeamElement.jl:196-203), evaluated in number type T
template <bool CHAIN, class T> struct RodriguesApply;
template <class T> struct RodriguesApply<false, T> { static M33<T> run(const V3<T>& v) { return rodrigues(v); } };     // apply{:direct}
template <int N> struct RodriguesApply<true, T2<N>> {                                                                    // apply{:chainrule}
    static M33<T2<N>> run(const V3<T2<N>>& v) {
        V3<T2<3>> tv; for (int i = 0; i < 3; ++i) tv[i] = revariate2<3>(VALUE(v[i]), i);
        M33<T2<3>> ty = rodrigues(tv);
        T2<N> dx[3]; for (int i = 0; i < 3; ++i) dx[i] = v[i] - VALUE(v[i]);
        return mclaurin(ty, dx);
    }
};
template <bool CHAIN, class T>
inline void corot_inner(const EulerBeam3D& o, const V3<T>& v1, const V3<T>& v2, V3<T>& dv, M33<T>& rsm, V3<T>& vsm) {
    M33<T> rs1 = RodriguesApply<CHAIN, T>::run(v1);
    M33<T> rs2 = RodriguesApply<CHAIN, T>::run(v2);
    V3<T> h = rodrigues_inv(matmul(rs2, transpose(rs1)));
    for (int i = 0; i < 3; ++i) dv[i] = 0.5 * h[i];
    rsm = matmul(matmul(RodriguesApply<CHAIN, T>::run(dv), rs1), o.rm);
    vsm = rodrigues_inv(rsm);
}
template <class T> struct Corot { V3<T> vsm; M33<T> rsm; V3<T> ul2, vl2, csm; };

template <bool CHAIN, class T> struct CorotApply;
template <class T> struct CorotApply<false, T> {
    static void run(const EulerBeam3D& o, const V3<T>& v1, const V3<T>& v2, V3<T>& dv, M33<T>& rsm, V3<T>& vsm) { corot_inner<false, T>(o, v1, v2, dv, rsm, vsm); }
};
template <> struct CorotApply<true, T2<12>> {
    static void run(const EulerBeam3D& o, const V3<T2<12>>& v1, const V3<T2<12>>& v2, V3<T2<12>>& dv, M33<T2<12>>& rsm, V3<T2<12>>& vsm) {
        V3<T2<6>> t1, t2;
        for (int i = 0; i < 3; ++i) { t1[i] = revariate2<6>(VALUE(v1[i]), i); t2[i] = revariate2<6>(VALUE(v2[i]), 3 + i); }
        V3<T2<6>> tdv, tvsm; M33<T2<6>> trsm;
        corot_inner<true, T2<6>>(o, t1, t2, tdv, trsm, tvsm);
        T2<12> dx[6];
        for (int i = 0; i < 3; ++i) { dx[i] = v1[i] - VALUE(v1[i]); dx[3 + i] = v2[i] - VALUE(v2[i]); }
        dv = mclaurin(tdv, dx); rsm = mclaurin(trsm, dx); vsm = mclaurin(tvsm, dx);
    }
};
template <bool CHAIN, class T> inline Corot<T> corotated(const EulerBeam3D& o, const T* X0) {
    V3<T> u1{X0[0], X0[1], X0[2]}, v1{X0[3], X0[4], X0[5]}, u2{X0[6], X0[7], X0[8]}, v2{X0[9], X0[10], X0[11]};
    Corot<T> c; V3<T> dv;
    CorotApply<CHAIN, T>::run(o, v1, v2, dv, c.rsm, c.vsm);
    V3<T> cs; for (int i = 0; i < 3; ++i) cs[i] = 0.5 * (u1[i] + u2[i]);
    V3<T> a; for (int i = 0; i < 3; ++i) a[i] = (u2[i] + o.tgm[i] * o.znod[1]) - cs[i];
    M33<T> rt = transpose(c.rsm);
    V3<T> ra = matvec(rt, a);
    for (int i = 0; i < 3; ++i) c.ul2[i] = ra[i] - o.tge[i] * o.znod[1];
    c.vl2 = matvec(rt, dv);
    for (int i = 0; i < 3; ++i) c.csm[i] = cs[i] + o.cm[i];
    return c;
}

// ---------------------------------------------------------------- kinematics{Mode}  BeamElement.jl:176-189
template <class T> struct Kin { V3<T> kap[ngp], x[ngp]; T eps; V3<T> vsm; M33<T> rsm; V3<T> vl2; };
template <bool CHAIN, class T> inline Kin<T> kinematics(const EulerBeam3D& o, const T* X0) {
    Corot<T> c = corotated<CHAIN, T>(o, X0);
    Kin<T> k;
    const double L = o.L;
    k.eps = dsqrt((ipow(c.ul2[0] + L / 2, 2) + ipow(c.ul2[1], 2)) + ipow(c.ul2[2], 2)) * 2. / L - 1.;
    for (int g = 0; g < ngp; ++g) {
        double ya = o.ya[g], yu = o.yu[g], yv = o.yv[g], ka = o.ka[g], ku = o.ku[g], kv = o.kv[g];
        k.kap[g][0] = ka * c.vl2[0];
        k.kap[g][1] = ku * c.ul2[1] + kv * c.vl2[2];
        k.kap[g][2] = ku * c.ul2[2] - kv * c.vl2[1];
        V3<T> y;
        y[0] = ya * c.ul2[0];
        y[1] = yu * c.ul2[1] + yv * c.vl2[2];
        y[2] = yu * c.ul2[2] - yv * c.vl2[1];
        V3<T> a; for (int i = 0; i < 3; ++i) a[i] = o.tge[i] * o.zgp[g] + y[i];
        V3<T> ra = matvec(c.rsm, a);
        for (int i = 0; i < 3; ++i) k.x[g][i] = ra[i] + c.csm[i];
    }
    k.vsm = c.vsm; k.rsm = c.rsm; k.vl2 = c.vl2;
    return k;
}

// ---------------------------------------------------------------- motion / motion⁻¹  Taylor.jl:24-65  (P=2)
template <int ND> struct MotionT;
template <> struct MotionT<1> { using type = DV; };
template <> struct MotionT<2> { using type = D<2, 1, DV>; };
template <> struct MotionT<3> { using type = D<3, 1, D<2, 1, DV>>; };
inline DV motion(const DV* a, std::integral_constant<int, 1>) { return a[0]; }
inline D<2, 1, DV> motion(const DV* a, std::integral_constant<int, 2>) { D<2, 1, DV> r; r.x = a[0]; r.dx[0] = a[1]; return r; }
inline D<3, 1, D<2, 1, DV>> motion(const DV* a, std::integral_constant<int, 3>) {
    D<3, 1, D<2, 1, DV>> r;
    r.x.x = a[0]; r.x.dx[0] = a[1];
    r.dx[0].x = a[1]; r.dx[0].dx[0] = a[2];
    return r;
}
// motion⁻¹{P,ND,OD}: returns (value, velocity, acceleration); missing orders are zero (ElementAPI.jl:24-48 ∂1/∂2 → zilch)
inline void motion_inv(const DV& a, DV out[3]) { out[0] = a; out[1] = DV(0.); out[2] = DV(0.); }
inline void motion_inv(const D<2, 1, DV>& a, DV out[3]) { out[0] = a.x; out[1] = a.dx[0]; out[2] = DV(0.); }
inline void motion_inv(const D<3, 1, D<2, 1, DV>>& a, DV out[3]) { out[0] = a.x.x; out[1] = a.dx[0].x; out[2] = a.dx[0].dx[0]; }

// ---------------------------------------------------------------- resultants  BeamElement.jl:28-64
struct Resultants { DV fi; V3<DV> mi, fe, me; };
inline Resultants resultants(const BeamCrossSection& o, const DV eps[3], const V3<DV> kap[3], const V3<DV> x[3], const M33<DV> rsm[3], const V3<DV> vi[3]) {
    const M33<DV>& r0 = rsm[0];
    const V3<DV>& vi2 = vi[2];
    const V3<DV>&x1 = x[1], &x2 = x[2];
    V3<DV> xl1 = vecmat(x1, r0);
    V3<DV> xl2 = vecmat(x2, r0);
    V3<DV> fi; for (int i = 0; i < 3; ++i) fi[i] = o.mu * x2[i];
    V3<DV> fal{o.Ca1 * xl2[0], o.Ca2 * xl2[1], o.Ca3 * xl2[2]};
    V3<DV> fag = matvec(r0, fal);
    V3<DV> fll{o.Cl1 * xl1[0], o.Cl2 * xl1[1], o.Cl3 * xl1[2]};
    V3<DV> flg = matvec(r0, fll);
    DV fq1 = o.Cq1 * ipow(xl1[0], 2); if (VALUE(xl1[0]) < 0) fq1 = -fq1;
    DV fq2 = o.Cq2 * ipow(xl1[1], 2); if (VALUE(xl1[1]) < 0) fq2 = -fq2;
    DV fq3 = o.Cq3 * ipow(xl1[2], 2); if (VALUE(xl1[2]) < 0) fq3 = -fq3;
    V3<DV> fql{fq1, fq2, fq3};
    V3<DV> fqg = matvec(r0, fql);
    Resultants R;
    const double fw[3] = {0., 0., o.w};
    for (int i = 0; i < 3; ++i) R.fe[i] = (((fi[i] + fag[i]) + flg[i]) + fqg[i]) + fw[i];
    DV m1l = o.iota1 * vi2[0];
    for (int i = 0; i < 3; ++i) R.me[i] = r0(i, 0) * m1l;
    R.fi = o.EA * eps[0];
    R.mi[0] = o.GJ * kap[0][0]; R.mi[1] = o.EI3 * kap[0][1]; R.mi[2] = o.EI2 * kap[0][2];
    return R;
}

// ---------------------------------------------------------------- residual  BeamElement.jl:151-174
// X[ider][12] as ∂ℝ{1,Np} (seeded by the caller = the solver's addin!), U0[3] likewise (zero partials when not seeded).
// Output R[12] as ∂ℝ{1,Np}.
// `espy` (77 doubles, optional): the ☼/♢ requestables of the element (@espy, BeamElement.jl:151-174 and :28-64), values only:
//   ε, rₛₘ (column-major 9), κ♢ (3), then per Gauss point x(3), κgp(3), fᵢ, mᵢ(3), fₑ(3) (before the −U of :169), mₑ(3)
template <int ND> inline void beam_residual(const EulerBeam3D& o, const DV (*X)[12], bool udof, const DV* U0, DV* R, double* espy = nullptr) {
    using T = typename MotionT<ND>::type;
    // motion{P}(X)
    T X_[12];
    for (int i = 0; i < 12; ++i) { DV a[3]; for (int d = 0; d < ND; ++d) a[d] = X[d][i]; X_[i] = motion(a, std::integral_constant<int, ND>()); }
    // kinematics{:direct}
    Kin<T> kd = kinematics<false, T>(o, X_);
    // motion⁻¹{P,ND}(gp_,ε_,rₛₘ_)
    DV eps[3]; motion_inv(kd.eps, eps);
    M33<DV> rsm[3]; for (int i = 0; i < 9; ++i) { DV t[3]; motion_inv(kd.rsm.m[i], t); for (int d = 0; d < 3; ++d) rsm[d].m[i] = t[d]; }
    V3<DV> gx[ngp][3], gk[ngp][3];
    for (int g = 0; g < ngp; ++g) for (int i = 0; i < 3; ++i) {
        DV t[3];
        motion_inv(kd.x[g][i], t);   for (int d = 0; d < 3; ++d) gx[g][d][i] = t[d];
        motion_inv(kd.kap[g][i], t); for (int d = 0; d < 3; ++d) gk[g][d][i] = t[d];
    }
    // intrinsicrotationrates  Rotations.jl:177-182
    V3<DV> vi[3];
    for (int i = 0; i < 3; ++i) { vi[0][i] = DV(0.); vi[1][i] = DV(0.); vi[2][i] = DV(0.); }
    if (ND >= 2) vi[1] = spin_inv(matmul(transpose(rsm[0]), rsm[1]));
    if (ND >= 3) vi[2] = spin_inv(matmul(transpose(rsm[1]), rsm[1]) + matmul(transpose(rsm[0]), rsm[2]));
    // Jacobians: TX₀ = revariate{P}(X₀); kinematics{:chainrule}; composeJacobian{P}
    T2<12> TX0[12];
    for (int i = 0; i < 12; ++i) TX0[i] = revariate2<12>(VALUE(X[0][i]), i);
    Kin<T2<12>> kc = kinematics<true, T2<12>>(o, TX0);
    DV dX[12]; for (int i = 0; i < 12; ++i) dX[i] = X[0][i] - VALUE(X[0][i]);
    DV eps_dX[12]; V3<DV> vsm_dX[12], x_dX[ngp][12], k_dX[ngp][12];     // [·][j] = ∂(·)/∂X₀[j]
    for (int j = 0; j < 12; ++j) {
        eps_dX[j] = mclaurin(kc.eps.dx[j], dX);
        for (int i = 0; i < 3; ++i) vsm_dX[j][i] = mclaurin(kc.vsm[i].dx[j], dX);
        for (int g = 0; g < ngp; ++g) for (int i = 0; i < 3; ++i) {
            x_dX[g][j][i] = mclaurin(kc.x[g][i].dx[j], dX);
            k_dX[g][j][i] = mclaurin(kc.kap[g][i].dx[j], dX);
        }
    }
    // quadrature loop
    DV Rg[ngp][12];
    for (int g = 0; g < ngp; ++g) {
        Resultants r = resultants(o.mat, eps, gk[g], gx[g], rsm, vi);
        if (espy) {
            double* q = espy + 13 + 16 * g;
            for (int i = 0; i < 3; ++i) { q[i] = VALUE(gx[g][0][i]); q[3 + i] = VALUE(gk[g][0][i]); q[7 + i] = VALUE(r.mi[i]); q[10 + i] = VALUE(r.fe[i]); q[13 + i] = VALUE(r.me[i]); }
            q[6] = VALUE(r.fi);
        }
        if (udof) for (int i = 0; i < 3; ++i) r.fe[i] = r.fe[i] - U0[i];
        for (int j = 0; j < 12; ++j) {
            DV t1 = r.fi * eps_dX[j];
            DV t2 = (r.mi[0] * k_dX[g][j][0] + r.mi[1] * k_dX[g][j][1]) + r.mi[2] * k_dX[g][j][2];
            DV t3 = (r.fe[0] * x_dX[g][j][0] + r.fe[1] * x_dX[g][j][1]) + r.fe[2] * x_dX[g][j][2];
            DV t4 = (r.me[0] * vsm_dX[j][0] + r.me[1] * vsm_dX[j][1]) + r.me[2] * vsm_dX[j][2];
            Rg[g][j] = (((t1 + t2) + t3) + t4) * o.dL[g];
        }
    }
    for (int j = 0; j < 12; ++j) R[j] = ((Rg[0][j] + Rg[1][j]) + Rg[2][j]) + Rg[3][j];
    if (espy) {
        espy[0] = VALUE(eps[0]);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) espy[1 + i + 3 * j] = VALUE(rsm[0](i, j));
        DV k0[3], k1[3], k2[3];                         // ♢κ = motion⁻¹{P,ND}(SVector(vₗ₂[1],vₗ₂[3],−vₗ₂[2])).*(2/L)
        motion_inv(kd.vl2[0], k0); motion_inv(kd.vl2[2], k1); motion_inv(kd.vl2[1], k2);
        espy[10] = VALUE(k0[0]) * (2 / o.L); espy[11] = VALUE(k1[0]) * (2 / o.L); espy[12] = -VALUE(k2[0]) * (2 / o.L);
    }
}

}  // namespace orc
