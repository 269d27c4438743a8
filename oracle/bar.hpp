// ORACLE — TEST INFRASTRUCTURE ONLY (see adiff.hpp).
//
// Literal restatement of
//   toolbox/BarElement.jl : AxisymmetricBarCrossSection (:36-48), Bar3D ctor (:117-133), resultants (:136-161), residual (:167-202)
//   toolbox/SoilContact.jl: residual (:10-20)
#pragma once
#include "beam.hpp"

namespace orc {

struct BarCrossSection { double EA, mu, w, Cat, Clt, Cqt, Can, Cln, Cqn; };       // BarElement.jl:36-46
struct Bar3D {                                                                      // BarElement.jl:89-101
    double cm[3], tgm[3], tge[3], L0, Ls;
    BarCrossSection mat;
    double wgp[ngp], zgp[ngp], znod[2], psi1[ngp], psi2[ngp];
};
constexpr int BAR_LEN = 3 + 3 + 3 + 1 + 1 + 9 + 4 + 4 + 2 + 4 + 4;   // 38 doubles

inline void bar_ctor(Bar3D& o, const double* c1, const double* c2, const BarCrossSection& mat, double eps_s) {
    for (int i = 0; i < 3; ++i) { o.cm[i] = (c1[i] + c2[i]) / 2; o.tgm[i] = c2[i] - c1[i]; o.tge[i] = 0.; }
    o.L0 = std::sqrt(o.tgm[0] * o.tgm[0] + o.tgm[1] * o.tgm[1] + o.tgm[2] * o.tgm[2]);
    o.Ls = (1 - eps_s) * o.L0;
    o.tge[0] = o.L0;
    const double s30 = std::sqrt(30.), s65 = std::sqrt(6. / 5);
    o.wgp[0] = o.L0 / 2 * (18 - s30) / 36; o.wgp[1] = o.L0 / 2 * (18 + s30) / 36; o.wgp[2] = o.L0 / 2 * (18 + s30) / 36; o.wgp[3] = o.L0 / 2 * (18 - s30) / 36;
    o.zgp[0] = -1. / 2 * std::sqrt(3. / 7 + 2. / 7 * s65); o.zgp[1] = -1. / 2 * std::sqrt(3. / 7 - 2. / 7 * s65);
    o.zgp[2] = +1. / 2 * std::sqrt(3. / 7 - 2. / 7 * s65); o.zgp[3] = +1. / 2 * std::sqrt(3. / 7 + 2. / 7 * s65);
    o.znod[0] = -1. / 2; o.znod[1] = 1. / 2;
    for (int g = 0; g < ngp; ++g) { o.psi1[g] = -o.zgp[g] + 1. / 2; o.psi2[g] = o.zgp[g] + 1. / 2; }
    o.mat = mat;
}
inline void pack_bar(const Bar3D& o, double* e) {
    int k = 0;
    for (int i = 0; i < 3; ++i) e[k++] = o.cm[i];
    for (int i = 0; i < 3; ++i) e[k++] = o.tgm[i];
    for (int i = 0; i < 3; ++i) e[k++] = o.tge[i];
    e[k++] = o.L0; e[k++] = o.Ls;
    std::memcpy(e + k, &o.mat, 9 * sizeof(double)); k += 9;
    for (int i = 0; i < 4; ++i) e[k++] = o.wgp[i];
    for (int i = 0; i < 4; ++i) e[k++] = o.zgp[i];
    for (int i = 0; i < 2; ++i) e[k++] = o.znod[i];
    for (int i = 0; i < 4; ++i) e[k++] = o.psi1[i];
    for (int i = 0; i < 4; ++i) e[k++] = o.psi2[i];
}
inline void unpack_bar(const double* e, Bar3D& o) {
    int k = 0;
    for (int i = 0; i < 3; ++i) o.cm[i] = e[k++];
    for (int i = 0; i < 3; ++i) o.tgm[i] = e[k++];
    for (int i = 0; i < 3; ++i) o.tge[i] = e[k++];
    o.L0 = e[k++]; o.Ls = e[k++];
    std::memcpy(&o.mat, e + k, 9 * sizeof(double)); k += 9;
    for (int i = 0; i < 4; ++i) o.wgp[i] = e[k++];
    for (int i = 0; i < 4; ++i) o.zgp[i] = e[k++];
    for (int i = 0; i < 2; ++i) o.znod[i] = e[k++];
    for (int i = 0; i < 4; ++i) o.psi1[i] = e[k++];
    for (int i = 0; i < 4; ++i) o.psi2[i] = e[k++];
}

// residual(o::Bar3D,X,U,A,t,SP,dbg)   BarElement.jl:167-202.  X[ider][6] as ∂ℝ{1,Np}; t as a plain number (its partials are zero)
template <int ND> inline void bar_residual(const Bar3D& o, const DV (*X)[6], bool udof, const DV* U0, double t, DV* R) {
    using T = typename MotionT<ND>::type;
    T x_[6];
    for (int i = 0; i < 6; ++i) { DV a[3]; for (int d = 0; d < ND; ++d) a[d] = X[d][i]; x_[i] = motion(a, std::integral_constant<int, ND>()); }
    V3<T> u1{x_[0], x_[1], x_[2]}, u2{x_[3], x_[4], x_[5]};
    V3<T> c, tg;
    for (int i = 0; i < 3; ++i) { c[i] = o.cm[i] + 0.5 * (u1[i] + u2[i]); tg[i] = (o.tgm[i] + u2[i]) - u1[i]; }
    T L = dsqrt((ipow(tg[0], 2) + ipow(tg[1], 2)) + ipow(tg[2], 2));
    V3<T> d_; for (int i = 0; i < 3; ++i) d_[i] = tg[i] / L;
    T eps_ = L / o.Ls - 1.;
    DV eps[3]; motion_inv(eps_, eps);
    V3<DV> dl[3]; for (int i = 0; i < 3; ++i) { DV tt[3]; motion_inv(d_[i], tt); for (int k = 0; k < 3; ++k) dl[k][i] = tt[k]; }
    const V3<DV>& d0 = dl[0];
    DV eps_dX[6];
    { const double s = 1 / o.L0; for (int i = 0; i < 3; ++i) { eps_dX[i] = s * (-d0[i]); eps_dX[3 + i] = s * d0[i]; } }
    DV Rg[ngp][6];
    for (int g = 0; g < ngp; ++g) {
        const double z = o.zgp[g];
        V3<DV> x[3];
        for (int i = 0; i < 3; ++i) { T xg = c[i] + tg[i] * z; DV tt[3]; motion_inv(xg, tt); for (int k = 0; k < 3; ++k) x[k][i] = tt[k]; }
        const double p1 = -z + 1. / 2, p2 = z + 1. / 2;                      // ψ₁(ζ), ψ₂(ζ)
        // resultants  BarElement.jl:136-161
        const V3<DV>&v = x[1], &a = x[2];
        V3<DV> fi; for (int i = 0; i < 3; ++i) fi[i] = o.mat.mu * a[i];
        const double fw[3] = {0., 0., (std::fmin(t, -5.) + 10) / 5 * o.mat.w};
        DV at = (a[0] * d0[0] + a[1] * d0[1]) + a[2] * d0[2];
        V3<DV> an; for (int i = 0; i < 3; ++i) an[i] = a[i] - at * d0[i];
        V3<DV> fa{o.mat.Cat * at, o.mat.Can * an[1], o.mat.Can * an[2]};
        DV vt = (v[0] * d0[0] + v[1] * d0[1]) + v[2] * d0[2];
        V3<DV> vn; for (int i = 0; i < 3; ++i) vn[i] = v[i] - vt * d0[i];
        DV fqt = o.mat.Cqt * ipow(vt, 2); if (VALUE(vt) < 0) fqt = -fqt;
        DV fqn2 = o.mat.Cqn * ipow(vn[1], 2); if (VALUE(vn[1]) < 0) fqn2 = -fqn2;
        DV fqn3 = o.mat.Cqn * ipow(vn[2], 2); if (VALUE(vn[2]) < 0) fqn3 = -fqn3;
        V3<DV> fd{o.mat.Clt * vt + fqt, o.mat.Cln * vn[1] + fqn2, o.mat.Cln * vn[2] + fqn3};
        V3<DV> fe; for (int i = 0; i < 3; ++i) fe[i] = ((fi[i] + fw[i]) + fa[i]) + fd[i];
        DV fint = o.mat.EA * eps[0];
        if (udof) for (int i = 0; i < 3; ++i) fe[i] = fe[i] - U0[i];
        for (int j = 0; j < 6; ++j) {
            // fₑ ∘₁ x∂X₀ with x∂X₀ = [ψ₁I ψ₂I] : the SMatrix product sums three terms, two of them with a zero factor
            const int i = j % 3; const double psi = j < 3 ? p1 : p2;
            DV t2 = fe[0] * (i == 0 ? psi : 0.);
            t2 = t2 + fe[1] * (i == 1 ? psi : 0.);
            t2 = t2 + fe[2] * (i == 2 ? psi : 0.);
            Rg[g][j] = (fint * eps_dX[j] + t2) * o.wgp[g];
        }
    }
    for (int j = 0; j < 6; ++j) R[j] = ((Rg[0][j] + Rg[1][j]) + Rg[2][j]) + Rg[3][j];
}

// residual(o::SoilContact,X,U,A,t,SP,dbg)   SoilContact.jl:10-20.  Returns false when z ≥ z₀ (R = SVector(0,0,0): integer zeros, no partials)
struct SoilContact { double z0, Kh, Kv, Ch, Cv; };
inline bool soil_residual(const SoilContact& o, int ND, const DV (*X)[3], DV* R) {
    DV z = X[0][2];
    if (!(VALUE(z) < o.z0)) { for (int i = 0; i < 3; ++i) R[i] = DV(0.); return false; }
    DV zero(0.);
    const DV &x = X[0][0], &y = X[0][1];
    const DV &xp = ND >= 2 ? X[1][0] : zero, &yp = ND >= 2 ? X[1][1] : zero, &zp = ND >= 2 ? X[1][2] : zero;
    R[0] = o.Kh * x + o.Ch * xp;
    R[1] = o.Kh * y + o.Ch * yp;
    R[2] = o.Kv * (z - o.z0) + o.Cv * zp;
    return true;
}

}  // namespace orc
