// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points (ctypes) over the literal restatement in adiff/rotations/beam/bar.hpp.
// Nothing under muscade.jl_b200/ may link or load this file; only tests/, __graft_entry__.smoke() and bench.py's CPU legs do.
//
// The assembly loops restate
//   src/Assemble.jl:470-487   assemble! / assemble_!  (serial loop over elements, gather of element dofs)
//   src/SweepX.jl:45-96       addin!{:step|:iter} for AssemblySweepX{0,1,2}
//   src/DirectXUA.jl:85-120   addin!{:matrices} for AssemblyDirect, no_second_order elements
//   src/Assemble.jl:559-608   add_value!, add_∂!{P,S,T}
#include "beam.hpp"
#include "bar.hpp"
#include <cstdint>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

static void unpack_beam(const double* e, EulerBeam3D& o) {
    int k = 0;
    for (int i = 0; i < 3; ++i) o.cm[i] = e[k++];
    for (int i = 0; i < 9; ++i) o.rm.m[i] = e[k++];
    for (int i = 0; i < 4; ++i) o.zgp[i] = e[k++];
    for (int i = 0; i < 2; ++i) o.znod[i] = e[k++];
    for (int i = 0; i < 3; ++i) o.tgm[i] = e[k++];
    for (int i = 0; i < 3; ++i) o.tge[i] = e[k++];
    for (int i = 0; i < 4; ++i) o.ya[i] = e[k++];
    for (int i = 0; i < 4; ++i) o.yu[i] = e[k++];
    for (int i = 0; i < 4; ++i) o.yv[i] = e[k++];
    for (int i = 0; i < 4; ++i) o.ka[i] = e[k++];
    for (int i = 0; i < 4; ++i) o.ku[i] = e[k++];
    for (int i = 0; i < 4; ++i) o.kv[i] = e[k++];
    o.L = e[k++];
    for (int i = 0; i < 4; ++i) o.dL[i] = e[k++];
    std::memcpy(&o.mat, e + k, 16 * sizeof(double));
}
static void pack_beam(const EulerBeam3D& o, double* e) {
    int k = 0;
    for (int i = 0; i < 3; ++i) e[k++] = o.cm[i];
    for (int i = 0; i < 9; ++i) e[k++] = o.rm.m[i];
    for (int i = 0; i < 4; ++i) e[k++] = o.zgp[i];
    for (int i = 0; i < 2; ++i) e[k++] = o.znod[i];
    for (int i = 0; i < 3; ++i) e[k++] = o.tgm[i];
    for (int i = 0; i < 3; ++i) e[k++] = o.tge[i];
    for (int i = 0; i < 4; ++i) e[k++] = o.ya[i];
    for (int i = 0; i < 4; ++i) e[k++] = o.yu[i];
    for (int i = 0; i < 4; ++i) e[k++] = o.yv[i];
    for (int i = 0; i < 4; ++i) e[k++] = o.ka[i];
    for (int i = 0; i < 4; ++i) e[k++] = o.ku[i];
    for (int i = 0; i < 4; ++i) e[k++] = o.kv[i];
    e[k++] = o.L;
    for (int i = 0; i < 4; ++i) e[k++] = o.dL[i];
    std::memcpy(e + k, &o.mat, 16 * sizeof(double));
}

template <class F> static void beam_dispatch(int nd, F&& f) {
    if (nd == 1) f(std::integral_constant<int, 1>());
    else if (nd == 2) f(std::integral_constant<int, 2>());
    else f(std::integral_constant<int, 3>());
}

// nested variate{k}(...variate{1}(x)) and extraction ∂{1}(∂{2}(...)) for TestRotations.jl:35-60
template <int K> struct Nest { using type = D<K, 1, typename Nest<K - 1>::type>; };
template <> struct Nest<0> { using type = double; };
template <int K> static typename Nest<K>::type nest_variate(double x) {
    if constexpr (K == 0) return x;
    else {
        typename Nest<K>::type r;
        r.x = nest_variate<K - 1>(x);
        r.dx[0] = typename Nest<K - 1>::type(1.);
        return r;
    }
}
template <int K> static double nest_extract(const typename Nest<K>::type& y) {
    if constexpr (K == 0) return y;
    else return nest_extract<K - 1>(y.dx[0]);
}

extern "C" {

int orc_beam_struct_len() { return 69; }

// EulerBeam3D constructor (BeamElement.jl:121-148). mat = 16 doubles in BeamCrossSection field order.
int orc_beam_ctor(const double* c1, const double* c2, const double* orient2, const double* mat, double* out69) {
    EulerBeam3D o; BeamCrossSection m; std::memcpy(&m, mat, sizeof m);
    int rc = beam_ctor(o, c1, c2, orient2, m);
    if (rc) return rc;
    pack_beam(o, out69);
    return 0;
}

// residual(o::EulerBeam3D, X,U,A,t,SP,dbg) with arbitrary first-order seeding (BeamElement.jl:151).
//   Xval[nd][12], Xseed[nd][12][np], Uval[3], Useed[3][np]  →  R[12], dR[12][np]
int orc_beam_residual(const double* elem69, int nd, int np, const double* Xval, const double* Xseed, int udof,
                      const double* Uval, const double* Useed, double* R, double* dR) {
    if (np > DV_MAX || nd < 1 || nd > 3) return -1;
    if (!DVctx::set(np)) return -1;
    EulerBeam3D o; unpack_beam(elem69, o);
    DV X[3][12], U[3], Rv[12];
    for (int d = 0; d < nd; ++d) for (int i = 0; i < 12; ++i) {
        X[d][i].x = Xval[d * 12 + i];
        for (int p = 0; p < np; ++p) X[d][i].dx[p] = Xseed ? Xseed[(d * 12 + i) * np + p] : 0.;
    }
    for (int i = 0; i < 3; ++i) {
        U[i].x = (udof && Uval) ? Uval[i] : 0.;
        for (int p = 0; p < np; ++p) U[i].dx[p] = (udof && Useed) ? Useed[i * np + p] : 0.;
    }
    beam_dispatch(nd, [&](auto ND) { beam_residual<decltype(ND)::value>(o, X, udof != 0, U, Rv); });
    int nan = 0;
    for (int i = 0; i < 12; ++i) {
        R[i] = Rv[i].x;
        for (int p = 0; p < np; ++p) dR[i * np + p] = Rv[i].dx[p];
        nan |= hasnan(Rv[i]);
    }
    return nan ? 2 : 0;
}

// getresult(state,req,els) for EulerBeam3D (src/Output.jl:131-181 → the @espy requestables): 77 values per element, see beam_residual
int orc_beam_results(const double* elem69, int nd, const double* Xval, int udof, const double* Uval, double* espy77) {
    if (nd < 1 || nd > 3) return -1;
    if (!DVctx::set(0)) return -1;
    EulerBeam3D o; unpack_beam(elem69, o);
    DV X[3][12], U[3], Rv[12];
    for (int d = 0; d < nd; ++d) for (int i = 0; i < 12; ++i) X[d][i].x = Xval[d * 12 + i];
    for (int i = 0; i < 3; ++i) U[i].x = (udof && Uval) ? Uval[i] : 0.;
    beam_dispatch(nd, [&](auto ND) { beam_residual<decltype(ND)::value>(o, X, udof != 0, U, Rv, espy77); });
    return 0;
}

double orc_sinc1k(int k, double x) {
    switch (k) { case 0: return sinc1k<0>(x); case 1: return sinc1k<1>(x); case 2: return sinc1k<2>(x); case 3: return sinc1k<3>(x); case 4: return sinc1k<4>(x); }
    return NAN;
}
// n-th derivative of scac by nested univariate duals (TestRotations.jl scac1..scac4)
double orc_scac_d(int n, double x) {
    switch (n) {
        case 0: return scac(x);
        case 1: return nest_extract<1>(scac(nest_variate<1>(x)));
        case 2: return nest_extract<2>(scac(nest_variate<2>(x)));
        case 3: return nest_extract<3>(scac(nest_variate<3>(x)));
        case 4: return nest_extract<4>(scac(nest_variate<4>(x)));
    }
    return NAN;
}
// M = Rodrigues(v) (col-major 9) ; w = Rodrigues⁻¹(M), dw[i][j] = ∂w_i/∂v_j by variate{1,3} (TestRotations.jl:87-110)
void orc_rodrigues_roundtrip(const double* v, double* M, double* w, double* dw) {
    using T = D<1, 3, double>;
    V3<T> tv; for (int i = 0; i < 3; ++i) { tv[i] = T(v[i]); tv[i].dx[i] = 1.; }
    M33<T> m = rodrigues(tv);
    V3<T> tw = rodrigues_inv(m);
    for (int i = 0; i < 9; ++i) M[i] = m.m[i].x;
    for (int i = 0; i < 3; ++i) { w[i] = tw[i].x; for (int j = 0; j < 3; ++j) dw[i * 3 + j] = tw[i].dx[j]; }
}

// ------------------------------------------------------------------------------------------------ SweepX assembly
// One element type of EulerBeam3D{Mat,false}. Arrays are the reference's: idx = dis.index[iele].X (1-based),
// asm1 = asm[1,ityp] (12×nele), asm2 = asm[2,ityp] (144×nele), column-major, 1-based, 0 = skip.
// newmark = (a1,a2,a3,b1,b2,b3,Δt) (SweepX.jl:3-15). mission: 0 = :step, 1 = :iter.
// Llambda / nzval are ACCUMULATED into (caller zeroes them: zero!(out), Assemble.jl:471).
// Returns 0, or 1+iele (0-based) of the first element returning NaN (Assemble.jl:630).
int orc_sweepx_assemble_beams(int64_t nele, const double* elems69, const int64_t* idx, const int64_t* asm1, const int64_t* asm2,
                              int OX, int mission, const double* X0, const double* X1, const double* X2,
                              const double* scaleX, const double* newmark, double* Llambda, double* nzval) {
    const int Nx = 12;
    const bool step = (mission == 0) && OX > 0;
    const int np = step ? Nx + 1 : Nx;
    if (!DVctx::set(np)) return -1;
    const double a1 = newmark[0], a2 = newmark[1], a3 = newmark[2], b1 = newmark[3], b2 = newmark[4], b3 = newmark[5];
    const int nd = OX + 1;
    for (int64_t e = 0; e < nele; ++e) {
        EulerBeam3D o; unpack_beam(elems69 + 69 * e, o);
        const int64_t* ix = idx + 12 * e;
        double x[12], xp[12] = {0}, xpp[12] = {0};
        for (int i = 0; i < 12; ++i) { x[i] = X0[ix[i] - 1]; if (OX >= 1) xp[i] = X1[ix[i] - 1]; if (OX >= 2) xpp[i] = X2[ix[i] - 1]; }
        DV dXs[12], dr;                                    // δX (and δr) : Adiff.jl:106 / Taylor.jl:168-170
        for (int i = 0; i < 12; ++i) { dXs[i] = DV(0.); dXs[i].dx[i] = scaleX[i]; }
        dr = DV(0.); if (step) dr.dx[Nx] = 1.;
        DV X[3][12], U[3], R[12];
        for (int i = 0; i < 12; ++i) {
            X[0][i] = x[i] + dXs[i];
            if (OX == 2) {
                if (step) {
                    double a = a2 * xp[i] + a3 * xpp[i], b = b2 * xp[i] + b3 * xpp[i];
                    X[1][i] = (xp[i] + a1 * dXs[i]) + a * dr;
                    X[2][i] = (xpp[i] + b1 * dXs[i]) + b * dr;
                } else {
                    X[1][i] = xp[i] + a1 * dXs[i];
                    X[2][i] = xpp[i] + b1 * dXs[i];
                }
            } else if (OX == 1) {
                if (step) { double a = a2 * xp[i]; X[1][i] = (xp[i] + a1 * dXs[i]) + a * dr; }
                else X[1][i] = xp[i] + a1 * dXs[i];
            }
        }
        beam_dispatch(nd, [&](auto ND) { beam_residual<decltype(ND)::value>(o, X, false, U, R); });
        for (int i = 0; i < 12; ++i) if (hasnan(R[i])) return int(1 + e);
        for (int i = 0; i < 12; ++i) R[i] = R[i] * scaleX[i];                       // Lλ .* scale.X
        const int64_t* m1 = asm1 + 12 * e;
        const int64_t* m2 = asm2 + 144 * e;
        for (int i = 0; i < 12; ++i) if (m1[i]) Llambda[m1[i] - 1] += R[i].x;       // add_value!
        if (step) for (int i = 0; i < 12; ++i) if (m1[i]) Llambda[m1[i] - 1] -= R[i].dx[Nx];   // add_∂!{1,:minus}
        for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) {                 // add_∂!{1}
            int64_t k = m2[i + 12 * j];
            if (k) nzval[k - 1] += R[i].dx[j];
        }
    }
    return 0;
}

// CPU baseline helper: per-element residual + tangent only (no scatter), optionally OpenMP-parallel over elements.
// Same seeding as orc_sweepx_assemble_beams with mission :iter. Writes Re[12*nele], Ke[144*nele] (col-major i+12j). Returns #NaN elements.
int orc_beams_iter_elementwise(int64_t nele, const double* elems69, const int64_t* idx, int OX,
                               const double* X0, const double* X1, const double* X2, const double* scaleX,
                               const double* newmark, double* Re, double* Ke, int nthreads) {
    const double a1 = newmark[0], b1 = newmark[3];
    const int nd = OX + 1;
    int bad = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static) reduction(+ : bad)
#endif
    for (int64_t e = 0; e < nele; ++e) {
        DVctx::set(12);
        EulerBeam3D o; unpack_beam(elems69 + 69 * e, o);
        const int64_t* ix = idx + 12 * e;
        DV X[3][12], U[3], R[12];
        for (int i = 0; i < 12; ++i) {
            DV d(0.); d.dx[i] = scaleX[i];
            X[0][i] = X0[ix[i] - 1] + d;
            if (OX >= 1) X[1][i] = X1[ix[i] - 1] + a1 * d;
            if (OX >= 2) X[2][i] = X2[ix[i] - 1] + b1 * d;
        }
        beam_dispatch(nd, [&](auto ND) { beam_residual<decltype(ND)::value>(o, X, false, U, R); });
        for (int i = 0; i < 12; ++i) {
            if (hasnan(R[i])) bad += 1;
            DV r = R[i] * scaleX[i];
            Re[12 * e + i] = r.x;
            for (int j = 0; j < 12; ++j) Ke[144 * e + i + 12 * j] = r.dx[j];
        }
    }
    return bad;
}

// Reference algorithm with the element loop parallelised over host threads (NOT something Muscade does: its loop
// src/Assemble.jl:479 is serial), followed by the serial scatter in element order. Used by `bench.py --impl reference`.
int orc_sweepx_assemble_beams_mt(int64_t nele, const double* elems69, const int64_t* idx, const int64_t* asm1, const int64_t* asm2,
                                 int OX, const double* X0, const double* X1, const double* X2, const double* scaleX,
                                 const double* newmark, double* Llambda, double* nzval, int nthreads) {
    std::vector<double> Re((size_t)nele * 12), Ke((size_t)nele * 144);
    int bad = orc_beams_iter_elementwise(nele, elems69, idx, OX, X0, X1, X2, scaleX, newmark, Re.data(), Ke.data(), nthreads);
    if (bad) return bad;
    for (int64_t e = 0; e < nele; ++e) {
        const int64_t* m1 = asm1 + 12 * e; const int64_t* m2 = asm2 + 144 * e;
        for (int i = 0; i < 12; ++i) if (m1[i]) Llambda[m1[i] - 1] += Re[12 * e + i];
        for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) { int64_t k = m2[i + 12 * j]; if (k) nzval[k - 1] += Ke[144 * e + i + 12 * j]; }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ DirectXUA assembly
// addin!{:matrices}(out::AssemblyDirect{OX,OU,0},asm,iele,scale,eleobj,no_second_order=Val(true),Λ,X,U,A,…) for one EulerBeam3D type,
// all elements of one time step (src/DirectXUA.jl:85-120; IA = 0).  Maps are asm[arrnum(·)] of prepare(AssemblyDirect) (1-based, 0 = skip):
//   asmL 12×nele, asmLX/asmXL 144×nele, asmLU/asmUL 36×nele.  Outputs are ACCUMULATED:
//   L1L[nΛ] ; L2LX[(OX+1)][nnzΛX], L2XL[(OX+1)][nnzXΛ], L2LU[(OU+1)][nnzΛU], L2UL[(OU+1)][nnzUΛ]  (row-major 2-D arrays)
int orc_direct_addin_beams(int64_t nele, const double* elems69, const int64_t* idxX, const int64_t* idxU, int udof, int OX, int OU,
                           const double* X0, const double* X1, const double* X2, const double* U0, const double* U1, const double* U2,
                           const double* scaleX, const double* scaleU,
                           const int64_t* asmL, const int64_t* asmLX, const int64_t* asmXL, const int64_t* asmLU, const int64_t* asmUL,
                           double* L1L, double* L2LX, int64_t nnzLX, double* L2XL, int64_t nnzXL, double* L2LU, int64_t nnzLU, double* L2UL, int64_t nnzUL) {
    const int nd = OX + 1, ndu = OU + 1;
    const int np = 12 * nd + (udof ? 3 * ndu : 0);
    if (np > DV_MAX) return -1;
    if (!DVctx::set(np)) return -1;
    const double* Xs[3] = {X0, X1, X2}; const double* Us[3] = {U0, U1, U2};
    for (int64_t e = 0; e < nele; ++e) {
        EulerBeam3D o; unpack_beam(elems69 + 69 * e, o);
        DV X[3][12], U[3], R[12];
        for (int d = 0; d < nd; ++d) for (int i = 0; i < 12; ++i) {       // revariate{1}((;X,U),(;X=scale.X,U=scale.U))  Taylor.jl:158-166
            X[d][i] = DV(Xs[d][idxX[12 * e + i] - 1]);
            X[d][i].dx[12 * d + i] = scaleX[i];
        }
        if (udof) for (int i = 0; i < 3; ++i) { U[i] = DV(Us[0][idxU[3 * e + i] - 1]); U[i].dx[12 * nd + i] = scaleU[i]; }
        beam_dispatch(nd, [&](auto ND) { beam_residual<decltype(ND)::value>(o, X, udof != 0, U, R); });
        for (int i = 0; i < 12; ++i) if (hasnan(R[i])) return int(1 + e);
        const int64_t* mL = asmL + 12 * e;
        for (int i = 0; i < 12; ++i) if (mL[i]) L1L[mL[i] - 1] += R[i].x;                       // add_value!(Lλ[1],…,R,iλ)
        for (int j = 0; j < nd; ++j) {                                                          // β = X
            const int64_t *mLX = asmLX + 144 * e, *mXL = asmXL + 144 * e;
            for (int i = 0; i < 12; ++i) for (int jj = 0; jj < 12; ++jj) {
                const double v = R[i].dx[12 * j + jj];
                int64_t k = mLX[i + 12 * jj]; if (k) L2LX[j * nnzLX + k - 1] += v;              // add_∂!{1}
                k = mXL[jj + 12 * i];        if (k) L2XL[j * nnzXL + k - 1] += v;               // add_∂!{1,:plus,:transpose}
            }
        }
        if (udof) for (int j = 0; j < ndu; ++j) {                                               // β = U
            const int64_t *mLU = asmLU + 36 * e, *mUL = asmUL + 36 * e;
            for (int i = 0; i < 12; ++i) for (int jj = 0; jj < 3; ++jj) {
                const double v = R[i].dx[12 * nd + 3 * j + jj];
                int64_t k = mLU[i + 12 * jj]; if (k) L2LU[j * nnzLU + k - 1] += v;
                k = mUL[jj + 3 * i];         if (k) L2UL[j * nnzUL + k - 1] += v;
            }
        }
    }
    return 0;
}

int orc_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"

#include "bar_api.inc"
#include "kat_api.inc"
