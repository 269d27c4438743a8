// Test harness: compiles the PRODUCT's device math header for the host (g++) so that the kernel's forward-over-reverse
// formulation can be checked against the oracle without a GPU.  Not shipped, not used by the product.
#include "../../muscade.jl_b200/csrc/beam_math.cuh"
#include <cstring>
using namespace mb;

template <int ND, int W>
static void run(const BeamGeo& g, const BeamMat& m, int np, const double* Xval, const double* Xseed, int udof, const double* Uval,
                const double* Useed, double* R, double* dR) {
    for (int p0 = 0; p0 < np || p0 == 0; p0 += W) {
        Dual<W> X[3][12], U[3], Rv[12];
        for (int d = 0; d < 3; ++d) for (int i = 0; i < 12; ++i) {
            X[d][i] = Make<Dual<W>>::c(d < ND ? Xval[d * 12 + i] : 0.);
            for (int k = 0; k < W; ++k) X[d][i].d[k] = (d < ND && p0 + k < np) ? Xseed[(d * 12 + i) * np + p0 + k] : 0.;
        }
        for (int i = 0; i < 3; ++i) {
            U[i] = Make<Dual<W>>::c(udof ? Uval[i] : 0.);
            for (int k = 0; k < W; ++k) U[i].d[k] = (udof && p0 + k < np) ? Useed[i * np + p0 + k] : 0.;
        }
        beam_residual<ND, W>(g, m, X, udof != 0, U, Rv);
        for (int i = 0; i < 12; ++i) {
            R[i] = Rv[i].v;
            for (int k = 0; k < W; ++k) if (p0 + k < np) dR[i * np + p0 + k] = Rv[i].d[k];
        }
        if (np == 0) break;
    }
}

extern "C" int mbh_beam_residual(const double* geo16, const double* mat16, int nd, int w, int np, const double* Xval, const double* Xseed,
                                 int udof, const double* Uval, const double* Useed, double* R, double* dR) {
    BeamGeo g; BeamMat m;
    for (int i = 0; i < 3; ++i) g.cm[i] = geo16[i];
    for (int i = 0; i < 9; ++i) g.rm.a[i] = geo16[3 + i];
    for (int i = 0; i < 3; ++i) g.tgm[i] = geo16[12 + i];
    g.L = geo16[15];
    std::memcpy(&m, mat16, sizeof m);
#define CASE(ND_, W_) if (nd == ND_ && w == W_) { run<ND_, W_>(g, m, np, Xval, Xseed, udof, Uval, Useed, R, dR); return 0; }
    CASE(1, 1) CASE(1, 4) CASE(2, 1) CASE(3, 1) CASE(3, 3) CASE(1, 12)
    return -1;
}

// SD path exactly as the kernel's lanes use it: lane l carries (rotation dof l, translation dof l) of the :iter seeding
// X₀+δX, X₁+a₁δX, X₂+b₁δX (src/SweepX.jl:63). Returns R[12], K[12][12] (K[i][p] = ∂R_i/∂seed_p), unscaled rows.
template <int ND> static void run_sd(const BeamGeo& g, const BeamMat& m, const double* Xval, const double* scale, double a1, double b1,
                                     int udof, const double* Uval, double* R, double* K) {
    using N = NumSD; using TR = N::TR; using TU = N::TU; using TS = N::TS;
    static const int rot[6] = {3, 4, 5, 9, 10, 11}, tra[6] = {0, 1, 2, 6, 7, 8};
    for (int l = 0; l < 6; ++l) {
        TU Xu[3][6], U[3]; TR Xv[3][6]; TS Rv[12];
        for (int d = 0; d < 3; ++d) for (int i = 0; i < 6; ++i) {
            const double coef = d == 0 ? 1.0 : (d == 1 ? a1 : b1);
            Xu[d][i].v = d < ND ? Xval[d * 12 + tra[i]] : 0.; Xu[d][i].d1 = (i == l) ? coef * scale[tra[i]] : 0.;
            Xv[d][i].v = d < ND ? Xval[d * 12 + rot[i]] : 0.; Xv[d][i].d0 = (i == l) ? coef * scale[rot[i]] : 0.;
        }
        for (int i = 0; i < 3; ++i) { U[i].v = udof ? Uval[i] : 0.; U[i].d1 = 0.; }
        beam_residual_n<ND, N>(g, m, Xu, Xv, udof != 0, U, Rv);
        for (int i = 0; i < 12; ++i) { R[i] = Rv[i].v; K[i * 12 + rot[l]] = Rv[i].d0; K[i * 12 + tra[l]] = Rv[i].d1; }
    }
}
extern "C" int mbh_beam_iter_sd(const double* geo16, const double* mat16, int nd, const double* Xval, const double* scale, double a1, double b1,
                                int udof, const double* Uval, double* R, double* K) {
    BeamGeo g; BeamMat m;
    for (int i = 0; i < 3; ++i) g.cm[i] = geo16[i];
    for (int i = 0; i < 9; ++i) g.rm.a[i] = geo16[3 + i];
    for (int i = 0; i < 3; ++i) g.tgm[i] = geo16[12 + i];
    g.L = geo16[15];
    std::memcpy(&m, mat16, sizeof m);
    if (nd == 1) run_sd<1>(g, m, Xval, scale, a1, b1, udof, Uval, R, K);
    else if (nd == 2) run_sd<2>(g, m, Xval, scale, a1, b1, udof, Uval, R, K);
    else run_sd<3>(g, m, Xval, scale, a1, b1, udof, Uval, R, K);
    return 0;
}

// statics, symmetric-tangent path exactly as beam_static_sym_kernel's lanes run it. K[i][p] scaled on rows AND columns (Ke of the kernel, transposed to row-major).
extern "C" int mbh_beam_static_sym(const double* geo16, const double* mat16, const double* Xval, const double* scale, int udof, const double* Uval,
                                   double* R, double* K) {
    BeamGeo g; BeamMat m;
    for (int i = 0; i < 3; ++i) g.cm[i] = geo16[i];
    for (int i = 0; i < 9; ++i) g.rm.a[i] = geo16[3 + i];
    for (int i = 0; i < 3; ++i) g.tgm[i] = geo16[12 + i];
    g.L = geo16[15];
    std::memcpy(&m, mat16, sizeof m);
    static const int rot[6] = {3, 4, 5, 9, 10, 11}, tra[6] = {0, 1, 2, 6, 7, 8};
    int bad = 0;
    for (int l = 0; l < 6; ++l) {
        SD<false, false> Xu[6], U[3]; SD<true, false> Xv[6], Rv[12];
        for (int i = 0; i < 6; ++i) { Xu[i].v = Xval[tra[i]]; Xv[i].v = Xval[rot[i]]; Xv[i].d0 = (i == l) ? scale[rot[i]] : 0.; }
        for (int i = 0; i < 3; ++i) U[i].v = udof ? Uval[i] : 0.;
        double Gc[3];
        beam_static_sym(g, m, Xu, Xv, udof != 0, U, l % 3, Rv, Gc);
        bad |= beam_static_sym_store(l, Rv, Gc, scale, [&](int k, double v) { K[(k % 12) * 12 + k / 12] = v; },
                                     [&](int k, double v0, double v1) { K[(k % 12) * 12 + k / 12] = v0; K[((k + 1) % 12) * 12 + (k + 1) / 12] = v1; });
        for (int i = 0; i < 12; ++i) R[i] = Rv[i].v * scale[i];
    }
    return bad;
}

// statics, active/passive formulation exactly as beam_static_ap_kernel's lanes run it (packs evaluated locally; the kernel exchanges them by shuffles)
extern "C" int mbh_beam_static_ap(const double* geo16, const double* mat16, const double* Xval, const double* scale, int udof, const double* Uval,
                                  double* R, double* K) {
    BeamGeo g; BeamMat m;
    for (int i = 0; i < 3; ++i) g.cm[i] = geo16[i];
    for (int i = 0; i < 9; ++i) g.rm.a[i] = geo16[3 + i];
    for (int i = 0; i < 3; ++i) g.tgm[i] = geo16[12 + i];
    g.L = geo16[15];
    std::memcpy(&m, mat16, sizeof m);
    static const int rot[6] = {3, 4, 5, 9, 10, 11}, tra[6] = {0, 1, 2, 6, 7, 8};
    int bad = 0;
    double xu[6], xv[6];
    for (int i = 0; i < 6; ++i) { xu[i] = Xval[tra[i]]; xv[i] = Xval[rot[i]]; }
    for (int l = 0; l < 6; ++l) {
        SD<false, false> U[3]; SD<true, false> Rv[12];
        for (int i = 0; i < 3; ++i) U[i].v = udof ? Uval[i] : 0.;
        double Gc[3];
        ApCollect col{l < 3, Rv, Gc};
        beam_static_ap_lane(g, m, xu, xv, scale[rot[l]], l, udof != 0, U, PacksLocal(), col);
        bad |= beam_static_sym_store(l, Rv, Gc, scale, [&](int k, double v) { K[(k % 12) * 12 + k / 12] = v; },
                                     [&](int k, double v0, double v1) { K[(k % 12) * 12 + k / 12] = v0; K[((k + 1) % 12) * 12 + (k + 1) / 12] = v1; });
        if (l == 0) for (int i = 0; i < 12; ++i) R[i] = Rv[i].v * scale[i];
    }
    return bad;
}

// getresult values (beam_results, the body of beam_results_kernel): X element dof order t1..r3 of node 1 then node 2, as the reference
extern "C" int mbh_beam_results(const double* geo16, const double* mat16, int nd, const double* Xval, double* out77) {
    BeamGeo g; BeamMat m;
    for (int i = 0; i < 3; ++i) g.cm[i] = geo16[i];
    for (int i = 0; i < 9; ++i) g.rm.a[i] = geo16[3 + i];
    for (int i = 0; i < 3; ++i) g.tgm[i] = geo16[12 + i];
    g.L = geo16[15];
    std::memcpy(&m, mat16, sizeof m);
    static const int rot[6] = {3, 4, 5, 9, 10, 11}, tra[6] = {0, 1, 2, 6, 7, 8};
    double Xu[3][6], Xv[3][6];
    for (int d = 0; d < 3; ++d) for (int i = 0; i < 6; ++i) { Xu[d][i] = d < nd ? Xval[d * 12 + tra[i]] : 0.; Xv[d][i] = d < nd ? Xval[d * 12 + rot[i]] : 0.; }
    if (nd == 1) beam_results<1>(g, m, Xu, Xv, out77);
    else if (nd == 2) beam_results<2>(g, m, Xu, Xv, out77);
    else beam_results<3>(g, m, Xu, Xv, out77);
    return 0;
}

// Two-phase :iter path exactly as beam_cot_kernel + beam_kernel_sd<ND,split> run it (phase A: time-jets → cotangents, phase B: order-0 forward + reverse),
// lane l = (rotation dof l, translation dof l); with MB_DYN_AP the lanes use the active/passive formulation. Returns R[12], K[12][12] like mbh_beam_iter_sd.
template <int ND> static void run_split(const BeamGeo& g, const BeamMat& m, const double* Xval, const double* scale, double a1, double b1,
                                        int udof, const double* Uval, double* R, double* K) {
    using N = NumSD; using TR = N::TR; using TU = N::TU; using TS = N::TS;
    static const int rot[6] = {3, 4, 5, 9, 10, 11}, tra[6] = {0, 1, 2, 6, 7, 8};
    for (int l = 0; l < 6; ++l) {
        TU Xu[3][6], U[3]; TR Xv[3][6]; TS Rv[12];
        for (int d = 0; d < 3; ++d) for (int i = 0; i < 6; ++i) {
            const double coef = d == 0 ? 1.0 : (d == 1 ? a1 : b1);
            Xu[d][i].v = d < ND ? Xval[d * 12 + tra[i]] : 0.; Xu[d][i].d1 = (i == l) ? coef * scale[tra[i]] : 0.;
            Xv[d][i].v = d < ND ? Xval[d * 12 + rot[i]] : 0.; Xv[d][i].d0 = (i == l) ? coef * scale[rot[i]] : 0.;
        }
        for (int i = 0; i < 3; ++i) { U[i].v = udof ? Uval[i] : 0.; U[i].d1 = 0.; }
        Vec3<TS> xb[NGP], vsmb;
        beam_dyn_cotangents<ND, N>(g, m, Xu, Xv, udof != 0, U, xb, vsmb, nullptr, nullptr, l < 3);
        beam_residual_cot<N, TS, (ND >= 3)>(g, m, Xu[0], Xv[0], xb, vsmb, Rv, l < 3);
        for (int i = 0; i < 12; ++i) { R[i] = Rv[i].v; K[i * 12 + rot[l]] = Rv[i].d0; K[i * 12 + tra[l]] = Rv[i].d1; }
    }
}
extern "C" int mbh_beam_iter_split(const double* geo16, const double* mat16, int nd, const double* Xval, const double* scale, double a1, double b1,
                                   int udof, const double* Uval, double* R, double* K) {
    BeamGeo g; BeamMat m;
    for (int i = 0; i < 3; ++i) g.cm[i] = geo16[i];
    for (int i = 0; i < 9; ++i) g.rm.a[i] = geo16[3 + i];
    for (int i = 0; i < 3; ++i) g.tgm[i] = geo16[12 + i];
    g.L = geo16[15];
    std::memcpy(&m, mat16, sizeof m);
    if (nd == 2) run_split<2>(g, m, Xval, scale, a1, b1, udof, Uval, R, K);
    else if (nd == 3) run_split<3>(g, m, Xval, scale, a1, b1, udof, Uval, R, K);
    else return -1;
    return 0;
}
