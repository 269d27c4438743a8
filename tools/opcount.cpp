// Op-counting build of the PRODUCT's kernel mathematics: every `double` of the device math headers
// becomes a counting scalar, the lanes of the element kernels are run on the host exactly as the kernels run them, and the FP64 operations are tallied.
//     F_alg (per element) = Σ_lanes count(lane) − (n_lanes − 1)·count(primal share of a lane)
// i.e. the forward-over-reverse mathematics with the value ("primal") sweep counted ONCE per element, although every lane of the kernel repeats it.
// Conventions: + − × ÷ √ = 1 flop, fma = 2, a libm call (sincos, acos, sinpi) is tallied separately and converted at 40 flop per call.
// Output: one JSON object on stdout → profiles/falg.json (frozen; bench.py reads it).   Build + run: tools/opcount.sh
#include <math.h>
#include <cmath>
#include <cstdio>
#include <cstring>

typedef double real_t;
struct Cnt { long long add = 0, mul = 0, fma_ = 0, div = 0, sqrt_ = 0, libm = 0, cmp = 0; };
static Cnt g_cnt;
struct CR {
    real_t v;
    CR() : v(0) {}
    CR(real_t x) : v(x) {}
    CR(int x) : v(x) {}
    explicit operator long long() const { return (long long)v; }
    explicit operator real_t() const { return v; }
};
inline CR operator+(CR a, CR b) { g_cnt.add++; return CR(a.v + b.v); }
inline CR operator-(CR a, CR b) { g_cnt.add++; return CR(a.v - b.v); }
inline CR operator*(CR a, CR b) { g_cnt.mul++; return CR(a.v * b.v); }
inline CR operator/(CR a, CR b) { g_cnt.div++; return CR(a.v / b.v); }
inline CR operator-(CR a) { return CR(-a.v); }
#define CR_MIX(op) \
    inline CR operator op(CR a, real_t b) { return a op CR(b); } inline CR operator op(real_t a, CR b) { return CR(a) op b; } \
    inline CR operator op(CR a, int b) { return a op CR((real_t)b); } inline CR operator op(int a, CR b) { return CR((real_t)a) op b; }
CR_MIX(+) CR_MIX(-) CR_MIX(*) CR_MIX(/)
#define CR_CMP(op) \
    inline bool operator op(CR a, CR b) { g_cnt.cmp++; return a.v op b.v; } inline bool operator op(CR a, real_t b) { g_cnt.cmp++; return a.v op b; } \
    inline bool operator op(real_t a, CR b) { g_cnt.cmp++; return a op b.v; } inline bool operator op(CR a, int b) { g_cnt.cmp++; return a.v op b; }
CR_CMP(<) CR_CMP(>) CR_CMP(<=) CR_CMP(>=) CR_CMP(==) CR_CMP(!=)
inline CR fma(CR a, CR b, CR c) { g_cnt.fma_++; return CR(std::fma(a.v, b.v, c.v)); }
inline CR fma(CR a, CR b, real_t c) { return fma(a, b, CR(c)); }
inline CR fma(CR a, real_t b, real_t c) { return fma(a, CR(b), CR(c)); }
inline CR fma(CR a, real_t b, CR c) { return fma(a, CR(b), c); }
inline CR sqrt(CR a) { g_cnt.sqrt_++; return CR(std::sqrt(a.v)); }
inline CR fabs(CR a) { return CR(std::fabs(a.v)); }
inline CR acos(CR a) { g_cnt.libm++; return CR(std::acos(a.v)); }
inline CR sin(CR a) { g_cnt.libm++; return CR(std::sin(a.v)); }
inline CR cos(CR a) { return CR(std::cos(a.v)); }                 // sin+cos of one argument = one sincos call (counted at sin)
inline CR nearbyint(CR a) { return CR(std::nearbyint(a.v)); }

#define double CR
#include "../muscade.jl_b200/csrc/beam_math.cuh"
#undef double
using namespace mb;

static void fill(BeamGeo& g, BeamMat& m, CR* xu, CR* xv, bool dyn) {
    // element 3 of the synthetic chain: direction (0.8,0.6,0), orient2 = (0,1,0), L = 1; state of amplitude 0.05 / 0.1
    const real_t t[3] = {0.8, 0.6, 0.}, n[3] = {-0.6, 0.8, 0.}, b[3] = {0., 0., 1.};
    for (int i = 0; i < 3; ++i) { g.cm[i] = CR(3.5 * t[i]); g.rm(i, 0) = CR(t[i]); g.rm(i, 1) = CR(n[i]); g.rm(i, 2) = CR(b[i]); g.tgm[i] = CR(t[i]); }
    g.L = CR(1.0);
    std::memset((void*)&m, 0, sizeof m);
    m.EA = CR(10.); m.EI2 = CR(3.); m.EI3 = CR(3.); m.GJ = CR(4.); m.mu = CR(1.); m.iota1 = CR(1.);
    if (dyn) { m.Ca2 = CR(169.6); m.Ca3 = CR(169.6); m.Cq2 = CR(235.2); m.Cq3 = CR(235.2); }
    const real_t u[12] = {0.031, -0.012, 0.044, -0.027, 0.008, 0.019, 0.071, -0.083, 0.052, -0.064, 0.037, 0.091};
    for (int i = 0; i < 6; ++i) { xu[i] = CR(u[i]); xv[i] = CR(u[6 + i]); }
}
struct NullOut {
    template <class T> void tt(const T*) {}
    template <class T> void trans(const T*, const T*) {}
    template <class T> void rot(const T*, const T*) {}
};
static Cnt take() { Cnt c = g_cnt; g_cnt = Cnt(); return c; }
static Cnt sub(Cnt a, const Cnt& b, long long k) { a.add -= k * b.add; a.mul -= k * b.mul; a.fma_ -= k * b.fma_; a.div -= k * b.div; a.sqrt_ -= k * b.sqrt_; a.libm -= k * b.libm; return a; }
static Cnt plus(Cnt a, const Cnt& b) { return sub(a, b, -1); }
static long long flops(const Cnt& c) { return c.add + c.mul + 2 * c.fma_ + c.div + c.sqrt_ + 40 * c.libm; }
static long long insts(const Cnt& c) { return c.add + c.mul + c.fma_ + c.div + c.sqrt_; }
static void show(const char* name, const Cnt& lanes, const Cnt& alg, int nlanes, bool last = false) {
    std::printf("  \"%s\": {\"lanes\": %d, \"executed\": {\"flop\": %lld, \"add\": %lld, \"mul\": %lld, \"fma\": %lld, \"div\": %lld, \"sqrt\": %lld, \"libm_calls\": %lld, \"arith_ops\": %lld},\n"
                "      \"algorithmic\": {\"flop\": %lld, \"add\": %lld, \"mul\": %lld, \"fma\": %lld, \"div\": %lld, \"sqrt\": %lld, \"libm_calls\": %lld, \"arith_ops\": %lld}}%s\n",
                name, nlanes, flops(lanes), lanes.add, lanes.mul, lanes.fma_, lanes.div, lanes.sqrt_, lanes.libm, insts(lanes),
                flops(alg), alg.add, alg.mul, alg.fma_, alg.div, alg.sqrt_, alg.libm, insts(alg), last ? "" : ",");
}

int main() {
    using V = SD<false, false>; using S = SD<true, false>;
    BeamGeo g; BeamMat m; CR xu[6], xv[6];
    std::printf("{\n  \"convention\": \"+ - * / sqrt = 1 flop, fma = 2, libm call (sincos, acos, sinpi) = 40; algorithmic = sum over the lanes minus (lanes-1) x the primal (value-only) share of a lane; "
                "counted on the product's device math headers compiled for the host with a counting scalar (tools/opcount.cpp)\",\n");
    // ---------------------------------------------------------------- statics: beam_static_ap_kernel (6 lanes, active/passive sweep, packs shared)
    {
        fill(g, m, xu, xv, false);
        V U[3]; for (int i = 0; i < 3; ++i) U[i].v = CR(0.);
        take();
        NullOut out;
        for (int l = 0; l < 6; ++l) beam_static_ap_lane(g, m, xu, xv, CR(1.0), l, false, U, PacksLocal(), out);
        Cnt lanes = take();
        // primal share of a lane: the same sweep in plain values
        V Xu0[6], va[3], vp[3];
        for (int i = 0; i < 6; ++i) Xu0[i].v = xu[i];
        for (int i = 0; i < 3; ++i) { va[i].v = xv[i]; vp[i].v = xv[3 + i]; }
        beam_static_ap<V>(g, m, Xu0, va, vp, CR(-1.0), false, U, 0, PacksLocal(), out);
        Cnt prim = take();
        // the six sinc1 packs of the passive nodes and of Δv′ are the same numbers in all lanes; the kernel evaluates each pack once per element (shuffles)
        show("static", lanes, sub(lanes, prim, 5), 6);
    }
    // ---------------------------------------------------------------- Newmark :iter (ND = 3): beam_cot_kernel + beam_kernel_sd<3,split>, 6 lanes of NumSD
    {
        fill(g, m, xu, xv, true);
        using N = NumSD; using TR = N::TR; using TU = N::TU; using TS = N::TS;
        const real_t a1 = 2. / 0.3, b1 = 4. / 0.09;
        take();
        for (int l = 0; l < 6; ++l) {
            TU Xu[3][6], U[3]; TR Xv[3][6]; TS R[12];
            for (int d = 0; d < 3; ++d) for (int i = 0; i < 6; ++i) {
                const real_t coef = d == 0 ? 1.0 : (d == 1 ? a1 : b1), amp = d == 0 ? 1.0 : 0.7;
                Xu[d][i].v = xu[i] * CR(amp); Xu[d][i].d1 = CR((i == l) ? coef : 0.);
                Xv[d][i].v = xv[i] * CR(amp); Xv[d][i].d0 = CR((i == l) ? coef : 0.);
            }
            for (int i = 0; i < 3; ++i) { U[i].v = CR(0.); U[i].d1 = CR(0.); }
            take();                                                 // (the amp products above are set-up, not kernel work)
            Vec3<TS> xb[NGP], vsmb;
            beam_dyn_cotangents<3, N>(g, m, Xu, Xv, false, U, xb, vsmb);
            beam_residual_cot<N, TS, true>(g, m, Xu[0], Xv[0], xb, vsmb, R);
            static Cnt acc; acc = plus(acc, take());
            if (l == 5) g_cnt = acc;
        }
        Cnt lanes = take();
        using NV = NumVal; using VV = SD<false, false>;
        VV Xu[3][6], U[3], Xv[3][6], R[12];
        for (int d = 0; d < 3; ++d) for (int i = 0; i < 6; ++i) { Xu[d][i].v = xu[i]; Xv[d][i].v = xv[i]; }
        for (int i = 0; i < 3; ++i) U[i].v = CR(0.);
        Vec3<VV> xb[NGP], vsmb;
        take();
        beam_dyn_cotangents<3, NV>(g, m, Xu, Xv, false, U, xb, vsmb);
        beam_residual_cot<NV, VV, true>(g, m, Xu[0], Xv[0], xb, vsmb, R);
        Cnt prim = take();
        show("newmark_iter", lanes, sub(lanes, prim, 5), 6);
    }
    // ---------------------------------------------------------------- DirectXUA{2,0,0}, Udof: cot (12 jet lanes) + b0 (6) + lin (12 X′/X″ lanes + 2 U lanes)
    {
        fill(g, m, xu, xv, true);
        using N = NumSD; using TR = N::TR; using TU = N::TU; using TS = N::TS; using VV = SD<false, false>;
        Cnt lanes, prim_jet, prim_0;
        // primal shares: jets forward (values), order-0 forward + reverse (values)
        {
            VV Xu[3][6], U[3], Xv[3][6], R[12];
            for (int d = 0; d < 3; ++d) for (int i = 0; i < 6; ++i) { Xu[d][i].v = xu[i]; Xv[d][i].v = xv[i]; }
            for (int i = 0; i < 3; ++i) U[i].v = CR(0.1);
            Vec3<VV> xb[NGP], vsmb;
            take();
            beam_dyn_cotangents<3, NumVal>(g, m, Xu, Xv, true, U, xb, vsmb);
            prim_jet = take();
            beam_residual_cot<NumVal, VV, true>(g, m, Xu[0], Xv[0], xb, vsmb, R);
            prim_0 = take();
        }
        Vec3<TS> xbk[2][6][NGP], vsk[2][6], xb2[6][NGP], vs2[6];
        for (int d = 0; d < 2; ++d) for (int l = 0; l < 6; ++l) {            // cot: seeds at X₀ (d = 0, also yields the X″ partials) or X′ (d = 1)
            TU Xu[3][6], U[3]; TR Xv[3][6];
            for (int k = 0; k < 3; ++k) for (int i = 0; i < 6; ++i) {
                Xu[k][i].v = xu[i]; Xu[k][i].d1 = CR((k == d && i == l) ? 1. : 0.);
                Xv[k][i].v = xv[i]; Xv[k][i].d0 = CR((k == d && i == l) ? 1. : 0.);
            }
            for (int i = 0; i < 3; ++i) { U[i].v = CR(0.1); U[i].d1 = CR(0.); }
            take();
            if (d == 0) beam_dyn_cotangents<3, N, true>(g, m, Xu, Xv, true, U, xbk[d][l], vsk[d][l], xb2[l], &vs2[l]);
            else beam_dyn_cotangents<3, N, false, true>(g, m, Xu, Xv, true, U, xbk[d][l], vsk[d][l]);          // X′ lanes: order-0 Taylor coefficients as plain values (beam_direct_cot_kernel<3,true>)
            lanes = plus(lanes, take());
        }
        for (int l = 0; l < 6; ++l) {                                        // b0: order-0 forward + reverse in SD arithmetic
            TU Xu0[6]; TR Xv0[6]; TS R[12];
            for (int i = 0; i < 6; ++i) { Xu0[i].v = xu[i]; Xu0[i].d1 = CR(i == l ? 1. : 0.); Xv0[i].v = xv[i]; Xv0[i].d0 = CR(i == l ? 1. : 0.); }
            take();
            beam_residual_cot<N, TS, true>(g, m, Xu0, Xv0, xbk[0][l], vsk[0][l], R);
            lanes = plus(lanes, take());
        }
        for (int q = 0; q < 14; ++q) {                                       // lin: 6 X′ lanes, 6 X″ lanes, 2 U lanes — forward in values, reverse on cotangent partials
            VV Xu0[6], Xv0[6]; TS R[12];
            for (int i = 0; i < 6; ++i) { Xu0[i].v = xu[i]; Xv0[i].v = xv[i]; }
            const Vec3<TS>* xb = q < 6 ? xbk[1][q] : (q < 12 ? xb2[q - 6] : xbk[1][0]);
            const Vec3<TS>& vs = q < 6 ? vsk[1][q] : (q < 12 ? vs2[q - 6] : vsk[1][0]);
            take();
            beam_residual_cot<NumVal, TS, true>(g, m, Xu0, Xv0, xb, vs, R);
            lanes = plus(lanes, take());
        }
        // algorithmic: jets' primal once (12 jet lanes repeat it), order-0 primal once (6 + 14 lanes repeat it)
        Cnt alg = sub(sub(lanes, prim_jet, 11), prim_0, 19);
        show("directxua_200_udof", lanes, alg, 32, true);
    }
    std::printf("}\n");
    return 0;
}
