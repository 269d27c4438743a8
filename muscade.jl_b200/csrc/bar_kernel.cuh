// K4/K5: Bar3D and SoilContact element kernels (sm_100a) — HBM-bound elements, one thread per element, dense Dual<NDIR>.
//   Bar3D       toolbox/BarElement.jl:136-202   (axial strain + inertia / added mass / drag at 4 Gauss points)
//   SoilContact toolbox/SoilContact.jl:10-20    (penalty springs/dampers below z₀)
#pragma once
#include "beam_kernel.cuh"

namespace mb {

struct BarMat { double EA, mu, w, Cat, Clt, Cqt, Can, Cln, Cqn; };      // AxisymmetricBarCrossSection, BarElement.jl:36-46
struct BarGroupDev {
    int64_t nele;
    const double* geo;        // [nele][8]  cₘ3 tgₘ3 L₀ Lₛ
    const BarMat* mats;
    const int32_t* mat_id;
    const int32_t* idxX;      // [nele][6]
    const int32_t* idxU;      // [nele][3] or nullptr
    double scaleX[6];
    int udof;
    double scaleU[3];
};

// residual(o::Bar3D,…) for one element. X[ider][6] carry the seeds; the Gauss point motion x = c + tg·ζ is linear in the nodal motion,
// so velocity/acceleration are ψ₁·u̇₁ + ψ₂·u̇₂ (the reference reaches the same through motion{P}/motion⁻¹, BarElement.jl:170,189).
template <int ND, class S> MB_HD void bar_residual(const double* geo8, const BarMat& m, const S (*X)[6], bool udof, const S* U0, double t, S* R) {
    const double L0 = geo8[6], Ls = geo8[7];
    Vec3<S> tg;
#pragma unroll
    for (int i = 0; i < 3; ++i) tg[i] = (geo8[3 + i] + X[0][3 + i]) - X[0][i];
    S L = mb_sqrt((tg[0] * tg[0] + tg[1] * tg[1]) + tg[2] * tg[2]);
    S iL = mb_rcp(L);
    Vec3<S> d0{tg[0] * iL, tg[1] * iL, tg[2] * iL};
    S fint = m.EA * (L * (1.0 / Ls) - 1.0);
    const double fwz = (fmin(t, -5.) + 10) / 5 * m.w;
    S z = Make<S>::c(0.);
#pragma unroll
    for (int j = 0; j < 6; ++j) R[j] = z;
    S sumw_fint = z;
#pragma unroll
    for (int g = 0; g < NGP; ++g) {
        const GpConst c = gp_const(g);
        const double p1 = -c.z + 0.5, p2 = c.z + 0.5, wg = c.w * L0;
        Vec3<S> fe{z, z, z};
        if (ND >= 2) {
            Vec3<S> v, a;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                v[i] = p1 * X[1][i] + p2 * X[1][3 + i];
                a[i] = (ND >= 3) ? (p1 * X[2][i] + p2 * X[2][3 + i]) : z;
            }
            S at = (a[0] * d0[0] + a[1] * d0[1]) + a[2] * d0[2];
            S vt = (v[0] * d0[0] + v[1] * d0[1]) + v[2] * d0[2];
            S an1 = a[1] - at * d0[1], an2 = a[2] - at * d0[2];
            S vn1 = v[1] - vt * d0[1], vn2 = v[2] - vt * d0[2];
            S fqt = m.Cqt * (vt * vt); if (value(vt) < 0) fqt = -fqt;
            S fq1 = m.Cqn * (vn1 * vn1); if (value(vn1) < 0) fq1 = -fq1;
            S fq2 = m.Cqn * (vn2 * vn2); if (value(vn2) < 0) fq2 = -fq2;
            fe[0] = (m.mu * a[0] + m.Cat * at) + (m.Clt * vt + fqt);
            fe[1] = (m.mu * a[1] + m.Can * an1) + (m.Cln * vn1 + fq1);
            fe[2] = (m.mu * a[2] + m.Can * an2) + (m.Cln * vn2 + fq2);
        }
        fe[2] = fe[2] + fwz;
        if (udof) for (int i = 0; i < 3; ++i) fe[i] = fe[i] - U0[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) { R[i] = R[i] + (wg * p1) * fe[i]; R[3 + i] = R[3 + i] + (wg * p2) * fe[i]; }
        sumw_fint = sumw_fint + wg * fint;
    }
    // fᵢ ∘₀ ε∂X₀ with ε∂X₀ = (−δ₀, δ₀)/L₀  (BarElement.jl:181)
    S k = sumw_fint * (1.0 / L0);
#pragma unroll
    for (int i = 0; i < 3; ++i) { S q = k * d0[i]; R[i] = R[i] - q; R[3 + i] = R[3 + i] + q; }
}

// One thread per element, dense Dual<6|7>.  A thread's 36 tangent entries are contiguous in Ke, so stored directly every warp store instruction would touch
// 32 different sectors (stride 288 bytes); the warp stages its 32 × 36 (+ 32 × 6 residual) values in shared memory and writes them out as contiguous 256-byte
// rows instead (2.43 → see profiles/r1_bar_soil_hbm.jsonl at 10 M elements).
constexpr int BAR_TS = 37;      // tile row stride in doubles (36 + 1: lanes hit different banks)
template <int ND, bool STEP>
__global__ void __launch_bounds__(128)
bar_kernel(BarGroupDev g, StateDev st, NewmarkDev nm, double t, double* __restrict__ Ke, double* __restrict__ Re, double* __restrict__ Rp,
           unsigned long long* nanflag, unsigned long long nanbase) {
    constexpr int W = 6 + (STEP ? 1 : 0);
    __shared__ double tileK[4][32 * BAR_TS];
    __shared__ double tileR[4][32 * 6];
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t e0 = e - lane;
    if (e0 >= g.nele) return;                                  // whole warp beyond the end
    const bool active = e < g.nele;
    using S = Dual<W>;
    bool bad = false;
    double rp[6] = {0., 0., 0., 0., 0., 0.};                   // :step only: −∂R/∂r, staged after the residual
    if (active) {
        double geo8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) geo8[k] = g.geo[e * 8 + k];
        const BarMat m = g.mats[g.mat_id ? g.mat_id[e] : 0];
        S X[3][6], U[3], R[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int32_t d = g.idxX[e * 6 + i];
            const double x0 = st.X0[d], x1 = (ND >= 2) ? st.X1[d] : 0., x2 = (ND >= 3) ? st.X2[d] : 0.;
            X[0][i] = Make<S>::c(x0); X[1][i] = Make<S>::c(x1); X[2][i] = Make<S>::c(x2);
            X[0][i].d[i] = g.scaleX[i]; X[1][i].d[i] = nm.a1 * g.scaleX[i]; X[2][i].d[i] = nm.b1 * g.scaleX[i];
            if (STEP) { X[1][i].d[W - 1] = nm.a2 * x1 + nm.a3 * x2; X[2][i].d[W - 1] = nm.b2 * x1 + nm.b3 * x2; }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) U[i] = Make<S>::c((g.udof && st.U0) ? st.U0[g.idxU[e * 3 + i]] : 0.);
        bar_residual<ND, S>(geo8, m, X, g.udof != 0, U, t, R);
        double* tk = tileK[warp] + lane * BAR_TS;
        double* tr = tileR[warp] + lane * 6;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const double s = g.scaleX[i];
            double v = R[i].v * s; bad |= (v != v); tr[i] = v;
#pragma unroll
            for (int j = 0; j < 6; ++j) { double k = R[i].d[j] * s; bad |= (k != k); tk[i + 6 * j] = k; }
            if (STEP) { double p = R[i].d[W - 1] * s; bad |= (p != p); rp[i] = p; }
        }
    }
    __syncwarp();
    const int cnt = (int)min((int64_t)32, g.nele - e0);
    {
        double* dst = Ke + e0 * 36; const double* src = tileK[warp];
        for (int q = lane; q < cnt * 36; q += 32) { const int el = q / 36; dst[q] = src[el * BAR_TS + (q - el * 36)]; }
        for (int q = lane; q < cnt * 6; q += 32) Re[e0 * 6 + q] = tileR[warp][q];
    }
    if (STEP) {
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 6; ++i) tileR[warp][lane * 6 + i] = rp[i];
        __syncwarp();
        for (int q = lane; q < cnt * 6; q += 32) Rp[e0 * 6 + q] = tileR[warp][q];
    }
    if (bad) atomicMin(nanflag, nanbase + (unsigned long long)e);
}

struct SoilGroupDev { int64_t nele; const double* par; const int32_t* idxX; double scaleX[3]; };   // par [nele][5] z₀ Kh Kv Ch Cv
template <int ND, bool STEP>
__global__ void __launch_bounds__(128)
soil_kernel(SoilGroupDev g, StateDev st, NewmarkDev nm, double* __restrict__ Ke, double* __restrict__ Re, double* __restrict__ Rp,
            unsigned long long* nanflag, unsigned long long nanbase) {
    // outputs staged per warp and written as contiguous rows, as in bar_kernel
    __shared__ double tileK[4][32 * 9];
    __shared__ double tileR[4][32 * 3 * 2];
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t e0 = e - lane;
    if (e0 >= g.nele) return;
    bool flag = false;
    if (e < g.nele) {
        const double z0 = g.par[e * 5], Kh = g.par[e * 5 + 1], Kv = g.par[e * 5 + 2], Ch = g.par[e * 5 + 3], Cv = g.par[e * 5 + 4];
        double x[3], xp[3], xpp[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int32_t d = g.idxX[e * 3 + i];
            x[i] = st.X0[d]; xp[i] = (ND >= 2) ? st.X1[d] : 0.; xpp[i] = (ND >= 3) ? st.X2[d] : 0.;
        }
        const bool contact = x[2] < z0;                                // if z < o.z₀  (SoilContact.jl:14); else R = SVector(0,0,0) and no partials
        const double K[3] = {Kh, Kh, Kv}, Cc[3] = {Ch, Ch, Cv};
        bool bad = (x[0] != x[0]) | (x[1] != x[1]) | (x[2] != x[2]);
        double* tk = tileK[warp] + lane * 9; double* tr = tileR[warp] + lane * 3;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double s = g.scaleX[i];
            const double r = contact ? (K[i] * (i == 2 ? x[2] - z0 : x[i]) + Cc[i] * xp[i]) : 0.;
            tr[i] = r * s; bad |= (r != r);
#pragma unroll
            for (int j = 0; j < 3; ++j) tk[i + 3 * j] = (contact && i == j) ? s * (K[i] + nm.a1 * Cc[i]) * s : 0.;
            if (STEP) tr[32 * 3 + i] = contact ? s * Cc[i] * (nm.a2 * xp[i] + nm.a3 * xpp[i]) : 0.;
        }
        flag = bad && contact;
    }
    __syncwarp();
    const int cnt = (int)min((int64_t)32, g.nele - e0);
    for (int q = lane; q < cnt * 9; q += 32) Ke[e0 * 9 + q] = tileK[warp][q];
    for (int q = lane; q < cnt * 3; q += 32) Re[e0 * 3 + q] = tileR[warp][q];
    if (STEP) for (int q = lane; q < cnt * 3; q += 32) Rp[e0 * 3 + q] = tileR[warp][32 * 3 + q];
    if (flag) atomicMin(nanflag, nanbase + (unsigned long long)e);
}

// ---------------------------------------------------------------------------------------------- DirectXUA first-order path (DirectXUA.jl:85-120)
// dR[e][p][i] = ∂R_i/∂seed_p, p over X₀(nx) X′(nx) X″(nx) U₀(nu) with seeds scale.X / scale.U; R[e][i] unscaled.
// Bar3D: lane d < ND carries the six partials of X_d, lane ND (Udof types) the three of U₀ — dense Dual<6>, one thread per (element, lane).
template <int ND>
__global__ void __launch_bounds__(128)
bar_direct_kernel(BarGroupDev g, DirectStateDev st, double t, double* __restrict__ dR, double* __restrict__ R, unsigned long long* nanflag, unsigned long long nanbase,
                  StepBatch sb) {
    batch_state(st, sb); t += blockIdx.y * sb.dt; dR += (int64_t)blockIdx.y * sb.sdR; R += (int64_t)blockIdx.y * sb.sR; nanbase += blockIdx.y * sb.snan;
    const int LPE = ND + (g.udof ? 1 : 0), NP = 6 * ND + (g.udof ? 3 : 0);
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = q / LPE;
    const int lane = (int)(q - e * LPE);
    if (e >= g.nele) return;
    using S = Dual<6>;
    double geo8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) geo8[k] = g.geo[e * 8 + k];
    const BarMat m = g.mats[g.mat_id ? g.mat_id[e] : 0];
    S X[3][6], U[3], Rv[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int32_t d = g.idxX[e * 6 + i];
#pragma unroll
        for (int k = 0; k < 3; ++k) { X[k][i] = Make<S>::c((k < ND) ? st.X[k][d] : 0.); if (k == lane) X[k][i].d[i] = g.scaleX[i]; }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { U[i] = Make<S>::c(g.udof ? st.U0[g.idxU[e * 3 + i]] : 0.); if (lane == ND) U[i].d[i] = g.scaleU[i]; }
    bar_residual<ND, S>(geo8, m, X, g.udof != 0, U, t, Rv);
    bool bad = false;
    double* out = dR + e * (int64_t)(6 * NP) + (int64_t)(6 * lane) * 6;
    const int ncol = (lane < ND) ? 6 : 3;
    for (int j = 0; j < ncol; ++j)
#pragma unroll
        for (int i = 0; i < 6; ++i) { const double v = Rv[i].d[j]; bad |= (v != v); out[6 * j + i] = v; }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) { const double v = Rv[i].v; bad |= (v != v); R[e * 6 + i] = v; }
    }
    if (bad) atomicMin(nanflag, nanbase + (unsigned long long)e);
}
// SoilContact declares no `no_second_order`, so in DirectXUA it takes the second-order path (DirectXUA.jl:152-171): L = Λ∘R with
// revariate{2}((;Λ,X,U),(;Λ=scale.Λ,X=scale.X,…)), scale.Λ = scale.X·Λscale (Assemble.jl:55).  R = K·(x − z₀e₃) + C·x′ below z₀, else 0 with no
// partials (SoilContact.jl:10-20), so in closed form:  ∂L/∂Λᵢ = Rᵢ·sΛᵢ → L1[Λ];  ∂L/∂X₀ᵢ = ΛᵢKᵢ·sXᵢ, ∂L/∂X′ᵢ = ΛᵢCᵢ·sXᵢ → L1[X][1], L1[X][2];
// ∂²L/∂Λᵢ∂X₀ᵢ = Kᵢ·sΛᵢ·sXᵢ, ∂²L/∂Λᵢ∂X′ᵢ = Cᵢ·sΛᵢ·sXᵢ → L2[Λ,X] and, transposed, L2[X,Λ];  L2[X,X] = L2[Λ,Λ] = 0 (R is linear).
// GX[(e·3+i)·ND+d] receives the L1[X][d+1] contribution of element dof i.
template <int ND>
__global__ void soil_direct_kernel(SoilGroupDev g, DirectStateDev st, const double* __restrict__ Lam, double lamscale, double* __restrict__ dR,
                                   double* __restrict__ R, double* __restrict__ GX, unsigned long long* nanflag, unsigned long long nanbase, StepBatch sb) {
    batch_state(st, sb); Lam += (int64_t)blockIdx.y * sb.sLam; dR += (int64_t)blockIdx.y * sb.sdR; R += (int64_t)blockIdx.y * sb.sR; GX += (int64_t)blockIdx.y * sb.sGX;
    nanbase += blockIdx.y * sb.snan;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= g.nele) return;
    const double z0 = g.par[e * 5], Kh = g.par[e * 5 + 1], Kv = g.par[e * 5 + 2], Ch = g.par[e * 5 + 3], Cv = g.par[e * 5 + 4];
    double x[3], xp[3], lam[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { const int32_t d = g.idxX[e * 3 + i]; x[i] = st.X[0][d]; xp[i] = (ND >= 2) ? st.X[1][d] : 0.; lam[i] = Lam[d]; }
    const bool contact = x[2] < z0;
    const double K[3] = {Kh, Kh, Kv}, Cc[3] = {Ch, Ch, Cv};
    bool bad = false;
    double* out = dR + e * (int64_t)(3 * 3 * ND);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double sL = g.scaleX[i] * lamscale;
        const double r = contact ? (K[i] * (i == 2 ? x[2] - z0 : x[i]) + Cc[i] * xp[i]) : 0.;
        R[e * 3 + i] = r * sL; bad |= (r != r);
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const double coef = (contact && d < 2) ? (d == 0 ? K[i] : Cc[i]) : 0.;
#pragma unroll
            for (int j = 0; j < 3; ++j) out[(3 * d + j) * 3 + i] = (i == j) ? coef * sL * g.scaleX[j] : 0.;
            const double gx = lam[i] * coef * g.scaleX[i];
            GX[(e * 3 + i) * ND + d] = gx; bad |= (gx != gx);
        }
    }
    if (bad && contact) atomicMin(nanflag, nanbase + (unsigned long long)e);
}

}  // namespace mb
