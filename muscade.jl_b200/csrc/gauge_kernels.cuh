// Strain gauges on EulerBeam3D for the ElementCost accelerator (src/DirectXUA.jl:172-198; toolbox/StrainGaugeOnBeamElement.jl): requestables (εₐₓ,κ) with their partials and
// the quadratic cost's gradient / Gauss-Newton block.  Shared by the general DirectXUA form (mb_xua.cu) and the beam-specialised, windowed path (mb_direct.cu).
#pragma once
#include "mb_internal.h"

namespace {

// requestables (εₐₓ, ♢κ) of EulerBeam3D (toolbox/BeamElement.jl:151-174) with their partials ∂/∂X₀ (scaled): six lanes per element through the forward kinematics only, lane l
// with the two seeds of the element kernels (slot 0 the rotation dof l, slot 1 the translation dof l, statically sparse duals: κ depends on rotations only).
// J[e][k][d], e4[e][k]  (k: εₐₓ, κ₁, κ₂, κ₃).  (A first version ran twelve one-direction lanes: 140 → 74 µs per step of 10⁵ elements, ncu.)
// Step batching (mb_direct.cu): grid row = time step; sX = stride of the state per step, sE = elements of the scratch arrays per step (0, 0 for one step).
__global__ void __launch_bounds__(128) beam_gauge_kernel(BeamGroupDev g, const double* __restrict__ X0, double* __restrict__ J, double* __restrict__ e4, int64_t sX = 0, int64_t sE = 0) {
    using N = NumSD; using TR = N::TR; using TU = N::TU; using TS = N::TS;
    X0 += (int64_t)blockIdx.y * sX; J += (int64_t)blockIdx.y * sE * 48; e4 += (int64_t)blockIdx.y * sE * 4;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t / 6; const int l = (int)(t - e * 6);
    if (e >= g.nele) return;
    BeamGeo geo; load_geo(g.geo + e * 16, geo);
    TU Xu[6]; TR Xv[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int iu = (i < 3) ? i : i + 3, iv = iu + 3;
        Xu[i].v = X0[g.idxX[e * 12 + iu]]; Xu[i].d1 = (i == l) ? g.scaleX[iu] : 0.;
        Xv[i].v = X0[g.idxX[e * 12 + iv]]; Xv[i].d0 = (i == l) ? g.scaleX[iv] : 0.;
    }
    BeamFwd<N> f;
    beam_forward<N, false>(geo, Vec3<TU>{Xu[0], Xu[1], Xu[2]}, Vec3<TR>{Xv[0], Xv[1], Xv[2]}, Vec3<TU>{Xu[3], Xu[4], Xu[5]}, Vec3<TR>{Xv[3], Xv[4], Xv[5]}, f);
    const double k = 2. / geo.L;
    const TR kap[3] = {f.vl[0] * k, f.vl[2] * k, -(f.vl[1] * k)};               // ♢κ (BeamElement.jl:164-166)
    const int cu = (l < 3) ? l : l + 3, cv = cu + 3;
    double* Je = J + e * 48;
    Je[cu] = sd1(f.eps); Je[cv] = sd0(f.eps);
#pragma unroll
    for (int i = 0; i < 3; ++i) { Je[(i + 1) * 12 + cu] = 0.; Je[(i + 1) * 12 + cv] = kap[i].d0; }
    if (l == 0) { e4[e * 4] = f.eps.v; for (int i = 0; i < 3; ++i) e4[e * 4 + 1 + i] = kap[i].v; }
}
// ElementCost accelerator for StrainGaugeOnEulerBeam3D (toolbox/StrainGaugeOnBeamElement.jl:70-76) under the quadratic cost Σ_g (ε_g − εm_g)²/(2σ²):
// ε_g = G[g]·(εₐₓ,κ); ∇cost = Jᵀ·Gᵀ·r/σ², ∇²cost = Jᵀ·GᵀG·J/σ² (chainrule of the second-order cost with the first-order eleres: to_order{2} adds no curvature, :190-196).
// One thread per (element, i): entry i of ∇cost → gX[e][12], row i of ∇²cost → HXX[e][12][12] (the whole X₀-X₀ block of the packet is the cost's: no Λ·∂²R/∂X²).
__global__ void gauge_cost_kernel(int64_t nele, int ng, const double* __restrict__ G, const double* __restrict__ epsm, bool per_element, double isig2,
                                  const double* __restrict__ J, const double* __restrict__ e4, double* __restrict__ gX, double* __restrict__ HXX, double* __restrict__ cost,
                                  int64_t sEps = 0, int64_t sE = 0) {
    { const int64_t b = blockIdx.y; epsm += b * sEps; J += b * sE * 48; e4 += b * sE * 4; gX += b * sE * 12; HXX += b * sE * 144; if (cost) cost += b * sE; }
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t / 12; const int i = (int)(t - e * 12);
    if (e >= nele) return;
    double q[4] = {0., 0., 0., 0.}, M[4][4] = {{0.}}, c = 0.;
    for (int a = 0; a < ng; ++a) {
        const double* Ga = G + a * 4;
        const double r = (((Ga[0] * e4[e * 4] + Ga[1] * e4[e * 4 + 1]) + Ga[2] * e4[e * 4 + 2]) + Ga[3] * e4[e * 4 + 3]) - epsm[(per_element ? e * ng : 0) + a];
        c += r * r;
        for (int k = 0; k < 4; ++k) { q[k] += Ga[k] * r; for (int l = 0; l < 4; ++l) M[k][l] += Ga[k] * Ga[l]; }
    }
    const double* Je = J + e * 48;
    double gi = 0., MJ[4];
    for (int k = 0; k < 4; ++k) { gi += Je[k * 12 + i] * q[k]; MJ[k] = ((M[k][0] * Je[i] + M[k][1] * Je[12 + i]) + M[k][2] * Je[24 + i]) + M[k][3] * Je[36 + i]; }
    gX[e * 12 + i] = gi * isig2;
    double* Hrow = HXX + (e * 12 + i) * 12;
    for (int j = 0; j < 12; ++j) Hrow[j] = (((Je[j] * MJ[0] + Je[12 + j] * MJ[1]) + Je[24 + j] * MJ[2]) + Je[36 + j] * MJ[3]) * isig2;
    if (i == 0 && cost) cost[e] = 0.5 * c * isig2;
}


}  // namespace
