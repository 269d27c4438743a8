// Strain gauges on EulerBeam3D for the ElementCost accelerator (src/DirectXUA.jl:172-198; toolbox/StrainGaugeOnBeamElement.jl): requestables (εₐₓ,κ) with their partials and
// the quadratic cost's gradient / Gauss-Newton block.  Shared by the general DirectXUA form (mb_xua.cu) and the beam-specialised, windowed path (mb_direct.cu).
#pragma once
#include "mb_internal.h"

namespace {

// requestables (εₐₓ, ♢κ) of EulerBeam3D (toolbox/BeamElement.jl:151-174) with their partials ∂/∂X₀ (scaled): one lane per (element, element dof), one-direction duals
// through the forward kinematics only.  J[e][k][d], e4[e][k]  (k: εₐₓ, κ₁, κ₂, κ₃)
// Step batching (mb_direct.cu): grid row = time step; sX = stride of the state per step, sE = elements of the scratch arrays per step (0, 0 for one step).
__global__ void __launch_bounds__(128) beam_gauge_kernel(BeamGroupDev g, const double* __restrict__ X0, double* __restrict__ J, double* __restrict__ e4, int64_t sX = 0, int64_t sE = 0) {
    using N = NumDual<1>; using T = Dual<1>;
    X0 += (int64_t)blockIdx.y * sX; J += (int64_t)blockIdx.y * sE * 48; e4 += (int64_t)blockIdx.y * sE * 4;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t / 12; const int d = (int)(t - e * 12);
    if (e >= g.nele) return;
    BeamGeo geo; load_geo(g.geo + e * 16, geo);
    T x[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) { x[i].v = X0[g.idxX[e * 12 + i]]; x[i].d[0] = (i == d) ? g.scaleX[i] : 0.; }
    BeamFwd<N> f;
    beam_forward<N, false>(geo, Vec3<T>{x[0], x[1], x[2]}, Vec3<T>{x[3], x[4], x[5]}, Vec3<T>{x[6], x[7], x[8]}, Vec3<T>{x[9], x[10], x[11]}, f);
    const double k = 2. / geo.L;
    const T q[4] = {f.eps, f.vl[0] * k, f.vl[2] * k, -(f.vl[1] * k)};          // ♢κ (BeamElement.jl:164-166)
#pragma unroll
    for (int i = 0; i < 4; ++i) { J[(e * 4 + i) * 12 + d] = q[i].d[0]; if (d == 0) e4[e * 4 + i] = q[i].v; }
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
