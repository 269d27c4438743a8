// K1/K2: EulerBeam3D residual + tangent kernel (sm_100a). One lane per W seed directions of one element.
#pragma once
#include "beam_math.cuh"
#include <stdint.h>
#include <cuda_runtime.h>

namespace mb {

struct NewmarkDev { double a1, a2, a3, b1, b2, b3, dt; };

struct BeamGroupDev {
    int64_t nele;
    const double* geo;        // [nele][16]  cₘ3 rₘ9 tgₘ3 L   — one 128-byte line per element
    const BeamMat* mats;      // material table
    const int32_t* mat_id;    // [nele] or nullptr (single material)
    const int32_t* idxX;      // [nele][12] 0-based model X-dof numbers
    const int32_t* idxU;      // [nele][3] or nullptr
    double scaleX[12];
    double scaleU[3];
    int udof;
};
struct StateDev { const double *X0, *X1, *X2, *U0; };

// Seed directions (src/SweepX.jl:46-53,62-63,92): p<12 → δX_p (scaled); p==12 (STEP) → δr.
// Thread t ↔ (element t / LPE, lane t % LPE); lane l owns directions l·W … l·W+W−1.
template <int ND, int W, bool STEP>
__global__ void __launch_bounds__(128)
beam_kernel(BeamGroupDev g, StateDev st, NewmarkDev nm, double* __restrict__ Ke, double* __restrict__ Re, double* __restrict__ Rp,
            unsigned long long* nanflag, unsigned long long nanbase) {
    constexpr int NDIR = 12 + (STEP ? 1 : 0);
    constexpr int LPE = (NDIR + W - 1) / W;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t / LPE;
    const int lane = (int)(t - e * LPE);
    if (e >= g.nele) return;

    BeamGeo geo;
    {
        const double2* p = reinterpret_cast<const double2*>(g.geo + e * 16);
        double buf[16];
#pragma unroll
        for (int k = 0; k < 8; ++k) { double2 v = __ldg(p + k); buf[2 * k] = v.x; buf[2 * k + 1] = v.y; }
        geo.cm[0] = buf[0]; geo.cm[1] = buf[1]; geo.cm[2] = buf[2];
#pragma unroll
        for (int k = 0; k < 9; ++k) geo.rm.a[k] = buf[3 + k];
        geo.tgm[0] = buf[12]; geo.tgm[1] = buf[13]; geo.tgm[2] = buf[14];
        geo.L = buf[15];
    }
    const BeamMat m = g.mats[g.mat_id ? g.mat_id[e] : 0];

    using S = Dual<W>;
    S X[3][12], U[3], R[12];
    {
        const int32_t* ix = g.idxX + e * 12;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const int32_t d = __ldg(ix + i);
            const double x0 = st.X0[d];
            const double x1 = (ND >= 2) ? st.X1[d] : 0.;
            const double x2 = (ND >= 3) ? st.X2[d] : 0.;
            X[0][i].v = x0; X[1][i].v = x1; X[2][i].v = x2;
            const double ar = nm.a2 * x1 + nm.a3 * x2;      // a = a₂x′ + a₃x″
            const double br = nm.b2 * x1 + nm.b3 * x2;      // b = b₂x′ + b₃x″
#pragma unroll
            for (int k = 0; k < W; ++k) {
                const int p = lane * W + k;
                const double s = (p == i) ? g.scaleX[i] : 0.;
                X[0][i].d[k] = s;
                X[1][i].d[k] = (STEP && p == 12) ? ar : nm.a1 * s;
                X[2][i].d[k] = (STEP && p == 12) ? br : nm.b1 * s;
            }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            U[i] = Make<S>::c((g.udof && st.U0) ? st.U0[g.idxU[e * 3 + i]] : 0.);
        }
    }
    beam_residual<ND, W>(geo, m, X, g.udof != 0, U, R);

    bool bad = false;
    double* ke = Ke + e * 144;
#pragma unroll
    for (int k = 0; k < W; ++k) {
        const int p = lane * W + k;
        if (p < 12) {
#pragma unroll
            for (int i = 0; i < 12; i += 2) {
                double2 v; v.x = R[i].d[k] * g.scaleX[i]; v.y = R[i + 1].d[k] * g.scaleX[i + 1];
                bad |= (v.x != v.x) | (v.y != v.y);
                *reinterpret_cast<double2*>(ke + 12 * p + i) = v;
            }
        } else if (STEP && p == 12) {
#pragma unroll
            for (int i = 0; i < 12; ++i) { double v = R[i].d[k] * g.scaleX[i]; bad |= (v != v); Rp[e * 12 + i] = v; }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 12; ++i) { double v = R[i].v * g.scaleX[i]; bad |= (v != v); Re[e * 12 + i] = v; }
    }
    if (bad) atomicMin(nanflag, nanbase + (unsigned long long)e);
}


// host-side launcher, one translation unit per (ND,STEP) so that the instantiations compile in parallel
struct BeamLaunch {
    BeamGroupDev g; StateDev st; NewmarkDev nm;
    double *Ke, *Re, *Rp; unsigned long long* nanflag; unsigned long long nanbase; int W; cudaStream_t stream;
};
template <int ND, bool STEP> void launch_beam(const BeamLaunch& a);
#define MB_INSTANTIATE_BEAM(ND_, STEP_)                                                                                          \
    template <> void launch_beam<ND_, STEP_>(const BeamLaunch& a) {                                                              \
        constexpr int NDIR = 12 + (STEP_ ? 1 : 0);                                                                               \
        const int64_t nt = a.g.nele * NDIR;                                                                                      \
        beam_kernel<ND_, 1, STEP_><<<(unsigned)((nt + 127) / 128), 128, 0, a.stream>>>(a.g, a.st, a.nm, a.Ke, a.Re, a.Rp, a.nanflag, a.nanbase); \
    }

}  // namespace mb
