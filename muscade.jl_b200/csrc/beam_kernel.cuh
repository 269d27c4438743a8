// K1/K2: EulerBeam3D residual + tangent kernel (sm_100a). One lane per W seed directions of one element.
#pragma once
#include "beam_math.cuh"
#include <stdint.h>
#include <cuda_runtime.h>
#include <algorithm>

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

#ifndef MB_MINB
#define MB_MINB 1
#endif
#ifndef MB_BLOCK
#define MB_BLOCK 128
#endif
#ifndef MB_SYM_MINB
#define MB_SYM_MINB 2     // resident CTAs per SM of the static symmetric-tangent kernel (2 × 128 threads × 255 registers)
#endif

// per-thread scratch in dynamic shared memory: slot k of thread t lives at base[k·blockDim + t] (conflict-free)
struct SmemScratch {
    static constexpr bool enabled = true;
    double* base; int stride;
    __device__ __forceinline__ void put(int k, double v) { base[k * stride] = v; }
    __device__ __forceinline__ double get(int k) const { return base[k * stride]; }
};
#ifndef MB_STASH
#define MB_STASH 0   // measured: parking r2/rd in shared memory made ptxas spill MORE (static 376→576 B, Newmark 3.1→5.3 KB); kept for experiments
#endif
constexpr int MB_STASH_SLOTS = 2 * 9 * 2;      // rₛ₂ and Rodrigues(Δvᵧ) as SD<1,0>

// two consecutive doubles: one 16-byte streaming store when the destination allows it.  Element types with an odd number of entries stored before
// the beams (SoilContact: 9 per element, host-evaluated DofLoad: 1) leave the beams' rows 8-byte aligned only; `al` is uniform over the launch.
__device__ __forceinline__ void store_pair_cs(double* p, double x, double y, bool al) {
    if (al) __stcs(reinterpret_cast<double2*>(p), make_double2(x, y));
    else { __stcs(p, x); __stcs(p + 1, y); }
}
__device__ __forceinline__ void store_pair(double* p, double x, double y, bool al) {
    if (al) *reinterpret_cast<double2*>(p) = make_double2(x, y);
    else { p[0] = x; p[1] = y; }
}
MB_HD void load_geo(const double* __restrict__ p16, BeamGeo& geo) {
    const double2* p = reinterpret_cast<const double2*>(p16);
    double buf[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#ifdef __CUDA_ARCH__
        double2 v = __ldg(p + k);
#else
        double2 v = p[k];
#endif
        buf[2 * k] = v.x; buf[2 * k + 1] = v.y;
    }
    geo.cm[0] = buf[0]; geo.cm[1] = buf[1]; geo.cm[2] = buf[2];
#pragma unroll
    for (int k = 0; k < 9; ++k) geo.rm.a[k] = buf[3 + k];
    geo.tgm[0] = buf[12]; geo.tgm[1] = buf[13]; geo.tgm[2] = buf[14];
    geo.L = buf[15];
}

// K1/K2 main kernel. Thread t ↔ (element t/6, lane t%6); lane l carries the seed directions
//   slot 0: rotation dof l  (element dofs r1,r2,r3 of node 1 then node 2)      slot 1: translation dof l (t1,t2,t3 of node 1, node 2)
// of δX (src/SweepX.jl:63,84,92): X₀+δX, X₁+a₁δX, X₂+b₁δX with δX_p = scale.X[p]·e_p, and writes the two columns of the element
// tangent they produce (rows scaled: Lλ .* scale.X, SweepX.jl:55,65). Lane 0 also writes the residual.
// lane seeds of δX (see beam_kernel_sd)
template <int ND> __device__ __forceinline__ void load_lane_state(const BeamGroupDev& g, const StateDev& st, const NewmarkDev& nm, int64_t e, int lane,
                                                                 NumSD::TU (*Xu)[6], NumSD::TR (*Xv)[6], NumSD::TU* U) {
    const int32_t* ix = g.idxX + e * 12;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int iu = (i < 3) ? i : i + 3, iv = iu + 3;            // element dof numbers of translation / rotation i
        const int32_t du = __ldg(ix + iu), dv = __ldg(ix + iv);
        const double su = (i == lane) ? g.scaleX[iu] : 0., sv = (i == lane) ? g.scaleX[iv] : 0.;
        Xu[0][i].v = st.X0[du]; Xu[0][i].d1 = su;
        Xv[0][i].v = st.X0[dv]; Xv[0][i].d0 = sv;
        Xu[1][i].v = (ND >= 2) ? st.X1[du] : 0.; Xu[1][i].d1 = nm.a1 * su;
        Xv[1][i].v = (ND >= 2) ? st.X1[dv] : 0.; Xv[1][i].d0 = nm.a1 * sv;
        Xu[2][i].v = (ND >= 3) ? st.X2[du] : 0.; Xu[2][i].d1 = nm.b1 * su;
        Xv[2][i].v = (ND >= 3) ? st.X2[dv] : 0.; Xv[2][i].d0 = nm.b1 * sv;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { U[i].v = (g.udof && st.U0) ? st.U0[g.idxU[e * 3 + i]] : 0.; U[i].d1 = 0.; }
}
constexpr int MB_NCOT = (NGP * 3 + 3) * 3;      // x̄_gp (4×3) and v̄ₛₘ (3) as SD<1,1>: 45 doubles per lane
// workspace tile of one warp: [45][32] doubles — every store instruction is one contiguous 256-byte line, the tile is 11.5 KB; streaming
// accesses (written once, read once) so that L2 keeps the spill lines of the resident threads
__device__ __forceinline__ void store_cot(double* __restrict__ Wc, int64_t t, const Vec3<NumSD::TS>* xb, const Vec3<NumSD::TS>& vsmb) {
    double* w = Wc + (t >> 5) * (int64_t)(MB_NCOT * 32) + (t & 31);
#pragma unroll
    for (int gp = 0; gp < NGP; ++gp)
#pragma unroll
        for (int i = 0; i < 3; ++i) { const int k = (gp * 3 + i) * 3; __stcs(w + (k) * 32, xb[gp][i].v); __stcs(w + (k + 1) * 32, xb[gp][i].d0); __stcs(w + (k + 2) * 32, xb[gp][i].d1); }
#pragma unroll
    for (int i = 0; i < 3; ++i) { const int k = (NGP * 3 + i) * 3; __stcs(w + (k) * 32, vsmb[i].v); __stcs(w + (k + 1) * 32, vsmb[i].d0); __stcs(w + (k + 2) * 32, vsmb[i].d1); }
}
__device__ __forceinline__ void load_cot(const double* __restrict__ Wc, int64_t t, Vec3<NumSD::TS>* xb, Vec3<NumSD::TS>& vsmb) {
    const double* w = Wc + (t >> 5) * (int64_t)(MB_NCOT * 32) + (t & 31);
#pragma unroll
    for (int gp = 0; gp < NGP; ++gp)
#pragma unroll
        for (int i = 0; i < 3; ++i) { const int k = (gp * 3 + i) * 3; xb[gp][i].v = __ldcs(w + (k) * 32); xb[gp][i].d0 = __ldcs(w + (k + 1) * 32); xb[gp][i].d1 = __ldcs(w + (k + 2) * 32); }
#pragma unroll
    for (int i = 0; i < 3; ++i) { const int k = (NGP * 3 + i) * 3; vsmb[i].v = __ldcs(w + (k) * 32); vsmb[i].d0 = __ldcs(w + (k + 1) * 32); vsmb[i].d1 = __ldcs(w + (k + 2) * 32); }
}
// K2 phase A (ND ≥ 2): time-jet forward → cotangent workspace Wc[k][thread] (k-major: coalesced)
template <int ND>
__global__ void __launch_bounds__(MB_BLOCK, MB_MINB)
beam_cot_kernel(BeamGroupDev g, StateDev st, NewmarkDev nm, double* __restrict__ Wc, int64_t nthreads) {
    using N = NumSD; using TR = N::TR; using TU = N::TU; using TS = N::TS;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t / 6;
    const int lane = (int)(t - e * 6);
    if (e >= g.nele) return;
    BeamGeo geo;
    load_geo(g.geo + e * 16, geo);
    const BeamMat m = g.mats[g.mat_id ? g.mat_id[e] : 0];
    TU Xu[3][6], U[3]; TR Xv[3][6];
    load_lane_state<ND>(g, st, nm, e, lane, Xu, Xv, U);
    Vec3<TS> xb[NGP], vsmb;
    beam_dyn_cotangents<ND, N>(geo, m, Xu, Xv, g.udof != 0, U, xb, vsmb, nullptr, nullptr, lane < 3);
    store_cot(Wc, t, xb, vsmb);
}

template <int ND, bool SPLIT>
__global__ void __launch_bounds__(MB_BLOCK, MB_MINB)
beam_kernel_sd(BeamGroupDev g, StateDev st, NewmarkDev nm, double* __restrict__ Ke, double* __restrict__ Re,
               unsigned long long* nanflag, unsigned long long nanbase, const double* __restrict__ Wc, int64_t nthreads) {
    using N = NumSD; using TR = N::TR; using TU = N::TU; using TS = N::TS;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t / 6;
    const int lane = (int)(t - e * 6);
    if (e >= g.nele) return;
    BeamGeo geo;
    load_geo(g.geo + e * 16, geo);
    const BeamMat m = g.mats[g.mat_id ? g.mat_id[e] : 0];
    TU Xu[3][6], U[3]; TR Xv[3][6]; TS R[12];
    load_lane_state<ND>(g, st, nm, e, lane, Xu, Xv, U);
    if (SPLIT) {
        Vec3<TS> xb[NGP], vsmb;
        load_cot(Wc, t, xb, vsmb);
        beam_residual_cot<N, TS, (ND >= 3)>(geo, m, Xu[0], Xv[0], xb, vsmb, R, lane < 3);      // v̄ₛₘ ≠ 0 only with accelerations
    } else {
        beam_residual_n<ND, N>(geo, m, Xu, Xv, g.udof != 0, U, R);
    }

    bool bad = false;
    const int cu = (lane < 3) ? lane : lane + 3, cv = cu + 3;          // tangent columns of this lane
    double* ke = Ke + e * 144;
    const bool al = (reinterpret_cast<uintptr_t>(ke) & 15) == 0;
#pragma unroll
    for (int i = 0; i < 12; i += 2) {
        double2 a, b;
        a.x = R[i].d1 * g.scaleX[i]; a.y = R[i + 1].d1 * g.scaleX[i + 1];
        b.x = R[i].d0 * g.scaleX[i]; b.y = R[i + 1].d0 * g.scaleX[i + 1];
        bad |= (a.x != a.x) | (a.y != a.y) | (b.x != b.x) | (b.y != b.y);
        store_pair_cs(ke + 12 * cu + i, a.x, a.y, al);                 // streaming: written once, read once by K6 — keep L2 for the spill lines
        store_pair_cs(ke + 12 * cv + i, b.x, b.y, al);
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 12; ++i) { double v = R[i].v * g.scaleX[i]; bad |= (v != v); Re[e * 12 + i] = v; }
    }
    if (bad) atomicMin(nanflag, nanbase + (unsigned long long)e);
}

// K1, statics (ND = 1): symmetric-tangent kernel (beam_static_sym, beam_math.cuh).  Thread t ↔ (element t/6, lane t%6); lane l sweeps the ROTATION
// dof l only (everything in SD<1,0>), writes that column of Ke, the transposed copies into the translation columns and six closed-form
// translation×translation entries.  Lane 0 also writes the residual.  −13 % FP64 instructions (16 348 → 14 274 per element) and a third of the
// spill traffic compared with beam_kernel_sd<1,false>; 19.8 → 17.3 ms at 10 M elements with the paired 16-byte stores.
// Measured and dropped (B200, 10 M elements): 3 / 4 CTAs per SM (168 / 128 registers: 19.6 / 23.8 ms, the spills cost more than the extra warps
// hide); a persistent CTA working on 21-element tiles with cp.async input staging and a shared-memory transpose for coalesced stores (21.8 ms:
// with two warps per scheduler every __syncthreads idles the FP64 pipe); a persistent, barrier-free variant with a private cp.async prefetch of the
// next element's dof indices / state / geometry (21.4 ms: long_scoreboard 1.49 → 1.00 per issue as intended, but no_instruction 0.37 → 1.12 —
// the 76 KB loop body thrashes the 32 KB instruction cache when every warp of the SM loops over it instead of streaming through it once);
// CTA shapes 64×4 / 256×1 / 96×2 / 96×3 / 288×1 (6.90 / 7.39 / 7.57 / 7.72 / 8.24 ms against 6.92 ms at 4 M elements); pulling the inputs of the element
// one grid wave ahead into L2 (prefetch.global.L2 of geometry / index lines at entry, of the far state entries at exit: 6.96 ms, no gain);
// persistent CTAs again, this time with ONE __syncthreads per tile: as a plain loop +3.7 %, with a two-stage cp.async prefetch of the next tiles' dof indices,
// geometry and gathered state (one shared-memory copy per element) +10 % — long_scoreboard 0.98 → 0.77 per issue as intended, but no_instruction 0.37 → 1.10 again:
// the back edge over the 76 KB body, not warp drift, is what the instruction fetch does not cope with;
// streaming CTAs of 21 whole elements that stage Ke/Re in shared memory and write contiguous rows after ONE barrier (what pays for the one-thread-per-element
// bar and soil kernels): 7.65 ms against 6.93 ms — the lanes' 16-byte column stores already merge in L2 and the barrier costs more than the LSU saves.
template <int MINB>          // resident CTAs per SM the register allocation aims at (template also so that only the ND = 1 translation unit compiles it)
__global__ void __launch_bounds__(MB_BLOCK, MINB)
beam_static_sym_kernel(BeamGroupDev g, StateDev st, double* __restrict__ Ke, double* __restrict__ Re, unsigned long long* nanflag, unsigned long long nanbase) {
    using V = SD<false, false>; using S = SD<true, false>;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t / 6;
    const int lane = (int)(t - e * 6);
    if (e >= g.nele) return;
    BeamGeo geo;
    load_geo(g.geo + e * 16, geo);
    const BeamMat m = g.mats[g.mat_id ? g.mat_id[e] : 0];
    V Xu[6], U[3]; S Xv[6], R[12];
    const int32_t* ix = g.idxX + e * 12;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int iu = (i < 3) ? i : i + 3, iv = iu + 3;
        Xu[i].v = st.X0[__ldg(ix + iu)];
        Xv[i].v = st.X0[__ldg(ix + iv)]; Xv[i].d0 = (i == lane) ? g.scaleX[iv] : 0.;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) U[i].v = (g.udof && st.U0) ? st.U0[g.idxU[e * 3 + i]] : 0.;
    double Gc[3];
    beam_static_sym(geo, m, Xu, Xv, g.udof != 0, U, lane % 3, R, Gc);
    double* ke = Ke + e * 144;
    const bool al = (reinterpret_cast<uintptr_t>(ke) & 15) == 0;
    bool bad = beam_static_sym_store(lane, R, Gc, g.scaleX, [&](int k, double v) { __stcs(ke + k, v); },
                                     [&](int k, double v0, double v1) { store_pair_cs(ke + k, v0, v1, al); });
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 12; ++i) { double v = R[i].v * g.scaleX[i]; bad |= (v != v); Re[e * 12 + i] = v; }
    }
    if (bad) atomicMin(nanflag, nanbase + (unsigned long long)e);
}

// K1, statics, active/passive formulation (beam_static_ap, beam_math.cuh).  Same lane ↔ rotation-dof mapping and same stores as beam_static_sym_kernel.
// SHARE: warps are cut into groups of 6 lanes = whole elements (30 working lanes, lanes 30/31 idle), and the sinc1 packs of the three Rodrigues maps —
// pure value computations that every lane of the element would otherwise repeat (sincos, reciprocals: a third of the primal FP64 work and most of its
// integer/branch instructions) — are evaluated once per element: lane l%6 == 0 takes (θ), lane 1 takes (θ/2) of whatever Rodrigues map the element is at,
// and the 4+4 doubles go round by warp shuffles.
#ifndef MB_AP_MINB
#define MB_AP_MINB 2
#endif
struct PacksShuffle {
    int sub;                 // lane within the element group (0..5); lanes 0-2: active node 1, lanes 3-5: active node 2
    int base;                // first warp lane of the group
    mutable RodPacks other;  // packs of the passive node, fetched together with the active ones
    __device__ __forceinline__ void operator()(int which, double th, RodPacks& pk) const {
        if (which == 1) { pk = other; return; }
        // which == 0: lanes 0,1 hold θ = |v₁| and take its packs at θ, θ/2; lanes 3,4 the same for |v₂|.  which == 2: |Δv′| is the same in all six lanes.
        const bool work = (which == 0) ? (sub != 2 && sub != 5) : (sub < 2);
        const bool half = (sub == 1) || (sub == 4);
        double S[4];
        if (work) sinc_pack(half ? 0.5 * th : th, S); else { S[0] = S[1] = S[2] = S[3] = 0.; }
        const int own = (which == 0 && sub >= 3) ? base + 3 : base, oth = (sub >= 3) ? base : base + 3;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            pk.Sth[k] = __shfl_sync(0xffffffffu, S[k], own);
            pk.Sh[k] = __shfl_sync(0xffffffffu, S[k], own + 1);
            if (which == 0) { other.Sth[k] = __shfl_sync(0xffffffffu, S[k], oth); other.Sh[k] = __shfl_sync(0xffffffffu, S[k], oth + 1); }
        }
    }
};
// Stores of one lane (rotation dof l of the element, tangent column cv) as the sweep produces them: column cv of the scaled element tangent
// Ke[i + 12j] = scale_i·∂R_i/∂X_j·scale_j (the seed carries scale_cv), the six translation rows of that column transposed into row cv of the translation
// columns (symmetric tangent), the translation rows of translation column cu = cv − 3 (±G/4), and, from lane 0, the residual.  Same values and
// positions as beam_static_sym_store; rows of the active / passive node are addressed, not selected.
// FUSE: `ke` is the element's 144 slots of the warp's SHARED-MEMORY tile (plain stores, same positions); the warp's epilogue sums them into nzval (kernels.cuh,
// "fused epilogue") and copies to Ke only what other warps contribute to.
template <bool FUSE> struct ApStoreT {
    double* ke; double* re; const double* sc; int lane; bool live, al, bad;
    __device__ __forceinline__ void put(int k, double v) { bad |= (v != v); if (live) { if constexpr (FUSE) ke[k] = v; else __stcs(ke + k, v); } }
    __device__ __forceinline__ void put2(int k, double v0, double v1) {
        bad |= (v0 != v0) | (v1 != v1);
        if (live) { if constexpr (FUSE) store_pair(ke + k, v0, v1, true); else store_pair_cs(ke + k, v0, v1, al); }
    }
    __device__ __forceinline__ void tt(const double* Gc) {
        const bool n1 = lane < 3;
        const int cu = n1 ? lane : lane + 3;
        const double s4 = (n1 ? 0.25 : -0.25) * sc[cu];
        const double v0 = Gc[0] * s4, v1 = Gc[1] * s4, v2 = Gc[2] * s4;
        put2(12 * cu, v0 * sc[0], v1 * sc[1]); put(12 * cu + 2, v2 * sc[2]);
        put2(12 * cu + 6, -v0 * sc[6], -v1 * sc[7]); put(12 * cu + 8, -v2 * sc[8]);
    }
    __device__ __forceinline__ void trans(const SD<true, false>* Ru1, const SD<true, false>* Ru2) {
        const int cv = ((lane < 3) ? lane : lane + 3) + 3;
        double a[3], b[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { a[i] = Ru1[i].d0 * sc[i]; b[i] = Ru2[i].d0 * sc[6 + i]; }
        put2(12 * cv, a[0], a[1]); put(12 * cv + 2, a[2]);
        put2(12 * cv + 6, b[0], b[1]); put(12 * cv + 8, b[2]);
#pragma unroll
        for (int i = 0; i < 3; ++i) { put(12 * i + cv, a[i]); put(12 * (6 + i) + cv, b[i]); }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 3; ++i) { const double x = Ru1[i].v * sc[i], y = Ru2[i].v * sc[6 + i]; bad |= (x != x) | (y != y); if (live) { re[i] = x; re[6 + i] = y; } }
        }
    }
    __device__ __forceinline__ void rot(const SD<true, false>* Rva, const SD<true, false>* Rvp) {
        const bool n1 = lane < 3;
        const int cv = (n1 ? lane : lane + 3) + 3, ra0 = n1 ? 3 : 9, rp0 = 12 - ra0;
        put(12 * cv + ra0, Rva[0].d0 * sc[ra0]); put2(12 * cv + ra0 + 1, Rva[1].d0 * sc[ra0 + 1], Rva[2].d0 * sc[ra0 + 2]);
        put(12 * cv + rp0, Rvp[0].d0 * sc[rp0]); put2(12 * cv + rp0 + 1, Rvp[1].d0 * sc[rp0 + 1], Rvp[2].d0 * sc[rp0 + 2]);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 3; ++i) { const double x = Rva[i].v * sc[3 + i], y = Rvp[i].v * sc[9 + i]; bad |= (x != x) | (y != y); if (live) { re[3 + i] = x; re[9 + i] = y; } }
        }
    }
};
// Fused epilogue (FUSE): a warp holds FIVE whole consecutive elements.  Their 720 tangent entries go to a shared-memory tile instead of Ke; after the sweep the warp walks
// its range of non-zeros [wbase, wbase + wlen) with its PATTERN (kernels.cuh, "fused epilogue": per non-zero the one or two tile entries that make it up, built at prepare
// from the segmented reduction's maps): such a non-zero is (0 + first) + second — the bit pattern of gather_one — and written to nzval, final; the entries of the other
// non-zeros (interface nodes shared with another warp, another group or type) are copied to Ke for gather_list_kernel.  Header and pattern arrive by cp.async while the sweep
// runs.  On a chain numbered along its axis 9/10 of the entries never reach HBM as Ke.
constexpr int MB_EPW = 5;                 // elements per warp
constexpr int MB_TILE = MB_EPW * 144;     // tile entries per warp (5.6 KB)
constexpr int MB_FUSE_RANGE = 768;        // non-zeros a warp walks at most (a chain numbered along its axis needs 648); what lies beyond stays with the segmented reduction
constexpr int MB_PAT_WORDS = MB_FUSE_RANGE + MB_TILE / 2;      // largest pattern, 32-bit words (multiple of 4); MB_FUSE_RANGE is a multiple of 256
constexpr uint32_t MB_PAT_NOTMINE = (uint32_t)(MB_TILE * 8) * 0x10001u;      // pattern word of a non-zero that is not this warp's: both byte offsets at the zero slot
struct FuseDev {
    const int4* whdr; const uint32_t* cpat; double* nzval;      // per warp of the group (wbase, wlen | nun << 16, pattern offset / 4 words, 0); the patterns; the CSC values
    int64_t warp0;                                              // first warp of this launch within the group
};
struct __align__(16) FuseSmem { double tile[MB_TILE + 2]; uint32_t pat[MB_PAT_WORDS]; int4 hdr; };      // tile[MB_TILE] = 0: the absent second contributor
template <bool SHARE, bool FUSE>
__device__ __forceinline__ void beam_static_ap_body(const BeamGroupDev& g, const StateDev& st, double* __restrict__ Ke, double* __restrict__ Re, unsigned long long* nanflag, unsigned long long nanbase, const FuseDev& fz) {
    static_assert(!FUSE || SHARE, "the fused epilogue needs whole elements per warp");
    using V = SD<false, false>;
    __shared__ FuseSmem fsm[FUSE ? MB_BLOCK / 32 : 1];
    int64_t e; int lane; bool live = true;
    int sub = 0, base = 0;
    const int wl = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (SHARE) {
        sub = wl % 6; base = wl - sub;
        e = warp * 5 + wl / 6; lane = sub;
        if (wl >= 30) { e = warp * 5 + 4; live = false; base = 24; }          // idle lanes shadow the last element of the warp (no stores)
        if (e >= g.nele) { e = g.nele - 1; live = false; }
    } else {
        const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        e = t / 6; lane = (int)(t - e * 6);
        if (e >= g.nele) return;
    }
    if constexpr (FUSE) {
        if (warp * 5 < g.nele) {
            FuseSmem& F = fsm[threadIdx.x >> 5];
            const int4 hd = __ldg(fz.whdr + fz.warp0 + warp);
            if (wl == 0) { F.hdr = hd; F.tile[MB_TILE] = 0.; }
            const int n4 = ((((hd.y & 0xFFFF) + 255) & ~255) + ((hd.y >> 16) + 1) / 2 + 3) / 4;
            const uint32_t* gp = fz.cpat + (int64_t)hd.z * 4;
            const unsigned sp = (unsigned)__cvta_generic_to_shared(F.pat);
            for (int i = wl; i < n4; i += 32) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sp + 16u * i), "l"(gp + 4 * i));
            asm volatile("cp.async.commit_group;");
        }
    }
    BeamGeo geo;
    load_geo(g.geo + e * 16, geo);
    const BeamMat m = g.mats[g.mat_id ? g.mat_id[e] : 0];
    double xu[6], xv[6]; V U[3];
    const int32_t* ix = g.idxX + e * 12;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int iu = (i < 3) ? i : i + 3;
        xu[i] = st.X0[__ldg(ix + iu)];
        xv[i] = st.X0[__ldg(ix + iu + 3)];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) U[i].v = (g.udof && st.U0) ? st.U0[g.idxU[e * 3 + i]] : 0.;
    const int cv = ((lane < 3) ? lane : lane + 3) + 3;
    double* ke = FUSE ? fsm[FUSE ? (threadIdx.x >> 5) : 0].tile + (e - warp * 5) * 144 : Ke + e * 144;
    ApStoreT<FUSE> out;
    out.ke = ke; out.re = Re + e * 12; out.sc = g.scaleX; out.lane = lane; out.live = live; out.al = (reinterpret_cast<uintptr_t>(ke) & 15) == 0; out.bad = false;
    if (SHARE) { PacksShuffle ps; ps.sub = sub; ps.base = base; beam_static_ap_lane(geo, m, xu, xv, g.scaleX[cv], lane, g.udof != 0, U, ps, out); }
    else beam_static_ap_lane(geo, m, xu, xv, g.scaleX[cv], lane, g.udof != 0, U, PacksLocal(), out);
    if (out.bad && live) atomicMin(nanflag, nanbase + (unsigned long long)e);
    if constexpr (FUSE) {
        // everything the epilogue needs is re-derived here (volatile reads of the special registers: nothing of it stays live across the sweep)
        unsigned bx, tx;
        asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bx));
        asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tx));
        const int el = (int)(tx & 31u);
        const int64_t wk = ((int64_t)bx * MB_BLOCK + tx) >> 5;
        if (wk * 5 >= g.nele) return;
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        const FuseSmem& F = fsm[tx >> 5];
        const int4 hd = F.hdr;
        const int tlen = hd.y & 0xFFFF, nun = hd.y >> 16;
        double* nz = fz.nzval + hd.x;
        // branch-free, eight non-zeros per lane at a time: all pattern words, then all tile entries, then the sums — the loads of a chunk are in flight together
        // (two warps per scheduler do not hide a dependent LDS → LDS → DADD → STG chain per non-zero).  An absent second contributor points at tile[MB_TILE] = 0:
        // 0 + first is never −0, so adding +0 leaves its bits alone.
        const char* tb = reinterpret_cast<const char*>(F.tile);
        for (int s0 = 0; s0 < tlen; s0 += 256) {               // the pattern is padded to whole chunks with MB_PAT_NOTMINE: no bounds checks
            uint32_t d[8]; double a[8], b[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = F.pat[s0 + 32 * j + el];
#pragma unroll
            for (int j = 0; j < 8; ++j) { a[j] = *reinterpret_cast<const double*>(tb + (d[j] & 0xFFFFu)); b[j] = *reinterpret_cast<const double*>(tb + (d[j] >> 16)); }
#pragma unroll
            for (int j = 0; j < 8; ++j) { const double acc = (0. + a[j]) + b[j]; if (d[j] != MB_PAT_NOTMINE) nz[s0 + 32 * j + el] = acc; }
        }
        const uint16_t* ul = reinterpret_cast<const uint16_t*>(F.pat + ((tlen + 255) & ~255));
        double* kw = Ke + wk * (int64_t)MB_TILE;
        for (int i0 = 0; i0 < nun; i0 += 128) {
            int q[4]; double v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const int i = i0 + 32 * j + el; q[j] = (i < nun) ? (int)ul[i] : -1; }
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = F.tile[max(q[j], 0)];
#pragma unroll
            for (int j = 0; j < 4; ++j) if (q[j] >= 0) __stcs(kw + q[j], v[j]);
        }
    }
}

template <int MINB, bool SHARE, bool FUSE = false>
__global__ void __launch_bounds__(MB_BLOCK, MINB)
beam_static_ap_kernel(BeamGroupDev g, StateDev st, double* __restrict__ Ke, double* __restrict__ Re, unsigned long long* nanflag, unsigned long long nanbase, FuseDev fz) {
    beam_static_ap_body<SHARE, FUSE>(g, st, Ke, Re, nanflag, nanbase, fz);
}
// :step mission only (src/SweepX.jl:46-57,69-78): the extra seed direction δr carries the Newmark predictor
//   vx′ = x′ + a₁δX + a·δr,  vx″ = x″ + b₁δX + b·δr,  a = a₂x′+a₃x″,  b = b₂x′+b₃x″ ;   Rp = ∂(Lλ)/∂r is subtracted from the rhs.
// One thread per element, dense one-direction dual.
// δr moves x′ and x″ only, and R is linear in the external-load cotangents: ∂R/∂r = J(X₀)ᵀ·∂c/∂r.  So: (A) time-jets with ONE dense direction
// (every number type carries slot 0) → ∂c/∂r, 15 doubles per element in Wd[k][e]; (B) order-0 forward in plain values and the reverse sweep
// on ∂c/∂r.  Two launches: fused they are 155 KB of SASS, beyond the instruction cache.
template <int ND>
__global__ void __launch_bounds__(MB_BLOCK, MB_MINB)
beam_dr_cot_kernel(BeamGroupDev g, StateDev st, NewmarkDev nm, double* __restrict__ Wd) {
    using N1 = Num<SD<true, false>, SD<true, false>, SD<true, false>>;
    using S = SD<true, false>;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= g.nele) return;
    BeamGeo geo;
    load_geo(g.geo + e * 16, geo);
    const BeamMat m = g.mats[g.mat_id ? g.mat_id[e] : 0];
    S Xu[3][6], Xv[3][6], U[3];
    const int32_t* ix = g.idxX + e * 12;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int iu = (i < 3) ? i : i + 3, iv = iu + 3;
        const int32_t du = __ldg(ix + iu), dv = __ldg(ix + iv);
        const double u0 = st.X0[du], u1 = (ND >= 2) ? st.X1[du] : 0., u2 = (ND >= 3) ? st.X2[du] : 0.;
        const double v0 = st.X0[dv], v1 = (ND >= 2) ? st.X1[dv] : 0., v2 = (ND >= 3) ? st.X2[dv] : 0.;
        Xu[0][i].v = u0; Xu[0][i].d0 = 0.; Xu[1][i].v = u1; Xu[1][i].d0 = nm.a2 * u1 + nm.a3 * u2; Xu[2][i].v = u2; Xu[2][i].d0 = nm.b2 * u1 + nm.b3 * u2;
        Xv[0][i].v = v0; Xv[0][i].d0 = 0.; Xv[1][i].v = v1; Xv[1][i].d0 = nm.a2 * v1 + nm.a3 * v2; Xv[2][i].v = v2; Xv[2][i].d0 = nm.b2 * v1 + nm.b3 * v2;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { U[i].v = (g.udof && st.U0) ? st.U0[g.idxU[e * 3 + i]] : 0.; U[i].d0 = 0.; }
    Vec3<S> xb[NGP], vsmb;
    beam_dyn_cotangents<(ND >= 2 ? ND : 2), N1>(geo, m, Xu, Xv, g.udof != 0, U, xb, vsmb);
    double* w = Wd + e;
#pragma unroll
    for (int gp = 0; gp < NGP; ++gp)
#pragma unroll
        for (int i = 0; i < 3; ++i) w[(gp * 3 + i) * g.nele] = xb[gp][i].d0;
#pragma unroll
    for (int i = 0; i < 3; ++i) w[(NGP * 3 + i) * g.nele] = vsmb[i].d0;
}
static __global__ void __launch_bounds__(MB_BLOCK, MB_MINB)
beam_dr_lin_kernel(BeamGroupDev g, StateDev st, const double* __restrict__ Wd, double* __restrict__ Rp, unsigned long long* nanflag, unsigned long long nanbase) {
    using S = SD<true, false>; using V = SD<false, false>;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= g.nele) return;
    BeamGeo geo;
    load_geo(g.geo + e * 16, geo);
    const BeamMat m = g.mats[g.mat_id ? g.mat_id[e] : 0];
    V Xu0[6], Xv0[6]; S R[12];
    const int32_t* ix = g.idxX + e * 12;
#pragma unroll
    for (int i = 0; i < 6; ++i) { const int iu = (i < 3) ? i : i + 3; Xu0[i].v = st.X0[__ldg(ix + iu)]; Xv0[i].v = st.X0[__ldg(ix + iu + 3)]; }
    Vec3<S> xb[NGP], vsmb;
    const double* w = Wd + e;
#pragma unroll
    for (int gp = 0; gp < NGP; ++gp)
#pragma unroll
        for (int i = 0; i < 3; ++i) { xb[gp][i].v = 0.; xb[gp][i].d0 = w[(gp * 3 + i) * g.nele]; }
#pragma unroll
    for (int i = 0; i < 3; ++i) { vsmb[i].v = 0.; vsmb[i].d0 = w[(NGP * 3 + i) * g.nele]; }
    beam_residual_cot<NumVal, S>(geo, m, Xu0, Xv0, xb, vsmb, R);
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 12; ++i) { double v = R[i].d0 * g.scaleX[i]; bad |= (v != v); Rp[e * 12 + i] = v; }
    if (bad) atomicMin(nanflag, nanbase + (unsigned long long)e);
}
// fused form (used when the workspace could not be allocated)
template <int ND>
__global__ void __launch_bounds__(MB_BLOCK, MB_MINB)
beam_dr_kernel(BeamGroupDev g, StateDev st, NewmarkDev nm, double* __restrict__ Rp, unsigned long long* nanflag, unsigned long long nanbase) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= g.nele) return;
    BeamGeo geo;
    load_geo(g.geo + e * 16, geo);
    const BeamMat m = g.mats[g.mat_id ? g.mat_id[e] : 0];
    using S = Dual<1>;
    S X[3][12], U[3], R[12];
    const int32_t* ix = g.idxX + e * 12;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const int32_t d = __ldg(ix + i);
        const double x0 = st.X0[d], x1 = (ND >= 2) ? st.X1[d] : 0., x2 = (ND >= 3) ? st.X2[d] : 0.;
        X[0][i].v = x0; X[0][i].d[0] = 0.;
        X[1][i].v = x1; X[1][i].d[0] = nm.a2 * x1 + nm.a3 * x2;
        X[2][i].v = x2; X[2][i].d[0] = nm.b2 * x1 + nm.b3 * x2;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) U[i] = Make<S>::c((g.udof && st.U0) ? st.U0[g.idxU[e * 3 + i]] : 0.);
    beam_residual<ND, 1>(geo, m, X, g.udof != 0, U, R);
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 12; ++i) { double v = R[i].d[0] * g.scaleX[i]; bad |= (v != v); Rp[e * 12 + i] = v; }
    if (bad) atomicMin(nanflag, nanbase + (unsigned long long)e);
}

// getresult: batched extraction of the element requestables, one thread per element (values only; HBM-bound: 616 B out per element)
// one thread per element; the 77 results of a thread are contiguous in `out`, so each warp stages its 32 × 77 values in shared memory and writes contiguous
// rows (as bar_kernel / soil_kernel do) — CTAs of two warps: 2 × 32 × 77 doubles = 39 KB
constexpr int MB_RES_BLOCK = 64;
template <int ND>
__global__ void __launch_bounds__(MB_RES_BLOCK)
beam_results_kernel(BeamGroupDev g, StateDev st, double* __restrict__ out) {
    __shared__ double tile[MB_RES_BLOCK / 32][32 * MB_NRES];
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t e0 = e - lane;
    if (e0 >= g.nele) return;
    if (e < g.nele) {
        BeamGeo geo;
        load_geo(g.geo + e * 16, geo);
        const BeamMat m = g.mats[g.mat_id ? g.mat_id[e] : 0];
        double Xu[3][6], Xv[3][6], r[MB_NRES];
        const int32_t* ix = g.idxX + e * 12;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int iu = (i < 3) ? i : i + 3;
            const int32_t du = __ldg(ix + iu), dv = __ldg(ix + iu + 3);
            Xu[0][i] = st.X0[du]; Xu[1][i] = (ND >= 2) ? st.X1[du] : 0.; Xu[2][i] = (ND >= 3) ? st.X2[du] : 0.;
            Xv[0][i] = st.X0[dv]; Xv[1][i] = (ND >= 2) ? st.X1[dv] : 0.; Xv[2][i] = (ND >= 3) ? st.X2[dv] : 0.;
        }
        beam_results<ND>(geo, m, Xu, Xv, r);
        double* tw = tile[warp] + lane * MB_NRES;
#pragma unroll
        for (int k = 0; k < MB_NRES; ++k) tw[k] = r[k];
    }
    __syncwarp();
    const int cnt = (int)min((int64_t)32, g.nele - e0);
    for (int q = lane; q < cnt * MB_NRES; q += 32) out[e0 * MB_NRES + q] = tile[warp][q];
}
void launch_beam_results(int ND, const BeamGroupDev& g, const StateDev& st, double* out, cudaStream_t s);

// K3: DirectXUA first-order path (src/DirectXUA.jl:85-120, no_second_order elements). Seeds are revariate{1}((;X,U),(;X=scale.X,U=scale.U))
// (src/Taylor.jl:158-166): one partial per X₀,X₁,…,X_OX dof and per U₀ dof.
// Output dR[e][p][i] = ∂R_i/∂seed_p with p in the reference's flat order X₀(12) X₁(12) X₂(12) U₀(3); R[e][i] unscaled (DirectXUA.jl:105).
// Three launches, each inside the instruction cache (the fused one-kernel form was 216 KB of SASS and instruction-fetch bound):
//   cot  (ND ≥ 2)  lanes (d,e,l), d ∈ {0,1}, d-major: time-jet forward seeded at X₀ or X′ → cotangents x̄_gp, v̄ₛₘ with partials → Wc; the
//                  partials with respect to X″ are the order-0 partials of the X₀ lane (beam_dyn_cotangents<…,DD>) and are written by it
//   b0             lanes (e,l): order-0 forward + reverse sweep in SD arithmetic → R and ∂R/∂X₀
//   lin            lanes (d≥1,e,l) and 2 U-lanes per Udof element: R is linear in the cotangents, so ∂R/∂X_d = J(X₀)ᵀ·∂c/∂X_d needs the forward
//                  sweep in plain values only and the reverse sweep on the partials of c (∂c/∂U is known in closed form: −dL·scale.U).
struct DirectStateDev { const double* X[3]; const double* U0; };
// Step batching (models with few elements and many time steps, BASELINE.json configs[4]: the steps of DirectXUA are independent, src/DirectXUA.jl:328-331):
// blockIdx.y = step within the batch; every per-step pointer advances by its stride.  One launch set then serves `nb` time steps instead of one.
struct StepBatch {
    int64_t sX = 0, sU = 0, sLam = 0, sdR = 0, sR = 0, sGX = 0, sWc = 0;     // strides in doubles per step
    unsigned long long snan = 0;                                              // increment of the NaN location code per step
    double dt = 0.;                                                           // state[step].time advances by Δt
};
__device__ __forceinline__ void batch_state(DirectStateDev& st, const StepBatch& sb) {
    const int64_t b = blockIdx.y;
    st.X[0] += b * sb.sX; st.X[1] += b * sb.sX; st.X[2] += b * sb.sX; if (st.U0) st.U0 += b * sb.sU;
}
// state of one lane: values of X₀..X_{ND-1}, U₀ and the lane's two seeds (rotation dof l, translation dof l) at derivative order d (d<0: no seed)
template <int ND> __device__ __forceinline__ void load_direct_state(const BeamGroupDev& g, const DirectStateDev& st, int64_t e, int d, int l,
                                               NumSD::TU (*Xu)[6], NumSD::TR (*Xv)[6], NumSD::TU* U) {
    const int32_t* ix = g.idxX + e * 12;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int iu = (i < 3) ? i : i + 3, iv = iu + 3;
        const int32_t du = __ldg(ix + iu), dv = __ldg(ix + iv);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            Xu[k][i].v = (k < ND) ? st.X[k][du] : 0.; Xu[k][i].d1 = (k == d && i == l) ? g.scaleX[iu] : 0.;
            Xv[k][i].v = (k < ND) ? st.X[k][dv] : 0.; Xv[k][i].d0 = (k == d && i == l) ? g.scaleX[iv] : 0.;
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { U[i].v = g.udof ? st.U0[g.idxU[e * 3 + i]] : 0.; U[i].d1 = 0.; }
}
// D1 (ND = 3): the launch of the X′-seeded lanes alone (thread t ↔ lane per + t), with the order-0 Taylor coefficients as plain values (NumJet1); the plain launch then covers
// the X₀ lanes only.  One instantiation of the jet code per KERNEL: each stays inside the instruction cache.
template <int ND, bool D1 = false>
__global__ void __launch_bounds__(MB_BLOCK, MB_MINB)
beam_direct_cot_kernel(BeamGroupDev g, DirectStateDev st, double* __restrict__ Wc, StepBatch sb) {
    using N = NumSD; using TR = N::TR; using TU = N::TU; using TS = N::TS;
    batch_state(st, sb); Wc += (int64_t)blockIdx.y * sb.sWc;
    constexpr int NDJ = (ND >= 3) ? 2 : ND;            // derivative orders that need their own time-jet lanes: X₀ and X′ (X″ comes with X₀)
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = g.nele * 6;
    if constexpr (ND >= 3) { if (t >= per) return; if (D1) t += per; }
    else if (t >= per * NDJ) return;
    const int d = (int)(t / per);
    const int64_t r = t - d * per, e = r / 6;
    const int l = (int)(r - e * 6);
    BeamGeo geo;
    load_geo(g.geo + e * 16, geo);
    const BeamMat m = g.mats[g.mat_id ? g.mat_id[e] : 0];
    TU Xu[3][6], U[3]; TR Xv[3][6];
    load_direct_state<ND>(g, st, e, d, l, Xu, Xv, U);
    Vec3<TS> xb[NGP], vsmb;
    if constexpr (ND >= 3 && D1) {
        beam_dyn_cotangents<ND, N, false, true>(geo, m, Xu, Xv, g.udof != 0, U, xb, vsmb, nullptr, nullptr, l < 3);
    } else if constexpr (ND >= 3) {
        Vec3<TS> xb2[NGP], vsmb2;
        beam_dyn_cotangents<ND, N, true>(geo, m, Xu, Xv, g.udof != 0, U, xb, vsmb, xb2, &vsmb2, l < 3);
        store_cot(Wc, 2 * per + t, xb2, vsmb2);                    // the tile the X″ lane (2,e,l) of the linear kernel reads
    } else
        beam_dyn_cotangents<ND, N>(geo, m, Xu, Xv, g.udof != 0, U, xb, vsmb, nullptr, nullptr, l < 3);
    store_cot(Wc, t, xb, vsmb);
}
template <int ND>
__global__ void __launch_bounds__(MB_BLOCK, MB_MINB)
beam_direct_b0_kernel(BeamGroupDev g, DirectStateDev st, double* __restrict__ dR, double* __restrict__ R, unsigned long long* nanflag,
                      unsigned long long nanbase, const double* __restrict__ Wc, StepBatch sb) {
    using N = NumSD; using TR = N::TR; using TU = N::TU; using TS = N::TS;
    batch_state(st, sb); Wc += (int64_t)blockIdx.y * sb.sWc; dR += (int64_t)blockIdx.y * sb.sdR; R += (int64_t)blockIdx.y * sb.sR; nanbase += blockIdx.y * sb.snan;
    const int NP = 12 * ND + (g.udof ? 3 : 0);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t / 6;
    const int l = (int)(t - e * 6);
    if (e >= g.nele) return;
    BeamGeo geo;
    load_geo(g.geo + e * 16, geo);
    const BeamMat m = g.mats[g.mat_id ? g.mat_id[e] : 0];
    TU Xu[3][6], U[3]; TR Xv[3][6]; TS Rv[12];
    load_direct_state<1>(g, st, e, 0, l, Xu, Xv, U);
    if (ND >= 2) {
        Vec3<TS> xb[NGP], vsmb;
        load_cot(Wc, t, xb, vsmb);
        beam_residual_cot<N, TS, (ND >= 3)>(geo, m, Xu[0], Xv[0], xb, vsmb, Rv, l < 3);
    } else {
        beam_residual_n<1, N>(geo, m, Xu, Xv, g.udof != 0, U, Rv);
    }
    bool bad = false;
    double* out = dR + e * (int64_t)(12 * NP);
    const bool al = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    const int cu = (l < 3) ? l : l + 3, cv = cu + 3;
#pragma unroll
    for (int i = 0; i < 12; i += 2) {
        double2 a, b; a.x = Rv[i].d1; a.y = Rv[i + 1].d1; b.x = Rv[i].d0; b.y = Rv[i + 1].d0;
        bad |= (a.x != a.x) | (a.y != a.y) | (b.x != b.x) | (b.y != b.y);
        store_pair(out + 12 * cu + i, a.x, a.y, al);
        store_pair(out + 12 * cv + i, b.x, b.y, al);
    }
    if (l == 0) {
#pragma unroll
        for (int i = 0; i < 12; ++i) { double v = Rv[i].v; bad |= (v != v); R[e * 12 + i] = v; }
    }
    if (bad) atomicMin(nanflag, nanbase + (unsigned long long)e);
}
template <int ND>
__global__ void __launch_bounds__(MB_BLOCK, MB_MINB)
beam_direct_lin_kernel(BeamGroupDev g, DirectStateDev st, double* __restrict__ dR, unsigned long long* nanflag, unsigned long long nanbase,
                       const double* __restrict__ Wc, StepBatch sb) {
    using NV = NumVal; using V = SD<false, false>; using S = NumSD::TS;
    batch_state(st, sb); Wc += (int64_t)blockIdx.y * sb.sWc; dR += (int64_t)blockIdx.y * sb.sdR; nanbase += blockIdx.y * sb.snan;
    const int NP = 12 * ND + (g.udof ? 3 : 0);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = g.nele * 6, nx = per * (ND - 1);
    int64_t e; int c0, c1;                       // columns of dR fed by slot 0 / slot 1 of this lane (c < 0: none)
    Vec3<S> xb[NGP], vsmb;
    BeamGeo geo;
    if (t < nx) {                                // X_d lane, d ≥ 1: cotangent partials from phase A (its thread index is per + t)
        const int d = 1 + (int)(t / per);
        const int64_t r = t - (d - 1) * per;
        e = r / 6;
        const int l = (int)(r - e * 6);
        c1 = 12 * d + ((l < 3) ? l : l + 3); c0 = c1 + 3;
        load_geo(g.geo + e * 16, geo);
        load_cot(Wc, per + t, xb, vsmb);
    } else {                                     // U lane u: slot 0 ← U dof 2u, slot 1 ← U dof 2u+1;  x̄_gp = dL·(fₑ − U) (BeamElement.jl:169)
        const int64_t tu = t - nx;
        e = tu >> 1;
        if (!g.udof || e >= g.nele) return;
        const int u = (int)(tu & 1);
        c0 = 12 * ND + 2 * u; c1 = (u == 0) ? 12 * ND + 1 : -1;
        load_geo(g.geo + e * 16, geo);
        const S z = Make<S>::c(0.);
#pragma unroll
        for (int gp = 0; gp < NGP; ++gp) {
            const double dL = gp_const(gp).w * geo.L;
            xb[gp] = Vec3<S>{z, z, z};
            if (u == 0) { xb[gp][0].d0 = -dL * g.scaleU[0]; xb[gp][1].d1 = -dL * g.scaleU[1]; }
            else xb[gp][2].d0 = -dL * g.scaleU[2];
        }
        vsmb = Vec3<S>{z, z, z};
    }
    const BeamMat m = g.mats[g.mat_id ? g.mat_id[e] : 0];
    V Xu0[6], Xv0[6]; S Rv[12];
    const int32_t* ix = g.idxX + e * 12;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int iu = (i < 3) ? i : i + 3;
        Xu0[i].v = st.X[0][__ldg(ix + iu)]; Xv0[i].v = st.X[0][__ldg(ix + iu + 3)];
    }
    beam_residual_cot<NV, S, (ND >= 3)>(geo, m, Xu0, Xv0, xb, vsmb, Rv);
    bool bad = false;
    double* out = dR + e * (int64_t)(12 * NP);
    const bool al = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
#pragma unroll
    for (int i = 0; i < 12; i += 2) {
        double2 a, b; a.x = Rv[i].d1; a.y = Rv[i + 1].d1; b.x = Rv[i].d0; b.y = Rv[i + 1].d0;
        bad |= (b.x != b.x) | (b.y != b.y);
        store_pair(out + 12 * c0 + i, b.x, b.y, al);
        if (c1 >= 0) { bad |= (a.x != a.x) | (a.y != a.y); store_pair(out + 12 * c1 + i, a.x, a.y, al); }
    }
    if (bad) atomicMin(nanflag, nanbase + (unsigned long long)e);
}
// returns the number of kernels launched; Wc must hold ((6·ND·nele+31)/32)·32·MB_NCOT doubles when ND ≥ 2
template <int ND> int launch_beam_direct(const BeamGroupDev& g, const DirectStateDev& st, double* dR, double* R, unsigned long long* nanflag,
                                         unsigned long long nanbase, double* Wc, cudaStream_t s, const StepBatch& sb, int nb);
#define MB_INSTANTIATE_BEAM_DIRECT(ND_)                                                                                                       \
    template <> int launch_beam_direct<ND_>(const BeamGroupDev& g, const DirectStateDev& st, double* dR, double* R,                           \
                                            unsigned long long* nanflag, unsigned long long nanbase, double* Wc, cudaStream_t s,              \
                                            const StepBatch& sb, int nb) {                                                                    \
        const int64_t per = g.nele * 6, nlin = per * (ND_ - 1) + (g.udof ? 2 * g.nele : 0);                                                  \
        int n = 1;                                                                                                                            \
        if constexpr (ND_ >= 3) {                                                                                                             \
            beam_direct_cot_kernel<(ND_ >= 3 ? ND_ : 3), false><<<dim3((unsigned)((per + MB_BLOCK - 1) / MB_BLOCK), nb), MB_BLOCK, 0, s>>>(g, st, Wc, sb);  \
            beam_direct_cot_kernel<(ND_ >= 3 ? ND_ : 3), true><<<dim3((unsigned)((per + MB_BLOCK - 1) / MB_BLOCK), nb), MB_BLOCK, 0, s>>>(g, st, Wc, sb);   \
            n += 2;                                                                                                                           \
        } else if constexpr (ND_ >= 2) { beam_direct_cot_kernel<(ND_ >= 2 ? ND_ : 2)><<<dim3((unsigned)((per * ND_ + MB_BLOCK - 1) / MB_BLOCK), nb), MB_BLOCK, 0, s>>>(g, st, Wc, sb); ++n; } \
        beam_direct_b0_kernel<ND_><<<dim3((unsigned)((per + MB_BLOCK - 1) / MB_BLOCK), nb), MB_BLOCK, 0, s>>>(g, st, dR, R, nanflag, nanbase, Wc, sb);      \
        if (nlin) { beam_direct_lin_kernel<ND_><<<dim3((unsigned)((nlin + MB_BLOCK - 1) / MB_BLOCK), nb), MB_BLOCK, 0, s>>>(g, st, dR, nanflag, nanbase, Wc, sb); ++n; } \
        return n;                                                                                                                             \
    }

// host-side launcher, one translation unit per (ND,STEP) so that the instantiations compile in parallel
struct BeamLaunch {
    BeamGroupDev g; StateDev st; NewmarkDev nm;
    double *Ke, *Re, *Rp; unsigned long long* nanflag; unsigned long long nanbase; int W; cudaStream_t stream; double* Wc;
    FuseDev fz{nullptr, nullptr, nullptr, 0};      // fused epilogue of the static kernel (statics, static_sym = 3): whdr = nullptr → off
    int static_sym = 3;     // statics: 3 = active/passive sweep with shuffle-shared packs (default), 2 = the same with local packs, 1 = symmetric-tangent kernel, 0 = two-direction SD kernel (A/B measurements)
    int nsm = 148;
};
template <int ND, bool STEP> void launch_beam(const BeamLaunch& a);
template <int ND> struct StaticSymLaunch { static void go(const BeamLaunch&, unsigned) {} };
template <> struct StaticSymLaunch<1> {
    static void go(const BeamLaunch& a, unsigned nb) {
        if (a.static_sym == 3 || a.static_sym == 13 || a.static_sym == 14) {             // active/passive sweep, packs shared by shuffles: 5 whole elements per warp
            const int64_t nw = (a.g.nele + 4) / 5;
            const unsigned nbw = (unsigned)((nw * 32 + MB_BLOCK - 1) / MB_BLOCK);
            const FuseDev nofz{nullptr, nullptr, nullptr, 0};
            if (a.static_sym == 13) beam_static_ap_kernel<3, true><<<nbw, MB_BLOCK, 0, a.stream>>>(a.g, a.st, a.Ke, a.Re, a.nanflag, a.nanbase, nofz);
            else if (a.static_sym == 14) beam_static_ap_kernel<4, true><<<nbw, MB_BLOCK, 0, a.stream>>>(a.g, a.st, a.Ke, a.Re, a.nanflag, a.nanbase, nofz);
            else if (a.fz.whdr) {
                // 40 KB of tiles and patterns per CTA, two CTAs per SM (register-bound): ask for a shared-memory carve-out that holds both
                static bool carved = false;
                if (!carved) { cudaFuncSetAttribute(beam_static_ap_kernel<MB_AP_MINB, true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 40); carved = true; }
                beam_static_ap_kernel<MB_AP_MINB, true, true><<<nbw, MB_BLOCK, 0, a.stream>>>(a.g, a.st, a.Ke, a.Re, a.nanflag, a.nanbase, a.fz);
            }
            else {
                // MB_AP_CARVE (experiment): shared-memory carve-out in % for a kernel that uses none — what is left of the 256 KB is L1 for the spill lines.  10 M elements:
                // default = 0 % = 6 % = 12 %: 14.25 ms; 100 % (smallest L1): 16.08 ms (tools/probe_carve.py)
                static int carve = -2;
                if (carve == -2) { const char* c = getenv("MB_AP_CARVE"); carve = c ? atoi(c) : -1; if (carve >= 0) cudaFuncSetAttribute(beam_static_ap_kernel<MB_AP_MINB, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve); }
                beam_static_ap_kernel<MB_AP_MINB, true><<<nbw, MB_BLOCK, 0, a.stream>>>(a.g, a.st, a.Ke, a.Re, a.nanflag, a.nanbase, nofz);
            }
        } else if (a.static_sym == 2)
            beam_static_ap_kernel<MB_AP_MINB, false><<<nb, MB_BLOCK, 0, a.stream>>>(a.g, a.st, a.Ke, a.Re, a.nanflag, a.nanbase, FuseDev{nullptr, nullptr, nullptr, 0});
        else
            beam_static_sym_kernel<MB_SYM_MINB><<<nb, MB_BLOCK, 0, a.stream>>>(a.g, a.st, a.Ke, a.Re, a.nanflag, a.nanbase);
    }
};
#define MB_INSTANTIATE_BEAM(ND_, STEP_)                                                                                               \
    template <> void launch_beam<ND_, STEP_>(const BeamLaunch& a) {                                                                   \
        const int64_t nt = a.g.nele * 6;                                                                                              \
        const unsigned nb = (unsigned)((nt + MB_BLOCK - 1) / MB_BLOCK);                                                               \
        if (ND_ == 1 && a.static_sym)                                                                                                 \
            StaticSymLaunch<ND_>::go(a, nb);                                                                                          \
        else if (ND_ >= 2 && a.Wc) {                                                                                                  \
            beam_cot_kernel<(ND_ >= 2 ? ND_ : 2)><<<nb, MB_BLOCK, 0, a.stream>>>(a.g, a.st, a.nm, a.Wc, nt);                         \
            beam_kernel_sd<ND_, true><<<nb, MB_BLOCK, 0, a.stream>>>(a.g, a.st, a.nm, a.Ke, a.Re, a.nanflag, a.nanbase, a.Wc, nt);    \
        } else                                                                                                                        \
            beam_kernel_sd<ND_, false><<<nb, MB_BLOCK, 0, a.stream>>>(a.g, a.st, a.nm, a.Ke, a.Re, a.nanflag, a.nanbase, nullptr, nt); \
        if (STEP_) {                                                                                                                  \
            const unsigned ne = (unsigned)((a.g.nele + MB_BLOCK - 1) / MB_BLOCK);                                                     \
            if (ND_ >= 2 && a.Wc) {       /* the cotangent workspace is free again once phase B has run (same stream) */               \
                beam_dr_cot_kernel<(ND_ >= 2 ? ND_ : 2)><<<ne, MB_BLOCK, 0, a.stream>>>(a.g, a.st, a.nm, a.Wc);                       \
                beam_dr_lin_kernel<<<ne, MB_BLOCK, 0, a.stream>>>(a.g, a.st, a.Wc, a.Rp, a.nanflag, a.nanbase);                       \
            } else                                                                                                                    \
                beam_dr_kernel<ND_><<<ne, MB_BLOCK, 0, a.stream>>>(a.g, a.st, a.nm, a.Rp, a.nanflag, a.nanbase);                      \
        }                                                                                                                             \
    }

}  // namespace mb
