// DirectXUA{OX,OU,0} on the device: per-step `assemble!{:matrices}(out::AssemblyDirect,…)` for no_second_order elements
// (src/DirectXUA.jl:85-120) and `assemblebig!` (src/DirectXUA.jl:316-356) of the owned time steps into the all-steps
// system Lvv / Lv (block layout of makepattern/SparseTools.prepare, src/DirectXUA.jl:245-315, src/SparseTools.jl:32-94).
//
// Layout: a handle owns the block columns of time steps [lo,hi) (0-based) of nstep; it evaluates and stores the per-step
// blocks L2[Λ,X][1,j], L2[X,Λ][j,1], L2[Λ,U][1,j], L2[U,Λ][j,1], L1[Λ] for the steps [elo,ehi) = [lo-2,hi+2)∩[0,nstep) whose
// finite-difference stencils (src/FiniteDifferences.jl) reach its columns.  Lvv is never assembled by scattered adds:
// one warp per Lvv column walks the blocks of that column and writes each value as the fixed-order weighted sum
// Σ_der w·Δt^(−der)·block_der[ilv] — deterministic, no atomics, no nnz-sized index map (the map is implicit in the block layout).
#include <algorithm>
#include <map>
#include "mb_internal.h"
#include <cub/cub.cuh>
#include "pattern_build.cuh"
#include "gauge_kernels.cuh"

namespace mb {
template <int ND> int launch_beam_direct(const BeamGroupDev& g, const DirectStateDev& st, double* dR, double* R, unsigned long long* nanflag,
                                         unsigned long long nanbase, double* Wc, cudaStream_t s, const StepBatch& sb, int nb);
int launch_bar_direct(int ND, const BarGroupDev& g, const DirectStateDev& st, double t, double* dR, double* R, unsigned long long* nanflag,
                      unsigned long long nanbase, cudaStream_t s, const StepBatch& sb, int nb);
int launch_soil_direct(int ND, const SoilGroupDev& g, const DirectStateDev& st, const double* Lam, double lamscale, double* dR, double* R, double* GX,
                       unsigned long long* nanflag, unsigned long long nanbase, cudaStream_t s, const StepBatch& sb, int nb);
}

namespace {

enum { P_XX = 0, P_XU = 1, P_UX = 2, P_UU = 3 };

constexpr int MAXG = 32;    // element types of one model on the DirectXUA path (the SCR riser of configs[4] has 11)
struct DirGroups { int n; uint32_t pbase[P_UU + 1][MAXG + 1]; int64_t drbase[MAXG]; int64_t rbase[MAXG]; int np[MAXG]; int udof[MAXG]; int nx[MAXG]; int64_t gxbase[MAXG];
                   int64_t hxxbase[MAXG]; };      // hxxbase: first element of a COSTED beam type in the strain-gauge scratch (gX [e][12], HXX [e][144], GU [e][3]), −1 otherwise

// ---------------------------------------------------------------------------------------------------------------- per-step gathers
__device__ __forceinline__ int find_group(const uint32_t* pbase, int n, uint32_t id) {
    int g = 0;
    while (g + 1 < n && id >= pbase[g + 1]) ++g;
    return g;
}
// XX-type pattern: L2[Λ,X][1,der] and L2[X,Λ][der,1] of one step (add_∂!{1} and add_∂!{1,:plus,:transpose}, DirectXUA.jl:114-115)
// Measured and dropped: one thread per pair of non-zeros with the pair / split descriptors of kernels.cuh instead of the cstart → src walk, all loads issued
// before the sums: 0.30 ms per step of 10⁵ elements either way — the strided loads of the transposed block and the 2·nd stores bound it, not the index chain.
// L2[X,Λ][der,1] is the transpose of L2[Λ,X][1,der] and the X-X class pattern is structurally symmetric (every element couples all its dofs both ways), with the contributors of
// (i,j) and (j,i) the same elements in the same order: the thread of non-zero k = (i,j) stores its sum at XL[permT[k]], permT[k] = the non-zero (j,i) — bit-identical to summing
// ∂R_j/∂X_i itself, without the strided loads of the transposed entries (a sector fetched per 8 bytes used)
__global__ void gather_xx_kernel(int64_t nnz, const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ src, DirGroups G, int nd,
                                 const double* __restrict__ dR, double* __restrict__ LX, double* __restrict__ XL, int64_t sdR, const uint32_t* __restrict__ permT,
                                 const double* __restrict__ slrow) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    dR += (int64_t)blockIdx.y * sdR; LX += (int64_t)blockIdx.y * nd * nnz; XL += (int64_t)blockIdx.y * nd * nnz;      // step batching: one grid row per time step
    double a[3] = {0., 0., 0.};
    for (uint32_t s = cstart[k]; s < cstart[k + 1]; ++s) {
        const uint32_t id = src[s];
        const int g = find_group(G.pbase[P_XX], G.n, id);
        const uint32_t loc = id - G.pbase[P_XX][g];
        const int nx = G.nx[g], n2 = nx * nx;
        const int64_t e = loc / n2; const int r = (int)(loc - e * n2); const int jj = r / nx, i = r - nx * jj;
        const double* d = dR + G.drbase[g] + e * (int64_t)(nx * G.np[g]);
        if (G.hxxbase[g] >= 0) {                   // costed beam type: the Λ rows are differentiated with respect to the SCALED Λ (rows of ∂R/∂seed times scale.Λ)
            const double sl = slrow[g * 12 + i];
            for (int der = 0; der < nd; ++der) a[der] += d[(nx * der + jj) * nx + i] * sl;
        } else
            for (int der = 0; der < nd; ++der) a[der] += d[(nx * der + jj) * nx + i];
    }
    const int64_t kt = permT[k];
    for (int der = 0; der < nd; ++der) { LX[der * nnz + k] = a[der]; XL[der * nnz + kt] = a[der]; }
}
// permT of a structurally symmetric CSC pattern: one thread per column j, binary search of row j in column i for each of its rows i
__global__ void transpose_perm_kernel(int64_t ncol, const int32_t* __restrict__ colptr, const int32_t* __restrict__ rowval, uint32_t* __restrict__ permT, unsigned long long* bad) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncol) return;
    for (int32_t k = colptr[j]; k < colptr[j + 1]; ++k) {
        const int32_t i = rowval[k];
        int32_t lo = colptr[i], hi = colptr[i + 1];
        while (lo < hi) { const int32_t mid = (lo + hi) >> 1; if (rowval[mid] < j) lo = mid + 1; else hi = mid; }
        if (lo < colptr[i + 1] && rowval[lo] == j) permT[k] = (uint32_t)lo; else { permT[k] = (uint32_t)k; atomicAdd(bad, 1ULL); }
    }
}
// XU-type (rows X dofs, cols U dofs): L2[Λ,U][1,1];  UX-type: L2[U,Λ][1,1].  Only ∂0(U) enters the toolbox elements.
__global__ void gather_xu_kernel(int64_t nnz, const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ src, DirGroups G, int nd, int transposed,
                                 const double* __restrict__ dR, double* __restrict__ out, int64_t sdR, const double* __restrict__ slrow) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    dR += (int64_t)blockIdx.y * sdR; out += (int64_t)blockIdx.y * nnz;
    const int pat = transposed ? P_UX : P_XU;
    double a = 0.;
    for (uint32_t s = cstart[k]; s < cstart[k + 1]; ++s) {
        const uint32_t id = src[s];
        const int g = find_group(G.pbase[pat], G.n, id);
        const uint32_t loc = id - G.pbase[pat][g];
        const int nx = G.nx[g], n2 = 3 * nx;
        const int64_t e = loc / n2; const int r = (int)(loc - e * n2);
        int ix, ju;
        if (!transposed) { ju = r / nx; ix = r - nx * ju; }      // rows X (nx), cols U (3): r = ix + nx·ju
        else { ix = r / 3; ju = r - 3 * ix; }                     // rows U (3), cols X (nx): r = ju + 3·ix
        const double v = dR[G.drbase[g] + e * (int64_t)(nx * G.np[g]) + (nx * nd + ju) * nx + ix];
        a += (G.hxxbase[g] >= 0) ? v * slrow[g * 12 + ix] : v;
    }
    out[k] = a;
}
__global__ void gather_l1_kernel(int64_t ndof, const uint32_t* __restrict__ vstart, const uint32_t* __restrict__ vsrc, const double* __restrict__ R,
                                 double* __restrict__ out, int64_t sR) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= ndof) return;
    R += (int64_t)blockIdx.y * sR; out += (int64_t)blockIdx.y * ndof;
    double acc = 0.;
    for (uint32_t s = vstart[d]; s < vstart[d + 1]; ++s) acc += R[vsrc[s]];
    out[d] = acc;
}
// L1[X][der] of one step from the second-order element types (∂L/∂X_der = Λᵀ∂R/∂X_der): GX[q·nd+der], q = contributor entry (element dof)
__global__ void gather_l1x_kernel(int64_t ndof, const uint32_t* __restrict__ vstart, const uint32_t* __restrict__ vsrc, const double* __restrict__ GX, int nd,
                                  double* __restrict__ out, int64_t sGX) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= ndof) return;
    GX += (int64_t)blockIdx.y * sGX; out += (int64_t)blockIdx.y * nd * ndof;
    double acc[3] = {0., 0., 0.};
    for (uint32_t s = vstart[d]; s < vstart[d + 1]; ++s) {
        const double* q = GX + (int64_t)vsrc[s] * nd;
        for (int der = 0; der < nd; ++der) acc[der] += q[der];
    }
    for (int der = 0; der < nd; ++der) out[der * ndof + d] = acc[der];
}
__global__ void vec_keys_kernel(int64_t n, const int32_t* __restrict__ idx, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t base) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    keys[base + q] = (uint32_t)idx[q]; vals[base + q] = base + (uint32_t)q;
}
__global__ void vstart2_kernel(int64_t nvec, const uint32_t* __restrict__ keys, int64_t ndof, uint32_t* __restrict__ vstart) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nvec) return;
    const int64_t d = keys[s];
    const int64_t prev = (s == 0) ? -1 : (int64_t)keys[s - 1];
    for (int64_t c = prev + 1; c <= d; ++c) vstart[c] = (uint32_t)s;
    if (s == nvec - 1) for (int64_t c = d + 1; c <= ndof; ++c) vstart[c] = (uint32_t)nvec;
}

// ---------------------------------------------------------------------------------------------------------------- ElementCost{StrainGaugeOnEulerBeam3D} on this path
// A costed beam type takes the accelerator's route (src/DirectXUA.jl:172-198 as it is meant, see mb_xua.cu): L = Λ∘₁R + cost with first-order R.  From the outputs of the
// first-order kernels (R unscaled, dR[e][p][i] with scaled seeds) of ONE element per CTA:
//   GX[(e·12+i)·nd+d] = Σₖ Λₖ·∂Rₖ/∂X_d,i (+ ∇cost_i for d = 0) → L1[X][d+1];   GU[e][j] = Σₖ Λₖ·∂Rₖ/∂U_j → L1[U][1];   then R and the rows of dR are scaled by scale.Λ
// (DirectXUA_lagrangian_addition! differentiates with respect to the scaled Λ), so that the reductions of the plain path give L1[Λ], L2[Λ,X], L2[X,Λ], L2[Λ,U], L2[U,Λ].
struct SL12 { double v[12]; };
__global__ void __launch_bounds__(128) costed_fill_kernel(int64_t nele, int nd, int npd, const int32_t* __restrict__ idxX, const double* __restrict__ Lam, SL12 sL,
                                                          const double* __restrict__ gX, double* __restrict__ R, const double* __restrict__ dR, double* __restrict__ GX, double* __restrict__ GU,
                                                          StepBatch sb, int64_t sE, int64_t sGU) {
    // one WARP per element (a CTA per element left 89 of its 128 threads idle behind a barrier): lane p < npd owns row p of ∂R/∂seed — 12 consecutive doubles, six 16-byte loads
    { const int64_t b = blockIdx.y; Lam += b * sb.sLam; gX += b * sE * 12; R += b * sb.sR; dR += b * sb.sdR; GX += b * sb.sGX; if (GU) GU += b * sGU; }      // step batching: grid row = time step
    const int lane = threadIdx.x & 31;
    const int64_t e = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (e >= nele) return;
    double lam = (lane < 12) ? Lam[idxX[e * 12 + lane]] : 0.;
    double l[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) l[k] = __shfl_sync(0xffffffffu, lam, k);
    const double* base = dR + e * (int64_t)(npd * 12);
    const bool al = (reinterpret_cast<uintptr_t>(base) & 15) == 0;
    for (int p = lane; p < npd; p += 32) {
        const double* row = base + p * 12;
        double acc = 0.;
        if (al) {
#pragma unroll
            for (int k = 0; k < 12; k += 2) { const double2 v = *reinterpret_cast<const double2*>(row + k); acc += l[k] * v.x; acc += l[k + 1] * v.y; }
        } else {
#pragma unroll
            for (int k = 0; k < 12; ++k) acc += l[k] * row[k];
        }
        if (p < 12 * nd) { const int d = p / 12, i = p - 12 * d; GX[(e * 12 + i) * nd + d] = acc + (d == 0 ? gX[e * 12 + i] : 0.); }
        else if (GU) GU[e * 3 + (p - 12 * nd)] = acc;
    }
    if (lane < 12) R[e * 12 + lane] *= sL.v[lane];      // the rows of ∂R/∂seed get their scale.Λ in the reductions (gather_xx / gather_xu_kernel): no rewrite of dR
}
// L2[X,X][1,1] of one step from the costs' Gauss-Newton blocks HXX[e][12][12] (costed beam types only; the reference's accumulation order over the X-X pattern)
__global__ void gather_xxc_kernel(int64_t nnz, const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ src, DirGroups G, const double* __restrict__ HXX, double* __restrict__ out,
                                  int64_t sE) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    HXX += (int64_t)blockIdx.y * sE * 144; out += (int64_t)blockIdx.y * nnz;
    double a = 0.;
    for (uint32_t s = cstart[k]; s < cstart[k + 1]; ++s) {
        const uint32_t id = src[s];
        const int g = find_group(G.pbase[P_XX], G.n, id);
        if (G.hxxbase[g] < 0) continue;
        const uint32_t loc = id - G.pbase[P_XX][g];
        const int64_t e = loc / 144; const int r = (int)(loc - e * 144); const int jj = r / 12, i = r - 12 * jj;
        a += HXX[(G.hxxbase[g] + e) * 144 + jj * 12 + i];
    }
    out[k] = a;
}
// L1[U][1] of one step: contributors GU[q] of every U-dof of the costed types
__global__ void gather_l1u_kernel(int64_t ndof, const uint32_t* __restrict__ vstart, const uint32_t* __restrict__ vsrc, const double* __restrict__ GU, double* __restrict__ out, int64_t sGU) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= ndof) return;
    GU += (int64_t)blockIdx.y * sGU; out += (int64_t)blockIdx.y * ndof;
    double acc = 0.;
    for (uint32_t s = vstart[d]; s < vstart[d + 1]; ++s) acc += GU[vsrc[s]];
    out[d] = acc;
}

// ---------------------------------------------------------------------------------------------------------------- all-steps system
struct BigDev {
    int64_t nX, nU, W;             // W = 2nX+nU rows per step
    int64_t nstep, lo, hi, elo;    // owned steps [lo,hi), stored steps start at elo (all 0-based)
    int OX, OU;
    double dt;
    const int32_t* bcolptr;        // block pattern of the owned block columns (CSC over blocks), from makepattern
    const int32_t* browval;        // global block row = 3·step + class (0 Λ, 1 X, 2 U)
    const int32_t* pc[4]; const int32_t* pr[4];   // class-pair patterns colptr0 / rowval0
    int64_t pnnz[4];
    const double *LX, *XL, *LU, *UL, *L1L;        // stored per-step blocks
    int64_t sLX, sLU, sUL;         // strides per stored step
    const double* hostc;           // host-evaluated single-dof costs per stored step: gX(nX) hX(nX) gU(nU) hU(nU), or nullptr
    const double* L1X;             // L1[X][der] per stored step [step][der][nX] from second-order element types, or nullptr
    int64_t ehi;
    const double *XXc, *L1U;       // costed beam types: L2[X,X][1,1] per stored step [step][nnz(X,X)], L1[U][1] per stored step [step][nU]; or nullptr
};
__device__ __forceinline__ int pat_of(int ca, int cb) { return (ca == 2 ? 2 : 0) + (cb == 2 ? 1 : 0); }   // class 2 = U
__device__ __forceinline__ void decode_col(const BigDev& B, int64_t c, int64_t& step, int& cls, int64_t& lc) {
    step = B.lo + c / B.W; const int64_t r = c % B.W;
    if (r < B.nX) { cls = 0; lc = r; } else if (r < 2 * B.nX) { cls = 1; lc = r - B.nX; } else { cls = 2; lc = r - 2 * B.nX; }
}
__global__ void big_count_kernel(BigDev B, int64_t ncol, int64_t* __restrict__ cnt) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    int64_t step, lc; int cb;
    decode_col(B, c, step, cb, lc);
    const int64_t bc = 3 * (step - B.lo) + cb;
    int64_t n = 0;
    for (int32_t q = B.bcolptr[bc]; q < B.bcolptr[bc + 1]; ++q) {
        const int p = pat_of(B.browval[q] % 3, cb);
        n += B.pc[p][lc + 1] - B.pc[p][lc];
    }
    cnt[c] = n;
}
// one warp per owned Lvv column: structure (rowval) and/or values
template <bool STRUCT>
__global__ void big_fill_kernel(BigDev B, int64_t ncol, const int64_t* __restrict__ colptr, int64_t* __restrict__ rowval, double* __restrict__ nzval) {
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= ncol) return;
    int64_t tcol, lc; int cb;
    decode_col(B, c, tcol, cb, lc);
    const int64_t bc = 3 * (tcol - B.lo) + cb;
    int64_t off = colptr[c];
    for (int32_t q = B.bcolptr[bc]; q < B.bcolptr[bc + 1]; ++q) {
        const int32_t br = B.browval[q];
        const int ca = br % 3; const int64_t trow = br / 3;
        const int p = pat_of(ca, cb);
        const int32_t p0 = B.pc[p][lc], p1 = B.pc[p][lc + 1];
        if (STRUCT) {
            const int64_t g0 = trow * B.W + (ca == 0 ? 0 : (ca == 1 ? B.nX : 2 * B.nX));
            for (int32_t k = p0 + lane; k < p1; k += 32) rowval[off + (k - p0)] = g0 + B.pr[p][k];
        } else {
            // which evaluated step feeds this block, and through which per-step array
            const double* arr = nullptr; int64_t stride = 0, s = 0, t = 0; int nder = 0; int64_t nnzp = B.pnnz[p];
            if (ca == 0 && cb != 0) { s = trow; t = tcol; if (cb == 1) { arr = B.LX; stride = B.sLX; nder = B.OX + 1; } else { arr = B.LU; stride = B.sLU; nder = 1; } }
            else if (cb == 0 && ca != 0) { s = tcol; t = trow; if (ca == 1) { arr = B.XL; stride = B.sLX; nder = B.OX + 1; } else { arr = B.UL; stride = B.sUL; nder = 1; } }
            else if (ca == 1 && cb == 1 && trow == tcol && B.XXc) { s = tcol; t = tcol; arr = B.XXc; stride = nnzp; nder = 1; }      // the costs' L2[X,X][1,1] of that step
            double wd[3]; bool on[3] = {false, false, false};
            double sc = 1.;
            for (int der = 0; der < nder; ++der) { double w = 0.; on[der] = fd_weight(der, B.nstep, s, t - s, w); wd[der] = w * sc; sc /= B.dt; }
            const double* a = arr ? arr + (s - B.elo) * stride : nullptr;
            for (int32_t k = p0 + lane; k < p1; k += 32) {
                double v = 0.;
                if (a) for (int der = 0; der < nder; ++der) if (on[der]) v += a[der * nnzp + k] * wd[der];
                nzval[off + (k - p0)] = v;
            }
        }
        off += p1 - p0;
    }
}
// Values of the owned Lvv columns, the hot form: a CTA works inside ONE block column (tcol,class) on a chunk of pattern columns, so the
// list of blocks of that block column — source array, finite-difference weights w_α·Δt^(-α) (DirectXUA.jl:347-349), class-pair pattern —
// is decoded once into shared memory.  Inside a block the source entries of consecutive pattern columns are contiguous (CSC), so the
// threads run over them flat: coalesced loads of ≤ 3 derivative arrays, destination = column base + offset of the block inside the
// column + position in the slice.  For a fixed column class only two class-pair patterns occur (rows Λ/X, rows U), so the offset of block
// j inside a column is nA_j·sizeA(col) + nB_j·sizeB(col).
constexpr int MB_BIG_MAXB = 48;          // ≥ 3 classes × 5 step offsets × (room for A-class blocks)
constexpr int MB_BIG_CPC = 128;          // pattern columns per CTA
constexpr int MB_BIG_CAP = 6144;         // pattern entries of one kind per CTA chunk held in the column-of-entry table
struct BlkDesc { const double* a; const int32_t* pc; double wd[3]; int64_t nnzp; int nder, kind, nA, nB; };
__global__ void __launch_bounds__(256) big_values_kernel(BigDev B, const int64_t* __restrict__ colptr, double* __restrict__ nzval, int big_sub) {
    __shared__ BlkDesc sd[MB_BIG_MAXB];
    __shared__ int64_t colbase[MB_BIG_CPC];
    __shared__ int32_t pst[2][MB_BIG_CPC + 1];           // pattern colptr of the chunk, kinds A (rows Λ/X) and B (rows U)
    __shared__ uint8_t colof[2][MB_BIG_CAP];
    extern __shared__ int32_t drel[];                    // [nb][MB_BIG_CPC], sized by the launch (maxb blocks)
    // block column fastest: CTAs running together read the same chunk of the per-step source arrays for all their time offsets (L2 reuse)
    const int nbc = 3 * (int)(B.hi - B.lo);
    const int bc = (int)(blockIdx.x % (unsigned)nbc), cb = bc % 3;
    const int64_t tcol = B.lo + bc / 3;
    const int64_t ncls = (cb == 2) ? B.nU : B.nX;
    const int64_t lc0 = (int64_t)(blockIdx.x / (unsigned)nbc) * MB_BIG_CPC;
    if (lc0 >= ncls) return;
    const int nc = (int)((lc0 + MB_BIG_CPC < ncls) ? MB_BIG_CPC : ncls - lc0);
    const int32_t q0 = B.bcolptr[bc];
    const int nb = B.bcolptr[bc + 1] - q0;
    const int pA = pat_of(0, cb), pB = pat_of(2, cb);
    if ((int)threadIdx.x < nb) {
        const int32_t br = B.browval[q0 + threadIdx.x];
        const int ca = br % 3; const int64_t trow = br / 3;
        const int p = pat_of(ca, cb);
        BlkDesc d; d.a = nullptr; d.pc = B.pc[p]; d.nnzp = B.pnnz[p]; d.nder = 0; d.wd[0] = d.wd[1] = d.wd[2] = 0.; d.kind = (p == pA) ? 0 : 1;
        d.nA = 0; d.nB = 0;
        for (int j = 0; j < (int)threadIdx.x; ++j) { if (pat_of(B.browval[q0 + j] % 3, cb) == pA) ++d.nA; else ++d.nB; }
        const double* arr = nullptr; int64_t stride = 0, s = 0, t = 0;
        if (ca == 0 && cb != 0) { s = trow; t = tcol; if (cb == 1) { arr = B.LX; stride = B.sLX; d.nder = B.OX + 1; } else { arr = B.LU; stride = B.sLU; d.nder = 1; } }
        else if (cb == 0 && ca != 0) { s = tcol; t = trow; if (ca == 1) { arr = B.XL; stride = B.sLX; d.nder = B.OX + 1; } else { arr = B.UL; stride = B.sUL; d.nder = 1; } }
        else if (ca == 1 && cb == 1 && trow == tcol && B.XXc) { s = tcol; t = tcol; arr = B.XXc; stride = d.nnzp; d.nder = 1; }      // the costs' L2[X,X][1,1] of that step
        double sc = 1.;
        for (int der = 0; der < d.nder; ++der) { double w = 0.; const bool on = fd_weight(der, B.nstep, s, t - s, w); d.wd[der] = on ? w * sc : 0.; sc /= B.dt; }
        if (arr) d.a = arr + (s - B.elo) * stride;
        sd[threadIdx.x] = d;
    }
    const int64_t cbase = (tcol - B.lo) * B.W + (cb == 0 ? 0 : (cb == 1 ? B.nX : 2 * B.nX));
    for (int c = threadIdx.x; c <= nc; c += blockDim.x) {
        pst[0][c] = B.pc[pA][lc0 + c]; pst[1][c] = B.pc[pB][lc0 + c];
        if (c < nc) colbase[c] = colptr[cbase + lc0 + c];
    }
    __syncthreads();
    const int nEA = pst[0][nc] - pst[0][0], nEB = pst[1][nc] - pst[1][0];
    if (nEA <= MB_BIG_CAP && nEB <= MB_BIG_CAP) {
        for (int c = threadIdx.x; c < nc; c += blockDim.x)
#pragma unroll
            for (int kd = 0; kd < 2; ++kd)
                for (int32_t k = pst[kd][c]; k < pst[kd][c + 1]; ++k) colof[kd][k - pst[kd][0]] = (uint8_t)c;
        __syncthreads();
        // Sub-chunks of MB_BIG_SUB columns, blocks inside: the Lvv columns of a sub-chunk (a contiguous destination range) are completed before the next one is started, so the
        // partially written sectors of all resident CTAs fit L2 (block-outer order left 276 KB per CTA in flight: written-back half-filled)
        // destination of pattern entry k of block j in column c: colbase[c] + nA_j·sizeA(c) + nB_j·sizeB(c) + (k − pst[kd][c]) = nzbase + drel[j][c] + k with drel tabulated once
        // per CTA (32-bit, relative to the chunk's first non-zero): the inner loop is one table look-up and an add
        const int64_t nzbase = colbase[0];
        for (int q = threadIdx.x; q < nb * nc; q += blockDim.x) {
            const int j = q / nc, c = q - j * nc;
            drel[j * MB_BIG_CPC + c] = (int32_t)(colbase[c] - nzbase) + sd[j].nA * (pst[0][c + 1] - pst[0][c]) + sd[j].nB * (pst[1][c + 1] - pst[1][c]) - pst[sd[j].kind][c];
        }
        __syncthreads();
        for (int cs = 0; cs < nc; cs += big_sub)
        for (int j = 0; j < nb; ++j) {
            const int kd = sd[j].kind;
            const double* a = sd[j].a;
            const int nder = sd[j].nder; const int64_t nnzp = sd[j].nnzp;
            const double w0 = sd[j].wd[0], w1 = sd[j].wd[1], w2 = sd[j].wd[2];
            const int ce = cs + big_sub < nc ? cs + big_sub : nc;
            const int32_t k0 = pst[kd][0], ka = pst[kd][cs], k1 = pst[kd][ce];
            const int32_t* dj = drel + j * MB_BIG_CPC;
            double* out = nzval + nzbase;
            const bool u0 = a && w0 != 0., u1 = a && nder > 1 && w1 != 0., u2 = a && nder > 2 && w2 != 0.;
            for (int32_t k = ka + threadIdx.x; k < k1; k += blockDim.x) {
                const int c = colof[kd][k - k0];
                double v = 0.;                                     // same order of additions as the reference (DirectXUA.jl:342-352)
                if (u0) v += a[k] * w0;
                if (u1) v += a[nnzp + k] * w1;
                if (u2) v += a[2 * nnzp + k] * w2;
                out[dj[c] + k] = v;
            }
        }
    } else {                                                       // very dense pattern columns: one warp per column
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int c = warp; c < nc; c += 8) {
            int64_t off = colbase[c];
            for (int j = 0; j < nb; ++j) {
                const int kd = sd[j].kind;
                const int32_t p0 = pst[kd][c], p1 = pst[kd][c + 1];
                const double* a = sd[j].a;
                const int nder = sd[j].nder; const int64_t nnzp = sd[j].nnzp;
                const double w0 = sd[j].wd[0], w1 = sd[j].wd[1], w2 = sd[j].wd[2];
                for (int32_t k = p0 + lane; k < p1; k += 32) {
                    double v = 0.;
                    if (a) {
                        if (w0 != 0.) v += a[k] * w0;
                        if (nder > 1 && w1 != 0.) v += a[nnzp + k] * w1;
                        if (nder > 2 && w2 != 0.) v += a[2 * nnzp + k] * w2;
                    }
                    nzval[off + (k - p0)] = v;
                }
                off += p1 - p0;
            }
        }
    }
}
static void launch_big_values(const BigDev& B, int64_t ncol, const int64_t* colptr, double* nzval, int maxb, cudaStream_t st) {
    const int64_t nbc = 3 * (B.hi - B.lo), ncls = B.nX > B.nU ? B.nX : B.nU;
    const int64_t nchunk = (ncls + MB_BIG_CPC - 1) / MB_BIG_CPC;
    if (maxb <= MB_BIG_MAXB && nchunk * nbc < (int64_t)INT32_MAX)
    {
        static int big_sub = -1;
        if (big_sub < 0) { const char* e = getenv("MB_BIG_SUB"); big_sub = e ? std::max(1, atoi(e)) : 32; }
        big_values_kernel<<<(unsigned)(nchunk * nbc), 256, (size_t)maxb * MB_BIG_CPC * sizeof(int32_t), st>>>(B, colptr, nzval, big_sub);
    }
    else
        big_fill_kernel<false><<<nblk(ncol * 32, 256), 256, 0, st>>>(B, ncol, colptr, nullptr, nzval);
}
__global__ void big_vec_kernel(BigDev B, int64_t ncol, double* __restrict__ Lv) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    int64_t step, lc; int cls;
    decode_col(B, c, step, cls, lc);
    double v = 0.;
    if (cls == 0) v = B.L1L[(step - B.elo) * B.nX + lc];               // addin!(Lvasm,Lv,L1[Λ][1],Λblk)  (DirectXUA.jl:332-341)
    else {
        if (B.hostc) v = B.hostc[(step - B.elo) * 2 * (B.nX + B.nU) + (cls == 1 ? 0 : 2 * B.nX) + lc];   // L1[X][1], L1[U][1] of cost elements
        if (cls == 1 && B.L1X) {
            // Lv[X block of step] += Σ_s Σ_der w·Δt^(−der)·L1[X][der](state s), w = finitediff(der,nstep,s) at offset step−s (DirectXUA.jl:332-341)
            double acc = 0.;
            for (int64_t s = step - 2; s <= step + 2; ++s) {
                if (s < B.elo || s >= B.ehi) continue;
                double sc = 1.;
                for (int der = 0; der <= B.OX; ++der) {
                    double w;
                    if (fd_weight(der, B.nstep, s, step - s, w)) acc += B.L1X[((s - B.elo) * (B.OX + 1) + der) * B.nX + lc] * (w * sc);
                    sc /= B.dt;
                }
            }
            v = (B.hostc ? v : 0.) + acc;
        }
        if (cls == 2 && B.L1U) v += B.L1U[(step - B.elo) * B.nU + lc];      // Λᵀ∂R/∂U of the costed types
    }
    Lv[c] = v;
}
// L2[X,X][1,1] / L2[U,U][1,1] of host-evaluated single-dof costs: one diagonal entry per dof, added to the (step,step) block
__global__ void big_diag_kernel(BigDev B, int64_t ncol, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval, double* __restrict__ nzval,
                                unsigned long long* missing, int64_t rowshift) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    int64_t step, lc; int cls;
    decode_col(B, c, step, cls, lc);
    if (cls == 0) return;
    const double hd = B.hostc[(step - B.elo) * 2 * (B.nX + B.nU) + (cls == 1 ? B.nX : 2 * B.nX + B.nU) + lc];
    if (hd == 0.) return;
    const int64_t row = step * B.W + (cls == 1 ? B.nX : 2 * B.nX) + lc - rowshift;
    int64_t lo = colptr[c], hi = colptr[c + 1];
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (rowval[mid] < row) lo = mid + 1; else hi = mid; }
    if (lo < colptr[c + 1] && rowval[lo] == row) nzval[lo] += hd; else atomicAdd(missing, 1ULL);
}

// L2[X,X][1,1] entries of host-evaluated element types (mb_direct_set_host_xx): unique (i,j) per step, added to the (step,step) X-X block
__global__ void big_xx_kernel(BigDev B, int64_t n, const int32_t* __restrict__ ii, const int32_t* __restrict__ jj, const double* __restrict__ v, int64_t step,
                              const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval, double* __restrict__ nzval, unsigned long long* missing, int64_t rowshift) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int64_t c = (step - B.lo) * B.W + B.nX + jj[q];
    const int64_t row = step * B.W + B.nX + ii[q] - rowshift;
    int64_t lo = colptr[c], hi = colptr[c + 1];
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (rowval[mid] < row) lo = mid + 1; else hi = mid; }
    if (lo < colptr[c + 1] && rowval[lo] == row) nzval[lo] += v[q]; else atomicAdd(missing, 1ULL);
}

// ---------------------------------------------------------------------------------------------------------------- sparser! / decrementbig!
// sparser!(T,S,rtol) (src/SparseTools.jl:172-199): keep |nzval| ≥ rtol·max|S|, order preserved, colptr shifted by the drops before it
// decrementbig! (src/DirectXUA.jl:357-383): state[step].β[βder] −= ((Δβ·w)·Δt^(1−βder))·scale for every stencil point (Δs,w) of
// finitediff(βder−1,nstep,step), Δβ = Δv block of step+Δs, class β — same sequence of roundings (no FMA contraction).
struct DecDev {
    int64_t nX, nU, W, nstep, elo, ehi, s0, s1; int OX; double dtp[3];
    const double* dv; double *Lam, *X, *U; const double *scL, *scX, *scU;
};
__global__ void decrementbig_kernel(DecDev D) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (D.ehi - D.elo) * D.W) return;
    const int64_t k = q / D.W, r = q - k * D.W, step = D.elo + k;
    int cls; int64_t i;
    if (r < D.nX) { cls = 0; i = r; } else if (r < 2 * D.nX) { cls = 1; i = r - D.nX; } else { cls = 2; i = r - 2 * D.nX; }
    const int nder = (cls == 1) ? D.OX + 1 : 1;
    const double sc = (cls == 0) ? (D.scL ? D.scL[i] : 1.) : (cls == 1 ? (D.scX ? D.scX[i] : 1.) : (D.scU ? D.scU[i] : 1.));
    for (int der = 0; der < nder; ++der) {
        double* x = (cls == 0) ? D.Lam + k * D.nX + i : (cls == 1 ? D.X + (k * 3 + der) * D.nX + i : D.U + k * D.nU + i);
        double v = *x;
        for (int64_t ds = -2; ds <= 2; ++ds) {                       // stencil points come in ascending Δs (FiniteDifferences.jl:8-31)
            double w;
            if (!fd_weight(der, D.nstep, step, ds, w)) continue;
            const double d = D.dv[(step + ds - D.s0) * D.W + r];
            v = __dsub_rn(v, __dmul_rn(__dmul_rn(__dmul_rn(d, w), D.dtp[der]), sc));
        }
        *x = v;
    }
}
// Σ Δβ² per (step, class) of the owned steps; the maximum over steps is Δ²[β] (DirectXUA.jl:369-371)
__global__ void __launch_bounds__(256) block_sumsq_kernel(int64_t nX, int64_t nU, int64_t W, const double* __restrict__ dv, double* __restrict__ out) {
    __shared__ double sh[256];
    const int64_t b = blockIdx.x; const int cls = (int)(b % 3); const int64_t k = b / 3;
    const int64_t n = (cls == 2) ? nU : nX, off = k * W + (cls == 0 ? 0 : (cls == 1 ? nX : 2 * nX));
    double a = 0.;
    for (int64_t i = threadIdx.x; i < n; i += 256) { const double d = dv[off + i]; a += d * d; }
    sh[threadIdx.x] = a; __syncthreads();
    for (int s = 128; s > 0; s >>= 1) { if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s]; __syncthreads(); }
    if (threadIdx.x == 0) out[b] = sh[0];
}

}  // namespace

// location of the first NaN, packed so that atomicMin keeps the FIRST one in (step, element type, element) order: step above bit 46, six bits of
// element-type index (MAXG = 32 types), 40 bits of element index — one encode / decode pair for every kernel of this file
static inline unsigned long long nan_pack(int64_t step, int ig) { return (((unsigned long long)step) << 46) | (((unsigned long long)ig) << 40); }
static inline void nan_unpack(unsigned long long f, mb_errinfo* where) {
    where->ieletyp = (int32_t)((f >> 40) & 0x3F) + 1; where->iele = (int64_t)(f & ((1ULL << 40) - 1)) + 1; where->step = (int64_t)(f >> 46) + 1;
}
static_assert(MAXG <= 64, "nan_pack gives the element-type index six bits");
struct DirectData {
    int OX = 0, OU = 0;
    int64_t nX = 0, nU = 0, nstep = 0, lo = 0, hi = 0, elo = 0, ehi = 0;
    double dt = 1., t0 = 0.;                            // state[step].time = t0 + step·dt (read by Bar3D's weight ramp, BarElement.jl:144)
    PairPat pat[4];
    uint32_t *vstart = nullptr, *vsrc = nullptr;
    DirGroups G;
    double *dR = nullptr, *R = nullptr;                 // element outputs of one BATCH of steps: [step in batch][ndr] / [nvec]
    int64_t ndr = 0, nvec = 0, batch = 1;               // batch: time steps evaluated by one launch set (step batching for small models)
    double *X = nullptr, *U = nullptr, *Lam = nullptr;  // stored states [step][3][nX], [step][nU], [step][nX] (Λ enters decrementbig! only)
    double *scL = nullptr, *scX = nullptr, *scU = nullptr, *dvbuf = nullptr; int64_t dvlen = 0;
    int64_t *ccolptr = nullptr, *crowval = nullptr; double* cnzval = nullptr; int64_t cnnz = -1;   // sparser! result
    double* hostc = nullptr; unsigned long long* missing = nullptr;
    struct HostXX { int64_t n = 0; int32_t *i = nullptr, *j = nullptr; double* v = nullptr; };
    std::map<int64_t, HostXX> hostxx;                   // per (absolute) step: pre-summed L2[X,X][1,1] entries of host-evaluated types that are not linear in X
    std::vector<double*> hoststore;                     // per host-evaluated type: [stored step][R(nele·nx) | dR(nele·nx·nx·nd) | GX(nele·nx·nd)] or nullptr
    double *GX = nullptr, *L1X = nullptr; uint32_t *vstart2 = nullptr, *vsrc2 = nullptr; int64_t ngx = 0; double lamscale = 1.;   // second-order types
    double *LX = nullptr, *XL = nullptr, *LU = nullptr, *UL = nullptr, *L1L = nullptr;
    int32_t *bcolptr = nullptr, *browval = nullptr;
    int64_t ncol = 0, nnzbig = 0; int maxb = 0;         // maxb: most blocks in one block column
    int64_t *colptr = nullptr, *rowval = nullptr;       // rowval: global rows of the window the structure was BUILT for; + rowshift after mb_direct_rebase
    int64_t rowshift = 0;
    bool elements_only = false;                         // timing aid: direct_eval_steps launches the element kernels without the per-step reductions
    // costed beam types (mb_direct_set_gauge_cost): per stored step L2[X,X][1,1] and L1[U][1]; scratch of one step (J, e4, gX, HXX, GU, costs); U-dof contributor lists;
    // per type the measurements of every stored step [step][ng] or [step][nele][ng]
    double* slrow = nullptr;                            // [group][12] scale.Λ of the residual rows of costed beam types (scale.X·Λscale), read by the reductions
    uint32_t* permT = nullptr;                          // X-X class pattern: position of the transposed non-zero (gather_xx_kernel)
    bool costed = false; int64_t ncost = 0, nqu = 0;
    double *XXc = nullptr, *L1U = nullptr, *cJ = nullptr, *ce4 = nullptr, *cgX = nullptr, *cHXX = nullptr, *cGU = nullptr, *ccost = nullptr, *csL = nullptr;
    uint32_t *vstartU = nullptr, *vsrcU = nullptr;
    struct Meas { double* eps = nullptr; int64_t stride = 0; bool per_element = false; std::vector<char> have; };
    std::map<int, Meas> meas;
    std::vector<int64_t> gubase;                        // first entry of a costed type with U-dofs in cGU, −1 otherwise
    double *nzval = nullptr, *Lv = nullptr;
};

static int32_t build_pattern(mb_handle* h, PairPat& P, bool rowU, bool colU, int64_t nrows, int64_t ncols) {
    std::vector<PatSide> rows, cols;
    for (const Group& g : h->groups) { rows.push_back({g.nele, rowU ? g.nu : g.nx, rowU ? g.idxU : g.idxX}); cols.push_back({g.nele, colU ? g.nu : g.nx, colU ? g.idxU : g.idxX}); }
    return build_pair_pattern(h, P, rows, cols, nrows, ncols);
}

// element types that contribute to L1[X][der]: the second-order branch (SoilContact, host-evaluated) and costed beams (Λᵀ∂R/∂X + ∇cost)
static inline bool has_gx(const Group& g) { return g.kind == G_SOIL || g.kind == G_HOST || (g.kind == G_BEAM && g.ng > 0); }
static BigDev make_bigdev(const DirectData* D) {
    BigDev B;
    B.nX = D->nX; B.nU = D->nU; B.W = 2 * D->nX + D->nU; B.nstep = D->nstep; B.lo = D->lo; B.hi = D->hi; B.elo = D->elo; B.OX = D->OX; B.OU = D->OU; B.dt = D->dt;
    B.bcolptr = D->bcolptr; B.browval = D->browval;
    for (int p = 0; p < 4; ++p) { B.pc[p] = D->pat[p].colptr0; B.pr[p] = D->pat[p].rowval0; B.pnnz[p] = D->pat[p].nnz; }
    B.LX = D->LX; B.XL = D->XL; B.LU = D->LU; B.UL = D->UL; B.L1L = D->L1L;
    B.sLX = (int64_t)(D->OX + 1) * D->pat[P_XX].nnz; B.sLU = D->pat[P_XU].nnz; B.sUL = D->pat[P_UX].nnz;
    B.hostc = D->hostc; B.L1X = D->L1X; B.ehi = D->ehi;
    B.XXc = D->XXc; B.L1U = D->L1U;
    return B;
}

void mb_direct_release(mb_handle* h) { if (h && h->direct) { delete h->direct; h->direct = nullptr; } }

extern "C" {

static int32_t upload_slrow(mb_handle* h);

int32_t mb_direct_prepare(mb_handle* h, int32_t OX, int32_t OU, int64_t ndofX, int64_t ndofU, int64_t nstep, int64_t step_lo, int64_t step_hi, double dt,
                          const int32_t* bcolptr, const int32_t* browval, int64_t* ncol_out, int64_t* nnz_out) {
    if (!h) return MB_ERR_ARG;
    ARG(!h->prepared && !h->direct, "already prepared");
    ARG(OX >= 0 && OX <= 2 && OU >= 0 && OU <= 2 && ndofX >= 1 && ndofU >= 0 && nstep >= 1 && step_lo >= 0 && step_hi > step_lo && step_hi <= nstep, "bad argument");
    ARG(bcolptr && browval, "block pattern missing");
    ARG((int)h->groups.size() <= MAXG, "too many element types");
    for (const Group& g : h->groups) ARG(g.kind == G_BEAM || g.kind == G_BAR || g.kind == G_SOIL || (g.kind == G_HOST && g.nu == 0), "DirectXUA on the device: EulerBeam3D, Bar3D, SoilContact and host-evaluated X-class element types");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    DirectData* D = new DirectData();
    h->direct = D;
    D->OX = OX; D->OU = OU; D->nX = ndofX; D->nU = ndofU; D->nstep = nstep; D->lo = step_lo; D->hi = step_hi; D->dt = dt;
    D->elo = step_lo - 2 < 0 ? 0 : step_lo - 2; D->ehi = step_hi + 2 > nstep ? nstep : step_hi + 2;
    { int32_t rcd = check_group_dofs(h, ndofX, ndofU); if (rcd) { delete D; h->direct = nullptr; return rcd; } }
    h->ndofX = ndofX; h->ndofU = ndofU;
    int32_t rc;
    if ((rc = build_pattern(h, D->pat[P_XX], false, false, ndofX, ndofX))) return rc;
    if ((rc = build_pattern(h, D->pat[P_XU], false, true, ndofX, ndofU))) return rc;
    if ((rc = build_pattern(h, D->pat[P_UX], true, false, ndofU, ndofX))) return rc;
    if ((rc = build_pattern(h, D->pat[P_UU], true, true, ndofU, ndofU))) return rc;
    if (D->pat[P_XX].nnz) {                              // transpose map of the (structurally symmetric) X-X pattern
        const PairPat& XX = D->pat[P_XX];
        CK(dalloc(h, &D->permT, XX.nnz));
        unsigned long long* bad = nullptr; CK(dalloc(h, &bad, 1)); CK(cudaMemsetAsync(bad, 0, 8, st));
        transpose_perm_kernel<<<nblk(ndofX, 256), 256, 0, st>>>(ndofX, XX.colptr0, XX.rowval0, D->permT, bad);
        h->launches++;
        unsigned long long nbad = 0; CK(cudaMemcpyAsync(&nbad, bad, 8, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); dfree(h, bad);
        ARG(nbad == 0, "internal: the X-X class pattern is not structurally symmetric");
    }
    // vector contributors (asmvec! for the Λ group)
    int64_t nvec = 0, ndr = 0;
    D->G.n = (int)h->groups.size();
    for (size_t ig = 0; ig < h->groups.size(); ++ig) {
        const Group& g = h->groups[ig];
        for (int p = 0; p < 4; ++p) D->G.pbase[p][ig] = (uint32_t)D->pat[p].gbase[ig];
        D->G.nx[ig] = g.nx; D->G.np[ig] = g.nx * (OX + 1) + (g.udof ? 3 : 0); D->G.udof[ig] = g.udof;
        D->G.drbase[ig] = ndr; D->G.rbase[ig] = nvec;
        ndr += g.nele * g.nx * D->G.np[ig]; nvec += g.nele * g.nx;
    }
    for (int p = 0; p < 4; ++p) D->G.pbase[p][h->groups.size()] = (uint32_t)D->pat[p].gbase[h->groups.size()];
    CK(dalloc(h, &D->vstart, ndofX + 1)); CK(dalloc(h, &D->vsrc, nvec));
    CK(cudaMemsetAsync(D->vstart, 0, (ndofX + 1) * sizeof(uint32_t), st));
    if (nvec > 0) {
        uint32_t *keys = nullptr, *keys2 = nullptr, *vals = nullptr;
        CK(dalloc(h, &keys, nvec)); CK(dalloc(h, &keys2, nvec)); CK(dalloc(h, &vals, nvec));
        for (size_t ig = 0; ig < h->groups.size(); ++ig) {
            const Group& g = h->groups[ig];
            if (g.nele == 0) continue;
            vec_keys_kernel<<<nblk(g.nele * g.nx, 256), 256, 0, st>>>(g.nele * g.nx, g.idxX, keys, vals, (uint32_t)D->G.rbase[ig]);
            h->launches++;
        }
        int end_bit = 1; while (end_bit < 32 && ((uint64_t)ndofX >> end_bit)) ++end_bit;
        void* tmp = nullptr; size_t tmpsz = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmpsz, keys, keys2, vals, D->vsrc, nvec, 0, end_bit, st));
        CK(cudaMalloc(&tmp, tmpsz ? tmpsz : 1));
        CK(cub::DeviceRadixSort::SortPairs(tmp, tmpsz, keys, keys2, vals, D->vsrc, nvec, 0, end_bit, st));
        vstart2_kernel<<<nblk(nvec, 256), 256, 0, st>>>(nvec, keys2, ndofX, D->vstart);
        h->launches++;
        CK(cudaStreamSynchronize(st)); cudaFree(tmp);
        dfree(h, keys); dfree(h, keys2); dfree(h, vals);
    }
    // second-order element types (SoilContact): contributor lists of their element dofs for L1[X][der]
    {
        int64_t nq = 0;
        for (size_t ig = 0; ig < h->groups.size(); ++ig) { D->G.gxbase[ig] = nq; if (has_gx(h->groups[ig])) nq += h->groups[ig].nele * h->groups[ig].nx; }
        D->ngx = nq;
        if (nq > 0) {
            CK(dalloc(h, &D->vstart2, ndofX + 1)); CK(dalloc(h, &D->vsrc2, nq)); CK(dalloc(h, &D->GX, nq * (OX + 1)));
            CK(cudaMemsetAsync(D->vstart2, 0, (ndofX + 1) * sizeof(uint32_t), st));
            uint32_t *keys = nullptr, *keys2 = nullptr, *vals = nullptr;
            CK(dalloc(h, &keys, nq)); CK(dalloc(h, &keys2, nq)); CK(dalloc(h, &vals, nq));
            for (size_t ig = 0; ig < h->groups.size(); ++ig) {
                const Group& g = h->groups[ig];
                if (!has_gx(g) || g.nele == 0) continue;
                vec_keys_kernel<<<nblk(g.nele * g.nx, 256), 256, 0, st>>>(g.nele * g.nx, g.idxX, keys, vals, (uint32_t)D->G.gxbase[ig]);
                h->launches++;
            }
            int end_bit = 1; while (end_bit < 32 && ((uint64_t)ndofX >> end_bit)) ++end_bit;
            void* tmp = nullptr; size_t tmpsz = 0;
            CK(cub::DeviceRadixSort::SortPairs(nullptr, tmpsz, keys, keys2, vals, D->vsrc2, nq, 0, end_bit, st));
            CK(cudaMalloc(&tmp, tmpsz ? tmpsz : 1));
            CK(cub::DeviceRadixSort::SortPairs(tmp, tmpsz, keys, keys2, vals, D->vsrc2, nq, 0, end_bit, st));
            vstart2_kernel<<<nblk(nq, 256), 256, 0, st>>>(nq, keys2, ndofX, D->vstart2);
            h->launches++;
            CK(cudaStreamSynchronize(st)); cudaFree(tmp);
            dfree(h, keys); dfree(h, keys2); dfree(h, vals);
            const int64_t nsx = D->ehi - D->elo;
            CK(dalloc(h, &D->L1X, nsx * (OX + 1) * ndofX));
            CK(cudaMemsetAsync(D->L1X, 0, (size_t)(nsx * (OX + 1) * ndofX) * sizeof(double), st));
        }
    }
    // buffers
    const int64_t ns = D->ehi - D->elo;
    // step batching: one launch set per `batch` time steps when the model is small (launch-bound otherwise: configs[4] has 100 beams × 4000 steps).
    // Aim at ≥ 65 536 element-steps per launch, within 1 GiB of extra element-output space; MB_DIRECT_BATCH overrides.
    {
        int64_t nel = 0; for (const Group& g : h->groups) nel += g.nele;
        int64_t want = nel > 0 ? (65536 + nel - 1) / nel : 1;
        int64_t ncst = 0; for (const Group& g : h->groups) if (g.kind == G_BEAM && g.ng > 0) ncst += g.nele;
        const int64_t bytes_per_step = 8 * (ndr + nvec + D->ngx * (OX + 1)) + 8 * 6 * (OX + 1) * nel * MB_NCOT + 8 * ncst * (48 + 4 + 12 + 144 + 1 + 3);
        const int64_t cap = std::max<int64_t>(1, (int64_t(1) << 30) / std::max<int64_t>(1, bytes_per_step));
        want = std::min<int64_t>(std::min<int64_t>(want, cap), std::min<int64_t>(ns, 65535));
        if (const char* eb = getenv("MB_DIRECT_BATCH")) want = std::max<int64_t>(1, std::min<int64_t>(atoll(eb), std::min<int64_t>(ns, 65535)));
        D->batch = std::max<int64_t>(1, want); D->ndr = (ndr + 1) & ~int64_t(1); D->nvec = (nvec + 1) & ~int64_t(1);
    }
    // costed beam types (ElementCost accelerator): scratch of one step, contributor lists of their U-dofs, per-stored-step L2[X,X][1,1] and L1[U][1]
    {
        int64_t nc = 0, nqu = 0;
        D->gubase.assign(h->groups.size(), -1);
        for (size_t ig = 0; ig < h->groups.size(); ++ig) {
            const Group& g = h->groups[ig];
            D->G.hxxbase[ig] = -1;
            if (g.kind != G_BEAM || g.ng == 0) continue;
            D->G.hxxbase[ig] = nc; nc += g.nele;
            if (g.udof) { D->gubase[ig] = nqu; nqu += g.nele * 3; }
        }
        D->ncost = nc; D->costed = nc > 0;
        if (D->costed) {
            const int64_t nb = D->batch;
            CK(dalloc(h, &D->cJ, nc * 48 * nb)); CK(dalloc(h, &D->ce4, nc * 4 * nb)); CK(dalloc(h, &D->cgX, nc * 12 * nb)); CK(dalloc(h, &D->cHXX, nc * 144 * nb)); CK(dalloc(h, &D->ccost, nc * nb));
            CK(dalloc(h, &D->XXc, ns * D->pat[P_XX].nnz)); CK(cudaMemsetAsync(D->XXc, 0, (size_t)(ns * D->pat[P_XX].nnz) * 8, st));
            if (nqu > 0 && ndofU > 0) {
                D->nqu = nqu;
                CK(dalloc(h, &D->cGU, nqu * D->batch)); CK(dalloc(h, &D->L1U, ns * ndofU)); CK(cudaMemsetAsync(D->L1U, 0, (size_t)(ns * ndofU) * 8, st));
                CK(dalloc(h, &D->vstartU, ndofU + 1)); CK(dalloc(h, &D->vsrcU, nqu));
                CK(cudaMemsetAsync(D->vstartU, 0, (ndofU + 1) * sizeof(uint32_t), st));
                uint32_t *keys = nullptr, *keys2 = nullptr, *vals = nullptr;
                CK(dalloc(h, &keys, nqu)); CK(dalloc(h, &keys2, nqu)); CK(dalloc(h, &vals, nqu));
                for (size_t ig = 0; ig < h->groups.size(); ++ig) {
                    const Group& g = h->groups[ig];
                    if (D->gubase[ig] < 0 || g.nele == 0) continue;
                    vec_keys_kernel<<<nblk(g.nele * 3, 256), 256, 0, st>>>(g.nele * 3, g.idxU, keys, vals, (uint32_t)D->gubase[ig]);
                    h->launches++;
                }
                int end_bit = 1; while (end_bit < 32 && ((uint64_t)ndofU >> end_bit)) ++end_bit;
                void* tmp = nullptr; size_t tmpsz = 0;
                CK(cub::DeviceRadixSort::SortPairs(nullptr, tmpsz, keys, keys2, vals, D->vsrcU, nqu, 0, end_bit, st));
                CK(cudaMalloc(&tmp, tmpsz ? tmpsz : 1));
                CK(cub::DeviceRadixSort::SortPairs(tmp, tmpsz, keys, keys2, vals, D->vsrcU, nqu, 0, end_bit, st));
                vstart2_kernel<<<nblk(nqu, 256), 256, 0, st>>>(nqu, keys2, ndofU, D->vstartU);
                h->launches++;
                CK(cudaStreamSynchronize(st)); cudaFree(tmp);
                dfree(h, keys); dfree(h, keys2); dfree(h, vals);
            }
        }
    }
    CK(dalloc(h, &D->dR, D->ndr * D->batch)); CK(dalloc(h, &D->R, D->nvec * D->batch));
    if (D->ngx > 0 && D->batch > 1) { dfree(h, D->GX); CK(dalloc(h, &D->GX, ((D->ngx * (OX + 1) + 1) & ~int64_t(1)) * D->batch)); }
    CK(dalloc(h, &D->X, ns * 3 * ndofX)); CK(dalloc(h, &D->U, ns * (ndofU > 0 ? ndofU : 1)));
    CK(dalloc(h, &D->Lam, ns * ndofX)); CK(cudaMemsetAsync(D->Lam, 0, (size_t)(ns * ndofX) * sizeof(double), st));
    CK(cudaMemsetAsync(D->X, 0, (size_t)(ns * 3 * ndofX) * sizeof(double), st));
    CK(cudaMemsetAsync(D->U, 0, (size_t)(ns * (ndofU > 0 ? ndofU : 1)) * sizeof(double), st));
    CK(dalloc(h, &D->LX, ns * (OX + 1) * D->pat[P_XX].nnz)); CK(dalloc(h, &D->XL, ns * (OX + 1) * D->pat[P_XX].nnz));
    CK(dalloc(h, &D->LU, ns * D->pat[P_XU].nnz)); CK(dalloc(h, &D->UL, ns * D->pat[P_UX].nnz)); CK(dalloc(h, &D->L1L, ns * ndofX));
    // block pattern of the owned block columns and the Lvv structure
    const int64_t nbc = 3 * (step_hi - step_lo);
    const int32_t nblocks = bcolptr[nbc];
    D->maxb = 0; for (int64_t b = 0; b < nbc; ++b) D->maxb = std::max<int>(D->maxb, bcolptr[b + 1] - bcolptr[b]);
    CK(dalloc(h, &D->bcolptr, nbc + 1)); CK(dalloc(h, &D->browval, nblocks));
    CK(cudaMemcpyAsync(D->bcolptr, bcolptr, (nbc + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(D->browval, browval, (size_t)nblocks * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    D->ncol = (step_hi - step_lo) * (2 * ndofX + ndofU);
    CK(dalloc(h, &D->colptr, D->ncol + 1));
    int64_t* cnt = nullptr; CK(dalloc(h, &cnt, D->ncol + 1));
    CK(cudaMemsetAsync(cnt, 0, (D->ncol + 1) * sizeof(int64_t), st));
    BigDev B = make_bigdev(D);
    big_count_kernel<<<nblk(D->ncol, 256), 256, 0, st>>>(B, D->ncol, cnt);
    h->launches++;
    void* tmp = nullptr; size_t tmpsz = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmpsz, cnt, D->colptr, D->ncol + 1, st));
    CK(cudaMalloc(&tmp, tmpsz ? tmpsz : 1));
    CK(cub::DeviceScan::ExclusiveSum(tmp, tmpsz, cnt, D->colptr, D->ncol + 1, st));
    CK(cudaMemcpyAsync(&D->nnzbig, D->colptr + D->ncol, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st)); cudaFree(tmp); dfree(h, cnt);
    CK(dalloc(h, &D->rowval, D->nnzbig)); CK(dalloc(h, &D->nzval, D->nnzbig)); CK(dalloc(h, &D->Lv, D->ncol));
    big_fill_kernel<true><<<nblk(D->ncol * 32, 256), 256, 0, st>>>(B, D->ncol, D->colptr, D->rowval, nullptr);
    h->launches++;
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    h->prepared = true;
    if (ncol_out) *ncol_out = D->ncol;
    if (nnz_out) *nnz_out = D->nnzbig;
    return upload_slrow(h);
}

int32_t mb_direct_class_pattern(mb_handle* h, int32_t which, int64_t* nnz, int64_t* colptr, int64_t* rowval) {
    if (!h || !h->direct) return MB_ERR_ARG;
    ARG(which >= 0 && which < 4, "pattern id: 0 XX, 1 XU, 2 UX, 3 UU");
    CK(cudaSetDevice(h->device));
    const PairPat& P = h->direct->pat[which];
    if (nnz) *nnz = P.nnz;
    if (colptr) { std::vector<int32_t> t((size_t)P.n + 1); CK(cudaMemcpy(t.data(), P.colptr0, t.size() * 4, cudaMemcpyDeviceToHost)); for (size_t i = 0; i < t.size(); ++i) colptr[i] = (int64_t)t[i] + 1; }
    if (rowval && P.nnz) { std::vector<int32_t> t((size_t)P.nnz); CK(cudaMemcpy(t.data(), P.rowval0, t.size() * 4, cudaMemcpyDeviceToHost)); for (size_t i = 0; i < t.size(); ++i) rowval[i] = (int64_t)t[i] + 1; }
    return MB_OK;
}
int32_t mb_direct_get_asm(mb_handle* h, int32_t ieletyp, int32_t which, int64_t* out) {
    if (!h || !h->direct) return MB_ERR_ARG;
    ARG(which >= 0 && which < 4 && ieletyp >= 1 && ieletyp <= (int32_t)h->groups.size() && out, "bad argument");
    CK(cudaSetDevice(h->device));
    const PairPat& P = h->direct->pat[which];
    const int64_t n = P.gbase[ieletyp] - P.gbase[ieletyp - 1];
    std::vector<int32_t> t((size_t)n);
    if (n) CK(cudaMemcpy(t.data(), P.asmK + P.gbase[ieletyp - 1], (size_t)n * 4, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; ++i) out[i] = t[(size_t)i];
    return MB_OK;
}
int32_t mb_direct_set_state(mb_handle* h, int64_t step, const double* X0, const double* X1, const double* X2, const double* U0) {
    if (!h || !h->direct) return MB_ERR_ARG;
    DirectData* D = h->direct;
    ARG(step >= D->elo && step < D->ehi && X0, "step not stored on this handle");
    CK(cudaSetDevice(h->device));
    double* x = D->X + (step - D->elo) * 3 * D->nX;
    const double* src[3] = {X0, X1, X2};
    for (int d = 0; d <= D->OX; ++d) { ARG(src[d], "state derivative missing"); CK(cudaMemcpyAsync(x + d * D->nX, src[d], (size_t)D->nX * 8, cudaMemcpyDefault, h->stream)); }
    if (U0 && D->nU) CK(cudaMemcpyAsync(D->U + (step - D->elo) * D->nU, U0, (size_t)D->nU * 8, cudaMemcpyDefault, h->stream));
    return MB_OK;
}

static int32_t direct_eval_steps(mb_handle* h, int64_t s0, int64_t s1) {
    DirectData* D = h->direct;
    cudaStream_t st = h->stream;
    const int nd = D->OX + 1;
    // Step batching: the steps are independent (src/DirectXUA.jl:328-331), so `nb` of them share one launch set (grid row = step): element outputs of the
    // batch lie [step][…] in dR / R / GX / Wc, the per-step blocks are contiguous over the stored steps anyway.  nb = 1 for large models (D->batch).
    const int64_t ndr = D->ndr, nvec = D->nvec, ngxd = (D->ngx * nd + 1) & ~int64_t(1);      // per-step strides, even: the kernels' 16-byte stores keep their alignment
    for (int64_t s = s0; s < s1; s += D->batch) {
        const int nb = (int)std::min<int64_t>(D->batch, s1 - s);
        const int64_t k = s - D->elo;
        DirectStateDev sd;
        for (int d = 0; d < 3; ++d) sd.X[d] = D->X + (k * 3 + d) * D->nX;
        sd.U0 = D->U + k * D->nU;
        StepBatch sb;
        sb.sX = 3 * D->nX; sb.sU = D->nU; sb.sLam = D->nX; sb.sdR = ndr; sb.sR = nvec; sb.sGX = ngxd; sb.snan = nan_pack(1, 0); sb.dt = D->dt;
        for (size_t ig = 0; ig < h->groups.size(); ++ig) {
            const Group& g = h->groups[ig];
            if (g.nele == 0) continue;
            const unsigned long long nanbase = nan_pack(s, (int)ig);
            double* dR = D->dR + D->G.drbase[ig]; double* R = D->R + D->G.rbase[ig];
            if (g.kind == G_BAR) {
                BarGroupDev gd; gd.nele = g.nele; gd.geo = g.geo; gd.mats = g.barmats; gd.mat_id = g.mat_id; gd.idxX = g.idxX; gd.idxU = g.idxU; gd.udof = g.udof;
                for (int i = 0; i < 6; ++i) gd.scaleX[i] = g.scaleX[i];
                for (int i = 0; i < 3; ++i) gd.scaleU[i] = g.scaleU[i];
                h->launches += launch_bar_direct(nd, gd, sd, D->t0 + (double)s * D->dt, dR, R, h->nanflag, nanbase, st, sb, nb);
                continue;
            }
            if (g.kind == G_HOST) {                // contributions uploaded by mb_direct_set_host_elements (zeros until then): [stored step][R | dR | GX]
                const int64_t nR = g.nele * g.nx, ndR = nR * g.nx * nd, nG = nR * nd, per = nR + ndR + nG;
                const double* src = (ig < D->hoststore.size() && D->hoststore[ig]) ? D->hoststore[ig] + k * per : nullptr;
                double* gx = D->GX + D->G.gxbase[ig] * nd;
                if (src) {
                    CK(cudaMemcpy2DAsync(R, (size_t)nvec * 8, src, (size_t)per * 8, (size_t)nR * 8, nb, cudaMemcpyDeviceToDevice, st));
                    CK(cudaMemcpy2DAsync(dR, (size_t)ndr * 8, src + nR, (size_t)per * 8, (size_t)ndR * 8, nb, cudaMemcpyDeviceToDevice, st));
                    CK(cudaMemcpy2DAsync(gx, (size_t)ngxd * 8, src + nR + ndR, (size_t)per * 8, (size_t)nG * 8, nb, cudaMemcpyDeviceToDevice, st));
                } else {
                    CK(cudaMemset2DAsync(R, (size_t)nvec * 8, 0, (size_t)nR * 8, nb, st)); CK(cudaMemset2DAsync(dR, (size_t)ndr * 8, 0, (size_t)ndR * 8, nb, st));
                    CK(cudaMemset2DAsync(gx, (size_t)ngxd * 8, 0, (size_t)nG * 8, nb, st));
                }
                continue;
            }
            if (g.kind == G_SOIL) {
                SoilGroupDev gd; gd.nele = g.nele; gd.par = g.geo; gd.idxX = g.idxX;
                for (int i = 0; i < 3; ++i) gd.scaleX[i] = g.scaleX[i];
                h->launches += launch_soil_direct(nd, gd, sd, D->Lam + k * D->nX, D->lamscale, dR, R, D->GX + D->G.gxbase[ig] * nd, h->nanflag, nanbase, st, sb, nb);
                continue;
            }
            BeamGroupDev gd;
            gd.nele = g.nele; gd.geo = g.geo; gd.mats = g.mats; gd.mat_id = g.mat_id; gd.idxX = g.idxX; gd.idxU = g.idxU; gd.udof = g.udof;
            for (int i = 0; i < 12; ++i) gd.scaleX[i] = g.scaleX[i];
            for (int i = 0; i < 3; ++i) gd.scaleU[i] = g.scaleU[i];
            double* Wc = nullptr;
            if (nd >= 2) {                         // cotangent workspace of the two-phase evaluation, shared with SweepX and sized lazily
                const int64_t one = ((g.nele * 6 * nd + 31) / 32) * 32 * MB_NCOT, need = one * D->batch;
                if (h->Wc_len < need) { if (h->Wc) dfree(h, h->Wc); h->Wc = nullptr; h->Wc_len = 0; CK(dalloc(h, &h->Wc, need)); h->Wc_len = need; }
                Wc = h->Wc; sb.sWc = one;
            }
            if (nd == 1) h->launches += launch_beam_direct<1>(gd, sd, dR, R, h->nanflag, nanbase, Wc, st, sb, nb);
            else if (nd == 2) h->launches += launch_beam_direct<2>(gd, sd, dR, R, h->nanflag, nanbase, Wc, st, sb, nb);
            else h->launches += launch_beam_direct<3>(gd, sd, dR, R, h->nanflag, nanbase, Wc, st, sb, nb);
            if (g.ng > 0) {                        // ElementCost accelerator: strains and their Jacobian, the cost's gradient and Gauss-Newton block, then the Lagrangian's own terms
                auto mit = D->meas.find((int)ig);
                ARG(mit != D->meas.end() && mit->second.eps, "strain-gauge measurements of a stored step are not set (mb_direct_set_gauge_measurements)");
                const DirectData::Meas& m = mit->second;
                for (int b = 0; b < nb; ++b) ARG(m.have[(size_t)(k + b)], "strain-gauge measurements of a stored step are not set (mb_direct_set_gauge_measurements)");
                const int64_t eb = D->G.hxxbase[ig], nc = D->ncost;
                beam_gauge_kernel<<<dim3(nblk(g.nele * 6, 128), nb), 128, 0, st>>>(gd, sd.X[0], D->cJ + eb * 48, D->ce4 + eb * 4, sb.sX, nc);
                gauge_cost_kernel<<<dim3(nblk(g.nele * 12, 128), nb), 128, 0, st>>>(g.nele, g.ng, g.gaugeG, m.eps + k * m.stride, m.per_element, g.isig2, D->cJ + eb * 48, D->ce4 + eb * 4,
                                                                                    D->cgX + eb * 12, D->cHXX + eb * 144, D->ccost + eb, m.stride, nc);
                SL12 sl; for (int i = 0; i < 12; ++i) sl.v[i] = g.scaleX[i] * D->lamscale;
                costed_fill_kernel<<<dim3((unsigned)((g.nele + 3) / 4), nb), 128, 0, st>>>(g.nele, nd, D->G.np[ig], g.idxX, D->Lam + k * D->nX, sl, D->cgX + eb * 12, R, dR, D->GX + D->G.gxbase[ig] * nd,
                                                                               D->gubase[ig] >= 0 && D->cGU ? D->cGU + D->gubase[ig] : nullptr, sb, nc, D->nqu);
                h->launches += 3;
            }
        }
        if (D->elements_only) continue;
        const PairPat& XX = D->pat[P_XX];
        if (XX.nnz) { gather_xx_kernel<<<dim3(nblk(XX.nnz, 256), nb), 256, 0, st>>>(XX.nnz, XX.cstart, XX.src, D->G, nd, D->dR, D->LX + k * nd * XX.nnz, D->XL + k * nd * XX.nnz, ndr, D->permT, D->slrow); h->launches++; }
        const PairPat& XU = D->pat[P_XU];
        if (XU.nnz) { gather_xu_kernel<<<dim3(nblk(XU.nnz, 256), nb), 256, 0, st>>>(XU.nnz, XU.cstart, XU.src, D->G, nd, 0, D->dR, D->LU + k * XU.nnz, ndr, D->slrow); h->launches++; }
        const PairPat& UX = D->pat[P_UX];
        if (UX.nnz) { gather_xu_kernel<<<dim3(nblk(UX.nnz, 256), nb), 256, 0, st>>>(UX.nnz, UX.cstart, UX.src, D->G, nd, 1, D->dR, D->UL + k * UX.nnz, ndr, D->slrow); h->launches++; }
        gather_l1_kernel<<<dim3(nblk(D->nX, 256), nb), 256, 0, st>>>(D->nX, D->vstart, D->vsrc, D->R, D->L1L + k * D->nX, nvec);
        h->launches++;
        if (D->ngx) { gather_l1x_kernel<<<dim3(nblk(D->nX, 256), nb), 256, 0, st>>>(D->nX, D->vstart2, D->vsrc2, D->GX, nd, D->L1X + k * nd * D->nX, ngxd); h->launches++; }
        if (D->costed) {
            if (XX.nnz) { gather_xxc_kernel<<<dim3(nblk(XX.nnz, 256), nb), 256, 0, st>>>(XX.nnz, XX.cstart, XX.src, D->G, D->cHXX, D->XXc + k * XX.nnz, D->ncost); h->launches++; }
            if (D->L1U) { gather_l1u_kernel<<<dim3(nblk(D->nU, 256), nb), 256, 0, st>>>(D->nU, D->vstartU, D->vsrcU, D->cGU, D->L1U + k * D->nU, D->nqu); h->launches++; }
        }
    }
    return MB_OK;
}

/* ElementCost{StrainGaugeOnEulerBeam3D} with cost(eleres,t) = Σ_g (ε_g − εm_g(t))²/(2σ²) on a beam type of the windowed path (the accelerator of src/DirectXUA.jl:172-198 as it is
 * meant, see mb_xua_set_gauge_cost): call between mb_add_eulerbeam3d and mb_direct_prepare.  G [ngauge][4]: ε_g = G[g]·(εₐₓ, κ₁, κ₂, κ₃) (toolbox/StrainGaugeOnBeamElement.jl:70-76). */
int32_t mb_direct_set_gauge_cost(mb_handle* h, int32_t ieletyp, int32_t ngauge, const double* G, double sigma) {
    if (!h) return MB_ERR_ARG;
    ARG(!h->prepared && !h->direct, "set the gauge cost before mb_direct_prepare");
    ARG(ieletyp >= 1 && ieletyp <= (int32_t)h->groups.size() && h->groups[(size_t)ieletyp - 1].kind == G_BEAM, "not an EulerBeam3D element type");
    ARG(ngauge >= 1 && ngauge <= 64 && G && sigma > 0., "bad gauge data");
    CK(cudaSetDevice(h->device));
    Group& g = h->groups[(size_t)ieletyp - 1];
    if (g.gaugeG) dfree(h, g.gaugeG);
    g.ng = ngauge; g.isig2 = 1. / (sigma * sigma);
    CK(dalloc(h, &g.gaugeG, (int64_t)ngauge * 4));
    CK(cudaMemcpy(g.gaugeG, G, (size_t)ngauge * 4 * 8, cudaMemcpyDefault));
    return MB_OK;
}
/* Measured strains of one stored step: epsm [ngauge] shared by all elements of the type, or [nele][ngauge] (per_element; the choice is fixed by the first call).
 * Kept per stored step (they move with mb_direct_rebase), to be set once for every step of a window before mb_direct_assemble. */
int32_t mb_direct_set_gauge_measurements(mb_handle* h, int64_t step, int32_t ieletyp, const double* epsm, int32_t per_element) {
    if (!h || !h->direct) return MB_ERR_ARG;
    DirectData* D = h->direct;
    ARG(ieletyp >= 1 && ieletyp <= (int32_t)h->groups.size() && h->groups[(size_t)ieletyp - 1].ng > 0 && epsm, "set the gauge cost of this element type first (mb_direct_set_gauge_cost)");
    ARG(step >= D->elo && step < D->ehi, "step not stored on this handle");
    CK(cudaSetDevice(h->device));
    const Group& g = h->groups[(size_t)ieletyp - 1];
    DirectData::Meas& m = D->meas[ieletyp - 1];
    const int64_t ns = D->ehi - D->elo;
    if (!m.eps) {
        m.per_element = per_element != 0; m.stride = (int64_t)g.ng * (m.per_element ? std::max<int64_t>(g.nele, 1) : 1);
        CK(dalloc(h, &m.eps, ns * m.stride)); m.have.assign((size_t)ns, 0);
    }
    ARG(m.per_element == (per_element != 0), "per_element differs from the first call for this element type");
    CK(cudaMemcpyAsync(m.eps + (step - D->elo) * m.stride, epsm, (size_t)m.stride * 8, cudaMemcpyDefault, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    m.have[(size_t)(step - D->elo)] = 1;
    return MB_OK;
}
static int32_t upload_slrow(mb_handle* h) {
    DirectData* D = h->direct;
    if (!D->costed) return MB_OK;
    std::vector<double> sl(h->groups.size() * 12, 1.);
    for (size_t ig = 0; ig < h->groups.size(); ++ig)
        if (h->groups[ig].kind == G_BEAM && h->groups[ig].ng > 0) for (int i = 0; i < 12; ++i) sl[ig * 12 + i] = h->groups[ig].scaleX[i] * D->lamscale;
    if (!D->slrow) CK(dalloc(h, &D->slrow, (int64_t)sl.size()));
    CK(cudaMemcpyAsync(D->slrow, sl.data(), sl.size() * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MB_OK;
}
int32_t mb_direct_set_lambda_scale(mb_handle* h, double lambda_scale) {
    if (!h || !h->direct) return MB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    h->direct->lamscale = lambda_scale;
    return upload_slrow(h);
}
int32_t mb_direct_set_time0(mb_handle* h, double t0) {
    if (!h || !h->direct) return MB_ERR_ARG;
    h->direct->t0 = t0;
    return MB_OK;
}

// assemblebig!{:matrices}: evaluate steps [eval_lo,eval_hi) (default: all stored steps, i.e. owned + halo), then build the owned columns of Lvv and Lv.
int32_t mb_direct_assemble(mb_handle* h, int64_t eval_lo, int64_t eval_hi, int32_t build_big, double* Lvv_nzval, double* Lv, mb_errinfo* where) {
    if (!h || !h->direct) return MB_ERR_ARG;
    DirectData* D = h->direct;
    CK(cudaSetDevice(h->device));
    if (eval_lo < 0) { eval_lo = D->elo; eval_hi = D->ehi; }
    ARG(eval_lo >= D->elo && eval_hi <= D->ehi && eval_lo <= eval_hi, "evaluation range outside the stored steps");
    CK(cudaMemsetAsync(h->nanflag, 0xFF, sizeof(unsigned long long), h->stream));
    int32_t rc = direct_eval_steps(h, eval_lo, eval_hi);
    if (rc) return rc;
    if (build_big) {
        BigDev B = make_bigdev(D);
        launch_big_values(B, D->ncol, D->colptr, D->nzval, D->maxb, h->stream);
        big_vec_kernel<<<nblk(D->ncol, 256), 256, 0, h->stream>>>(B, D->ncol, D->Lv);
        if (D->hostc) { big_diag_kernel<<<nblk(D->ncol, 256), 256, 0, h->stream>>>(B, D->ncol, D->colptr, D->rowval, D->nzval, D->missing, D->rowshift); h->launches++; }
        for (const auto& kv : D->hostxx) {
            if (kv.first < D->lo || kv.first >= D->hi || kv.second.n == 0) continue;       // halo steps: their columns belong to a neighbour
            big_xx_kernel<<<nblk(kv.second.n, 256), 256, 0, h->stream>>>(B, kv.second.n, kv.second.i, kv.second.j, kv.second.v, kv.first, D->colptr, D->rowval, D->nzval,
                                                                          D->missing, D->rowshift);
            h->launches++;
        }
        h->launches += 2;
        if (Lvv_nzval) CK(cudaMemcpyAsync(Lvv_nzval, D->nzval, (size_t)D->nnzbig * 8, cudaMemcpyDeviceToHost, h->stream));
        if (Lv) CK(cudaMemcpyAsync(Lv, D->Lv, (size_t)D->ncol * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaGetLastError());
    rc = mb_sync(h, where);
    if (rc == MB_OK && build_big && D->missing) {
        unsigned long long miss = 0;
        CK(cudaMemcpy(&miss, D->missing, 8, cudaMemcpyDeviceToHost));
        if (miss) { CK(cudaMemset(D->missing, 0, 8)); h->err = "a host-evaluated cost or X-X entry is not in the Lvv pattern (no element added with mb_add_* couples these dofs)"; return MB_ERR_ARG; }
    }
    if (rc == MB_ERR_NAN && where) {     // nanbase packs (step, ieletyp, iele)
        const unsigned long long f = *h->nanflag_host;
        nan_unpack(f, where);
    }
    return rc;
}
int32_t mb_direct_big_pattern(mb_handle* h, int64_t* colptr, int64_t* rowval) {
    if (!h || !h->direct) return MB_ERR_ARG;
    DirectData* D = h->direct;
    CK(cudaSetDevice(h->device));
    if (colptr) { CK(cudaMemcpy(colptr, D->colptr, (size_t)(D->ncol + 1) * 8, cudaMemcpyDeviceToHost)); for (int64_t i = 0; i <= D->ncol; ++i) colptr[i] += 1; }
    if (rowval) { CK(cudaMemcpy(rowval, D->rowval, (size_t)D->nnzbig * 8, cudaMemcpyDeviceToHost)); for (int64_t i = 0; i < D->nnzbig; ++i) rowval[i] += 1 + D->rowshift; }
    return MB_OK;
}
// Host-evaluated X-class element types (Hold, DofLoad, DofConstraint: user closures, default no_second_order = Val(false)) in the second-order
// branch of DirectXUA (src/DirectXUA.jl:152-171), for residuals that are LINEAR in X (L2[X,X] = Λ·∂²R/∂X² = 0). Per stored step and element:
//   R  [nele][nx]            Rᵢ·scale.Λᵢ                          → L1[Λ]
//   dR [nele][nd·nx][nx]     ∂Rᵢ/∂X_der,ⱼ·scale.Λᵢ·scale.Xⱼ  (entry [(nx·der+j)][i]) → L2[Λ,X][1,der+1] and, transposed, L2[X,Λ][der+1,1]
//   GX [nele][nx][nd]        Σₖ Λₖ·∂Rₖ/∂X_der,ᵢ·scale.Xᵢ        → L1[X][der+1]
int32_t mb_direct_set_host_elements(mb_handle* h, int64_t step, int32_t ieletyp, const double* R, const double* dR, const double* GX) {
    if (!h || !h->direct) return MB_ERR_ARG;
    DirectData* D = h->direct;
    ARG(step >= D->elo && step < D->ehi, "step not stored on this handle");
    ARG(ieletyp >= 1 && ieletyp <= (int32_t)h->groups.size() && h->groups[ieletyp - 1].kind == G_HOST, "not a host-evaluated element type");
    CK(cudaSetDevice(h->device));
    const size_t ig = (size_t)ieletyp - 1;
    const Group& g = h->groups[ig];
    const int nd = D->OX + 1;
    const int64_t nR = g.nele * g.nx, ndR = nR * g.nx * nd, nG = nR * nd, per = nR + ndR + nG, ns = D->ehi - D->elo;
    if (D->hoststore.size() < h->groups.size()) D->hoststore.resize(h->groups.size(), nullptr);
    if (!D->hoststore[ig]) { CK(dalloc(h, &D->hoststore[ig], ns * per)); CK(cudaMemsetAsync(D->hoststore[ig], 0, (size_t)(ns * per) * 8, h->stream)); }
    double* dst = D->hoststore[ig] + (step - D->elo) * per;
    const double* src[3] = {R, dR, GX}; const int64_t n[3] = {nR, ndR, nG}; const int64_t off[3] = {0, nR, nR + ndR};
    for (int q = 0; q < 3; ++q) {
        if (n[q] == 0) continue;
        if (src[q]) CK(cudaMemcpyAsync(dst + off[q], src[q], (size_t)n[q] * 8, cudaMemcpyDefault, h->stream));
        else CK(cudaMemsetAsync(dst + off[q], 0, (size_t)n[q] * 8, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return MB_OK;
}
// Host-evaluated element types whose residual is NOT linear in X (DofConstraint in `positive` mode, curved gaps): the X-X part of the second-order branch
// (src/DirectXUA.jl:121-150), L2[X,X][1,1](i,j) += Σₖ Λₖ·∂²Rₖ/∂Xᵢ∂Xⱼ·scale.Xᵢ·scale.Xⱼ, as triplets with 1-based model X dofs — ONE triplet per (i,j), the caller
// sums the contributions of its elements in element order.  Replaces what was set for this step; n = 0 clears it.
int32_t mb_direct_set_host_xx(mb_handle* h, int64_t step, int64_t n, const int64_t* i, const int64_t* j, const double* v) {
    if (!h || !h->direct) return MB_ERR_ARG;
    DirectData* D = h->direct;
    ARG(step >= D->elo && step < D->ehi && n >= 0 && (n == 0 || (i && j && v)), "step not stored on this handle / bad triplets");
    CK(cudaSetDevice(h->device));
    DirectData::HostXX& x = D->hostxx[step];
    if (x.i) { dfree(h, x.i); dfree(h, x.j); dfree(h, x.v); x = DirectData::HostXX(); }
    if (n == 0) { D->hostxx.erase(step); return MB_OK; }
    std::vector<int32_t> ii((size_t)n), jj((size_t)n);
    for (int64_t q = 0; q < n; ++q) {
        ARG(i[q] >= 1 && i[q] <= D->nX && j[q] >= 1 && j[q] <= D->nX, "X dof out of range");
        ii[(size_t)q] = (int32_t)(i[q] - 1); jj[(size_t)q] = (int32_t)(j[q] - 1);
    }
    CK(dalloc(h, &x.i, n)); CK(dalloc(h, &x.j, n)); CK(dalloc(h, &x.v, n));
    CK(cudaMemcpyAsync(x.i, ii.data(), (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(x.j, jj.data(), (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(x.v, v, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    x.n = n;
    if (!D->missing) { CK(dalloc(h, &D->missing, 1)); CK(cudaMemsetAsync(D->missing, 0, 8, h->stream)); }
    CK(cudaStreamSynchronize(h->stream));
    return MB_OK;
}
// host-evaluated single-dof costs of one stored step (SingleDofCost on X or U dofs, src/BasicElements.jl:198-208): gradient → L1[X][1] / L1[U][1],
// second derivative → the diagonal of L2[X,X][1,1] / L2[U,U][1,1]; dense per-dof vectors, already multiplied by the dof scales; NULL = zeros
int32_t mb_direct_set_host_cost(mb_handle* h, int64_t step, const double* gX, const double* hX, const double* gU, const double* hU) {
    if (!h || !h->direct) return MB_ERR_ARG;
    DirectData* D = h->direct;
    ARG(step >= D->elo && step < D->ehi, "step not stored on this handle");
    CK(cudaSetDevice(h->device));
    const int64_t per = 2 * (D->nX + D->nU), ns = D->ehi - D->elo;
    if (!D->hostc) {
        CK(dalloc(h, &D->hostc, ns * per)); CK(cudaMemsetAsync(D->hostc, 0, (size_t)(ns * per) * 8, h->stream));
        if (!D->missing) { CK(dalloc(h, &D->missing, 1)); CK(cudaMemsetAsync(D->missing, 0, 8, h->stream)); }
    }
    double* base = D->hostc + (step - D->elo) * per;
    const double* src[4] = {gX, hX, gU, hU}; const int64_t off[4] = {0, D->nX, 2 * D->nX, 2 * D->nX + D->nU}; const int64_t n[4] = {D->nX, D->nX, D->nU, D->nU};
    for (int k = 0; k < 4; ++k) {
        if (n[k] == 0) continue;
        if (src[k]) CK(cudaMemcpyAsync(base + off[k], src[k], (size_t)n[k] * 8, cudaMemcpyDefault, h->stream));
        else CK(cudaMemsetAsync(base + off[k], 0, (size_t)n[k] * 8, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return MB_OK;
}
// ---- Newton update of the all-steps problem on the device
int32_t mb_direct_set_lambda(mb_handle* h, int64_t step, const double* Lambda) {
    if (!h || !h->direct) return MB_ERR_ARG;
    DirectData* D = h->direct;
    ARG(step >= D->elo && step < D->ehi && Lambda, "step not stored on this handle");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(D->Lam + (step - D->elo) * D->nX, Lambda, (size_t)D->nX * 8, cudaMemcpyDefault, h->stream));
    return MB_OK;
}
int32_t mb_direct_get_state(mb_handle* h, int64_t step, double* X0, double* X1, double* X2, double* U0, double* Lambda) {
    if (!h || !h->direct) return MB_ERR_ARG;
    DirectData* D = h->direct;
    ARG(step >= D->elo && step < D->ehi, "step not stored on this handle");
    CK(cudaSetDevice(h->device));
    const int64_t k = step - D->elo;
    double* dst[3] = {X0, X1, X2};
    for (int d = 0; d <= D->OX; ++d) if (dst[d]) CK(cudaMemcpyAsync(dst[d], D->X + (k * 3 + d) * D->nX, (size_t)D->nX * 8, cudaMemcpyDefault, h->stream));
    if (U0 && D->nU) CK(cudaMemcpyAsync(U0, D->U + k * D->nU, (size_t)D->nU * 8, cudaMemcpyDefault, h->stream));
    if (Lambda) CK(cudaMemcpyAsync(Lambda, D->Lam + k * D->nX, (size_t)D->nX * 8, cudaMemcpyDefault, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MB_OK;
}
int32_t mb_direct_set_dof_scale(mb_handle* h, const double* scaleL, const double* scaleX, const double* scaleU) {
    if (!h || !h->direct) return MB_ERR_ARG;
    DirectData* D = h->direct;
    CK(cudaSetDevice(h->device));
    const double* src[3] = {scaleL, scaleX, scaleU}; double** dst[3] = {&D->scL, &D->scX, &D->scU}; const int64_t n[3] = {D->nX, D->nX, D->nU};
    for (int c = 0; c < 3; ++c) {
        if (!src[c] || n[c] == 0) continue;
        if (!*dst[c]) CK(dalloc(h, dst[c], n[c]));
        CK(cudaMemcpyAsync(*dst[c], src[c], (size_t)n[c] * 8, cudaMemcpyDefault, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return MB_OK;
}
int32_t mb_direct_decrement(mb_handle* h, int64_t s0, int64_t s1, const double* dv, double* delta2) {
    if (!h || !h->direct) return MB_ERR_ARG;
    DirectData* D = h->direct;
    ARG(dv && s0 >= 0 && s1 <= D->nstep && s0 < s1, "bad step range");
    ARG(D->OU == 0, "device decrementbig! stores ∂0(U) only (OU = 0)");
    const int64_t need0 = D->elo - 2 < 0 ? 0 : D->elo - 2, need1 = D->ehi + 2 > D->nstep ? D->nstep : D->ehi + 2;
    ARG(s0 <= need0 && s1 >= need1, "Δv must cover the stored steps and two steps on either side (finite-difference stencils)");
    CK(cudaSetDevice(h->device));
    const int64_t W = 2 * D->nX + D->nU, n = (s1 - s0) * W;
    cudaPointerAttributes at;
    const bool on_dev = cudaPointerGetAttributes(&at, dv) == cudaSuccess && at.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    const double* d = dv;
    if (!on_dev) {
        if (D->dvlen < n) { if (D->dvbuf) dfree(h, D->dvbuf); D->dvbuf = nullptr; D->dvlen = 0; CK(dalloc(h, &D->dvbuf, n)); D->dvlen = n; }
        CK(cudaMemcpyAsync(D->dvbuf, dv, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
        d = D->dvbuf;
    }
    DecDev Q;
    Q.nX = D->nX; Q.nU = D->nU; Q.W = W; Q.nstep = D->nstep; Q.elo = D->elo; Q.ehi = D->ehi; Q.s0 = s0; Q.s1 = s1; Q.OX = D->OX;
    Q.dtp[0] = 1.; Q.dtp[1] = 1. / D->dt; Q.dtp[2] = (1. / D->dt) * (1. / D->dt);          // Δt^(1−βder), Julia's power_by_squaring of inv(Δt)
    Q.dv = d; Q.Lam = D->Lam; Q.X = D->X; Q.U = D->U; Q.scL = D->scL; Q.scX = D->scX; Q.scU = D->scU;
    decrementbig_kernel<<<nblk((D->ehi - D->elo) * W, 256), 256, 0, h->stream>>>(Q);
    h->launches++;
    if (delta2) {
        const int64_t nown = D->hi - D->lo;
        double* sums = nullptr; CK(dalloc(h, &sums, 3 * nown));
        block_sumsq_kernel<<<(unsigned)(3 * nown), 256, 0, h->stream>>>(D->nX, D->nU, W, d + (D->lo - s0) * W, sums);
        h->launches++;
        std::vector<double> hs((size_t)(3 * nown));
        CK(cudaMemcpyAsync(hs.data(), sums, hs.size() * 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        dfree(h, sums);
        delta2[0] = delta2[1] = delta2[2] = 0.;
        for (int64_t k = 0; k < nown; ++k) for (int c = 0; c < 3; ++c) delta2[c] = std::max(delta2[c], hs[(size_t)(3 * k + c)]);
    } else CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    return MB_OK;
}
int32_t mb_direct_sparser(mb_handle* h, double rtol, int64_t* nnz_out) {
    if (!h || !h->direct) return MB_ERR_ARG;
    DirectData* D = h->direct;
    ARG(D->nzval && D->nnzbig > 0, "assemble first");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int64_t nnz = D->nnzbig;
    if (!D->ccolptr) { CK(dalloc(h, &D->ccolptr, D->ncol + 1)); CK(dalloc(h, &D->crowval, nnz)); CK(dalloc(h, &D->cnzval, nnz)); }
    double* dmax = nullptr; int64_t* pos = nullptr;
    CK(dalloc(h, &dmax, 1)); CK(dalloc(h, &pos, nnz));
    void* tmp = nullptr; size_t t1 = 0, t2 = 0;
    cub::TransformInputIterator<double, AbsF, const double*> absit(D->nzval, AbsF{});
    CK(cub::DeviceReduce::Max(nullptr, t1, absit, dmax, nnz, st));
    cub::CountingInputIterator<int64_t> ids(0);
    CK(cub::DeviceScan::ExclusiveSum(nullptr, t2, cub::TransformInputIterator<int64_t, KeepF, cub::CountingInputIterator<int64_t>>(ids, KeepF{D->nzval, 0.}), pos, nnz, st));
    CK(cudaMalloc(&tmp, std::max(t1, t2) + 1));
    size_t tsz = std::max(t1, t2) + 1;
    CK(cub::DeviceReduce::Max(tmp, tsz, absit, dmax, nnz, st));
    double hmax = 0.;
    CK(cudaMemcpyAsync(&hmax, dmax, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const double atol = rtol * hmax;
    tsz = std::max(t1, t2) + 1;
    CK(cub::DeviceScan::ExclusiveSum(tmp, tsz, cub::TransformInputIterator<int64_t, KeepF, cub::CountingInputIterator<int64_t>>(ids, KeepF{D->nzval, atol}), pos, nnz, st));
    int64_t lastpos = 0; double lastv = 0.;
    CK(cudaMemcpyAsync(&lastpos, pos + (nnz - 1), 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&lastv, D->nzval + (nnz - 1), 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int64_t nkeep = lastpos + (fabs(lastv) >= atol ? 1 : 0);
    sparser_scatter_kernel<<<nblk(nnz, 256), 256, 0, st>>>(nnz, D->nzval, D->rowval, pos, atol, D->cnzval, D->crowval);
    sparser_colptr_kernel<<<nblk(D->ncol + 1, 256), 256, 0, st>>>(D->ncol, nnz, D->colptr, pos, nkeep, D->ccolptr);
    h->launches += 4;
    CK(cudaStreamSynchronize(st));
    cudaFree(tmp); dfree(h, dmax); dfree(h, pos);
    D->cnnz = nkeep;
    if (nnz_out) *nnz_out = nkeep;
    CK(cudaGetLastError());
    return MB_OK;
}
int32_t mb_direct_get_sparse(mb_handle* h, int64_t* colptr, int64_t* rowval, double* nzval) {
    if (!h || !h->direct) return MB_ERR_ARG;
    DirectData* D = h->direct;
    ARG(D->cnnz >= 0, "call mb_direct_sparser first");
    CK(cudaSetDevice(h->device));
    if (colptr) { CK(cudaMemcpy(colptr, D->ccolptr, (size_t)(D->ncol + 1) * 8, cudaMemcpyDeviceToHost)); for (int64_t i = 0; i <= D->ncol; ++i) colptr[i] += 1; }
    if (rowval && D->cnnz) { CK(cudaMemcpy(rowval, D->crowval, (size_t)D->cnnz * 8, cudaMemcpyDeviceToHost)); for (int64_t i = 0; i < D->cnnz; ++i) rowval[i] += 1 + D->rowshift; }
    if (nzval && D->cnnz) CK(cudaMemcpy(nzval, D->cnzval, (size_t)D->cnnz * 8, cudaMemcpyDeviceToHost));
    return MB_OK;
}
// out.L1 / out.L2 blocks of one stored step: which = 0 L1[Λ] (nX), 1 L2[Λ,X][1,der] , 2 L2[X,Λ][der,1], 3 L2[Λ,U][1,1], 4 L2[U,Λ][1,1]
int32_t mb_direct_get_step_block(mb_handle* h, int64_t step, int32_t which, int32_t der, double* out) {
    if (!h || !h->direct) return MB_ERR_ARG;
    DirectData* D = h->direct;
    ARG(step >= D->elo && step < D->ehi && out && which >= 0 && which <= 4 && der >= 0 && der <= D->OX, "bad argument");
    CK(cudaSetDevice(h->device));
    const int64_t k = step - D->elo, nd = D->OX + 1;
    const double* src = nullptr; int64_t n = 0;
    if (which == 0) { src = D->L1L + k * D->nX; n = D->nX; }
    else if (which == 1) { n = D->pat[P_XX].nnz; src = D->LX + (k * nd + der) * n; }
    else if (which == 2) { n = D->pat[P_XX].nnz; src = D->XL + (k * nd + der) * n; }
    else if (which == 3) { n = D->pat[P_XU].nnz; src = D->LU + k * n; }
    else { n = D->pat[P_UX].nnz; src = D->UL + k * n; }
    if (n) CK(cudaMemcpy(out, src, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return MB_OK;
}
// device pointers and sizes of the per-step blocks, for the halo exchange between time-shards (NCCL send/recv of whole steps)
// Slide the owned window of an INTERIOR handle to [new_lo, new_lo + (hi−lo)).  Away from the first and the last time step every finite-difference
// stencil is the central one (src/FiniteDifferences.jl:8-31), so the block pattern of makepattern (src/DirectXUA.jl:245-307) and the CSC structure of the
// owned columns repeat from window to window up to a shift of the global row numbers by (new_lo−lo)·(2·ndofX+ndofU): nothing is rebuilt, the shift is
// applied when the structure is exported.  States and per-step blocks of the steps stored both before and after the move keep their values (a
// window that advances by its own length keeps the four steps around its old right edge: only the new steps need set_state + evaluation).
static __global__ void shift_i32_kernel(int64_t n, int32_t* __restrict__ a, int32_t d) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += d;
}
int32_t mb_direct_rebase(mb_handle* h, int64_t new_lo, int64_t* row_shift_out) {
    if (!h || !h->direct) return MB_ERR_ARG;
    DirectData* D = h->direct;
    const int64_t len = D->hi - D->lo, delta = new_lo - D->lo, ns = D->ehi - D->elo;
    ARG(D->lo >= 3 && D->hi <= D->nstep - 3, "only a handle whose stored steps exclude the first and the last time step can be moved");
    ARG(new_lo >= 3 && new_lo + len <= D->nstep - 3, "the new window must stay clear of the first and the last time step (their stencils differ)");
    if (row_shift_out) *row_shift_out = D->rowshift + delta * (2 * D->nX + D->nU);
    if (delta == 0) return MB_OK;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int nd = D->OX + 1;
    const int64_t nbc = 3 * len;
    int32_t nblocks = 0;
    CK(cudaMemcpyAsync(&nblocks, D->bcolptr + nbc, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    shift_i32_kernel<<<nblk(nblocks, 256), 256, 0, st>>>(nblocks, D->browval, (int32_t)(3 * delta));
    h->launches++;
    // per-step arrays: slot of step s moves from s−elo to s−(elo+delta)
    struct Arr { double* p; int64_t stride; };
    std::vector<Arr> arrs = {{D->X, 3 * D->nX}, {D->U, D->nU}, {D->Lam, D->nX}, {D->LX, nd * D->pat[P_XX].nnz}, {D->XL, nd * D->pat[P_XX].nnz},
                             {D->LU, D->pat[P_XU].nnz}, {D->UL, D->pat[P_UX].nnz}, {D->L1L, D->nX}, {D->L1X, nd * D->nX},
                             {D->hostc, 2 * (D->nX + D->nU)}};
    for (size_t ig = 0; ig < D->hoststore.size(); ++ig) {
        if (!D->hoststore[ig]) continue;
        const Group& g = h->groups[ig];
        const int64_t nR = g.nele * g.nx;
        arrs.push_back({D->hoststore[ig], nR + nR * g.nx * nd + nR * nd});
    }
    arrs.push_back({D->XXc, D->pat[P_XX].nnz}); arrs.push_back({D->L1U, D->nU});
    for (auto& kv : D->meas) {
        arrs.push_back({kv.second.eps, kv.second.stride});
        std::vector<char> hv((size_t)ns, 0);
        for (int64_t k = 0; k < ns; ++k) if (k + delta >= 0 && k + delta < ns) hv[(size_t)k] = kv.second.have[(size_t)(k + delta)];
        kv.second.have.swap(hv);
    }
    const int64_t k0 = delta > 0 ? delta : 0, k1 = delta > 0 ? ns : ns + delta;      // old slots that stay stored
    for (const Arr& a : arrs) {
        if (!a.p || a.stride == 0) continue;
        if (delta > 0) for (int64_t k = k0; k < k1; ++k) CK(cudaMemcpyAsync(a.p + (k - delta) * a.stride, a.p + k * a.stride, (size_t)a.stride * 8, cudaMemcpyDeviceToDevice, st));
        else for (int64_t k = k1 - 1; k >= k0; --k) CK(cudaMemcpyAsync(a.p + (k - delta) * a.stride, a.p + k * a.stride, (size_t)a.stride * 8, cudaMemcpyDeviceToDevice, st));
    }
    D->lo += delta; D->hi += delta; D->elo += delta; D->ehi += delta;
    for (auto it = D->hostxx.begin(); it != D->hostxx.end();) {
        if (it->first < D->elo || it->first >= D->ehi) { dfree(h, it->second.i); dfree(h, it->second.j); dfree(h, it->second.v); it = D->hostxx.erase(it); } else ++it;
    }
    D->rowshift += delta * (2 * D->nX + D->nU);
    D->cnnz = -1;
    CK(cudaGetLastError());
    return MB_OK;
}
int32_t mb_direct_step_ptrs(mb_handle* h, int64_t step, double** LX, int64_t* nLX, double** LU, int64_t* nLU, double** L1L, int64_t* nL1) {
    if (!h || !h->direct) return MB_ERR_ARG;
    DirectData* D = h->direct;
    ARG(step >= D->elo && step < D->ehi, "step not stored on this handle");
    const int64_t k = step - D->elo, nd = D->OX + 1;
    if (LX) *LX = D->LX + k * nd * D->pat[P_XX].nnz; if (nLX) *nLX = nd * D->pat[P_XX].nnz;
    if (LU) *LU = D->LU + k * D->pat[P_XU].nnz; if (nLU) *nLU = D->pat[P_XU].nnz;
    if (L1L) *L1L = D->L1L + k * D->nX; if (nL1) *nL1 = D->nX;
    return MB_OK;
}
// Time shards over NCCL: rank r owns the steps [lo,hi) and stores [lo−2,hi+2)∩[0,nstep).  After its own steps are evaluated
// (mb_direct_assemble(eval_lo = lo, eval_hi = hi, build_big = 0)) the per-step blocks its neighbours' stencils reach go out and the halo blocks come in:
// L2[Λ,X][1,:] (the only per-step block read at a foreign step by the columns of an owned step, src/DirectXUA.jl:342-352 with src/FiniteDifferences.jl:2-4)
// and, when second-order element types are present, L1[X][:] .  Ranks are consecutive time shards of equal length: rank−1 owns the steps before lo.
// Asynchronous on the handle's stream; follow with mb_direct_assemble(eval_lo = eval_hi, build_big = 1).
int32_t mb_direct_halo_exchange(mb_handle* h) {
    if (!h || !h->direct) return MB_ERR_ARG;
    DirectData* D = h->direct;
    ARG(h->comm, "call mb_comm_init first");
    CK(cudaSetDevice(h->device));
    const int64_t nd = D->OX + 1, nLX = nd * D->pat[P_XX].nnz, nL1X = D->L1X ? nd * D->nX : 0;
    const int64_t nl = D->lo - D->elo, nr = D->ehi - D->hi;              // halo steps stored on the left / right (0, 1 or 2)
    const int64_t own = D->hi - D->lo;
    ARG(own >= 2, "a time shard needs at least two steps");
    std::vector<MbXfer> s, r;
    auto at = [&](double* base, int64_t per, int64_t step) { return base + (step - D->elo) * per; };
    if (D->lo > 0) {                                                      // left neighbour: my first two steps out, its last nl steps in
        ARG(h->rank > 0, "steps before this shard but no rank to the left");
        s.push_back({at(D->LX, nLX, D->lo), 2 * nLX, h->rank - 1});
        r.push_back({at(D->LX, nLX, D->elo), nl * nLX, h->rank - 1});
        if (nL1X) { s.push_back({at(D->L1X, nL1X, D->lo), 2 * nL1X, h->rank - 1}); r.push_back({at(D->L1X, nL1X, D->elo), nl * nL1X, h->rank - 1}); }
    }
    if (D->hi < D->nstep) {                                               // right neighbour: my last two steps out, its first nr steps in
        ARG(h->rank + 1 < h->world, "steps after this shard but no rank to the right");
        s.push_back({at(D->LX, nLX, D->hi - 2), 2 * nLX, h->rank + 1});
        r.push_back({at(D->LX, nLX, D->hi), nr * nLX, h->rank + 1});
        if (nL1X) { s.push_back({at(D->L1X, nL1X, D->hi - 2), 2 * nL1X, h->rank + 1}); r.push_back({at(D->L1X, nL1X, D->hi), nr * nL1X, h->rank + 1}); }
    }
    return mb_comm_sendrecv(h, s, r);
}
int32_t mb_direct_time_dev(mb_handle* h, int32_t reps, float* ms) {
    if (!h || !h->direct || !ms) return MB_ERR_ARG;
    DirectData* D = h->direct;
    CK(cudaSetDevice(h->device));
    cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    float a = 0, b = 0;
    CK(cudaMemsetAsync(h->nanflag, 0xFF, sizeof(unsigned long long), h->stream));
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0, h->stream));
        direct_eval_steps(h, D->lo, D->hi);
        CK(cudaEventRecord(e1, h->stream));
        BigDev B = make_bigdev(D);
        launch_big_values(B, D->ncol, D->colptr, D->nzval, D->maxb, h->stream);
        big_vec_kernel<<<nblk(D->ncol, 256), 256, 0, h->stream>>>(B, D->ncol, D->Lv);
        if (D->hostc) { big_diag_kernel<<<nblk(D->ncol, 256), 256, 0, h->stream>>>(B, D->ncol, D->colptr, D->rowval, D->nzval, D->missing, D->rowshift); h->launches++; }
        h->launches += 2;
        CK(cudaEventRecord(e2, h->stream));
        CK(cudaEventSynchronize(e2));
        float x, y; CK(cudaEventElapsedTime(&x, e0, e1)); CK(cudaEventElapsedTime(&y, e1, e2)); a += x; b += y;
    }
    ms[0] = a / reps; ms[1] = b / reps;
    // ms[2]: the element kernels of the owned steps alone (they write the one-step scratch dR/R only: the stored blocks stay valid)
    D->elements_only = true;
    CK(cudaEventRecord(e0, h->stream));
    direct_eval_steps(h, D->lo, D->hi);
    CK(cudaEventRecord(e1, h->stream));
    D->elements_only = false;
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms[2], e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    return MB_OK;
}

}  // extern "C"
