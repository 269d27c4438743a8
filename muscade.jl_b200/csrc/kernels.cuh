// Device kernels of the assembly engine (sm_100a).
//   beam_kernel      : K1/K2  EulerBeam3D residual + tangent, one lane per W seed directions of one element
//   gather_nz_kernel : K6     deterministic segmented reduction of element tangents into the CSC value array
//   gather_vec_kernel: K6     same for the gradient vector Lλ
//   pattern kernels  : K8     asmmat!-identical pattern / index-map construction (setup)
#pragma once
#include "beam_kernel.cuh"
#include <stdint.h>

namespace mb {

// nzval[k] = Σ_{s ∈ [cstart[k],cstart[k+1])} Ke[src[s]]  — contributions added in the reference's element order
// (src/Assemble.jl:472,479 → add_∂! :572-588), no atomics.
// Pair descriptors.  Consecutive rows of an element-matrix column land on consecutive non-zeros of a CSC column whenever a node's dofs are numbered
// consecutively, so for most PAIRS of non-zeros (2p, 2p+1) the contributor lists are "the same one or two elements, entry and entry+1".  prepare
// stores such a pair as (b0,b1): contributors Ke[b0] (+ Ke[b1]) for non-zero 2p and Ke[b0+1] (+ Ke[b1+1]) for 2p+1, b1 = NONE for a single contributor;
// b0 = NONE marks an irregular pair, which walks cstart/src as before.  Index traffic 4 B per non-zero instead of 9.3 (cstart 4 + src 4·144/108).
// A thread owns two pairs (four non-zeros): one 16-byte descriptor load, then up to eight independent value loads.
constexpr uint32_t MB_NONE = 0xFFFFFFFFu;
// Pairs that are not "entry and entry+1 of the same elements" but whose two non-zeros have at most two contributors each (three dofs per node: the pair
// straddles two node blocks; a spring on a beam node) are SPLIT pairs: (NONE, q) points to xdesc[q] = (a0,a1,c0,c1), contributors of 2p and of 2p+1 (NONE =
// absent) — one more dependent load, still no walk through cstart/src.  Everything else is (NONE, NONE).
static __global__ void pair_desc_kernel(int64_t nnz, int64_t npairs_padded, const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ src,
                                        uint32_t* __restrict__ pdesc, uint32_t* __restrict__ split) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs_padded) return;
    const int64_t k = 2 * p;
    uint32_t b0 = MB_NONE, b1 = MB_NONE, sp = 0;
    if (k + 1 < nnz) {
        const uint32_t c0 = cstart[k], c1 = cstart[k + 1], c2 = cstart[k + 2];
        const uint32_t n0 = c1 - c0, n1 = c2 - c1;
        if (n0 == n1 && (n0 == 1 || n0 == 2)) {
            const uint32_t s0 = src[c0], t0 = src[c1];
            if (t0 == s0 + 1) {
                if (n0 == 1) b0 = s0;
                else { const uint32_t s1 = src[c0 + 1], t1 = src[c1 + 1]; if (t1 == s1 + 1) { b0 = s0; b1 = s1; } }
            }
        }
        if (b0 == MB_NONE && n0 >= 1 && n0 <= 2 && n1 >= 1 && n1 <= 2) sp = 1;
    }
    pdesc[2 * p] = b0; pdesc[2 * p + 1] = b1;
    split[p] = sp;
}
static __global__ void split_desc_kernel(int64_t npairs_padded, const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ src, const uint32_t* __restrict__ split,
                                         const uint32_t* __restrict__ pos, uint32_t* __restrict__ pdesc, uint4* __restrict__ xdesc) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs_padded || !split[p]) return;
    const int64_t k = 2 * p;
    const uint32_t c0 = cstart[k], c1 = cstart[k + 1], c2 = cstart[k + 2];
    uint4 x;
    x.x = src[c0]; x.y = (c1 - c0 == 2) ? src[c0 + 1] : MB_NONE;
    x.z = src[c1]; x.w = (c2 - c1 == 2) ? src[c1 + 1] : MB_NONE;
    xdesc[pos[p]] = x;
    pdesc[2 * p + 1] = pos[p];
}
__device__ __forceinline__ double gather_one(const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ src, const double* __restrict__ Ke, int64_t k) {
    double acc = 0.;
    for (uint32_t s = cstart[k]; s < cstart[k + 1]; ++s) acc += __ldg(Ke + src[s]);
    return acc;
}
__device__ __forceinline__ void gather_pair(const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ src, const uint4* __restrict__ xdesc, uint32_t q,
                                            const double* __restrict__ Ke, int64_t k, double& a, double& b) {
    if (q != MB_NONE) {
        const uint4 x = __ldg(xdesc + q);
        const double va = __ldg(Ke + x.x), vb = __ldg(Ke + x.z);
        const double wa = __ldg(Ke + (x.y != MB_NONE ? x.y : 0u)), wb = __ldg(Ke + (x.w != MB_NONE ? x.w : 0u));
        a = 0. + va; if (x.y != MB_NONE) a += wa;
        b = 0. + vb; if (x.w != MB_NONE) b += wb;
    } else { a = gather_one(cstart, src, Ke, k); b = gather_one(cstart, src, Ke, k + 1); }
}
// wflag (or nullptr): one byte per non-zero, 1 = already final in nzval — written by the element kernel's fused epilogue (beam_static_ap_kernel<…,FUSE>),
// whose warp held every contributor of that non-zero; such non-zeros are neither read nor written here.
static __global__ void gather_nz_kernel(int64_t nnz, const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ src, const uint32_t* __restrict__ pdesc,
                                 const uint4* __restrict__ xdesc, const double* __restrict__ Ke, double* __restrict__ nzval, const uint8_t* __restrict__ wflag) {
    const int64_t k0 = 4 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (k0 >= nnz) return;
    uint32_t fl = 0;
    if (wflag) {
        if (k0 + 4 <= nnz) fl = __ldg(reinterpret_cast<const uint32_t*>(wflag + k0));
        else for (int64_t k = k0; k < nnz; ++k) fl |= (uint32_t)wflag[k] << (8 * (k - k0));
        if (fl == 0x01010101u) return;
    }
    if (fl != 0) {                                   // some of the four are final already: the others one by one
        for (int64_t k = k0; k < nnz && k < k0 + 4; ++k) if (!((fl >> (8 * (k - k0))) & 1u)) nzval[k] = gather_one(cstart, src, Ke, k);
        return;
    }
    if (k0 + 4 <= nnz) {
        const uint4 d = __ldg(reinterpret_cast<const uint4*>(pdesc + k0));          // (b0,b1) of the pairs (k0,k0+1) and (k0+2,k0+3)
        const bool r0 = d.x != MB_NONE, r1 = d.z != MB_NONE, t0 = r0 && d.y != MB_NONE, t1 = r1 && d.w != MB_NONE;
        // all value loads up front (a safe entry stands in where there is nothing to load)
        const uint32_t i0 = r0 ? d.x : 0u, j0 = t0 ? d.y : 0u, i1 = r1 ? d.z : 0u, j1 = t1 ? d.w : 0u;
        const double v00 = __ldg(Ke + i0), v01 = __ldg(Ke + i0 + 1), w00 = __ldg(Ke + j0), w01 = __ldg(Ke + j0 + 1);
        const double v10 = __ldg(Ke + i1), v11 = __ldg(Ke + i1 + 1), w10 = __ldg(Ke + j1), w11 = __ldg(Ke + j1 + 1);
        double a0, a1, a2, a3;
        if (r0) { a0 = 0. + v00; a1 = 0. + v01; if (t0) { a0 += w00; a1 += w01; } }
        else gather_pair(cstart, src, xdesc, d.y, Ke, k0, a0, a1);
        if (r1) { a2 = 0. + v10; a3 = 0. + v11; if (t1) { a2 += w10; a3 += w11; } }
        else gather_pair(cstart, src, xdesc, d.w, Ke, k0 + 2, a2, a3);
        double2* out = reinterpret_cast<double2*>(nzval + k0);
        out[0] = make_double2(a0, a1); out[1] = make_double2(a2, a3);
    } else {
        for (int64_t k = k0; k < nnz; ++k) nzval[k] = gather_one(cstart, src, Ke, k);
    }
}
// ---------------------------------------------------------------------------------------------- fused epilogue: which non-zeros a warp can finish itself
// The static element kernel gives a warp FIVE whole consecutive elements and keeps their 720 tangent entries in a shared-memory tile (beam_kernel.cuh).  A non-zero whose
// contributors (one or two) all lie in one such tile is summed there, in element order, and written to nzval by the element kernel; the others go through Ke and
// gather_list_kernel.  Built once at prepare, per beam group and per warp:
//   header  (wbase, wlen | nun << 16, offset of its PATTERN in 16-byte units, 0)
//   pattern wlen words, one per non-zero of the range [wbase, wbase + wlen): BYTE offsets (o0 | o1 << 16) of its contributors in the tile, o1 = 8·MB_TILE (a zero slot) = none, both there = not this warp's; padded to whole chunks of 256;
//           then nun 16-bit tile entries that are NOT finished here (ascending) and are copied to Ke.
// Patterns are relative to the tile and to wbase, so on a regular mesh they repeat: a warp whose pattern equals its predecessor's points to the same copy
// (a chain numbered along its axis keeps three patterns in all, resident in L2).  The kernel pulls header and pattern into shared memory with cp.async at entry —
// no register is held across the sweep for the epilogue, and no load of the epilogue waits on HBM.
__device__ __forceinline__ bool fuse_owned(int64_t k, uint32_t q0, uint32_t q1, const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ src) {
    const uint32_t c0 = cstart[k], c1 = cstart[k + 1];
    if (c1 == c0 || c1 - c0 > 2) return false;
    for (uint32_t c = c0; c < c1; ++c) { const uint32_t q = src[c]; if (q < q0 || q >= q1) return false; }
    return true;
}
static __global__ void fuse_mark_kernel(int64_t npair_g, uint32_t pair_base, const int32_t* __restrict__ asm2, const uint32_t* __restrict__ cstart,
                                        const uint32_t* __restrict__ src, int32_t* __restrict__ wbase) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= npair_g) return;
    const int64_t k = (int64_t)asm2[pair_base + q] - 1;
    if (k < 0) return;
    const int64_t w = q / MB_TILE;
    const uint32_t q0 = pair_base + (uint32_t)(w * MB_TILE), q1 = (uint32_t)min((int64_t)q0 + MB_TILE, (int64_t)pair_base + npair_g);
    if (fuse_owned(k, q0, q1, cstart, src)) atomicMin(&wbase[w], (int32_t)k);
}
static __global__ void fuse_dest_kernel(int64_t npair_g, uint32_t pair_base, const int32_t* __restrict__ asm2, const uint32_t* __restrict__ cstart,
                                        const uint32_t* __restrict__ src, const int32_t* __restrict__ wbase, int32_t* __restrict__ wlen, uint32_t* __restrict__ ebits,
                                        uint8_t* __restrict__ wflag) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= npair_g) return;
    const int64_t k = (int64_t)asm2[pair_base + q] - 1, w = q / MB_TILE;
    const int i = (int)(q - w * MB_TILE);
    bool done = (k < 0);                              // an entry that goes nowhere needs no store either
    if (!done) {
        const uint32_t q0 = pair_base + (uint32_t)(w * MB_TILE), q1 = (uint32_t)min((int64_t)q0 + MB_TILE, (int64_t)pair_base + npair_g);
        const int64_t slot = k - wbase[w];
        if (fuse_owned(k, q0, q1, cstart, src) && slot < MB_FUSE_RANGE) {       // the same decision for every contributor of k: they share the warp, hence wbase
            done = true; wflag[k] = 1; atomicMax(&wlen[w], (int32_t)slot + 1);
        }
    }
    if (done) atomicOr(&ebits[w * 23 + (i >> 5)], 1u << (i & 31));
}
// words of warp w's pattern (multiple of 4) and its count of unfinished entries
static __global__ void fuse_size_kernel(int64_t nw, int64_t npair_g, const int32_t* __restrict__ wlen, const uint32_t* __restrict__ ebits, int32_t* __restrict__ wnun, int64_t* __restrict__ psz) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nw) return;
    const int nq = (int)min((int64_t)MB_TILE, npair_g - w * MB_TILE);
    int done = 0;
    for (int j = 0; j < 23; ++j) done += __popc(ebits[w * 23 + j]);
    const int nun = nq - done;
    wnun[w] = nun;
    psz[w] = ((((wlen[w] + 255) & ~255) + (nun + 1) / 2 + 3) / 4) * 4;
}
// one CUDA warp per element-kernel warp: writes its pattern at pat + poff[w]
static __global__ void fuse_fill_kernel(int64_t nw, int64_t npair_g, uint32_t pair_base, const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ src,
                                        const int32_t* __restrict__ wbase, const int32_t* __restrict__ wlen, const uint32_t* __restrict__ ebits,
                                        const int64_t* __restrict__ poff, const int64_t* __restrict__ psz, uint32_t* __restrict__ pat) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int l = threadIdx.x & 31;
    if (w >= nw) return;
    uint32_t* P = pat + poff[w];
    const int n = wlen[w];
    const int64_t kb = wbase[w];
    const uint32_t q0 = pair_base + (uint32_t)(w * MB_TILE), q1 = (uint32_t)min((int64_t)q0 + MB_TILE, (int64_t)pair_base + npair_g);
    for (int sl = l; sl < n; sl += 32) {
        const int64_t k = kb + sl;
        uint32_t d = MB_PAT_NOTMINE;
        if (fuse_owned(k, q0, q1, cstart, src)) {
            const uint32_t c0 = cstart[k], c1 = cstart[k + 1];
            d = 8u * ((src[c0] - q0) | ((c1 - c0 == 2 ? src[c0 + 1] - q0 : (uint32_t)MB_TILE) << 16));
        }
        P[sl] = d;
    }
    for (int64_t i = n + l; i < psz[w]; i += 32) P[i] = MB_PAT_NOTMINE;        // chunk padding (read by the kernel), list area and tail: defined, so that equal patterns compare equal
    __syncwarp();
    uint16_t* U = reinterpret_cast<uint16_t*>(P + ((n + 255) & ~255));
    const int nq = (int)(q1 - q0);
    int base = 0;
    for (int j = 0; j < 23; ++j) {
        const int i = j * 32 + l;
        const bool un = i < nq && !((ebits[w * 23 + j] >> l) & 1u);
        const unsigned m = __ballot_sync(0xffffffffu, un);
        if (un) U[base + __popc(m & ((1u << l) - 1u))] = (uint16_t)i;
        base += __popc(m);
    }
}
// eq[w] = 0 where warp w's header sizes or pattern differ from warp w−1's (w is the head of a run), else 1; hd[w] = w or 0 for the max-scan that finds the run's head
static __global__ void fuse_eq_kernel(int64_t nw, const int32_t* __restrict__ wlen, const int32_t* __restrict__ wnun, const int64_t* __restrict__ poff, const int64_t* __restrict__ psz,
                                      const uint32_t* __restrict__ pat, int64_t* __restrict__ hd, int64_t* __restrict__ hsz) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int l = threadIdx.x & 31;
    if (w >= nw) return;
    bool same = w > 0 && wlen[w] == wlen[w - 1] && wnun[w] == wnun[w - 1] && psz[w] == psz[w - 1];
    if (same) {
        const uint32_t *A = pat + poff[w], *B = pat + poff[w - 1];
        bool ok = true;
        for (int64_t i = l; i < psz[w]; i += 32) ok &= (A[i] == B[i]);
        same = __all_sync(0xffffffffu, ok);
    }
    if (l == 0) { hd[w] = same ? 0 : w; hsz[w] = same ? 0 : psz[w]; }
}
static __global__ void fuse_compact_kernel(int64_t nw, const int64_t* __restrict__ head, const int64_t* __restrict__ poff, const int64_t* __restrict__ psz, const int64_t* __restrict__ noff,
                                           const uint32_t* __restrict__ pat, uint32_t* __restrict__ cpat, const int32_t* __restrict__ wbase, const int32_t* __restrict__ wlen,
                                           const int32_t* __restrict__ wnun, int4* __restrict__ whdr) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int l = threadIdx.x & 31;
    if (w >= nw) return;
    const int64_t hw = head[w];
    if (hw == w) for (int64_t i = l; i < psz[w]; i += 32) cpat[noff[w] + i] = pat[poff[w] + i];
    if (l == 0) whdr[w] = make_int4(wbase[w], wlen[w] | (wnun[w] << 16), (int)(noff[hw] / 4), 0);
}
struct MaxI64 { __host__ __device__ int64_t operator()(int64_t a, int64_t b) const { return a > b ? a : b; } };

// the non-zeros the element kernel did NOT finish, as a compact ascending list (built once at prepare) of (k, first, second contributor | NONE, ·): one thread per
// non-zero, one 16-byte descriptor load, then the value loads side by side; first = NONE: more than two contributors, walk cstart / src
static __global__ void list_desc_kernel(int64_t n, const int32_t* __restrict__ list, const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ src, uint4* __restrict__ udesc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t k = list[i];
    const uint32_t c0 = cstart[k], c1 = cstart[k + 1];
    uint4 d = make_uint4((uint32_t)k, MB_NONE, MB_NONE, 0u);
    if (c1 - c0 >= 1 && c1 - c0 <= 2) { d.y = src[c0]; if (c1 - c0 == 2) d.z = src[c0 + 1]; }
    else if (c1 == c0) d.w = 1u;                     // no contributor: the value is 0
    udesc[i] = d;
}
static __global__ void gather_list_kernel(int64_t n, const uint4* __restrict__ udesc, const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ src,
                                          const double* __restrict__ Ke, double* __restrict__ nzval) {
    const int64_t i0 = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);       // two listed non-zeros per thread: four value loads in flight
    if (i0 >= n) return;
    const bool two = i0 + 1 < n;
    const uint4 d = __ldg(udesc + i0), e = two ? __ldg(udesc + i0 + 1) : d;
    const bool fd = d.y != MB_NONE, fe = e.y != MB_NONE;
    const double a0 = __ldcs(Ke + (fd ? d.y : 0u)), a1 = __ldcs(Ke + (fd ? (d.z != MB_NONE ? d.z : d.y) : 0u));
    const double b0 = __ldcs(Ke + (fe ? e.y : 0u)), b1 = __ldcs(Ke + (fe ? (e.z != MB_NONE ? e.z : e.y) : 0u));
    double x = 0. + a0; if (d.z != MB_NONE) x += a1;
    double y = 0. + b0; if (e.z != MB_NONE) y += b1;
    if (!fd) x = d.w ? 0. : gather_one(cstart, src, Ke, d.x);
    if (!fe) y = e.w ? 0. : gather_one(cstart, src, Ke, e.x);
    nzval[d.x] = x;
    if (two) nzval[e.x] = y;
}
struct NotFlag { const uint8_t* f; __host__ __device__ bool operator()(int32_t k) const { return f[k] == 0; } };

// Lλ[d] = Σ (Re[q] − Rp[q])  (add_value! then add_∂!{1,:minus}, src/SweepX.jl:56-57)
static __global__ void gather_vec_kernel(int64_t d0, int64_t ndof, const uint32_t* __restrict__ vstart, const uint32_t* __restrict__ vsrc,
                                  const double* __restrict__ Re, const double* __restrict__ Rp, double* __restrict__ out) {
    const int64_t d = d0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= ndof) return;
    const uint32_t s0 = vstart[d], s1 = vstart[d + 1];
    double acc = 0.;
    for (uint32_t s = s0; s < s1; ++s) {
        const uint32_t q = vsrc[s];
        acc += Re[q];
        if (Rp) acc -= Rp[q];
    }
    out[d] = acc;
}

// ---------------------------------------------------------------------------------------------- completion prefixes (host-buffer pipeline)
// last[k] = 1 + largest contributor index of segment k, ignoring contributors inside [skip0[r],skip1[r]) (host-evaluated element types, whose
// contributions are uploaded before the device kernels start); 0 when nothing else contributes.
struct SkipRanges { int n; uint32_t lo[8], hi[8]; };
static __global__ void seg_last_kernel(int64_t nseg, const uint32_t* __restrict__ start, const uint32_t* __restrict__ src, SkipRanges sk, uint32_t* __restrict__ last) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nseg) return;
    uint32_t m = 0;
    for (uint32_t s = start[k]; s < start[k + 1]; ++s) {
        const uint32_t id = src[s];
        bool skip = false;
        for (int r = 0; r < sk.n; ++r) skip |= (id >= sk.lo[r] && id < sk.hi[r]);
        if (!skip && id + 1 > m) m = id + 1;
    }
    last[k] = m;
}
// cnt[j] = number of leading segments whose running maximum of `last` is ≤ bound[j]  (pm is non-decreasing)
static __global__ void prefix_count_kernel(int64_t nseg, const uint32_t* __restrict__ pm, int nb, const int64_t* __restrict__ bound, int64_t* __restrict__ cnt) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb) return;
    int64_t lo = 0, hi = nseg;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if ((int64_t)pm[mid] <= bound[j]) lo = mid + 1; else hi = mid; }
    cnt[j] = lo;
}

// ---------------------------------------------------------------------------------------------- pattern construction
// (jmoddof,imoddof) pairs in the reference's enumeration order: element, jeledof, ieledof (src/Assemble.jl:385-395),
// packed as j·ndof + i so that an ascending sort is the lexicographic sortperm of (j,i) (:397).
static __global__ void make_pair_keys_kernel(int64_t nele, int nx, const int32_t* __restrict__ idx, uint64_t ndof, uint64_t* __restrict__ keys,
                                      uint32_t* __restrict__ vals, uint32_t base) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n2 = (int64_t)nx * nx;
    if (p >= nele * n2) return;
    const int64_t e = p / n2;
    const int r = (int)(p - e * n2);
    const int j = r / nx, i = r - j * nx;
    const uint64_t dj = (uint64_t)idx[e * nx + j], di = (uint64_t)idx[e * nx + i];
    keys[base + p] = dj * ndof + di;
    vals[base + p] = base + (uint32_t)p;
}
static __global__ void make_vec_keys_kernel(int64_t n, const int32_t* __restrict__ idx, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t base) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    keys[base + q] = (uint32_t)idx[q];
    vals[base + q] = base + (uint32_t)q;
}
struct KeyFlag {                      // 1 where a sorted key differs from its predecessor (new non-zero, src/Assemble.jl:406)
    const uint64_t* k;
    __host__ __device__ uint32_t operator()(int64_t s) const { return (s == 0 || k[s] != k[s - 1]) ? 1u : 0u; }
};
// For every sorted pair s: asm2[vals[s]] = inz (1-based, src/Assemble.jl:432,444); at the first pair of a non-zero: rowval, cstart.
static __global__ void finish_pattern_kernel(int64_t npair, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                      const uint32_t* __restrict__ inz, uint64_t ndof, int32_t* __restrict__ asm2,
                                      int32_t* __restrict__ rowval0, uint32_t* __restrict__ cstart) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= npair) return;
    const uint32_t k = inz[s];                       // 1-based
    asm2[vals[s]] = (int32_t)k;
    const uint64_t key = keys[s];
    if (s == 0 || key != keys[s - 1]) {
        rowval0[k - 1] = (int32_t)(key % ndof);
        cstart[k - 1] = (uint32_t)s;
    }
    if (s == npair - 1) cstart[k] = (uint32_t)npair;
}
// colptr0[c] = number of non-zeros in columns < c  (src/Assemble.jl:415-430)
static __global__ void colptr_kernel(int64_t nnz, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ cstart, uint64_t ndof,
                              int32_t* __restrict__ colptr0) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const int64_t col = (int64_t)(keys[cstart[k]] / ndof);
    const int64_t prev = (k == 0) ? -1 : (int64_t)(keys[cstart[k - 1]] / ndof);
    for (int64_t c = prev + 1; c <= col; ++c) colptr0[c] = (int32_t)k;
    if (k == nnz - 1) for (int64_t c = col + 1; c <= (int64_t)ndof; ++c) colptr0[c] = (int32_t)nnz;
}
struct KeyFlag32 {
    const uint32_t* k;
    __host__ __device__ uint32_t operator()(int64_t s) const { return (s == 0 || k[s] != k[s - 1]) ? 1u : 0u; }
};
// vstart[d] = first sorted contributor of dof d (dofs without contributors get an empty range)
static __global__ void vstart_kernel(int64_t nvec, const uint32_t* __restrict__ keys, int64_t ndof, uint32_t* __restrict__ vstart) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nvec) return;
    const int64_t d = keys[s];
    const int64_t prev = (s == 0) ? -1 : (int64_t)keys[s - 1];
    for (int64_t c = prev + 1; c <= d; ++c) vstart[c] = (uint32_t)s;
    if (s == nvec - 1) for (int64_t c = d + 1; c <= ndof; ++c) vstart[c] = (uint32_t)nvec;
}

static __global__ void widen_plus1_kernel(int64_t n, const int32_t* __restrict__ in, int64_t* __restrict__ out, int add) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int64_t)in[i] + add;
}

// ---------------------------------------------------------------------------------------------- state update
// Newmarkβdecrement!{OX} over all X dofs (src/SweepX.jl:98-132): x′,x″ ← getdof! (divide by scale), a = a₂x′+a₃x″, b = b₂x′+b₃x″ on the first
// iteration, Δx′ = a₁Δx(+a), Δx″ = b₁Δx(+b), then decrement! X_d −= Δ·scale.  Explicit _rn intrinsics: no FMA contraction, i.e. the
// reference's (Julia broadcast) rounding sequence.
template <int OX>
static __global__ void newmark_decrement_kernel(int64_t n, const double* __restrict__ dx, const double* __restrict__ scale, double* __restrict__ X0,
                                                double* __restrict__ X1, double* __restrict__ X2, NewmarkDev c, int firstiter) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d = dx[i], s = scale ? scale[i] : 1.0;
    if (OX >= 1) {
        double d1 = __dmul_rn(c.a1, d), d2 = (OX >= 2) ? __dmul_rn(c.b1, d) : 0.;
        if (firstiter) {
            const double xp = __ddiv_rn(X1[i], s), xpp = (OX >= 2) ? __ddiv_rn(X2[i], s) : 0.;
            if (OX >= 2) {
                d1 = __dadd_rn(d1, __dadd_rn(__dmul_rn(c.a2, xp), __dmul_rn(c.a3, xpp)));
                d2 = __dadd_rn(d2, __dadd_rn(__dmul_rn(c.b2, xp), __dmul_rn(c.b3, xpp)));
            } else d1 = __dadd_rn(d1, __dmul_rn(c.a2, xp));
        }
        X1[i] = __dsub_rn(X1[i], __dmul_rn(d1, s));
        if (OX >= 2) X2[i] = __dsub_rn(X2[i], __dmul_rn(d2, s));
    }
    X0[i] = __dsub_rn(X0[i], __dmul_rn(d, s));
}
struct Square { __host__ __device__ double operator()(double x) const { return x * x; } };

// ---------------------------------------------------------------------------------------------- measurement
// FP64 FMA peak: 8 independent chains per thread, no memory traffic.
static __global__ void fp64_peak_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
static __global__ void copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i];
}

}  // namespace mb
