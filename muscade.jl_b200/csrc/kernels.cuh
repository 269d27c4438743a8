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
// (src/Assemble.jl:472,479 → add_∂! :572-588), one thread per non-zero, no atomics.
// Four consecutive non-zeros per thread: their contributor lists are one contiguous range of src, so the index and value loads of up to
// four contributors are issued together (the one-non-zero-per-thread form was latency-bound on the cstart → src → Ke chain at 44 % of DRAM).
static __global__ void gather_nz_kernel(int64_t nnz, const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ src,
                                 const double* __restrict__ Ke, double* __restrict__ nzval) {
    const int64_t k0 = 4 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (k0 >= nnz) return;
    if (k0 + 4 <= nnz) {
        const uint4 c = __ldg(reinterpret_cast<const uint4*>(cstart + k0));
        const uint32_t c4 = __ldg(cstart + k0 + 4);
        double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
        for (uint32_t s = c.x; s < c4; s += 4) {
            uint32_t i[4]; double v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) i[j] = (s + j < c4) ? __ldg(src + s + j) : 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (s + j < c4) ? __ldg(Ke + i[j]) : 0.;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t sj = s + j;
                if (sj < c4) { if (sj < c.y) a0 += v[j]; else if (sj < c.z) a1 += v[j]; else if (sj < c.w) a2 += v[j]; else a3 += v[j]; }
            }
        }
        double2* out = reinterpret_cast<double2*>(nzval + k0);
        out[0] = make_double2(a0, a1); out[1] = make_double2(a2, a3);
    } else {
        for (int64_t k = k0; k < nnz; ++k) {
            double acc = 0.;
            for (uint32_t s = cstart[k]; s < cstart[k + 1]; ++s) acc += __ldg(Ke + src[s]);
            nzval[k] = acc;
        }
    }
}
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
