// asmmat! on the device (src/Assemble.jl:373-448): the sparsity pattern of one class pair (rows = model dofs of class α, columns = of class β), the map
// element entry → non-zero (asm[arrnum(α,β)]) and, for the atomics-free reductions, the contributors of every non-zero in the reference's accumulation order
// (element type, element, jeledof, ieledof).  Shared by the beam-specialised DirectXUA path (mb_direct.cu) and the general one (mb_xua.cu), with the kernels of
// sparser! and the finite-difference stencils.
#pragma once
#include <vector>
#include "mb_internal.h"
#include <cub/cub.cuh>

namespace {

struct PairPat {                  // one class-pair pattern of prepare(AssemblyDirect): asmmat! (src/Assemble.jl:373-448)
    int64_t m = 0, n = 0, nnz = 0, npair = 0;
    int32_t *colptr0 = nullptr, *rowval0 = nullptr, *asmK = nullptr;
    uint32_t *cstart = nullptr, *src = nullptr;
    std::vector<int64_t> gbase;   // per group: first pair id
};

struct PatSide { int64_t nele; int n; const int32_t* idx; };      // one element type's dofs of one class: [nele][n], 0-based model dof numbers (device)

// ---------------------------------------------------------------------------------------------------------------- pattern build
__global__ void pair_keys_kernel(int64_t nele, int ni, const int32_t* __restrict__ idxR, int nj, const int32_t* __restrict__ idxC, uint64_t nrows,
                                 uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t base) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n2 = (int64_t)ni * nj;
    if (p >= nele * n2) return;
    const int64_t e = p / n2;
    const int r = (int)(p - e * n2);
    const int j = r / ni, i = r - j * ni;                      // jeledof outer, ieledof inner (src/Assemble.jl:389)
    keys[base + p] = (uint64_t)idxC[e * nj + j] * nrows + (uint64_t)idxR[e * ni + i];
    vals[base + p] = base + (uint32_t)p;
}
__global__ void finish_pat_kernel(int64_t npair, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ inz,
                                  uint64_t nrows, int32_t* __restrict__ asmK, int32_t* __restrict__ rowval0, uint32_t* __restrict__ cstart) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= npair) return;
    const uint32_t k = inz[s];
    asmK[vals[s]] = (int32_t)k;
    const uint64_t key = keys[s];
    if (s == 0 || key != keys[s - 1]) { rowval0[k - 1] = (int32_t)(key % nrows); cstart[k - 1] = (uint32_t)s; }
    if (s == npair - 1) cstart[k] = (uint32_t)npair;
}
__global__ void colptr_pat_kernel(int64_t nnz, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ cstart, uint64_t nrows, int64_t ncols,
                                  int32_t* __restrict__ colptr0) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const int64_t col = (int64_t)(keys[cstart[k]] / nrows);
    const int64_t prev = (k == 0) ? -1 : (int64_t)(keys[cstart[k - 1]] / nrows);
    for (int64_t c = prev + 1; c <= col; ++c) colptr0[c] = (int32_t)k;
    if (k == nnz - 1) for (int64_t c = col + 1; c <= ncols; ++c) colptr0[c] = (int32_t)nnz;
}
struct KeyFlagD {
    const uint64_t* k;
    __host__ __device__ uint32_t operator()(int64_t s) const { return (s == 0 || k[s] != k[s - 1]) ? 1u : 0u; }
};

static int32_t build_pair_pattern(mb_handle* h, PairPat& P, const std::vector<PatSide>& rows, const std::vector<PatSide>& cols, int64_t nrows, int64_t ncols) {
    cudaStream_t st = h->stream;
    P.m = nrows; P.n = ncols; P.gbase.clear();
    int64_t npair = 0;
    for (size_t ig = 0; ig < rows.size(); ++ig) { P.gbase.push_back(npair); npair += rows[ig].nele * rows[ig].n * cols[ig].n; }
    P.gbase.push_back(npair);
    P.npair = npair;
    if (npair >= (int64_t)UINT32_MAX) { h->err = "more than 2^32 element-matrix entries on one device"; return MB_ERR_TOOBIG; }
    CK(dalloc(h, &P.colptr0, ncols + 1));
    CK(cudaMemsetAsync(P.colptr0, 0, (ncols + 1) * sizeof(int32_t), st));
    if (npair == 0 || nrows == 0 || ncols == 0) { P.nnz = 0; CK(dalloc(h, &P.rowval0, 1)); CK(dalloc(h, &P.cstart, 1)); CK(dalloc(h, &P.src, 1)); CK(dalloc(h, &P.asmK, 1));
        CK(cudaMemsetAsync(P.cstart, 0, sizeof(uint32_t), st)); return MB_OK; }
    uint64_t *keys = nullptr, *keys2 = nullptr; uint32_t *vals = nullptr, *inz = nullptr;
    CK(dalloc(h, &keys, npair)); CK(dalloc(h, &keys2, npair)); CK(dalloc(h, &vals, npair)); CK(dalloc(h, &inz, npair));
    CK(dalloc(h, &P.src, npair)); CK(dalloc(h, &P.asmK, npair));
    for (size_t ig = 0; ig < rows.size(); ++ig) {
        const int ni = rows[ig].n, nj = cols[ig].n;
        const int64_t n = rows[ig].nele * ni * nj;
        if (n == 0) continue;
        pair_keys_kernel<<<nblk(n, 256), 256, 0, st>>>(rows[ig].nele, ni, rows[ig].idx, nj, cols[ig].idx, (uint64_t)nrows, keys, vals, (uint32_t)P.gbase[ig]);
        h->launches++;
    }
    int end_bit = 1; while (end_bit < 64 && ((uint64_t)nrows * (uint64_t)ncols) >> end_bit) ++end_bit;
    void* tmp = nullptr; size_t tmpsz = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmpsz, keys, keys2, vals, P.src, npair, 0, end_bit, st));
    CK(cudaMalloc(&tmp, tmpsz ? tmpsz : 1));
    CK(cub::DeviceRadixSort::SortPairs(tmp, tmpsz, keys, keys2, vals, P.src, npair, 0, end_bit, st));
    CK(cudaStreamSynchronize(st)); cudaFree(tmp);
    cub::CountingInputIterator<int64_t> cnt(0);
    cub::TransformInputIterator<uint32_t, KeyFlagD, cub::CountingInputIterator<int64_t>> flags(cnt, KeyFlagD{keys2});
    tmp = nullptr; tmpsz = 0;
    CK(cub::DeviceScan::InclusiveSum(nullptr, tmpsz, flags, inz, npair, st));
    CK(cudaMalloc(&tmp, tmpsz ? tmpsz : 1));
    CK(cub::DeviceScan::InclusiveSum(tmp, tmpsz, flags, inz, npair, st));
    uint32_t last = 0;
    CK(cudaMemcpyAsync(&last, inz + (npair - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st)); cudaFree(tmp);
    P.nnz = last;
    if (P.nnz > (int64_t)INT32_MAX) { h->err = "more than 2^31-1 non-zeros in a class-pair pattern on one device"; dfree(h, keys); dfree(h, keys2); dfree(h, vals); dfree(h, inz); return MB_ERR_TOOBIG; }
    CK(dalloc(h, &P.rowval0, P.nnz)); CK(dalloc(h, &P.cstart, P.nnz + 1));
    finish_pat_kernel<<<nblk(npair, 256), 256, 0, st>>>(npair, keys2, P.src, inz, (uint64_t)nrows, P.asmK, P.rowval0, P.cstart);
    colptr_pat_kernel<<<nblk(P.nnz, 256), 256, 0, st>>>(P.nnz, keys2, P.cstart, (uint64_t)nrows, ncols, P.colptr0);
    h->launches += 2;
    CK(cudaStreamSynchronize(st));
    dfree(h, keys); dfree(h, keys2); dfree(h, vals); dfree(h, inz);
    return MB_OK;
}

// sparser!(T,S,rtol) (src/SparseTools.jl:172-199): keep |nzval| ≥ rtol·max|S|, order preserved, colptr shifted by the drops before it
struct AbsF { __host__ __device__ double operator()(double x) const { return fabs(x); } };
struct KeepF { const double* v; double atol; __host__ __device__ int64_t operator()(int64_t i) const { return fabs(v[i]) >= atol ? 1 : 0; } };
__global__ void sparser_scatter_kernel(int64_t nnz, const double* __restrict__ v, const int64_t* __restrict__ rv, const int64_t* __restrict__ pos, double atol,
                                       double* __restrict__ v2, int64_t* __restrict__ rv2) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    if (fabs(v[i]) >= atol) { const int64_t q = pos[i]; v2[q] = v[i]; rv2[q] = rv[i]; }
}
__global__ void sparser_colptr_kernel(int64_t ncol, int64_t nnz, const int64_t* __restrict__ colptr, const int64_t* __restrict__ pos, int64_t nkeep,
                                      int64_t* __restrict__ colptr2) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > ncol) return;
    const int64_t p = colptr[c];
    colptr2[c] = (p < nnz) ? pos[p] : nkeep;
}
// finitediff(order,n,s) (src/FiniteDifferences.jl:2-31): weight of offset ds at 0-based step s, or 0 with found=false
__device__ __forceinline__ bool fd_weight(int order, int64_t n, int64_t s, int64_t ds, double& w) {
    if (order == 0) { w = 1.; return ds == 0; }
    const int pos = (s == 0) ? 0 : (s == n - 1 ? 1 : 2);
    if (order == 1) {
        if (pos == 0) { if (ds == 0) { w = -1.; return true; } if (ds == 1) { w = 1.; return true; } return false; }
        if (pos == 1) { if (ds == -1) { w = -1.; return true; } if (ds == 0) { w = 1.; return true; } return false; }
        if (ds == -1) { w = -.5; return true; } if (ds == 1) { w = .5; return true; } return false;
    }
    if (pos == 0) { if (ds == 0) { w = 1.; return true; } if (ds == 1) { w = -2.; return true; } if (ds == 2) { w = 1.; return true; } return false; }
    if (pos == 1) { if (ds == -2) { w = 1.; return true; } if (ds == -1) { w = -2.; return true; } if (ds == 0) { w = 1.; return true; } return false; }
    if (ds == -1) { w = 1.; return true; } if (ds == 0) { w = -2.; return true; } if (ds == 1) { w = 1.; return true; } return false;
}

}  // namespace
