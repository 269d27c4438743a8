// Internal declarations shared by the translation units of libmuscade_b200.so (not part of the C ABI).
#pragma once
#include "../../include/muscade_b200.h"
#include "kernels.cuh"
#include "bar_kernel.cuh"
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace mb;

struct mb_handle;
void mb_direct_release(mb_handle* h);   // mb_direct.cu
void mb_xua_release(mb_handle* h);      // mb_xua.cu
struct MbXfer { double* p; int64_t n; int peer; };
int32_t mb_comm_sendrecv(mb_handle* h, const std::vector<MbXfer>& sends, const std::vector<MbXfer>& recvs);   // mb_comm.cu: one NCCL group on h->stream

enum GroupKind { G_BEAM = 1, G_BAR = 2, G_SOIL = 3, G_HOST = 4 };

struct Group {
    int kind = 0;
    int64_t nele = 0;
    int nx = 0, nu = 0, udof = 0;
    double* geo = nullptr;
    BeamMat* mats = nullptr;
    BarMat* barmats = nullptr;
    int32_t* mat_id = nullptr;
    int32_t* idxX = nullptr;     // [nele][nx] 0-based
    int32_t* idxU = nullptr;
    double scaleX[12] = {0}, scaleU[3] = {0};
    int64_t pair_base = 0;       // offset of this group's nx²·nele tangent entries in Ke_all
    int64_t vec_base = 0;        // offset of this group's nx·nele residual entries in Re_all
    int64_t maxX = 0, maxU = 0;  // largest 1-based dof numbers of the group (checked against ndofX / ndofU at prepare)
    // ElementCost{StrainGaugeOnEulerBeam3D} with the quadratic strain cost on this beam type, DirectXUA (mb_direct_set_gauge_cost): gauge matrix [ng][4] on the device, 1/σ²
    int ng = 0; double* gaugeG = nullptr; double isig2 = 0.;
};

struct mb_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    std::vector<Group> groups;
    // model sizes
    int64_t ndofX = 0, ndofU = 0, nnz = 0, npair = 0, nvec = 0;
    bool prepared = false;
    // state and outputs
    double *X0 = nullptr, *X1 = nullptr, *X2 = nullptr, *U0 = nullptr, *Ll = nullptr, *nzval = nullptr;
    // element outputs
    double *Ke = nullptr, *Re = nullptr, *Rp = nullptr;
    // maps
    int32_t *asm2 = nullptr, *colptr0 = nullptr, *rowval0 = nullptr;
    uint32_t *cstart = nullptr, *src = nullptr, *vstart = nullptr, *vsrc = nullptr;
    uint32_t* pdesc = nullptr;      // pair descriptors of the non-zero reduction (kernels.cuh), 2 words per pair of non-zeros
    uint4* xdesc = nullptr;         // split pairs: contributors of the two non-zeros of an irregular pair (≤ 2 each)
    double *dofscale = nullptr, *dxbuf = nullptr, *red = nullptr; void* redtmp = nullptr; size_t redtmp_sz = 0;   // device-resident Newton update
    unsigned long long* nanflag = nullptr;
    unsigned long long* nanflag_host = nullptr;   // pinned
    int64_t launches = 0;
    int beamW = 1;
    std::vector<void*> owned;
    double* Wc = nullptr; int64_t Wc_len = 0; int split_dyn = 1; int static_sym = 3; int nsm = 148;   // cotangent workspace of the two-phase Newmark kernel
    bool own_stream = true;
    // fused epilogue of the static beam kernel (kernels.cuh): per beam group the non-zero range of every warp and the entries it finishes in its tile
    struct FuseGroup { int4* whdr = nullptr; uint32_t* cpat = nullptr; };
    std::vector<FuseGroup> fuse_groups; uint8_t* wflag = nullptr; int32_t* ulist = nullptr; uint4* udesc = nullptr; int64_t nulist = 0; int fuse = 0; bool fused_now = false;   // fused_now: the assembly being issued is a static one (OX = 0)
    int32_t *if_send = nullptr, *if_recv = nullptr;   // interface index lists (0-based into [nzval | Lλ], −1 = ghost)
    int64_t if_nsend_nz = 0, if_nsend_v = 0, if_nrecv_nz = 0, if_nrecv_v = 0;
    double *if_sendbuf = nullptr, *if_recvbuf = nullptr;   // device buffers of mb_iface_exchange
    void* comm = nullptr; bool comm_owned = true; int rank = 0, world = 1; double* comm_scratch = nullptr;   // NCCL communicator of this handle (mb_comm.cu)
    struct XuaData* xua = nullptr;            // DirectXUA{OX,OU,IA}, general form (mb_xua.cu)
    struct DirectData* direct = nullptr;      // DirectXUA state (mb_direct.cu)
    // host-buffer path (mb_sweepx_assemble): element ranges are evaluated chunk by chunk and every prefix of nzval / Lλ whose contributors
    // are all done is reduced and copied to the host while the next chunk computes
    struct PipeItem { int ig; int64_t e0, e1; int chunk; };
    std::vector<PipeItem> pipe_items;
    std::vector<int64_t> pipe_nz_end, pipe_vec_end;      // per chunk: non-zeros / dofs complete after it
    int pipe_state = 0;                                  // 0 not built, 1 ready, -1 not applicable
    int pipe_chunks = 16; int64_t pipe_min_nnz = 1 << 22;
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> pipe_ev, pipe_evx; std::vector<int64_t> pipe_x_need; cudaStream_t h2d_stream = nullptr;      // host-state-in pipeline: events of the H2D pieces, dofs each chunk needs
    // device-resident path (mb_sweepx_assemble_dev), optional (MB_DEV_OVERLAP=1): the same chunk plan; the segmented reduction of chunk j on a
    // high-priority stream while the element kernels of chunk j+1 run.  Measured on B200 (10 M elements): 23.1 ms vs 23.0 ms serial — the
    // element CTAs hold the whole register file, the two kernels time-share the SMs instead of overlapping — so it is off by default.
    cudaStream_t gather_stream = nullptr; cudaEvent_t gather_done = nullptr; int dev_overlap = 0; int gather_block = 256;
};

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                 \
            return MB_ERR_CUDA;                                                                          \
        }                                                                                                \
    } while (0)
#define ARG(cond, msg)                                                                                   \
    do {                                                                                                 \
        if (!(cond)) { h->err = msg; return MB_ERR_ARG; }                                                \
    } while (0)

template <class T> static inline cudaError_t dalloc(mb_handle* h, T** p, int64_t n) {
    *p = nullptr;
    if (n <= 0) n = 1;
    cudaError_t e = cudaMalloc((void**)p, (size_t)n * sizeof(T));
    if (e == cudaSuccess) h->owned.push_back(*p);
    return e;
}
static inline void dfree(mb_handle* h, void* p) {
    if (!p) return;
    for (size_t i = 0; i < h->owned.size(); ++i)
        if (h->owned[i] == p) { h->owned.erase(h->owned.begin() + i); break; }
    cudaFree(p);
}
static inline unsigned nblk(int64_t n, int b) { return (unsigned)((n + b - 1) / b); }
// every group's dof numbers against the model sizes given to prepare (an index beyond them would scatter outside colptr / the state vectors)
static inline int32_t check_group_dofs(mb_handle* h, int64_t ndofX, int64_t ndofU) {
    for (size_t ig = 0; ig < h->groups.size(); ++ig) {
        const Group& g = h->groups[ig];
        if (g.maxX > ndofX) { h->err = "element type " + std::to_string(ig + 1) + ": X-dof index " + std::to_string(g.maxX) + " exceeds ndofX = " + std::to_string(ndofX); return MB_ERR_ARG; }
        if (g.maxU > ndofU) { h->err = "element type " + std::to_string(ig + 1) + ": U-dof index " + std::to_string(g.maxU) + " exceeds ndofU = " + std::to_string(ndofU); return MB_ERR_ARG; }
    }
    return MB_OK;
}

// Int64 1-based [nele][n] (reference memory order) → int32 0-based on the device
// maxout: largest 1-based index seen (the model sizes are not known yet when groups are added: prepare checks it against ndofX / ndofU)
static inline int32_t upload_index(mb_handle* h, const int64_t* src, int64_t n, int64_t limit, int32_t** dst, int64_t* maxout = nullptr) {
    std::vector<int32_t> tmp((size_t)n);
    int64_t mx = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t v = src[i];
        if (v < 1 || (limit > 0 && v > limit) || v > INT32_MAX) { h->err = "dof index out of range (expects 1-based Int64)"; return MB_ERR_ARG; }
        tmp[(size_t)i] = (int32_t)(v - 1);
        if (v > mx) mx = v;
    }
    if (maxout) *maxout = mx;
    CK(dalloc(h, dst, n));
    CK(cudaMemcpy(*dst, tmp.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
    return MB_OK;
}

// pair / split descriptors of a contributor-list structure (kernels.cuh), padded to whole threads of four non-zeros
static inline int32_t build_pair_descriptors(mb_handle* h, int64_t nnz, const uint32_t* cstart, const uint32_t* src, uint32_t** pdesc, uint4** xdesc) {
    cudaStream_t st = h->stream;
    const int64_t npp = 2 * ((nnz + 3) / 4);
    CK(dalloc(h, pdesc, npp > 0 ? 2 * npp : 4));
    *xdesc = nullptr;
    if (npp == 0) return MB_OK;
    uint32_t *split = nullptr, *pos = nullptr;
    CK(dalloc(h, &split, npp)); CK(dalloc(h, &pos, npp));
    mb::pair_desc_kernel<<<nblk(npp, 256), 256, 0, st>>>(nnz, npp, cstart, src, *pdesc, split);
    void* tmp = nullptr; size_t tmpsz = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmpsz, split, pos, npp, st));
    CK(cudaMalloc(&tmp, tmpsz ? tmpsz : 1));
    CK(cub::DeviceScan::ExclusiveSum(tmp, tmpsz, split, pos, npp, st));
    uint32_t lastpos = 0, lastflag = 0;
    CK(cudaMemcpyAsync(&lastpos, pos + (npp - 1), 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&lastflag, split + (npp - 1), 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st)); cudaFree(tmp);
    const int64_t nsplit = (int64_t)lastpos + lastflag;
    CK(dalloc(h, xdesc, nsplit > 0 ? nsplit : 1));
    if (nsplit > 0) mb::split_desc_kernel<<<nblk(npp, 256), 256, 0, st>>>(npp, cstart, src, split, pos, *pdesc, *xdesc);
    h->launches += 2;
    CK(cudaStreamSynchronize(st));
    dfree(h, split); dfree(h, pos);
    return MB_OK;
}
