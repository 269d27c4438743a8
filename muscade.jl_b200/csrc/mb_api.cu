// C ABI of the B200 assembly engine (include/muscade_b200.h). Owns all device memory; no PyTorch, no CPU fallback.
#include "mb_internal.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace mb {
void launch_bar(int ND, bool step, const BarGroupDev& g, const StateDev& st, const NewmarkDev& nm, double t, double* Ke, double* Re, double* Rp,
                unsigned long long* nanflag, unsigned long long nanbase, cudaStream_t s);
void launch_soil(int ND, bool step, const SoilGroupDev& g, const StateDev& st, const NewmarkDev& nm, double* Ke, double* Re, double* Rp,
                 unsigned long long* nanflag, unsigned long long nanbase, cudaStream_t s);
}

extern "C" {

int32_t mb_version(void) { return 100; }

int32_t mb_create(int32_t device, mb_handle** out) {
    if (!out) return MB_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return MB_ERR_CUDA;
    mb_handle* h = new mb_handle();
    h->device = device;
    int prlo = 0, prhi = 0;                              // the engine's stream at the greatest priority: the background reduction (MB_DEV_OVERLAP=2) runs below it
    if (cudaSetDevice(device) != cudaSuccess || cudaDeviceGetStreamPriorityRange(&prlo, &prhi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prhi) != cudaSuccess) { delete h; return MB_ERR_CUDA; }
    if (cudaMalloc((void**)&h->nanflag, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMallocHost((void**)&h->nanflag_host, sizeof(unsigned long long)) != cudaSuccess) { delete h; return MB_ERR_CUDA; }
    { cudaDeviceProp prop; if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->nsm = prop.multiProcessorCount; }
    const char* w = getenv("MB_BEAM_W");
    if (w) h->beamW = atoi(w);
    const char* sd = getenv("MB_SPLIT_DYN");
    if (sd) h->split_dyn = atoi(sd);
    const char* fu = getenv("MB_FUSE");                  // 1: static kernel with the fused epilogue (beam_kernel.cuh; bit-identical, half the DRAM traffic, 2 % slower: off by default)
    if (fu) h->fuse = atoi(fu);
    const char* ss = getenv("MB_STATIC_SYM");            // 0: statics through the two-direction SD kernel (A/B measurements)
    if (ss) h->static_sym = atoi(ss);
    const char* pc = getenv("MB_E2E_CHUNKS");            // host-buffer pipeline: number of element chunks (<2: one shot) and size threshold
    if (pc) h->pipe_chunks = std::min(64, atoi(pc));
    const char* pm = getenv("MB_E2E_MIN_NNZ");
    if (pm) h->pipe_min_nnz = atoll(pm);
    // 1 / 2: reduction of element chunk j on a second stream (above / below the engine's priority) while the element kernels of chunk j+1 run.  No gain measured, either way
    // (10 M elements, 8-32 chunks: 18.2-18.8 ms against 18.2), also not with the static kernel held to 240 / 224 / 208 registers so that the reduction's CTAs fit BESIDE
    // its two CTAs per SM (18.5 / 19.1 / 19.6 ms: the spills cost, the overlap does not pay) — profiles/r2_probe_overlap.txt
    const char* ov = getenv("MB_DEV_OVERLAP");
    if (ov) h->dev_overlap = atoi(ov);
    const char* gb = getenv("MB_GATHER_BLOCK");
    if (gb) h->gather_block = atoi(gb);
    *out = h;
    return MB_OK;
}

int32_t mb_destroy(mb_handle* h) {
    if (!h) return MB_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    mb_comm_destroy(h);
    mb_direct_release(h);
    mb_xua_release(h);
    for (void* p : h->owned) cudaFree(p);
    cudaFree(h->nanflag);
    if (h->redtmp) cudaFree(h->redtmp);
    cudaFreeHost(h->nanflag_host);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
    for (cudaEvent_t ev : h->pipe_evx) cudaEventDestroy(ev);
    for (cudaEvent_t ev : h->pipe_ev) cudaEventDestroy(ev);
    if (h->gather_done) cudaEventDestroy(h->gather_done);
    if (h->gather_stream) cudaStreamDestroy(h->gather_stream);
    delete h;
    return MB_OK;
}

const char* mb_last_error(const mb_handle* h) { return h ? h->err.c_str() : "null handle"; }
int64_t mb_launch_count(const mb_handle* h) { return h ? h->launches : 0; }

int32_t mb_add_eulerbeam3d(mb_handle* h, int64_t nele, const double* eleobj, int32_t udof, const int64_t* idxX, const int64_t* idxU,
                           const double* scaleX, const double* scaleU, int32_t* ieletyp_out) {
    if (!h) return MB_ERR_ARG;
    ARG(!h->prepared, "groups must be added before prepare");
    ARG(nele >= 0 && eleobj && idxX && scaleX, "null argument");
    ARG(!udof || (idxU && scaleU), "Udof element type needs idxU and scaleU");
    CK(cudaSetDevice(h->device));
    Group g;
    g.kind = G_BEAM; g.nele = nele; g.nx = 12; g.nu = udof ? 3 : 0; g.udof = udof;
    // split the reference AoS struct (69 doubles, toolbox/BeamElement.jl:87-103) into geometry lines + a de-duplicated material table
    std::vector<double> geo((size_t)nele * 16);
    std::vector<BeamMat> mats;
    std::vector<int32_t> mat_id((size_t)nele);
    std::map<std::string, int32_t> seen;
    for (int64_t e = 0; e < nele; ++e) {
        const double* o = eleobj + e * 69;
        double* q = geo.data() + e * 16;
        for (int k = 0; k < 12; ++k) q[k] = o[k];          // cₘ, rₘ
        for (int k = 0; k < 3; ++k) q[12 + k] = o[18 + k];  // tgₘ
        q[15] = o[48];                                      // L
        std::string key((const char*)(o + 53), 16 * sizeof(double));
        auto it = seen.find(key);
        if (it == seen.end()) {
            BeamMat m; std::memcpy(&m, o + 53, sizeof m);
            it = seen.emplace(key, (int32_t)mats.size()).first;
            mats.push_back(m);
        }
        mat_id[(size_t)e] = it->second;
    }
    CK(dalloc(h, &g.geo, nele * 16));
    CK(cudaMemcpy(g.geo, geo.data(), geo.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(dalloc(h, &g.mats, (int64_t)mats.size()));
    CK(cudaMemcpy(g.mats, mats.data(), mats.size() * sizeof(BeamMat), cudaMemcpyHostToDevice));
    if (mats.size() > 1) {
        CK(dalloc(h, &g.mat_id, nele));
        CK(cudaMemcpy(g.mat_id, mat_id.data(), mat_id.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    int32_t rc = upload_index(h, idxX, nele * 12, 0, &g.idxX, &g.maxX);
    if (rc) return rc;
    if (udof) { rc = upload_index(h, idxU, nele * 3, 0, &g.idxU, &g.maxU); if (rc) return rc; }
    for (int i = 0; i < 12; ++i) g.scaleX[i] = scaleX[i];
    if (udof) for (int i = 0; i < 3; ++i) g.scaleU[i] = scaleU[i];
    h->groups.push_back(g);
    if (ieletyp_out) *ieletyp_out = (int32_t)h->groups.size();
    return MB_OK;
}

int32_t mb_add_host_elements(mb_handle* h, int64_t nele, int32_t nx, const int64_t* idxX, int32_t* ieletyp_out) {
    if (!h) return MB_ERR_ARG;
    ARG(!h->prepared, "groups must be added before prepare");
    ARG(nele >= 0 && nx >= 0 && nx <= 64 && (idxX || nele * nx == 0), "bad argument");
    CK(cudaSetDevice(h->device));
    Group g;
    g.kind = G_HOST; g.nele = nele; g.nx = nx;
    int32_t rc = upload_index(h, idxX, nele * nx, 0, &g.idxX, &g.maxX);
    if (rc) return rc;
    h->groups.push_back(g);
    if (ieletyp_out) *ieletyp_out = (int32_t)h->groups.size();
    return MB_OK;
}

int32_t mb_set_host_elements(mb_handle* h, int32_t ieletyp, const double* Re, const double* Rp, const double* Ke) {
    if (!h) return MB_ERR_ARG;
    ARG(h->prepared, "call mb_sweepx_prepare first");
    ARG(ieletyp >= 1 && ieletyp <= (int32_t)h->groups.size() && h->groups[ieletyp - 1].kind == G_HOST, "not a host-evaluated element type");
    CK(cudaSetDevice(h->device));
    const Group& g = h->groups[ieletyp - 1];
    const size_t nv = (size_t)g.nele * g.nx, nk = nv * g.nx;
    if (Re) CK(cudaMemcpyAsync(h->Re + g.vec_base, Re, nv * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    else CK(cudaMemsetAsync(h->Re + g.vec_base, 0, nv * sizeof(double), h->stream));
    if (Rp) CK(cudaMemcpyAsync(h->Rp + g.vec_base, Rp, nv * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    else CK(cudaMemsetAsync(h->Rp + g.vec_base, 0, nv * sizeof(double), h->stream));
    if (Ke) CK(cudaMemcpyAsync(h->Ke + g.pair_base, Ke, nk * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    else CK(cudaMemsetAsync(h->Ke + g.pair_base, 0, nk * sizeof(double), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MB_OK;
}

int32_t mb_set_ndofU(mb_handle* h, int64_t ndofU) {
    if (!h) return MB_ERR_ARG;
    ARG(!h->prepared && ndofU >= 0, "set ndofU before prepare");
    h->ndofU = ndofU;
    return MB_OK;
}

// ------------------------------------------------------------------------------------------------------------ prepare
int32_t mb_sweepx_prepare(mb_handle* h, int64_t ndofX, int64_t* nnz_out) {
    if (!h) return MB_ERR_ARG;
    ARG(!h->prepared, "already prepared");
    ARG(ndofX >= 1 && ndofX < INT32_MAX, "ndofX out of range");
    { int32_t rc = check_group_dofs(h, ndofX, h->ndofU > 0 ? h->ndofU : INT64_MAX); if (rc) return rc; }     // ndofU may be left unset for SweepX (U is read through idxU only when given)
    CK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    h->ndofX = ndofX;
    int64_t npair = 0, nvec = 0;
    for (Group& g : h->groups) { g.pair_base = npair; g.vec_base = nvec; npair += g.nele * g.nx * g.nx; nvec += g.nele * g.nx; }
    if (npair >= (int64_t)UINT32_MAX) { h->err = "more than 2^32 element-matrix entries on one device"; return MB_ERR_TOOBIG; }
    h->npair = npair; h->nvec = nvec;

    CK(dalloc(h, &h->X0, ndofX)); CK(dalloc(h, &h->X1, ndofX)); CK(dalloc(h, &h->X2, ndofX)); CK(dalloc(h, &h->U0, h->ndofU));
    CK(cudaMemsetAsync(h->X0, 0, ndofX * sizeof(double), st)); CK(cudaMemsetAsync(h->X1, 0, ndofX * sizeof(double), st));
    CK(cudaMemsetAsync(h->X2, 0, ndofX * sizeof(double), st)); CK(cudaMemsetAsync(h->U0, 0, (h->ndofU > 0 ? h->ndofU : 1) * sizeof(double), st));
    CK(dalloc(h, &h->Ll, ndofX));
    CK(dalloc(h, &h->Ke, npair)); CK(dalloc(h, &h->Re, nvec)); CK(dalloc(h, &h->Rp, nvec));
    CK(cudaMemsetAsync(h->Ke, 0, (npair > 0 ? npair : 1) * sizeof(double), st));
    CK(cudaMemsetAsync(h->Re, 0, (nvec > 0 ? nvec : 1) * sizeof(double), st));
    CK(cudaMemsetAsync(h->Rp, 0, (nvec > 0 ? nvec : 1) * sizeof(double), st));
    CK(dalloc(h, &h->asm2, npair)); CK(dalloc(h, &h->src, npair));
    CK(dalloc(h, &h->colptr0, ndofX + 1));
    CK(dalloc(h, &h->vstart, ndofX + 1)); CK(dalloc(h, &h->vsrc, nvec));
    CK(cudaMemsetAsync(h->colptr0, 0, (ndofX + 1) * sizeof(int32_t), st));
    CK(cudaMemsetAsync(h->vstart, 0, (ndofX + 1) * sizeof(uint32_t), st));

    // ---- matrix pattern: sort (j,i) pairs, unique, number them (asmmat!, src/Assemble.jl:373-448)
    int64_t nnz = 0;
    if (npair > 0) {
        uint64_t *keys = nullptr, *keys2 = nullptr; uint32_t *vals = nullptr, *inz = nullptr;
        CK(dalloc(h, &keys, npair)); CK(dalloc(h, &keys2, npair)); CK(dalloc(h, &vals, npair)); CK(dalloc(h, &inz, npair));
        for (const Group& g : h->groups) {
            const int64_t n = g.nele * g.nx * g.nx;
            if (n == 0) continue;
            make_pair_keys_kernel<<<nblk(n, 256), 256, 0, st>>>(g.nele, g.nx, g.idxX, (uint64_t)ndofX, keys, vals, (uint32_t)g.pair_base);
            h->launches++;
        }
        int end_bit = 1; while (end_bit < 64 && ((uint64_t)ndofX * (uint64_t)ndofX) >> end_bit) ++end_bit;
        void* tmp = nullptr; size_t tmpsz = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmpsz, keys, keys2, vals, h->src, npair, 0, end_bit, st));
        CK(cudaMalloc(&tmp, tmpsz ? tmpsz : 1));
        CK(cub::DeviceRadixSort::SortPairs(tmp, tmpsz, keys, keys2, vals, h->src, npair, 0, end_bit, st));
        CK(cudaStreamSynchronize(st)); cudaFree(tmp);
        // inclusive scan of "new non-zero" flags → inz per sorted pair
        cub::CountingInputIterator<int64_t> cnt(0);
        cub::TransformInputIterator<uint32_t, KeyFlag, cub::CountingInputIterator<int64_t>> flags(cnt, KeyFlag{keys2});
        tmp = nullptr; tmpsz = 0;
        CK(cub::DeviceScan::InclusiveSum(nullptr, tmpsz, flags, inz, npair, st));
        CK(cudaMalloc(&tmp, tmpsz ? tmpsz : 1));
        CK(cub::DeviceScan::InclusiveSum(tmp, tmpsz, flags, inz, npair, st));
        uint32_t last = 0;
        CK(cudaMemcpyAsync(&last, inz + (npair - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st)); cudaFree(tmp);
        nnz = last;
        // asm2 / colptr0 / rowval0 / the pair descriptors index non-zeros with 32 bits (signed in asm2): an EulerBeam3D chain has 108 non-zeros per element,
        // so the npair < 2^32 guard above (29.8 M beams) is not enough — 19.9 M elements already pass 2^31 non-zeros
        if (nnz > (int64_t)INT32_MAX) { h->err = "more than 2^31-1 non-zeros on one device: shard the element range (mb_iface_*)"; dfree(h, keys); dfree(h, keys2); dfree(h, vals); dfree(h, inz); return MB_ERR_TOOBIG; }
        CK(dalloc(h, &h->rowval0, nnz)); CK(dalloc(h, &h->cstart, nnz + 1));
        finish_pattern_kernel<<<nblk(npair, 256), 256, 0, st>>>(npair, keys2, h->src, inz, (uint64_t)ndofX, h->asm2, h->rowval0, h->cstart);
        colptr_kernel<<<nblk(nnz, 256), 256, 0, st>>>(nnz, keys2, h->cstart, (uint64_t)ndofX, h->colptr0);
        h->launches += 2;
        CK(cudaStreamSynchronize(st));
        dfree(h, keys); dfree(h, keys2); dfree(h, vals); dfree(h, inz);
    } else {
        CK(dalloc(h, &h->rowval0, 1)); CK(dalloc(h, &h->cstart, 1));
        CK(cudaMemsetAsync(h->cstart, 0, sizeof(uint32_t), st));
    }
    h->nnz = nnz;
    CK(dalloc(h, &h->nzval, nnz));
    { int32_t rcpd = build_pair_descriptors(h, nnz, h->cstart, h->src, &h->pdesc, &h->xdesc); if (rcpd) return rcpd; }
    // ---- fused epilogue of the static beam kernel: which element entries a warp of five whole elements finishes itself (kernels.cuh)
    h->fuse_groups.assign(h->groups.size(), mb_handle::FuseGroup());
    if (h->fuse && h->static_sym == 3 && nnz > 0) {
        bool any = false;
        for (const Group& g : h->groups) any |= (g.kind == G_BEAM && g.nele > 0);
        if (any) {
            CK(dalloc(h, &h->wflag, nnz + 4)); CK(cudaMemsetAsync(h->wflag, 0, (size_t)(nnz + 4), st));
            for (size_t ig = 0; ig < h->groups.size(); ++ig) {
                const Group& g = h->groups[ig];
                if (g.kind != G_BEAM || g.nele == 0) continue;
                mb_handle::FuseGroup& F = h->fuse_groups[ig];
                const int64_t nw = (g.nele + MB_EPW - 1) / MB_EPW, np = g.nele * 144;
                int32_t *wbase = nullptr, *wlen = nullptr, *wnun = nullptr; uint32_t *ebits = nullptr, *pat = nullptr; int64_t *psz = nullptr, *poff = nullptr, *hd = nullptr, *hsz = nullptr, *noff = nullptr;
                CK(dalloc(h, &wbase, nw)); CK(dalloc(h, &wlen, nw)); CK(dalloc(h, &wnun, nw)); CK(dalloc(h, &ebits, nw * 23));
                CK(dalloc(h, &psz, nw)); CK(dalloc(h, &poff, nw)); CK(dalloc(h, &hd, nw)); CK(dalloc(h, &hsz, nw)); CK(dalloc(h, &noff, nw));
                CK(cudaMemsetAsync(wbase, 0x7F, (size_t)nw * 4, st)); CK(cudaMemsetAsync(wlen, 0, (size_t)nw * 4, st));
                CK(cudaMemsetAsync(ebits, 0, (size_t)nw * 23 * 4, st));
                fuse_mark_kernel<<<nblk(np, 256), 256, 0, st>>>(np, (uint32_t)g.pair_base, h->asm2, h->cstart, h->src, wbase);
                fuse_dest_kernel<<<nblk(np, 256), 256, 0, st>>>(np, (uint32_t)g.pair_base, h->asm2, h->cstart, h->src, wbase, wlen, ebits, h->wflag);
                fuse_size_kernel<<<nblk(nw, 256), 256, 0, st>>>(nw, np, wlen, ebits, wnun, psz);
                void* tmp = nullptr; size_t t1 = 0, t2 = 0;
                CK(cub::DeviceScan::ExclusiveSum(nullptr, t1, psz, poff, (int)nw, st));
                CK(cub::DeviceScan::InclusiveScan(nullptr, t2, hd, hd, MaxI64(), (int)nw, st));
                CK(cudaMalloc(&tmp, std::max(t1, t2) + 1));
                CK(cub::DeviceScan::ExclusiveSum(tmp, t1, psz, poff, (int)nw, st));
                int64_t last[2] = {0, 0};
                CK(cudaMemcpyAsync(&last[0], poff + nw - 1, 8, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(&last[1], psz + nw - 1, 8, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                CK(dalloc(h, &pat, last[0] + last[1]));
                fuse_fill_kernel<<<nblk(nw * 32, 256), 256, 0, st>>>(nw, np, (uint32_t)g.pair_base, h->cstart, h->src, wbase, wlen, ebits, poff, psz, pat);
                fuse_eq_kernel<<<nblk(nw * 32, 256), 256, 0, st>>>(nw, wlen, wnun, poff, psz, pat, hd, hsz);
                CK(cub::DeviceScan::InclusiveScan(tmp, t2, hd, hd, MaxI64(), (int)nw, st));          // hd[w] = first warp of the run of equal patterns w belongs to
                CK(cub::DeviceScan::ExclusiveSum(tmp, t1, hsz, noff, (int)nw, st));
                CK(cudaMemcpyAsync(&last[0], noff + nw - 1, 8, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(&last[1], hsz + nw - 1, 8, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                if ((last[0] + last[1]) / 4 >= INT32_MAX) { cudaFree(tmp); h->err = "fused epilogue: pattern table too large"; return MB_ERR_TOOBIG; }
                CK(dalloc(h, &F.cpat, last[0] + last[1])); CK(dalloc(h, &F.whdr, nw));
                fuse_compact_kernel<<<nblk(nw * 32, 256), 256, 0, st>>>(nw, hd, poff, psz, noff, pat, F.cpat, wbase, wlen, wnun, F.whdr);
                h->launches += 9;
                CK(cudaStreamSynchronize(st)); CK(cudaGetLastError());
                cudaFree(tmp);
                dfree(h, wbase); dfree(h, wlen); dfree(h, wnun); dfree(h, ebits); dfree(h, pat); dfree(h, psz); dfree(h, poff); dfree(h, hd); dfree(h, hsz); dfree(h, noff);
            }
            // compact ascending list of the non-zeros that still go through the segmented reduction
            int64_t* dn = nullptr; CK(dalloc(h, &dn, 1));
            cub::CountingInputIterator<int32_t> it0(0);
            void* tmp = nullptr; size_t tmpsz = 0, tmpsz2 = 0;
            CK(cub::DeviceSelect::If(nullptr, tmpsz, it0, (int32_t*)nullptr, dn, (int)nnz, NotFlag{h->wflag}, st));
            CK(cub::DeviceSelect::If(nullptr, tmpsz2, it0, cub::DiscardOutputIterator<int32_t>(), dn, (int)nnz, NotFlag{h->wflag}, st));
            tmpsz = std::max(tmpsz, tmpsz2);
            CK(cudaMalloc(&tmp, tmpsz ? tmpsz : 1));
            CK(cub::DeviceSelect::If(tmp, tmpsz2, it0, cub::DiscardOutputIterator<int32_t>(), dn, (int)nnz, NotFlag{h->wflag}, st));      // count first: the list is a tenth of nnz on a chain
            CK(cudaMemcpyAsync(&h->nulist, dn, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            CK(dalloc(h, &h->ulist, std::max<int64_t>(h->nulist, 1)));
            CK(cub::DeviceSelect::If(tmp, tmpsz, it0, h->ulist, dn, (int)nnz, NotFlag{h->wflag}, st));
            CK(dalloc(h, &h->udesc, std::max<int64_t>(h->nulist, 1)));
            if (h->nulist > 0) { list_desc_kernel<<<nblk(h->nulist, 256), 256, 0, st>>>(h->nulist, h->ulist, h->cstart, h->src, h->udesc); h->launches++; }
            CK(cudaStreamSynchronize(st)); cudaFree(tmp); dfree(h, dn); dfree(h, h->ulist); h->ulist = nullptr;
        }
    }

    // ---- vector map: contributors of every dof in element order (asmvec!, src/Assemble.jl:340-357)
    if (nvec > 0) {
        uint32_t *keys = nullptr, *keys2 = nullptr, *vals = nullptr;
        CK(dalloc(h, &keys, nvec)); CK(dalloc(h, &keys2, nvec)); CK(dalloc(h, &vals, nvec));
        for (const Group& g : h->groups) {
            const int64_t n = g.nele * g.nx;
            if (n == 0) continue;
            make_vec_keys_kernel<<<nblk(n, 256), 256, 0, st>>>(n, g.idxX, keys, vals, (uint32_t)g.vec_base);
            h->launches++;
        }
        int end_bit = 1; while (end_bit < 32 && ((uint64_t)ndofX >> end_bit)) ++end_bit;
        void* tmp = nullptr; size_t tmpsz = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmpsz, keys, keys2, vals, h->vsrc, nvec, 0, end_bit, st));
        CK(cudaMalloc(&tmp, tmpsz ? tmpsz : 1));
        CK(cub::DeviceRadixSort::SortPairs(tmp, tmpsz, keys, keys2, vals, h->vsrc, nvec, 0, end_bit, st));
        vstart_kernel<<<nblk(nvec, 256), 256, 0, st>>>(nvec, keys2, ndofX, h->vstart);
        h->launches++;
        CK(cudaStreamSynchronize(st)); cudaFree(tmp);
        dfree(h, keys); dfree(h, keys2); dfree(h, vals);
    }
    CK(cudaGetLastError());
    h->prepared = true;
    if (nnz_out) *nnz_out = nnz;
    return MB_OK;
}

int32_t mb_sweepx_get_pattern(mb_handle* h, int64_t* colptr, int64_t* rowval) {
    if (!h) return MB_ERR_ARG;
    ARG(h->prepared && colptr && rowval, "not prepared / null");
    CK(cudaSetDevice(h->device));
    // widen on the device in chunks, 1-based like SparseMatrixCSC
    const int64_t CH = 1 << 24;
    int64_t* buf = nullptr;
    CK(cudaMalloc((void**)&buf, CH * sizeof(int64_t)));
    auto widen = [&](const int32_t* in, int64_t n, int64_t* out) -> cudaError_t {
        for (int64_t o = 0; o < n; o += CH) {
            const int64_t m = (n - o < CH) ? n - o : CH;
            widen_plus1_kernel<<<nblk(m, 256), 256, 0, h->stream>>>(m, in + o, buf, 1);
            h->launches++;
            cudaError_t e = cudaMemcpyAsync(out + o, buf, m * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream);
            if (e != cudaSuccess) return e;
            e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    cudaError_t e1 = widen(h->colptr0, h->ndofX + 1, colptr);
    cudaError_t e2 = (e1 == cudaSuccess) ? widen(h->rowval0, h->nnz, rowval) : e1;
    cudaFree(buf);
    CK(e2);
    return MB_OK;
}

int32_t mb_sweepx_get_asm_range(mb_handle* h, int32_t ieletyp, int64_t e0, int64_t e1, int64_t* asm1, int64_t* asm2) {
    if (!h) return MB_ERR_ARG;
    ARG(h->prepared && ieletyp >= 1 && ieletyp <= (int32_t)h->groups.size(), "bad element type / not prepared");
    const Group& g = h->groups[ieletyp - 1];
    ARG(e0 >= 0 && e1 >= e0 && e1 <= g.nele, "element range");
    CK(cudaSetDevice(h->device));
    const int64_t n1 = (e1 - e0) * g.nx, n2 = n1 * g.nx;
    std::vector<int32_t> t((size_t)(n2 > n1 ? n2 : n1));
    if (asm1 && n1) { CK(cudaMemcpy(t.data(), g.idxX + e0 * g.nx, (size_t)n1 * 4, cudaMemcpyDeviceToHost)); for (int64_t i = 0; i < n1; ++i) asm1[i] = (int64_t)t[(size_t)i] + 1; }
    if (asm2 && n2) { CK(cudaMemcpy(t.data(), h->asm2 + g.pair_base + e0 * g.nx * g.nx, (size_t)n2 * 4, cudaMemcpyDeviceToHost)); for (int64_t i = 0; i < n2; ++i) asm2[i] = t[(size_t)i]; }
    return MB_OK;
}

int32_t mb_sweepx_get_asm(mb_handle* h, int32_t ieletyp, int64_t* asm1, int64_t* asm2) {
    if (!h) return MB_ERR_ARG;
    ARG(h->prepared && ieletyp >= 1 && ieletyp <= (int32_t)h->groups.size(), "bad element type / not prepared");
    CK(cudaSetDevice(h->device));
    const Group& g = h->groups[ieletyp - 1];
    const int64_t CH = 1 << 24;
    int64_t* buf = nullptr;
    CK(cudaMalloc((void**)&buf, CH * sizeof(int64_t)));
    auto widen = [&](const int32_t* in, int64_t n, int64_t* out, int add) -> cudaError_t {
        for (int64_t o = 0; o < n; o += CH) {
            const int64_t m = (n - o < CH) ? n - o : CH;
            widen_plus1_kernel<<<nblk(m, 256), 256, 0, h->stream>>>(m, in + o, buf, add);
            h->launches++;
            cudaError_t e = cudaMemcpyAsync(out + o, buf, m * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream);
            if (e != cudaSuccess) return e;
            e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    cudaError_t e = cudaSuccess;
    if (asm1) e = widen(g.idxX, g.nele * g.nx, asm1, 1);                                   // asm[1]: X-dof group is the identity map
    if (asm2 && e == cudaSuccess) e = widen(h->asm2 + g.pair_base, g.nele * g.nx * g.nx, asm2, 0);
    cudaFree(buf);
    CK(e);
    return MB_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------ assemble
template <int ND, bool STEP> static void launch_beam_w(mb_handle* h, const Group& g, const BeamGroupDev& gd, const StateDev& sd, const NewmarkDev& nm,
                                                       unsigned long long nanbase, int64_t e0) {
    double* Wc = nullptr;
    if (ND >= 2 && h->split_dyn) {                 // two-phase Newmark kernel: 45 doubles per lane of workspace (lazily sized to the largest launch)
        const int64_t need = ((gd.nele * 6 + 31) / 32) * 32 * MB_NCOT;
        if (h->Wc_len < need) { if (h->Wc) dfree(h, h->Wc); h->Wc = nullptr; h->Wc_len = 0; if (dalloc(h, &h->Wc, need) == cudaSuccess) h->Wc_len = need; else cudaGetLastError(); }
        Wc = h->Wc_len >= need ? h->Wc : nullptr;  // falls back to the fused kernel when the workspace does not fit
    }
    BeamLaunch a{gd, sd, nm, h->Ke + g.pair_base + e0 * 144, h->Re + g.vec_base + e0 * 12, h->Rp + g.vec_base + e0 * 12, h->nanflag, nanbase + (unsigned long long)e0,
                 h->beamW, h->stream, Wc};
    a.static_sym = h->static_sym; a.nsm = h->nsm;
    if (ND == 1 && h->wflag) {                     // statics: entries a warp can finish go straight to nzval (e0 is a multiple of MB_EPW: the chunk plan keeps it so)
        const size_t ig = (size_t)(&g - h->groups.data());
        const mb_handle::FuseGroup& F = h->fuse_groups[ig];
        if (F.whdr && e0 % MB_EPW == 0) a.fz = FuseDev{F.whdr, F.cpat, h->nzval, e0 / MB_EPW};
    }
    launch_beam<ND, STEP>(a);
    h->launches += 1 + (Wc ? 1 : 0) + (STEP ? (Wc ? 2 : 1) : 0);
}

// elements [e0,e1) of group ig (bar / soil groups are always launched whole)
static int32_t launch_group_range(mb_handle* h, size_t ig, int64_t e0, int64_t e1, int OX, int mission, const NewmarkDev& nm, double tnow) {
    const bool step = (mission == 0) && OX > 0;
    StateDev sd{h->X0, h->X1, h->X2, h->ndofU > 0 ? h->U0 : nullptr};
    const Group& g = h->groups[ig];
    h->fused_now = (OX == 0) && h->wflag != nullptr;        // the segmented reduction that follows skips what the static kernel finishes itself
    if (g.nele == 0 || e1 <= e0) return MB_OK;
    if (h->fused_now && g.kind == G_BEAM && e0 % MB_EPW != 0) { h->err = "element range of a static launch does not start at a whole warp"; return MB_ERR_STATE; }
    const unsigned long long nanbase = ((unsigned long long)ig) << 40;
    if (g.kind == G_BEAM) {
        BeamGroupDev gd;
        gd.nele = e1 - e0; gd.geo = g.geo + e0 * 16; gd.mats = g.mats; gd.mat_id = g.mat_id ? g.mat_id + e0 : nullptr;
        gd.idxX = g.idxX + e0 * 12; gd.idxU = g.idxU ? g.idxU + e0 * 3 : nullptr; gd.udof = g.udof;
        for (int i = 0; i < 12; ++i) gd.scaleX[i] = g.scaleX[i];
        for (int i = 0; i < 3; ++i) gd.scaleU[i] = g.scaleU[i];
        if (OX == 0) launch_beam_w<1, false>(h, g, gd, sd, nm, nanbase, e0);
        else if (OX == 1) { if (step) launch_beam_w<2, true>(h, g, gd, sd, nm, nanbase, e0); else launch_beam_w<2, false>(h, g, gd, sd, nm, nanbase, e0); }
        else { if (step) launch_beam_w<3, true>(h, g, gd, sd, nm, nanbase, e0); else launch_beam_w<3, false>(h, g, gd, sd, nm, nanbase, e0); }
    }
    else if (g.kind == G_BAR) {
        BarGroupDev gd; gd.nele = g.nele; gd.geo = g.geo; gd.mats = g.barmats; gd.mat_id = g.mat_id; gd.idxX = g.idxX; gd.idxU = g.idxU; gd.udof = g.udof;
        for (int i = 0; i < 6; ++i) gd.scaleX[i] = g.scaleX[i];
        for (int i = 0; i < 3; ++i) gd.scaleU[i] = g.scaleU[i];
        launch_bar(OX + 1, step, gd, sd, nm, tnow, h->Ke + g.pair_base, h->Re + g.vec_base, h->Rp + g.vec_base, h->nanflag, nanbase, h->stream);
        h->launches++;
    } else if (g.kind == G_SOIL) {
        SoilGroupDev gd; gd.nele = g.nele; gd.par = g.geo; gd.idxX = g.idxX;
        for (int i = 0; i < 3; ++i) gd.scaleX[i] = g.scaleX[i];
        launch_soil(OX + 1, step, gd, sd, nm, h->Ke + g.pair_base, h->Re + g.vec_base, h->Rp + g.vec_base, h->nanflag, nanbase, h->stream);
        h->launches++;
    }
    // G_HOST: contributions were uploaded by mb_set_host_elements
    return MB_OK;
}
static int32_t launch_elements(mb_handle* h, int OX, int mission, const NewmarkDev& nm, double tnow) {
    for (size_t ig = 0; ig < h->groups.size(); ++ig) { int32_t rc = launch_group_range(h, ig, 0, h->groups[ig].nele, OX, mission, nm, tnow); if (rc) return rc; }
    return MB_OK;
}
// segmented reductions of non-zeros [k0,k1) (k0 a multiple of 4) and dofs [d0,d1)
static void launch_gather_range(mb_handle* h, bool step, int64_t k0, int64_t k1, int64_t d0, int64_t d1, cudaStream_t st = nullptr) {
    if (!st) st = h->stream;
    const uint8_t* wf = (h->fused_now && h->wflag) ? h->wflag + k0 : nullptr;      // non-zeros the static element kernel has finished already
    if (k1 > k0) { gather_nz_kernel<<<nblk((k1 - k0 + 3) / 4, h->gather_block), h->gather_block, 0, st>>>(k1 - k0, h->cstart + k0, h->src, h->pdesc + k0, h->xdesc, h->Ke, h->nzval + k0, wf); h->launches++; }
    if (d1 > d0) { gather_vec_kernel<<<nblk(d1 - d0, h->gather_block), h->gather_block, 0, st>>>(d0, d1, h->vstart, h->vsrc, h->Re, step ? h->Rp : nullptr, h->Ll); h->launches++; }
}
static void launch_gather(mb_handle* h, bool step) {
    if (h->fused_now && h->udesc) {                  // statics with the fused epilogue: only the listed non-zeros are still to be summed
        if (h->nulist > 0) { gather_list_kernel<<<nblk((h->nulist + 1) / 2, 256), 256, 0, h->stream>>>(h->nulist, h->udesc, h->cstart, h->src, h->Ke, h->nzval); h->launches++; }
        launch_gather_range(h, step, 0, 0, 0, h->ndofX);
        return;
    }
    launch_gather_range(h, step, 0, h->nnz, 0, h->ndofX);
}

// ---- host-buffer pipeline: chunk plan + completion prefixes (built once, at the first large mb_sweepx_assemble)
static int32_t build_pipeline(mb_handle* h) {
    h->pipe_state = -1;
    cudaStream_t st = h->stream;
    int64_t work = 0;
    for (const Group& g : h->groups) if (g.kind != G_HOST) work += g.nele * g.nx * g.nx;
    if (work == 0 || h->nnz == 0 || h->pipe_chunks < 2) return MB_OK;
    const int C = h->pipe_chunks;
    std::vector<int64_t> pair_end((size_t)C, 0), vec_end((size_t)C, 0);
    h->pipe_items.clear();
    int64_t done = 0; int chunk = 0;
    for (size_t ig = 0; ig < h->groups.size(); ++ig) {
        const Group& g = h->groups[ig];
        if (g.kind == G_HOST || g.nele == 0) continue;
        const int64_t per = (int64_t)g.nx * g.nx;
        int64_t e = 0;
        while (e < g.nele) {
            int64_t e1 = g.nele;
            if (g.kind == G_BEAM) {                  // fill the current chunk up to its share of the work
                const int64_t target = (work * (chunk + 1) + C - 1) / C;
                const int64_t room = (target - done + per - 1) / per;
                e1 = std::min(g.nele, e + std::max<int64_t>(room, 1));
                if (e1 < g.nele) e1 = std::min(g.nele, e + std::max<int64_t>(MB_EPW, ((e1 - e) / MB_EPW) * MB_EPW));      // whole warps of the static kernel (fused epilogue)
            }
            h->pipe_items.push_back({(int)ig, e, e1, chunk});
            done += (e1 - e) * per;
            for (int j = chunk; j < C; ++j) { pair_end[(size_t)j] = g.pair_base + e1 * per; vec_end[(size_t)j] = g.vec_base + e1 * g.nx; }
            e = e1;
            while (chunk < C - 1 && done >= (work * (chunk + 1) + C - 1) / C) ++chunk;
        }
    }
    // completion prefixes on the device
    SkipRanges skp{0, {0}, {0}}, skv{0, {0}, {0}};
    for (const Group& g : h->groups)
        if (g.kind == G_HOST && g.nele > 0) {
            if (skp.n == 8) return MB_OK;            // too many host-evaluated types: keep the one-shot path
            skp.lo[skp.n] = (uint32_t)g.pair_base; skp.hi[skp.n] = (uint32_t)(g.pair_base + g.nele * g.nx * g.nx); ++skp.n;
            skv.lo[skv.n] = (uint32_t)g.vec_base; skv.hi[skv.n] = (uint32_t)(g.vec_base + g.nele * g.nx); ++skv.n;
        }
    uint32_t* last = nullptr; int64_t *bound = nullptr, *cnt = nullptr;
    CK(dalloc(h, &last, std::max(h->nnz, h->ndofX))); CK(dalloc(h, &bound, C)); CK(dalloc(h, &cnt, C));
    h->pipe_nz_end.assign((size_t)C, 0); h->pipe_vec_end.assign((size_t)C, 0);
    for (int pass = 0; pass < 2; ++pass) {
        const int64_t nseg = pass == 0 ? h->nnz : h->ndofX;
        seg_last_kernel<<<nblk(nseg, 256), 256, 0, st>>>(nseg, pass == 0 ? h->cstart : h->vstart, pass == 0 ? h->src : h->vsrc, pass == 0 ? skp : skv, last);
        void* tmp = nullptr; size_t tmpsz = 0;
        CK(cub::DeviceScan::InclusiveScan(nullptr, tmpsz, last, last, cub::Max(), nseg, st));
        CK(cudaMalloc(&tmp, tmpsz ? tmpsz : 1));
        CK(cub::DeviceScan::InclusiveScan(tmp, tmpsz, last, last, cub::Max(), nseg, st));
        CK(cudaMemcpyAsync(bound, pass == 0 ? pair_end.data() : vec_end.data(), (size_t)C * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        prefix_count_kernel<<<1, 64, 0, st>>>(nseg, last, C, bound, cnt);
        CK(cudaMemcpyAsync(pass == 0 ? h->pipe_nz_end.data() : h->pipe_vec_end.data(), cnt, (size_t)C * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st)); cudaFree(tmp);
        h->launches += 3;
    }
    dfree(h, last); dfree(h, bound); dfree(h, cnt);
    for (int j = 0; j < C - 1; ++j) h->pipe_nz_end[(size_t)j] &= ~(int64_t)3;     // ranges of the 4-per-thread reduction start at multiples of 4
    h->pipe_nz_end[(size_t)C - 1] = h->nnz; h->pipe_vec_end[(size_t)C - 1] = h->ndofX;
    if (!h->copy_stream) CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (!h->h2d_stream) CK(cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
    while ((int)h->pipe_ev.size() < C) { cudaEvent_t ev; CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); h->pipe_ev.push_back(ev); }
    while ((int)h->pipe_evx.size() < C + 1) { cudaEvent_t ev; CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); h->pipe_evx.push_back(ev); }
    // dofs of the state an element chunk reads: one past the largest X-dof number of the chunks up to it (host-evaluated types do not read the device state)
    {
        h->pipe_x_need.assign((size_t)C, 0);
        int32_t* dmax = nullptr; CK(dalloc(h, &dmax, 1));
        void* tmp = nullptr; size_t tsz = 0;
        for (const mb_handle::PipeItem& w : h->pipe_items) {
            const Group& g = h->groups[(size_t)w.ig];
            const int64_t n = (w.e1 - w.e0) * g.nx;
            if (n <= 0 || !g.idxX) continue;
            size_t need = 0;
            CK(cub::DeviceReduce::Max(nullptr, need, g.idxX + w.e0 * g.nx, dmax, n, st));
            if (need > tsz) { if (tmp) cudaFree(tmp); CK(cudaMalloc(&tmp, need)); tsz = need; }
            CK(cub::DeviceReduce::Max(tmp, need, g.idxX + w.e0 * g.nx, dmax, n, st));
            int32_t m = 0; CK(cudaMemcpyAsync(&m, dmax, 4, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
            h->pipe_x_need[(size_t)w.chunk] = std::max<int64_t>(h->pipe_x_need[(size_t)w.chunk], (int64_t)m + 1);
        }
        if (tmp) cudaFree(tmp);
        dfree(h, dmax);
        for (int j = 1; j < C; ++j) h->pipe_x_need[(size_t)j] = std::max(h->pipe_x_need[(size_t)j], h->pipe_x_need[(size_t)j - 1]);
    }
    h->pipe_state = 1;
    return MB_OK;
}

extern "C" {

int32_t mb_sweepx_assemble_dev(mb_handle* h, int32_t OX, int32_t mission, double t, const double* newmark) {
    if (!h) return MB_ERR_ARG;
    ARG(h->prepared, "call mb_sweepx_prepare first");
    ARG(OX >= 0 && OX <= 2 && (mission == 0 || mission == 1) && newmark, "bad OX / mission / newmark");
    CK(cudaSetDevice(h->device));
    NewmarkDev nm{newmark[0], newmark[1], newmark[2], newmark[3], newmark[4], newmark[5], newmark[6]};
    CK(cudaMemsetAsync(h->nanflag, 0xFF, sizeof(unsigned long long), h->stream));
    if (h->dev_overlap && h->pipe_state == 0 && h->nnz >= h->pipe_min_nnz) { int32_t rc = build_pipeline(h); if (rc) return rc; }
    if (h->dev_overlap && h->pipe_state == 1) {
        // chunk j's non-zeros are reduced on the high-priority stream while the element kernels of chunk j+1 run; results are visible to
        // h->stream when the call returns (event wait), so callers keep ordering against h->stream only
        if (!h->gather_stream) {
            int lo = 0, hi = 0;
            CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CK(cudaStreamCreateWithPriority(&h->gather_stream, cudaStreamNonBlocking, h->dev_overlap == 2 ? lo : hi));      // 2: below the element kernels — the reduction takes what they leave
            CK(cudaEventCreateWithFlags(&h->gather_done, cudaEventDisableTiming));
        }
        const bool step = mission == 0 && OX > 0;
        const int C = h->pipe_chunks;
        size_t it = 0; int64_t k0 = 0, d0 = 0;
        for (int j = 0; j < C; ++j) {
            for (; it < h->pipe_items.size() && h->pipe_items[it].chunk == j; ++it) {
                const mb_handle::PipeItem& w = h->pipe_items[it];
                int32_t rc = launch_group_range(h, (size_t)w.ig, w.e0, w.e1, OX, mission, nm, t);
                if (rc) return rc;
            }
            const int64_t k1 = h->pipe_nz_end[(size_t)j], d1 = h->pipe_vec_end[(size_t)j];
            if (k1 > k0 || d1 > d0) {
                CK(cudaEventRecord(h->pipe_ev[(size_t)j], h->stream));
                CK(cudaStreamWaitEvent(h->gather_stream, h->pipe_ev[(size_t)j], 0));
                launch_gather_range(h, step, k0, k1, d0, d1, h->gather_stream);
            }
            k0 = k1; d0 = d1;
        }
        CK(cudaEventRecord(h->gather_done, h->gather_stream));
        CK(cudaStreamWaitEvent(h->stream, h->gather_done, 0));
        CK(cudaGetLastError());
        return MB_OK;
    }
    int32_t rc = launch_elements(h, OX, mission, nm, t);
    if (rc) return rc;
    launch_gather(h, mission == 0 && OX > 0);
    CK(cudaGetLastError());
    return MB_OK;
}

// whole device-resident steps (mb_sweepx_assemble_dev as configured, overlapped or not), CUDA events on the engine's stream
int32_t mb_sweepx_time_step_dev(mb_handle* h, int32_t OX, int32_t mission, double t, const double* newmark, int32_t reps, float* ms) {
    if (!h) return MB_ERR_ARG;
    ARG(h->prepared && reps >= 1 && ms && newmark, "bad argument");
    CK(cudaSetDevice(h->device));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, h->stream));
    const bool exch = h->comm && (h->if_nsend_nz + h->if_nsend_v + h->if_nrecv_nz + h->if_nrecv_v) > 0;      // element-range shard: a step ends with the interface exchange
    for (int r = 0; r < reps; ++r) {
        int32_t rc = mb_sweepx_assemble_dev(h, OX, mission, t, newmark);
        if (!rc && exch) rc = mb_iface_exchange(h);
        if (rc) { cudaEventDestroy(e0); cudaEventDestroy(e1); return rc; }
    }
    CK(cudaEventRecord(e1, h->stream));
    CK(cudaEventSynchronize(e1));
    float a; CK(cudaEventElapsedTime(&a, e0, e1));
    *ms = a / reps;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return MB_OK;
}

int32_t mb_sync(mb_handle* h, mb_errinfo* where) {
    if (!h) return MB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->nanflag_host, h->nanflag, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (where) { where->kind = 0; where->ieletyp = 0; where->iele = 0; where->step = 0; }
    const unsigned long long f = *h->nanflag_host;
    if (f != ~0ULL) {
        if (where) { where->kind = MB_ERR_NAN; where->ieletyp = (int32_t)(f >> 40) + 1; where->iele = (int64_t)(f & ((1ULL << 40) - 1)) + 1; }
        h->err = "residual(...) returned NaN in R, FB or derivatives";
        return MB_ERR_NAN;
    }
    return MB_OK;
}

int32_t mb_sweepx_assemble(mb_handle* h, int32_t OX, int32_t mission, const double* X0, const double* X1, const double* X2,
                           const double* U0, double t, const double* newmark, double* Llambda, double* nzval, mb_errinfo* where) {
    if (!h) return MB_ERR_ARG;
    ARG(h->prepared, "call mb_sweepx_prepare first");
    ARG(!X0 || ((OX < 1 || X1) && (OX < 2 || X2)), "state vectors missing for this OX");
    CK(cudaSetDevice(h->device));
    const size_t nb = (size_t)h->ndofX * sizeof(double);
    // chunked pipeline: with host nzval (D2H of the CSC values overlapped), or with host state in and host Lλ out (H2D of the state, the element kernels and the
    // D2H of Lλ overlapped: the state arrives in ascending dof order and element chunk j starts as soon as the dofs it reads are there)
    const bool want_pipe = (nzval != nullptr) || (X0 != nullptr && Llambda != nullptr);
    if (h->pipe_state == 0 && want_pipe && h->nnz >= h->pipe_min_nnz) { int32_t rc = build_pipeline(h); if (rc) return rc; }
    const bool piped = h->pipe_state == 1 && want_pipe;
    const bool piped_in = piped && X0 != nullptr && !h->pipe_x_need.empty();
    if (X0 && !piped_in) {                         // X0 == NULL: assemble at the device-resident state (mb_sweepx_set_state / _newmark_decrement)
        CK(cudaMemcpyAsync(h->X0, X0, nb, cudaMemcpyHostToDevice, h->stream));
        if (OX >= 1) CK(cudaMemcpyAsync(h->X1, X1, nb, cudaMemcpyHostToDevice, h->stream));
        if (OX >= 2) CK(cudaMemcpyAsync(h->X2, X2, nb, cudaMemcpyHostToDevice, h->stream));
    }
    if (U0 && h->ndofU > 0) CK(cudaMemcpyAsync(h->U0, U0, (size_t)h->ndofU * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (piped) {
        // chunked: evaluate element range j, reduce the non-zeros / dofs it completes, ship them while range j+1 computes
        ARG(OX >= 0 && OX <= 2 && (mission == 0 || mission == 1) && newmark, "bad OX / mission / newmark");
        NewmarkDev nm{newmark[0], newmark[1], newmark[2], newmark[3], newmark[4], newmark[5], newmark[6]};
        const bool step = mission == 0 && OX > 0;
        CK(cudaMemsetAsync(h->nanflag, 0xFF, sizeof(unsigned long long), h->stream));
        const int C = h->pipe_chunks;
        size_t it = 0; int64_t k0 = 0, d0 = 0, x0 = 0;
        if (piped_in) {                            // the state, in C pieces on its own stream (after whatever the engine's stream still does with the old state)
            CK(cudaEventRecord(h->pipe_evx[(size_t)C], h->stream));
            CK(cudaStreamWaitEvent(h->h2d_stream, h->pipe_evx[(size_t)C], 0));
            const double* src[3] = {X0, X1, X2}; double* dst[3] = {h->X0, h->X1, h->X2};
            for (int j = 0; j < C; ++j) {
                const int64_t x1 = (j == C - 1) ? h->ndofX : std::min(h->ndofX, h->pipe_x_need[(size_t)j]);
                if (x1 > x0) for (int d = 0; d <= OX; ++d) CK(cudaMemcpyAsync(dst[d] + x0, src[d] + x0, (size_t)(x1 - x0) * sizeof(double), cudaMemcpyHostToDevice, h->h2d_stream));
                CK(cudaEventRecord(h->pipe_evx[(size_t)j], h->h2d_stream));
                x0 = std::max(x0, x1);
            }
        }
        for (int j = 0; j < C; ++j) {
            if (piped_in) CK(cudaStreamWaitEvent(h->stream, h->pipe_evx[(size_t)j], 0));
            for (; it < h->pipe_items.size() && h->pipe_items[it].chunk == j; ++it) {
                const mb_handle::PipeItem& w = h->pipe_items[it];
                int32_t rc = launch_group_range(h, (size_t)w.ig, w.e0, w.e1, OX, mission, nm, t);
                if (rc) return rc;
            }
            const int64_t k1 = h->pipe_nz_end[(size_t)j], d1 = h->pipe_vec_end[(size_t)j];
            launch_gather_range(h, step, k0, k1, d0, d1);
            CK(cudaEventRecord(h->pipe_ev[(size_t)j], h->stream));
            CK(cudaStreamWaitEvent(h->copy_stream, h->pipe_ev[(size_t)j], 0));
            if (nzval && k1 > k0) CK(cudaMemcpyAsync(nzval + k0, h->nzval + k0, (size_t)(k1 - k0) * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
            if (Llambda && d1 > d0) CK(cudaMemcpyAsync(Llambda + d0, h->Ll + d0, (size_t)(d1 - d0) * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
            k0 = k1; d0 = d1;
        }
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(h->copy_stream));
        return mb_sync(h, where);
    }
    int32_t rc = mb_sweepx_assemble_dev(h, OX, mission, t, newmark);
    if (rc) return rc;
    if (Llambda) CK(cudaMemcpyAsync(Llambda, h->Ll, nb, cudaMemcpyDeviceToHost, h->stream));
    if (nzval && h->nnz > 0) CK(cudaMemcpyAsync(nzval, h->nzval, (size_t)h->nnz * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    return mb_sync(h, where);
}

int32_t mb_sweepx_set_state(mb_handle* h, int32_t OX, const double* X0, const double* X1, const double* X2, const double* U0) {
    if (!h) return MB_ERR_ARG;
    ARG(h->prepared && OX >= 0 && OX <= 2, "not prepared / bad OX");
    ARG(X0 && (OX < 1 || X1) && (OX < 2 || X2), "state vectors missing for this OX");
    CK(cudaSetDevice(h->device));
    const size_t nb = (size_t)h->ndofX * sizeof(double);
    const double* src[3] = {X0, X1, X2}; double* dst[3] = {h->X0, h->X1, h->X2};
    for (int d = 0; d <= OX; ++d) CK(cudaMemcpyAsync(dst[d], src[d], nb, cudaMemcpyDefault, h->stream));
    if (U0 && h->ndofU > 0) CK(cudaMemcpyAsync(h->U0, U0, (size_t)h->ndofU * sizeof(double), cudaMemcpyDefault, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MB_OK;
}
int32_t mb_sweepx_get_state(mb_handle* h, int32_t OX, double* X0, double* X1, double* X2) {
    if (!h) return MB_ERR_ARG;
    ARG(h->prepared && OX >= 0 && OX <= 2, "not prepared / bad OX");
    CK(cudaSetDevice(h->device));
    const size_t nb = (size_t)h->ndofX * sizeof(double);
    double* dst[3] = {X0, X1, X2}; const double* src[3] = {h->X0, h->X1, h->X2};
    for (int d = 0; d <= OX; ++d) if (dst[d]) CK(cudaMemcpyAsync(dst[d], src[d], nb, cudaMemcpyDefault, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MB_OK;
}
int32_t mb_sweepx_set_dof_scale(mb_handle* h, const double* scaleX) {
    if (!h) return MB_ERR_ARG;
    ARG(h->prepared && scaleX, "not prepared / null");
    CK(cudaSetDevice(h->device));
    if (!h->dofscale) CK(dalloc(h, &h->dofscale, h->ndofX));
    CK(cudaMemcpyAsync(h->dofscale, scaleX, (size_t)h->ndofX * sizeof(double), cudaMemcpyDefault, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MB_OK;
}
int32_t mb_sweepx_newmark_decrement(mb_handle* h, int32_t OX, int32_t firstiter, const double* dx, const double* newmark, double* dx2, double* Ll2) {
    if (!h) return MB_ERR_ARG;
    ARG(h->prepared && OX >= 0 && OX <= 2 && dx && (OX == 0 || newmark), "not prepared / bad OX / null");
    CK(cudaSetDevice(h->device));
    const int64_t n = h->ndofX;
    cudaPointerAttributes at;
    const bool on_dev = cudaPointerGetAttributes(&at, dx) == cudaSuccess && at.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    const double* d = dx;
    if (!on_dev) {
        if (!h->dxbuf) CK(dalloc(h, &h->dxbuf, n));
        CK(cudaMemcpyAsync(h->dxbuf, dx, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        d = h->dxbuf;
    }
    NewmarkDev c{0, 0, 0, 0, 0, 0, 0};
    if (newmark) c = NewmarkDev{newmark[0], newmark[1], newmark[2], newmark[3], newmark[4], newmark[5], newmark[6]};
    if (dx2 || Ll2) {                              // Σ Δx², Σ Lλ² (src/SweepX.jl:209) before the state moves
        if (!h->red) CK(dalloc(h, &h->red, 2));
        for (int k = 0; k < 2; ++k) {
            if (!(k == 0 ? dx2 : Ll2)) continue;
            cub::TransformInputIterator<double, Square, const double*> it(k == 0 ? d : h->Ll, Square{});
            size_t need = 0;
            CK(cub::DeviceReduce::Sum(nullptr, need, it, h->red + k, n, h->stream));
            if (need > h->redtmp_sz) { if (h->redtmp) cudaFree(h->redtmp); CK(cudaMalloc(&h->redtmp, need)); h->redtmp_sz = need; }
            CK(cub::DeviceReduce::Sum(h->redtmp, need, it, h->red + k, n, h->stream));
            h->launches += 1;
        }
    }
    if (OX == 0) newmark_decrement_kernel<0><<<nblk(n, 256), 256, 0, h->stream>>>(n, d, h->dofscale, h->X0, h->X1, h->X2, c, firstiter);
    else if (OX == 1) newmark_decrement_kernel<1><<<nblk(n, 256), 256, 0, h->stream>>>(n, d, h->dofscale, h->X0, h->X1, h->X2, c, firstiter);
    else newmark_decrement_kernel<2><<<nblk(n, 256), 256, 0, h->stream>>>(n, d, h->dofscale, h->X0, h->X1, h->X2, c, firstiter);
    h->launches++;
    double r[2] = {0., 0.};
    if (dx2 || Ll2) CK(cudaMemcpyAsync(r, h->red, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (dx2) *dx2 = r[0];
    if (Ll2) *Ll2 = r[1];
    CK(cudaGetLastError());
    return MB_OK;
}

// getresult(state,req,els) for one EulerBeam3D element type (src/Output.jl:131-181): 77 values per element, layout in include/muscade_b200.h
int32_t mb_beam_results(mb_handle* h, int32_t ieletyp, int32_t OX, const double* X0, const double* X1, const double* X2, double* out) {
    if (!h) return MB_ERR_ARG;
    ARG(h->prepared && OX >= 0 && OX <= 2 && out && ieletyp >= 1 && ieletyp <= (int32_t)h->groups.size(), "bad argument");
    const Group& g = h->groups[(size_t)ieletyp - 1];
    ARG(g.kind == G_BEAM, "element type is not EulerBeam3D");
    ARG(!X0 || ((OX < 1 || X1) && (OX < 2 || X2)), "state vectors missing for this OX");
    CK(cudaSetDevice(h->device));
    const size_t nb = (size_t)h->ndofX * sizeof(double);
    if (X0) {
        CK(cudaMemcpyAsync(h->X0, X0, nb, cudaMemcpyDefault, h->stream));
        if (OX >= 1) CK(cudaMemcpyAsync(h->X1, X1, nb, cudaMemcpyDefault, h->stream));
        if (OX >= 2) CK(cudaMemcpyAsync(h->X2, X2, nb, cudaMemcpyDefault, h->stream));
    }
    BeamGroupDev gd;
    gd.nele = g.nele; gd.geo = g.geo; gd.mats = g.mats; gd.mat_id = g.mat_id; gd.idxX = g.idxX; gd.idxU = g.idxU; gd.udof = g.udof;
    for (int i = 0; i < 12; ++i) gd.scaleX[i] = g.scaleX[i];
    for (int i = 0; i < 3; ++i) gd.scaleU[i] = g.scaleU[i];
    StateDev sd{h->X0, h->X1, h->X2, nullptr};
    double* dout = nullptr;
    CK(dalloc(h, &dout, g.nele * MB_NRES));
    launch_beam_results(OX + 1, gd, sd, dout, h->stream);
    h->launches++;
    cudaError_t e = cudaMemcpyAsync(out, dout, (size_t)(g.nele * MB_NRES) * sizeof(double), cudaMemcpyDefault, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    dfree(h, dout);
    CK(e);
    CK(cudaGetLastError());
    return MB_OK;
}

int32_t mb_get_device_ptrs(mb_handle* h, mb_dev_ptrs* out) {
    if (!h || !out) return MB_ERR_ARG;
    ARG(h->prepared, "not prepared");
    out->X0 = h->X0; out->X1 = h->X1; out->X2 = h->X2; out->U0 = h->U0; out->Llambda = h->Ll; out->nzval = h->nzval;
    out->colptr0 = h->colptr0; out->rowval0 = h->rowval0; out->ndofX = h->ndofX; out->ndofU = h->ndofU; out->nnz = h->nnz;
    return MB_OK;
}

// ------------------------------------------------------------------------------------------------------------ measurement
int32_t mb_sweepx_time_dev(mb_handle* h, int32_t OX, int32_t mission, double t, const double* newmark, int32_t reps, float* ms) {
    if (!h) return MB_ERR_ARG;
    ARG(h->prepared && reps >= 1 && ms && newmark, "bad argument");
    CK(cudaSetDevice(h->device));
    NewmarkDev nm{newmark[0], newmark[1], newmark[2], newmark[3], newmark[4], newmark[5], newmark[6]};
    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    float tel = 0, tga = 0;
    CK(cudaMemsetAsync(h->nanflag, 0xFF, sizeof(unsigned long long), h->stream));
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0, h->stream));
        launch_elements(h, OX, mission, nm, t);
        CK(cudaEventRecord(e1, h->stream));
        launch_gather(h, mission == 0 && OX > 0);
        CK(cudaEventRecord(e2, h->stream));
        CK(cudaEventSynchronize(e2));
        float a, b;
        CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b, e1, e2));
        tel += a; tga += b;
    }
    ms[0] = tel / reps; ms[1] = tga / reps;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    CK(cudaGetLastError());
    return MB_OK;
}

int32_t mb_measure_fp64_tflops(mb_handle* h, double* tflops) {
    if (!h || !tflops) return MB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, h->device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 1 << 16;
    double* out = nullptr;
    CK(cudaMalloc((void**)&out, (size_t)threads * blocks * sizeof(double)));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0, h->stream));
        fp64_peak_kernel<<<blocks, threads, 0, h->stream>>>(out, iters);
        CK(cudaEventRecord(e1, h->stream));
        CK(cudaEventSynchronize(e1));
        float msv; CK(cudaEventElapsedTime(&msv, e0, e1));
        if (r > 0 && msv < best) best = msv;
        h->launches++;
    }
    *tflops = 2.0 * 8.0 * iters * (double)threads * blocks / (best * 1e-3) / 1e12;
    cudaFree(out); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return MB_OK;
}

// the e2e path's ceiling: nbytes host→device and nbytes device→host at once (two copy engines, the streams of the host-state pipeline), pinned host buffers; wall-clock ms
// of `reps` rounds after one warm-up.  With several ranks calling it together it measures what the HOST gives N GPUs in aggregate.
int32_t mb_measure_host_copy_ms(mb_handle* h, const void* host_in, void* host_out, int64_t nbytes, int32_t reps, double* ms) {
    if (!h || !host_in || !host_out || nbytes <= 0 || reps < 1 || !ms) return MB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    void *din = nullptr, *dout = nullptr;
    CK(cudaMalloc(&din, (size_t)nbytes)); CK(cudaMalloc(&dout, (size_t)nbytes));
    cudaStream_t s1, s2; CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t a1, b1, a2, b2; CK(cudaEventCreate(&a1)); CK(cudaEventCreate(&b1)); CK(cudaEventCreate(&a2)); CK(cudaEventCreate(&b2));
    double tot = 0.;
    for (int r = 0; r <= reps; ++r) {
        CK(cudaEventRecord(a1, s1)); CK(cudaEventRecord(a2, s2));
        CK(cudaMemcpyAsync(din, host_in, (size_t)nbytes, cudaMemcpyHostToDevice, s1));
        CK(cudaMemcpyAsync(host_out, dout, (size_t)nbytes, cudaMemcpyDeviceToHost, s2));
        CK(cudaEventRecord(b1, s1)); CK(cudaEventRecord(b2, s2));
        CK(cudaEventSynchronize(b1)); CK(cudaEventSynchronize(b2));
        float m1, m2; CK(cudaEventElapsedTime(&m1, a1, b1)); CK(cudaEventElapsedTime(&m2, a2, b2));
        if (r > 0) tot += (double)std::max(m1, m2);
    }
    *ms = tot / reps;
    cudaFree(din); cudaFree(dout); cudaStreamDestroy(s1); cudaStreamDestroy(s2);
    cudaEventDestroy(a1); cudaEventDestroy(b1); cudaEventDestroy(a2); cudaEventDestroy(b2);
    return MB_OK;
}

int32_t mb_measure_copy_gbs(mb_handle* h, double* gbs) {
    if (!h || !gbs) return MB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    const int64_t n = (int64_t)1 << 27;   // 2 GiB of double2 in, 2 GiB out
    double2 *a = nullptr, *b = nullptr;
    CK(cudaMalloc((void**)&a, n * sizeof(double2))); CK(cudaMalloc((void**)&b, n * sizeof(double2)));
    CK(cudaMemsetAsync(a, 0, n * sizeof(double2), h->stream));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, h->device));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 6; ++r) {
        CK(cudaEventRecord(e0, h->stream));
        copy_kernel<<<prop.multiProcessorCount * 16, 512, 0, h->stream>>>(a, b, n);
        CK(cudaEventRecord(e1, h->stream));
        CK(cudaEventSynchronize(e1));
        float msv; CK(cudaEventElapsedTime(&msv, e0, e1));
        if (r > 0 && msv < best) best = msv;
        h->launches++;
    }
    *gbs = 2.0 * n * sizeof(double2) / (best * 1e-3) / 1e9;
    cudaFree(a); cudaFree(b); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return MB_OK;
}

int32_t mb_set_stream(mb_handle* h, void* cuda_stream) {
    if (!h) return MB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    if (h->own_stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)cuda_stream; h->own_stream = false;
    return MB_OK;
}
static __global__ void iface_pack_kernel(int64_t nnz_part, int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ nzval,
                                         const double* __restrict__ Ll, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (i < nnz_part) ? nzval[idx[i]] : Ll[idx[i]];
}
static __global__ void iface_unpack_kernel(int64_t nnz_part, int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ in,
                                           double* __restrict__ nzval, double* __restrict__ Ll) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // positions are distinct: plain read-modify-write, no atomics
    if (i < n && idx[i] >= 0) { if (i < nnz_part) nzval[idx[i]] += in[i]; else Ll[idx[i]] += in[i]; }
}
int32_t mb_iface_setup(mb_handle* h, int64_t n_send_nz, const int64_t* send_nz, int64_t n_send_v, const int64_t* send_v,
                       int64_t n_recv_nz, const int64_t* recv_nz, int64_t n_recv_v, const int64_t* recv_v) {
    if (!h) return MB_ERR_ARG;
    ARG(h->prepared, "call mb_sweepx_prepare first");
    CK(cudaSetDevice(h->device));
    auto up = [&](int64_t n1, const int64_t* a, int64_t lim1, int64_t n2, const int64_t* b, int64_t lim2, int32_t** dst) -> int32_t {
        std::vector<int32_t> t((size_t)(n1 + n2));
        for (int64_t i = 0; i < n1; ++i) { if (a[i] < 0 || a[i] > lim1) { h->err = "interface index out of range"; return MB_ERR_ARG; } t[(size_t)i] = (int32_t)(a[i] - 1); }
        for (int64_t i = 0; i < n2; ++i) { if (b[i] < 0 || b[i] > lim2) { h->err = "interface index out of range"; return MB_ERR_ARG; } t[(size_t)(n1 + i)] = (int32_t)(b[i] - 1); }
        CK(dalloc(h, dst, n1 + n2));
        if (n1 + n2) CK(cudaMemcpy(*dst, t.data(), t.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        return MB_OK;
    };
    for (int64_t i = 0; i < n_send_nz; ++i) ARG(send_nz[i] >= 1, "send positions are 1-based (0 = ghost is for receive lists only)");
    for (int64_t i = 0; i < n_send_v; ++i) ARG(send_v[i] >= 1, "send positions are 1-based (0 = ghost is for receive lists only)");
    if (h->if_sendbuf) { dfree(h, h->if_sendbuf); h->if_sendbuf = nullptr; }
    if (h->if_recvbuf) { dfree(h, h->if_recvbuf); h->if_recvbuf = nullptr; }
    int32_t rc = up(n_send_nz, send_nz, h->nnz, n_send_v, send_v, h->ndofX, &h->if_send);
    if (rc) return rc;
    rc = up(n_recv_nz, recv_nz, h->nnz, n_recv_v, recv_v, h->ndofX, &h->if_recv);
    if (rc) return rc;
    h->if_nsend_nz = n_send_nz; h->if_nsend_v = n_send_v; h->if_nrecv_nz = n_recv_nz; h->if_nrecv_v = n_recv_v;
    return MB_OK;
}
int32_t mb_iface_pack_dev(mb_handle* h, double* sendbuf_dev) {
    if (!h) return MB_ERR_ARG;
    ARG(h->prepared, "call mb_sweepx_prepare first");
    const int64_t n = h->if_nsend_nz + h->if_nsend_v;
    if (n == 0) return MB_OK;
    ARG(sendbuf_dev, "null buffer");
    CK(cudaSetDevice(h->device));
    iface_pack_kernel<<<nblk(n, 128), 128, 0, h->stream>>>(h->if_nsend_nz, n, h->if_send, h->nzval, h->Ll, sendbuf_dev);
    CK(cudaGetLastError());
    h->launches++;
    return MB_OK;
}
int32_t mb_iface_unpack_add_dev(mb_handle* h, const double* recvbuf_dev) {
    if (!h) return MB_ERR_ARG;
    ARG(h->prepared, "call mb_sweepx_prepare first");
    const int64_t n = h->if_nrecv_nz + h->if_nrecv_v;
    if (n == 0) return MB_OK;
    ARG(recvbuf_dev, "null buffer");
    CK(cudaSetDevice(h->device));
    iface_unpack_kernel<<<nblk(n, 128), 128, 0, h->stream>>>(h->if_nrecv_nz, n, h->if_recv, recvbuf_dev, h->nzval, h->Ll);
    CK(cudaGetLastError());
    h->launches++;
    return MB_OK;
}

int32_t mb_host_register(mb_handle* h, void* p, int64_t bytes) {
    if (!h) return MB_ERR_ARG;
    ARG(p && bytes > 0, "bad argument");
    CK(cudaSetDevice(h->device));
    CK(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault));
    return MB_OK;
}
int32_t mb_host_unregister(mb_handle* h, void* p) {
    if (!h) return MB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaHostUnregister(p));
    return MB_OK;
}

int32_t mb_add_bar3d(mb_handle* h, int64_t nele, const double* eleobj, int32_t udof, const int64_t* idxX, const int64_t* idxU,
                     const double* scaleX, const double* scaleU, int32_t* ieletyp_out) {
    if (!h) return MB_ERR_ARG;
    ARG(!h->prepared, "groups must be added before prepare");
    ARG(nele >= 0 && eleobj && idxX && scaleX, "null argument");
    ARG(!udof || (idxU && scaleU), "Udof element type needs idxU and scaleU");
    CK(cudaSetDevice(h->device));
    Group g;
    g.kind = G_BAR; g.nele = nele; g.nx = 6; g.nu = udof ? 3 : 0; g.udof = udof;
    // reference struct (38 doubles, toolbox/BarElement.jl:89-101): cₘ3 tgₘ3 tgₑ3 L₀ Lₛ mat9 wgp4 ζgp4 ζnod2 ψ₁4 ψ₂4
    std::vector<double> geo((size_t)nele * 8);
    std::vector<BarMat> mats; std::vector<int32_t> mat_id((size_t)nele);
    std::map<std::string, int32_t> seen;
    for (int64_t e = 0; e < nele; ++e) {
        const double* o = eleobj + e * 38; double* q = geo.data() + e * 8;
        for (int k = 0; k < 6; ++k) q[k] = o[k];
        q[6] = o[9]; q[7] = o[10];
        std::string key((const char*)(o + 11), 9 * sizeof(double));
        auto it = seen.find(key);
        if (it == seen.end()) { BarMat m; std::memcpy(&m, o + 11, sizeof m); it = seen.emplace(key, (int32_t)mats.size()).first; mats.push_back(m); }
        mat_id[(size_t)e] = it->second;
    }
    CK(dalloc(h, &g.geo, nele * 8));
    CK(cudaMemcpy(g.geo, geo.data(), geo.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(dalloc(h, &g.barmats, (int64_t)mats.size()));
    CK(cudaMemcpy(g.barmats, mats.data(), mats.size() * sizeof(BarMat), cudaMemcpyHostToDevice));
    if (mats.size() > 1) { CK(dalloc(h, &g.mat_id, nele)); CK(cudaMemcpy(g.mat_id, mat_id.data(), mat_id.size() * sizeof(int32_t), cudaMemcpyHostToDevice)); }
    int32_t rc = upload_index(h, idxX, nele * 6, 0, &g.idxX, &g.maxX);
    if (rc) return rc;
    if (udof) { rc = upload_index(h, idxU, nele * 3, 0, &g.idxU, &g.maxU); if (rc) return rc; }
    for (int i = 0; i < 6; ++i) g.scaleX[i] = scaleX[i];
    if (udof) for (int i = 0; i < 3; ++i) g.scaleU[i] = scaleU[i];
    h->groups.push_back(g);
    if (ieletyp_out) *ieletyp_out = (int32_t)h->groups.size();
    return MB_OK;
}
int32_t mb_add_soilcontact(mb_handle* h, int64_t nele, const double* eleobj, const int64_t* idxX, const double* scaleX, int32_t* ieletyp_out) {
    if (!h) return MB_ERR_ARG;
    ARG(!h->prepared, "groups must be added before prepare");
    ARG(nele >= 0 && eleobj && idxX && scaleX, "null argument");
    CK(cudaSetDevice(h->device));
    Group g;
    g.kind = G_SOIL; g.nele = nele; g.nx = 3;
    CK(dalloc(h, &g.geo, nele * 5));
    CK(cudaMemcpy(g.geo, eleobj, (size_t)nele * 5 * sizeof(double), cudaMemcpyHostToDevice));
    int32_t rc = upload_index(h, idxX, nele * 3, 0, &g.idxX, &g.maxX);
    if (rc) return rc;
    for (int i = 0; i < 3; ++i) g.scaleX[i] = scaleX[i];
    h->groups.push_back(g);
    if (ieletyp_out) *ieletyp_out = (int32_t)h->groups.size();
    return MB_OK;
}

}  // extern "C"
