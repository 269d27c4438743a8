// Multi-GPU plumbing INSIDE the shim: one handle per GPU / process, NCCL over NVLink for the two exchanges the path has
// (SURVEY.md §8e): the interface rows of element-range shards (SweepX) and the halo blocks of time shards (DirectXUA).
// A Julia host binds these with ccall like everything else — no PyTorch, no MPI: the 128-byte NCCL unique id is created by
// one process (mb_comm_unique_id) and handed to the others by whatever the host has (a file, a socket, MPI, a TCP store).
// libnccl.so.2 is opened at run time (dlopen) so that the single-GPU library has no link-time dependency on NCCL; if another
// copy of that soname is already mapped in the process (torch ships one) the loader hands back the same library.
#include "mb_internal.h"
#include <dlfcn.h>

namespace {
// the few NCCL declarations used (nccl.h: 2.27/2.28 ABI, stable since 2.7)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclSum = 0, ncclMax = 2, ncclMin = 3 };
enum { ncclFloat64 = 8 };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    std::string err;
};
NcclApi* nccl_api() {
    static NcclApi api;
    if (api.lib || !api.err.empty()) return &api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
    if (!api.lib) { api.err = std::string("cannot open libnccl.so.2: ") + dlerror(); return &api; }
#define MB_SYM(field, name) *(void**)(&api.field) = dlsym(api.lib, name); if (!api.field) { api.err = std::string("missing symbol ") + name; return &api; }
    MB_SYM(GetUniqueId, "ncclGetUniqueId") MB_SYM(CommInitRank, "ncclCommInitRank") MB_SYM(CommDestroy, "ncclCommDestroy")
    MB_SYM(Send, "ncclSend") MB_SYM(Recv, "ncclRecv") MB_SYM(AllReduce, "ncclAllReduce") MB_SYM(GroupStart, "ncclGroupStart")
    MB_SYM(GroupEnd, "ncclGroupEnd") MB_SYM(GetErrorString, "ncclGetErrorString") MB_SYM(GetVersion, "ncclGetVersion")
#undef MB_SYM
    return &api;
}
}  // namespace

#define NK(call)                                                                                           \
    do {                                                                                                   \
        int r_ = (call);                                                                                   \
        if (r_ != ncclSuccess) { h->err = std::string(#call) + ": " + api->GetErrorString(r_); return MB_ERR_NCCL; } \
    } while (0)

static int32_t need_comm(mb_handle* h, NcclApi*& api) {
    api = nccl_api();
    if (!api->lib) { h->err = api->err; return MB_ERR_NCCL; }
    if (!h->comm) { h->err = "call mb_comm_init first"; return MB_ERR_STATE; }
    return MB_OK;
}

// neighbour exchange used by both shardings: whatever is in `sends` / `recvs` goes out / comes in as ONE NCCL group on the handle's stream
int32_t mb_comm_sendrecv(mb_handle* h, const std::vector<MbXfer>& sends, const std::vector<MbXfer>& recvs) {
    NcclApi* api; int32_t rc = need_comm(h, api); if (rc) return rc;
    if (sends.empty() && recvs.empty()) return MB_OK;
    NK(api->GroupStart());
    for (const MbXfer& x : sends) if (x.n > 0) NK(api->Send(x.p, (size_t)x.n, ncclFloat64, x.peer, (ncclComm_t)h->comm, h->stream));
    for (const MbXfer& x : recvs) if (x.n > 0) NK(api->Recv(x.p, (size_t)x.n, ncclFloat64, x.peer, (ncclComm_t)h->comm, h->stream));
    NK(api->GroupEnd());
    return MB_OK;
}

extern "C" {

int32_t mb_comm_unique_id(uint8_t* id128) {
    NcclApi* api = nccl_api();
    if (!api->lib || !id128) return MB_ERR_NCCL;
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return MB_ERR_NCCL;
    std::memcpy(id128, id.internal, 128);
    return MB_OK;
}
int32_t mb_comm_init(mb_handle* h, const uint8_t* id128, int32_t rank, int32_t world) {
    if (!h) return MB_ERR_ARG;
    ARG(id128 && world >= 1 && rank >= 0 && rank < world, "bad rank / world / id");
    ARG(!h->comm, "communicator already initialised");
    NcclApi* api = nccl_api();
    if (!api->lib) { h->err = api->err; return MB_ERR_NCCL; }
    CK(cudaSetDevice(h->device));
    ncclUniqueId id; std::memcpy(id.internal, id128, 128);
    ncclComm_t c = nullptr;
    NK(api->CommInitRank(&c, world, id, rank));
    h->comm = c; h->rank = rank; h->world = world;
    CK(dalloc(h, &h->comm_scratch, 256));
    return MB_OK;
}
// A second handle on the same GPU (say the DirectXUA handle next to the SweepX one) borrows the owner's communicator instead of creating its own;
// the owner must outlive it.  Calls on the two handles must be issued in the same order on every rank (NCCL's rule for one communicator).
int32_t mb_comm_share(mb_handle* h, mb_handle* owner) {
    if (!h || !owner) return MB_ERR_ARG;
    ARG(!h->comm, "communicator already initialised");
    ARG(owner->comm && owner->device == h->device, "owner has no communicator on this device");
    CK(cudaSetDevice(h->device));
    h->comm = owner->comm; h->rank = owner->rank; h->world = owner->world; h->comm_owned = false;
    CK(dalloc(h, &h->comm_scratch, 256));
    return MB_OK;
}
int32_t mb_comm_destroy(mb_handle* h) {
    if (!h) return MB_ERR_ARG;
    if (!h->comm) return MB_OK;
    NcclApi* api = nccl_api();
    cudaSetDevice(h->device); cudaStreamSynchronize(h->stream);
    if (api->lib && h->comm_owned) api->CommDestroy((ncclComm_t)h->comm);
    h->comm = nullptr; h->world = 1; h->rank = 0; h->comm_owned = true;
    return MB_OK;
}
int32_t mb_comm_info(const mb_handle* h, int32_t* rank, int32_t* world, int32_t* nccl_version) {
    if (!h) return MB_ERR_ARG;
    if (rank) *rank = h->rank;
    if (world) *world = h->comm ? h->world : 1;
    if (nccl_version) { NcclApi* api = nccl_api(); int v = 0; if (api->lib) api->GetVersion(&v); *nccl_version = v; }
    return MB_OK;
}
// Small host vectors (timings, norms, L1[A] / L2[A,A] partial sums of time shards): all-reduce in place, blocking.  op: 0 sum, 1 max, 2 min.
int32_t mb_comm_allreduce(mb_handle* h, double* buf, int64_t n, int32_t op) {
    if (!h) return MB_ERR_ARG;
    NcclApi* api; int32_t rc = need_comm(h, api); if (rc) return rc;
    ARG(buf && n >= 0 && op >= 0 && op <= 2, "bad argument");
    if (n == 0) return MB_OK;
    CK(cudaSetDevice(h->device));
    double* d = h->comm_scratch; bool own = false;
    if (n > 256) { CK(cudaMalloc((void**)&d, (size_t)n * sizeof(double))); own = true; }
    CK(cudaMemcpyAsync(d, buf, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    const int ops[3] = {ncclSum, ncclMax, ncclMin};
    NK(api->AllReduce(d, d, (size_t)n, ncclFloat64, ops[op], (ncclComm_t)h->comm, h->stream));
    CK(cudaMemcpyAsync(buf, d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (own) cudaFree(d);
    return MB_OK;
}
// the same on device memory, asynchronous on the handle's stream (L1[A], L2[A,A] of the time shards stay in HBM)
int32_t mb_comm_allreduce_dev(mb_handle* h, double* dev, int64_t n, int32_t op) {
    if (!h) return MB_ERR_ARG;
    NcclApi* api; int32_t rc = need_comm(h, api); if (rc) return rc;
    ARG(dev && n >= 0 && op >= 0 && op <= 2, "bad argument");
    if (n == 0) return MB_OK;
    CK(cudaSetDevice(h->device));
    const int ops[3] = {ncclSum, ncclMax, ncclMin};
    NK(api->AllReduce(dev, dev, (size_t)n, ncclFloat64, ops[op], (ncclComm_t)h->comm, h->stream));
    return MB_OK;
}
int32_t mb_comm_barrier(mb_handle* h) {
    double x = 0.;
    return mb_comm_allreduce(h, &x, 1, 0);
}

// SweepX element-range shards (mb_iface_setup): pack the interface entries of nzval / Lλ, send them to rank+1, receive those of rank−1 and
// add them in — one call per assembly, asynchronous on the handle's stream.  The couplings to the left neighbour's interior node (positions 0 in
// recv_nz) stay in the receive buffer: mb_iface_ghost.
int32_t mb_iface_exchange(mb_handle* h) {
    if (!h) return MB_ERR_ARG;
    ARG(h->prepared, "call mb_sweepx_prepare first");
    CK(cudaSetDevice(h->device));
    const int64_t ns = h->if_nsend_nz + h->if_nsend_v, nr = h->if_nrecv_nz + h->if_nrecv_v;
    if (ns && !h->if_sendbuf) CK(dalloc(h, &h->if_sendbuf, ns));
    if (nr && !h->if_recvbuf) CK(dalloc(h, &h->if_recvbuf, nr));
    int32_t rc;
    if (ns) { rc = mb_iface_pack_dev(h, h->if_sendbuf); if (rc) return rc; }
    std::vector<MbXfer> s, r;
    if (ns) { ARG(h->rank + 1 < h->world, "send list on the last rank"); s.push_back({h->if_sendbuf, ns, h->rank + 1}); }
    if (nr) { ARG(h->rank > 0, "receive list on rank 0"); r.push_back({h->if_recvbuf, nr, h->rank - 1}); }
    rc = mb_comm_sendrecv(h, s, r); if (rc) return rc;
    if (nr) { rc = mb_iface_unpack_add_dev(h, h->if_recvbuf); if (rc) return rc; }
    return MB_OK;
}
int32_t mb_iface_get_recvbuf(mb_handle* h, double* out) {
    if (!h) return MB_ERR_ARG;
    const int64_t nr = h->if_nrecv_nz + h->if_nrecv_v;
    if (nr == 0) return MB_OK;
    ARG(out && h->if_recvbuf, "no receive buffer yet");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(out, h->if_recvbuf, (size_t)nr * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MB_OK;
}
int32_t mb_iface_buffers(mb_handle* h, double** sendbuf_dev, int64_t* nsend, double** recvbuf_dev, int64_t* nrecv) {
    if (!h) return MB_ERR_ARG;
    CK(cudaSetDevice(h->device));
    const int64_t ns = h->if_nsend_nz + h->if_nsend_v, nr = h->if_nrecv_nz + h->if_nrecv_v;
    if (ns && !h->if_sendbuf) CK(dalloc(h, &h->if_sendbuf, ns));
    if (nr && !h->if_recvbuf) CK(dalloc(h, &h->if_recvbuf, nr));
    if (sendbuf_dev) *sendbuf_dev = h->if_sendbuf; if (nsend) *nsend = ns;
    if (recvbuf_dev) *recvbuf_dev = h->if_recvbuf; if (nrecv) *nrecv = nr;
    return MB_OK;
}

}  // extern "C"
