"""Sharding of the hot path over the GPUs of one box (one process per GPU, torch.distributed for the plumbing).

 * SweepX, large meshes — contiguous element ranges per rank.  Every cut shares one node; the rank to the right owns it.  After the
   local assembly the left rank sends the contributions of its last element to the rows of that node (72 tangent entries + 6 gradient
   entries); the owner adds the 36+6 that fall on its own pattern and keeps the 36 couplings to the neighbour's interior node as a ghost
   block (what a row-distributed sparse solver stores for off-process columns).  One neighbour send/recv of 78 doubles per cut.
 * DirectXUA — contiguous time-step ranges per rank (src/DirectXUA.jl:328-353: steps are independent); the finite-difference stencils
   reach ±2 steps, so each rank receives L2[Λ,X] of two steps from each neighbour (directxua.exchange_halo).

The index logic here is backend-neutral (numpy in, numpy out) so that the world_size-2 gloo tests exercise it on CPU.
"""
import numpy as np

from . import synthetic


def chain_shard(M, rank, world, h=1.0, dynamic=False):
    """rank's part of the synthetic chain of world·M elements: elements [rank·M,(rank+1)·M), its M+1 nodes renumbered locally.
    Returns (eleobj (M,69), idx (M,12) local 1-based, ndof_local, first global dof (0-based) of the local numbering)."""
    eleobj, idx, ndof = synthetic.chain(M, h=h, dynamic=dynamic)
    off = np.array([0.8, 0.6, 0.0]) * h * rank * M
    eleobj = eleobj.copy()
    eleobj[:, 0:3] += off[None, :]          # cₘ is the only field that depends on the absolute position
    return eleobj, idx, ndof, 6 * rank * M


def interface_indices(asm2_last, asm2_first, ndof_local, rank, world):
    """index lists for mb_iface_setup. asm2_last/first: asm[2] column (144,) of the local last / first element (1-based nz numbers).
    Send (to rank+1): rows 7..12 of the last element, all 12 columns, column-major (i fastest), then the 6 gradient entries.
    Recv (from rank-1): same ordering; columns 1..6 are the neighbour's interior node → ghost (0); columns 7..12 are my node 1."""
    send_nz = send_v = recv_nz = recv_v = np.zeros(0, np.int64)
    if rank < world - 1:
        send_nz = np.array([asm2_last[i + 12 * j] for j in range(12) for i in range(6, 12)], np.int64)
        send_v = np.arange(ndof_local - 5, ndof_local + 1, dtype=np.int64)
    if rank > 0:
        recv_nz = np.array([0 if j < 6 else asm2_first[(i - 6) + 12 * (j - 6)] for j in range(12) for i in range(6, 12)], np.int64)
        recv_v = np.arange(1, 7, dtype=np.int64)
    return send_nz, send_v, recv_nz, recv_v


def pack(nzval, Ll, send_nz, send_v):
    return np.concatenate([nzval[send_nz - 1], Ll[send_v - 1]])


def unpack_add(nzval, Ll, recv_nz, recv_v, buf):
    """adds the received contributions in place; returns the ghost coupling block (entries whose index is 0)"""
    a = buf[: len(recv_nz)]; b = buf[len(recv_nz):]
    own = recv_nz > 0
    nzval[recv_nz[own] - 1] += a[own]
    Ll[recv_v - 1] += b
    return a[~own]


def exchange_neighbours(dist, sendbuf, recvbuf, rank, world):
    """left→right neighbour exchange (torch tensors on the backend's device); no-op at the ends of the chain"""
    ops = []
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, sendbuf, rank + 1))
    if rank > 0:
        ops.append(dist.P2POp(dist.irecv, recvbuf, rank - 1))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()


def directxua_halo_plan(nstep, lo, hi):
    """which evaluated steps a time-shard [lo,hi) must send to / receive from its neighbours: L2[Λ,X](s) feeds block columns X_t, |t−s| ≤ 2
    (src/FiniteDifferences.jl:2-4).  Returns dict(send_left, send_right, recv_left, recv_right) of 0-based step lists."""
    return dict(send_left=[s for s in (lo, lo + 1) if lo > 0 and s < hi],
                send_right=[s for s in (hi - 2, hi - 1) if hi < nstep and s >= lo],
                recv_left=[s for s in (lo - 2, lo - 1) if s >= 0],
                recv_right=[s for s in (hi, hi + 1) if s < nstep])


def directxua_windows(nstep, rank, world, window):
    """Time-shard of rank `rank` and its sliding windows (BASELINE configs[3] streamed through mb_direct_rebase): the rank owns the steps
    [L,H) = [rank·nstep/world,(rank+1)·nstep/world) and walks them in windows of Wn steps, Wn the largest divisor of H−L not above `window`.
    Returns (L, H, Wn, windows, interior): `windows` = [(lo,hi)…] in order; `interior` = those with 3 ≤ lo and hi ≤ nstep−3, whose Lvv structure repeats up to a
    row shift (central finite-difference stencils only, src/FiniteDifferences.jl:8-31) — one handle serves them all; the others contain the first or the last
    step and need a handle of their own."""
    if nstep % world != 0 or nstep // world < 6:
        raise ValueError("the number of time steps must be a multiple of the number of ranks, at least 6 steps each")
    L, H = rank * nstep // world, (rank + 1) * nstep // world
    Wn = max(w for w in range(1, min(window, H - L) + 1) if (H - L) % w == 0)
    windows = [(a, a + Wn) for a in range(L, H, Wn)]
    interior = [w for w in windows if w[0] >= 3 and w[1] <= nstep - 3]
    return L, H, Wn, windows, interior


class CudaView:
    """expose a raw device pointer to torch through __cuda_array_interface__ (zero-copy view for NCCL send/recv)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = dict(shape=(int(n),), typestr="<f8", data=(int(ptr), False), version=2)
