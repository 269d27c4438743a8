"""N>1 host logic on CPU: world_size-2 `gloo` runs of the sharding plumbing (no GPU).

 * SweepX element-range sharding: per-rank local assembly (here by the oracle), interface pack → neighbour send/recv → unpack-add must
   reproduce the single-process assembly of the whole chain on the owned rows, and the ghost block must be the missing coupling.
 * DirectXUA time sharding: the halo plan covers exactly the steps whose stencils reach the owned columns."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    import muscade_b200 as mb
    from oracle import elements as OE, pattern as OP
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    M = 7
    def local_assembly(eleobj, idx, ndof, X):
        dis = [dict(X=idx, U=np.zeros((len(idx), 0), np.int64), A=np.zeros((len(idx), 0), np.int64))]
        a1, a2, cp, rv = OP.prepare_sweepx(dis, ndof, 0, 0)
        L = np.zeros(ndof); nz = np.zeros(len(rv))
        OE.sweepx_assemble_beams(eleobj, idx, a1[0].T, a2[0].T, 0, "iter", [X], np.ones(12), OE.newmark_coefficients(0, 0.), L, nz)
        return a2[0], cp, rv, L, nz
    # global reference on every rank
    ge, gi, gnd = mb.synthetic.chain(world * M)
    GX = mb.synthetic.state(gnd)[0]
    ga2, gcp, grv, GL, Gnz = local_assembly(ge, gi, gnd, GX)
    import scipy.sparse as sp
    G = sp.csc_matrix((Gnz, grv - 1, gcp - 1), shape=(gnd, gnd)).toarray()
    # my shard
    eleobj, idx, ndof, dof0 = mb.sharding.chain_shard(M, rank, world)
    X = GX[dof0:dof0 + ndof]
    a2, cp, rv, L, nz = local_assembly(eleobj, idx, ndof, X)
    snz, sv, rnz, rvv = mb.sharding.interface_indices(a2[:, -1], a2[:, 0], ndof, rank, world)
    sendbuf = torch.from_numpy(mb.sharding.pack(nz, L, snz, sv)) if len(snz) else torch.zeros(0, dtype=torch.float64)
    recvbuf = torch.zeros(len(rnz) + len(rvv), dtype=torch.float64)
    mb.sharding.exchange_neighbours(dist, sendbuf, recvbuf, rank, world)
    ghost = mb.sharding.unpack_add(nz, L, rnz, rvv, recvbuf.numpy())
    Kloc = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(ndof, ndof)).toarray()
    # owned rows: every local node except the last one of a non-final rank (owned by the right neighbour)
    nown = ndof - (6 if rank < world - 1 else 0)
    ok = np.abs(L[:nown] - GL[dof0:dof0 + nown]).max() <= 1e-13 * np.abs(Gnz).max()
    ok &= np.abs(Kloc[:nown, :] - G[dof0:dof0 + nown, dof0:dof0 + ndof]).max() <= 1e-13 * np.abs(Gnz).max()
    if rank > 0:   # ghost block = rows of my first node × columns of the neighbour's interior node, column-major
        ok &= np.abs(ghost.reshape(6, 6).T - G[dof0:dof0 + 6, dof0 - 6:dof0]).max() <= 1e-13 * np.abs(Gnz).max()
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARD_OK" if flag.item() == 1 else "SHARD_FAIL")
    dist.destroy_process_group()
''')


def test_sweepx_element_sharding_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29731", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert "SHARD_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


def test_directxua_halo_plan(mb):
    nstep = 12
    for world in (2, 3, 4):
        S = nstep // world
        for r in range(world):
            lo, hi = r * S, (r + 1) * S
            plan = mb.sharding.directxua_halo_plan(nstep, lo, hi)
            need = [s for s in range(max(0, lo - 2), min(nstep, hi + 2)) if not (lo <= s < hi)]
            assert sorted(plan["recv_left"] + plan["recv_right"]) == need
            # what I send right is what my right neighbour expects from its left, and vice versa
            if r < world - 1:
                assert plan["send_right"] == mb.sharding.directxua_halo_plan(nstep, hi, hi + S)["recv_left"]
            if r > 0:
                assert plan["send_left"] == mb.sharding.directxua_halo_plan(nstep, lo - S, lo)["recv_right"]
    # the block pattern of a shard only references stored steps
    bc, br = mb.directxua.block_pattern(2, 0, nstep, 4, 8)
    assert br.min() // 3 >= 2 and br.max() // 3 <= 9 and len(bc) == 3 * 4 + 1


def test_directxua_window_plan(mb):
    """the sliding-window plan of the configs[3] bench: the ranks' windows tile the time axis exactly once; only the windows holding the first / last step
    fall outside the set that one moved handle can serve"""
    for nstep, window in [(2000, 10), (2000, 25), (96, 7), (48, 16)]:
        for world in (1, 2, 4, 8):
            if nstep % world or nstep // world < 6:
                with pytest.raises(ValueError):
                    mb.sharding.directxua_windows(nstep, 0, world, window)
                continue
            seen = []
            for rank in range(world):
                L, H, Wn, windows, interior = mb.sharding.directxua_windows(nstep, rank, world, window)
                assert (H - L) % Wn == 0 and Wn <= window and windows[0][0] == L and windows[-1][1] == H
                assert all(hi - lo == Wn for lo, hi in windows) and all(a[1] == b[0] for a, b in zip(windows, windows[1:]))
                for w in windows:
                    if w in interior:
                        assert w[0] - 2 >= 1 and w[1] + 2 <= nstep - 1        # its stored steps exclude step 0 and step nstep-1
                    else:
                        assert w[0] < 3 or w[1] > nstep - 3
                assert len(windows) - len(interior) <= (rank == 0) + (rank == world - 1)
                seen += [s for lo, hi in windows for s in range(lo, hi)]
            assert seen == list(range(nstep))
