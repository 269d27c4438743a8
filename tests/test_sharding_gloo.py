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


XUA_WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
    import torch, torch.distributed as dist
    import muscade_b200 as mb
    from muscade_b200 import xua
    from oracle import pattern as OP
    import xua_models as XM
    from test_host_xua import _assemble_outs
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(5)                       # the same model and states on every rank
    OX, OU, IA, nsteps, dts = 2, 0, 1, [6, 7], [1., 0.5]
    m = XM.model_testdirectxua(); s0 = mb.initialize(m); dis = s0.dis
    st = s0.with_orders(1, OX + 1, OU + 1)
    A = rng.normal(0, 0.1, st.A.shape)
    states = [[mb.State(3. + dt * k, [rng.normal(0, .3, v.shape) for v in st.Λ], [rng.normal(0, .3, v.shape) for v in st.X], [rng.normal(0, .3, v.shape) for v in st.U], A, None, m, dis)
               for k in range(n)] for n, dt in zip(nsteps, dts)]
    P = OP.prepare_direct(XM.dis_lists(dis), 2, 4, 6, OX, OU, IA)
    big, basm, pgr, _ = OP.preparebig(IA, nsteps, P["nL2"], P["pat"])
    outA, outs = _assemble_outs(m, dis, P, OX, OU, IA, states)
    nz_full, Lv_full = OP.assemblebig_general(IA, nsteps, dts, P, big, basm, pgr, outA, outs)
    # my share: the outs of the other ranks' steps (and, off rank 0, of assembleA!) zeroed — what XUAEngine.assemblebig_sharded adds on this rank
    mine = set(xua.shard_steps(nsteps, rank, world))
    zero = OP.out_zeros(P)
    outs_r = [[o if (ie, k) in mine else zero for k, o in enumerate(row)] for ie, row in enumerate(outs)]
    nz, Lv = OP.assemblebig_general(IA, nsteps, dts, P, big, basm, pgr, outA if rank == 0 else zero, outs_r)
    tnz, tLv = torch.from_numpy(nz), torch.from_numpy(Lv)
    dist.all_reduce(tnz); dist.all_reduce(tLv)            # mb_xua_allreduce_big
    ok = np.abs(tnz.numpy() - nz_full).max() <= 1e-14 * np.abs(nz_full).max() and np.abs(tLv.numpy() - Lv_full).max() <= 1e-14 * np.abs(nz_full).max()
    cnt = torch.tensor([len(mine)]); dist.all_reduce(cnt)
    ok &= cnt.item() == sum(nsteps)
    flag = torch.tensor([1 if ok else 0]); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("XUA_SHARD_OK" if flag.item() == 1 else "XUA_SHARD_FAIL")
    dist.destroy_process_group()
''')


def test_general_form_time_shards_world2(tmp_path):
    """XUAEngine.assemblebig_sharded on two ranks, with the oracle in place of the device: each rank's share of the steps (the A step on rank 0) summed by all_reduce is the
    whole assemblebig! (IA = 1, two experiments) — the plan mb_xua_allreduce_big rests on"""
    script = tmp_path / "xua_worker.py"
    script.write_text(XUA_WORKER % (ROOT, ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29733", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert "XUA_SHARD_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
