"""GPU parity of the multi-GPU DATA paths, i.e. the device code that runs when bench.py is launched with N > 1:

 * SweepX element-range shards: mb_iface_setup → mb_iface_pack_dev → (transfer) → mb_iface_unpack_add_dev must turn the shards' local nzval / Lλ into
   the unsharded assembly on the owned rows, bit for bit (two contributors per interface entry: the sum commutes), and leave the ghost couplings
   in the receive buffer.
 * DirectXUA time shards: the per-step blocks mb_direct_step_ptrs exposes (what mb_direct_halo_exchange sends) are all a shard needs from its
   neighbours to build its block columns of Lvv / rows of Lv, bit-identical to the one-handle result.

On one GPU the transfer between two handles is a device copy; with two or more GPUs (gpurun --gpus 2) the same checks run with one process per GPU
and the transfer done by NCCL inside the shim (mb_comm_init, mb_iface_exchange, mb_direct_halo_exchange)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dev_view(mb, ptr, n):
    import torch
    return torch.as_tensor(mb.sharding.CudaView(ptr, n), device="cuda")


def fetch(mb, eng):
    """Lλ and nzval of an engine, read from its device buffers"""
    p = eng.device_ptrs()
    eng.sync()
    return dev_view(mb, p.Llambda, p.ndofX).cpu().numpy().copy(), dev_view(mb, p.nzval, p.nnz).cpu().numpy().copy()


@pytest.mark.parametrize("OX", [0, 2])
def test_sweepx_interface_exchange_kernels(mb, OX):
    import torch
    world, M = 3, 37
    nm = mb.synthetic.newmark_coefficients(OX, 0.3)
    # unsharded reference on the device
    ge, gi, gnd = mb.synthetic.chain(world * M, dynamic=OX > 0)
    GX = mb.synthetic.state(gnd, nder=OX + 1)
    g = mb.Engine(0)
    g.add_eulerbeam3d(ge, gi, np.ones(12)); g.sweepx_prepare(gnd)
    g.set_state(GX); g.sweepx_assemble_dev(OX, "iter", nm)
    GL, Gnz = fetch(mb, g)
    gcp, grv = g.sweepx_pattern()
    import scipy.sparse as sp
    G = sp.csc_matrix((Gnz, grv - 1, gcp - 1), shape=(gnd, gnd)).toarray()
    shards = []
    for r in range(world):
        _, idx, ndof, dof0 = mb.sharding.chain_shard(M, r, world, dynamic=OX > 0)
        eleobj = ge[r * M:(r + 1) * M]                        # the very element objects of the unsharded model (bit-exact comparison)
        e = mb.Engine(0)
        ityp = e.add_eulerbeam3d(eleobj, idx, np.ones(12)); e.sweepx_prepare(ndof)
        e.set_state([x[dof0:dof0 + ndof] for x in GX])
        a2f = e.sweepx_asm_range(ityp, 0, 1)[1][0]; a2l = e.sweepx_asm_range(ityp, M - 1, M)[1][0]
        snz, sv, rnz, rvv = mb.sharding.interface_indices(a2l, a2f, ndof, r, world)
        e.iface_setup(snz, sv, rnz, rvv)
        e.sweepx_assemble_dev(OX, "iter", nm)
        shards.append((e, ndof, dof0, rnz))
    # the exchange: every shard packs, the buffer travels to the right neighbour, which adds it in
    for r in range(world - 1):
        src, dst = shards[r][0], shards[r + 1][0]
        ps, ns, _, _ = src.iface_buffers(); _, _, pr, nr = dst.iface_buffers()
        assert ns == nr == 78
        src.iface_pack(ps); src.sync()
        dev_view(mb, pr, nr).copy_(dev_view(mb, ps, ns)); torch.cuda.synchronize()
        dst.iface_unpack_add(pr); dst.sync()
    for r, (e, ndof, dof0, rnz) in enumerate(shards):
        L, nz = fetch(mb, e)
        cp, rv = e.sweepx_pattern()
        K = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(ndof, ndof)).toarray()
        nown = ndof - (6 if r < world - 1 else 0)        # the last node of a non-final shard is owned by the right neighbour
        assert np.array_equal(L[:nown], GL[dof0:dof0 + nown]), r
        assert np.array_equal(K[:nown, :], G[dof0:dof0 + nown, dof0:dof0 + ndof]), r
        if r > 0:
            ghost = e.iface_recvbuf()[: len(rnz)][rnz == 0]
            assert np.array_equal(ghost.reshape(6, 6).T, G[dof0:dof0 + 6, dof0 - 6:dof0]), r
        e.close()
    g.close()


def test_iface_setup_rejects_ghost_in_send_lists(mb):
    eleobj, idx, ndof = mb.synthetic.chain(4)
    e = mb.Engine(0)
    e.add_eulerbeam3d(eleobj, idx, np.ones(12)); e.sweepx_prepare(ndof)
    z = np.zeros(0, np.int64)
    with pytest.raises(mb.MuscadeB200Error):
        e.iface_setup(np.array([0, 3]), z, z, z)
    with pytest.raises(mb.MuscadeB200Error):
        e.iface_setup(z, np.array([ndof + 1]), z, z)
    e.iface_setup(z, z, np.array([0, 5]), np.array([1]))          # ghosts are fine on the receive side
    e.close()


def _direct_setup(mb, N=7, nstep=12):
    from test_gpu_directxua import udof_chain, states
    model = udof_chain(mb, N, np.random.default_rng(5))
    st0 = mb.initialize(model)
    nX, nU = model.getndof("X"), model.getndof("U")
    return model, st0.dis, nX, nU, states(mb, nX, nU, nstep)


def _gauged_setup(mb, n=6, nstep=12, dt=0.05):
    """chain of Udof beams wrapped in ElementCost{StrainGaugeOnEulerBeam3D} (5 gauges, quadratic strain cost), non-unit scales; random Λ, X, X′, X″, U per step"""
    rng = np.random.default_rng(8)
    P5 = np.array([[0., .5, 0.], [0., 0, .5], [0., -.5, 0.], [0., 0, -.5], [0., .5, 0.]]).T
    D5 = np.array([[1., 0., 0.], [1., 0., 0.], [1., 0., 0.], [1., 0., 0.], [1 / np.sqrt(2), 0, 1 / np.sqrt(2)]]).T
    m = mb.Model("gauged chain")
    nod = mb.addnode(m, np.cumsum(rng.uniform(0.5, 1.5, (n + 1, 3)), axis=0))
    un = np.array([mb.addnode(m, np.zeros(3)) for _ in range(n)])
    nodes = np.concatenate([np.stack([nod[:-1], nod[1:]], axis=1), un[:, None]], axis=1)
    tgt = rng.normal(0., 1e-3, 5)
    cost = mb.QuadraticGaugeCost(2e-3, lambda t: tgt * np.cos(t))
    mb.addelement(m, mb.ElementCost, nodes, req=("ε",), cost=cost, ElementType=mb.StrainGaugeOnEulerBeam3D,
                  elementkwargs=dict(P=P5, D=D5, elementkwargs=dict(mat=mb.BeamCrossSection(EA=1e3, EI2=30., EI3=20., GJ=40., mu=1.5, iota1=0.7), orient2=(0., 1., 0.), Udof=True)))
    mb.setscale(m, scale=dict(X=dict(t1=2., t2=2., t3=2.), U=dict(t1=3., t2=3., t3=3.)), Λscale=3.)
    s0 = mb.initialize(m)
    nX, nU = m.getndof("X"), m.getndof("U")
    st = [([rng.normal(0, 0.05, nX) for _ in range(3)], rng.normal(0, 0.5, nU), rng.normal(0, 1., nX)) for _ in range(nstep)]
    return m, s0.dis, nX, nU, st, dt * np.arange(nstep)


def _gauged_engine(mb, m, dis, st, time, nstep, dt, lo, hi, device, own_only):
    e = mb.directxua.prepare(2, 0, m, dis, nstep, dt, lo, hi, device=device)
    a, b = (lo, hi) if own_only else e.stored_range()
    for s in range(a, b):
        e.set_state(s, st[s][0], st[s][1]); e.set_lambda(s, st[s][2])
        e.set_gauge_measurements(s, 1, e.gauge_costs[0][1].measured(time[s]))
    return e


def test_directxua_time_shard_halo_blocks(mb):
    import torch
    OX, OU, nstep, dt, world = 2, 0, 12, 0.05, 3
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    model, dis, nX, nU, st = _direct_setup(mb, nstep=nstep)
    W = 2 * nX + nU
    whole = mb.directxua.prepare(OX, OU, model, dis, nstep, dt)
    for s in range(nstep):
        whole.set_state(s, st[s][0], st[s][1])
    Lvv = np.zeros(whole.nnzbig); Lv = np.zeros(whole.ncol)
    whole.direct_assemble(Lvv=Lvv, Lv=Lv)
    cpw, rvw = whole.big_pattern()
    S = nstep // world
    shards = []
    for r in range(world):
        lo, hi = r * S, (r + 1) * S
        e = mb.directxua.prepare(OX, OU, model, dis, nstep, dt, lo, hi)
        for s in range(lo, hi):                               # ONLY its own steps: the halo states are never given to this handle
            e.set_state(s, st[s][0], st[s][1])
        e.direct_assemble(eval_range=(lo, hi), build_big=False)
        shards.append((e, lo, hi))
    # halo: L2[Λ,X] of the two steps at each end of every shard travels to the neighbour that stores them (what mb_direct_halo_exchange sends)
    for r, (e, lo, hi) in enumerate(shards):
        e.sync()
        for nb, steps in ((r - 1, (lo, lo + 1)), (r + 1, (hi - 2, hi - 1))):
            if 0 <= nb < world:
                for s in steps:
                    (ps, n), _, _ = e.step_ptrs(s)
                    (pd, n2), _, _ = shards[nb][0].step_ptrs(s)
                    assert n == n2
                    dev_view(mb, pd, n).copy_(dev_view(mb, ps, n))
    torch.cuda.synchronize()
    for e, lo, hi in shards:
        a = np.zeros(e.nnzbig); b = np.zeros(e.ncol)
        e.direct_assemble(eval_range=(lo, lo), build_big=True, Lvv=a, Lv=b)
        c0, c1 = lo * W, hi * W
        p0, p1 = cpw[c0] - 1, cpw[c1] - 1
        cp, rv = e.big_pattern()
        assert np.array_equal(cp - 1, cpw[c0:c1 + 1] - 1 - p0) and np.array_equal(rv, rvw[p0:p1])
        assert np.array_equal(a, Lvv[p0:p1]) and np.array_equal(b, Lv[c0:c1])
        e.close()
    whole.close()


WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
    import muscade_b200 as mb
    from test_gpu_multigpu import fetch, _direct_setup
    rank, world, idfile = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    import torch
    torch.cuda.set_device(rank)
    # ---- SweepX element-range shards over NCCL inside the shim
    OX, M = 2, 41
    nm = mb.synthetic.newmark_coefficients(OX, 0.3)
    ge, gi, gnd = mb.synthetic.chain(world * M, dynamic=True)
    GX = mb.synthetic.state(gnd, nder=OX + 1)
    g = mb.Engine(rank); g.add_eulerbeam3d(ge, gi, np.ones(12)); g.sweepx_prepare(gnd); g.set_state(GX); g.sweepx_assemble_dev(OX, "iter", nm)
    GL, Gnz = fetch(mb, g); gcp, grv = g.sweepx_pattern(); g.close()
    import scipy.sparse as sp
    G = sp.csc_matrix((Gnz, grv - 1, gcp - 1), shape=(gnd, gnd)).toarray()
    _, idx, ndof, dof0 = mb.sharding.chain_shard(M, rank, world, dynamic=True)
    eleobj = ge[rank * M:(rank + 1) * M]
    e = mb.Engine(rank)
    if rank == 0:
        uid = mb.Engine.comm_unique_id(); uid.tofile(idfile + ".tmp"); os.replace(idfile + ".tmp", idfile)
    else:
        import time
        while not os.path.exists(idfile): time.sleep(0.05)
        uid = np.fromfile(idfile, np.uint8)
    e.comm_init(uid, rank, world)
    assert e.comm_info()[:2] == (rank, world)
    ityp = e.add_eulerbeam3d(eleobj, idx, np.ones(12)); e.sweepx_prepare(ndof)
    e.set_state([x[dof0:dof0 + ndof] for x in GX])
    a2f = e.sweepx_asm_range(ityp, 0, 1)[1][0]; a2l = e.sweepx_asm_range(ityp, M - 1, M)[1][0]
    snz, sv, rnz, rvv = mb.sharding.interface_indices(a2l, a2f, ndof, rank, world)
    e.iface_setup(snz, sv, rnz, rvv)
    for _ in range(2):                                       # twice: the exchange must not accumulate across assemblies
        e.sweepx_assemble_dev(OX, "iter", nm); e.iface_exchange()
    L, nz = fetch(mb, e); cp, rv = e.sweepx_pattern()
    K = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(ndof, ndof)).toarray()
    nown = ndof - (6 if rank < world - 1 else 0)
    ok = np.array_equal(L[:nown], GL[dof0:dof0 + nown]) and np.array_equal(K[:nown, :], G[dof0:dof0 + nown, dof0:dof0 + ndof])
    if rank > 0:
        ghost = e.iface_recvbuf()[: len(rnz)][rnz == 0]
        ok = ok and np.array_equal(ghost.reshape(6, 6).T, G[dof0:dof0 + 6, dof0 - 6:dof0])
    assert e.comm_allreduce([1.0 if ok else 0.0], "min")[0] == 1.0, "sweepx shard differs on rank %%d" %% rank
    assert e.comm_allreduce([float(rank)], "sum")[0] == world * (world - 1) / 2
    # ---- DirectXUA time shards: halo blocks over NCCL
    OX, OU, nstep, dt = 2, 0, 6 * world, 0.05
    model, dis, nX, nU, st = _direct_setup(mb, nstep=nstep)
    W = 2 * nX + nU
    whole = mb.directxua.prepare(OX, OU, model, dis, nstep, dt, device=rank)
    for s in range(nstep): whole.set_state(s, st[s][0], st[s][1])
    Lvv = np.zeros(whole.nnzbig); Lv = np.zeros(whole.ncol)
    whole.direct_assemble(Lvv=Lvv, Lv=Lv); cpw, rvw = whole.big_pattern(); whole.close()
    lo, hi = rank * 6, (rank + 1) * 6
    d = mb.directxua.prepare(OX, OU, model, dis, nstep, dt, lo, hi, device=rank)
    d.comm_init_from(e)
    for s in range(lo, hi): d.set_state(s, st[s][0], st[s][1])
    d.direct_assemble(eval_range=(lo, hi), build_big=False)
    d.halo_exchange()
    a = np.zeros(d.nnzbig); b = np.zeros(d.ncol)
    d.direct_assemble(eval_range=(lo, lo), build_big=True, Lvv=a, Lv=b)
    c0, c1 = lo * W, hi * W; p0, p1 = cpw[c0] - 1, cpw[c1] - 1
    ok = np.array_equal(a, Lvv[p0:p1]) and np.array_equal(b, Lv[c0:c1])
    assert e.comm_allreduce([1.0 if ok else 0.0], "min")[0] == 1.0, "directxua shard differs on rank %%d" %% rank
    d.close()
    # ---- the same with a strain-gauge cost on every beam (ElementCost accelerator in the windowed path): L1[X] of the costed beams travels with the halo, the costs' X-X blocks
    # and L1[U] are read at the owned steps only
    from test_gpu_multigpu import _gauged_setup, _gauged_engine
    m, gdis, nX, nU, gst, gtime = _gauged_setup(mb, nstep=nstep, dt=dt)
    W = 2 * nX + nU
    whole = _gauged_engine(mb, m, gdis, gst, gtime, nstep, dt, 0, nstep, rank, False)
    Lvv = np.zeros(whole.nnzbig); Lv = np.zeros(whole.ncol)
    whole.direct_assemble(Lvv=Lvv, Lv=Lv); cpw, rvw = whole.big_pattern(); whole.close()
    d = _gauged_engine(mb, m, gdis, gst, gtime, nstep, dt, lo, hi, rank, True)
    d.comm_init_from(e)
    d.direct_assemble(eval_range=(lo, hi), build_big=False)
    d.halo_exchange()
    a = np.zeros(d.nnzbig); b = np.zeros(d.ncol)
    d.direct_assemble(eval_range=(lo, lo), build_big=True, Lvv=a, Lv=b)
    c0, c1 = lo * W, hi * W; p0, p1 = cpw[c0] - 1, cpw[c1] - 1
    ok = np.array_equal(a, Lvv[p0:p1]) and np.array_equal(b, Lv[c0:c1]) and np.abs(b).max() > 0
    assert e.comm_allreduce([1.0 if ok else 0.0], "min")[0] == 1.0, "gauged directxua shard differs on rank %%d" %% rank
    d.close(); e.close()
    print("RANK_OK %%d" %% rank, flush=True)
''')


def test_nccl_in_shim_two_or_more_gpus(tmp_path):
    """one process per GPU, NCCL inside the shim (no torch.distributed): sharded SweepX assembly + mb_iface_exchange and DirectXUA time shards +
    mb_direct_halo_exchange reproduce the unsharded results bit for bit"""
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs (gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT))
    idfile = str(tmp_path / "nccl_id.bin")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), idfile], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            p.kill(); o, _ = p.communicate()
        outs.append(o)
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "RANK_OK %d" % r in o, "rank %d:\n%s" % (r, o[-3000:])
