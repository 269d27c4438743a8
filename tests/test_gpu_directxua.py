"""GPU parity, DirectXUA{2,0,0} path (src/DirectXUA.jl:22-120, 245-356): class-pair patterns and maps, per-step out.L1/L2 blocks, the
all-steps Lvv structure and values and Lv — whole problem on one handle and a time-shard of it — against the oracle."""
import numpy as np
import pytest

from oracle import elements as OE
from oracle import pattern as OP

pytestmark = pytest.mark.gpu
TOL = 1e-12


def rel(a, b, floor=0.):
    return np.abs(a - b).max() / max(np.abs(b).max(), floor, 1e-300)


def udof_chain(mb, N, rng=None):
    model = mb.Model()
    coord = np.arange(N + 1)[:, None] * np.array([.8, .6, 0.])[None, :]
    if rng is not None:
        coord = coord + rng.uniform(-.1, .1, coord.shape)
    nod = mb.addnode(model, coord)
    unod = mb.addnode(model, np.zeros((N, 0)))
    mat = mb.BeamCrossSection(EA=10., EI2=3., EI3=2.5, GJ=4., mu=1., iota1=1.2, w=.3, Ca2=2., Ca3=1.5, Cq2=1., Cq3=.7, Cl1=.2)
    mb.addelement(model, mb.EulerBeam3D, np.stack([nod[:-1], nod[1:], unod], axis=1), mat=mat, Udof=True)
    return model


def states(mb, nX, nU, nstep):
    return [([mb.synthetic.uniform_pm1(10 + 3 * s + d, nX) * (0.1 if d == 0 else 0.3) for d in range(3)], mb.synthetic.uniform_pm1(99 + s, nU)) for s in range(nstep)]


def oracle_all(mb, model, dis, OX, OU, nstep, dt, st):
    nX, nU, nA = model.getndof(("X", "U", "A"))
    odis = [dict(X=d.X, U=d.U, A=d.A) for d in dis.dis]
    P = OP.prepare_direct(odis, nX, nU, nA, OX, OU, 0)
    big, bigasm, pgr, pgc = OP.preparebig(0, [nstep], P["nL2"], P["pat"])
    outs = [OE.direct_assemble_step_beams(model.ele[0].eleobj, dis.dis[0].X, dis.dis[0].U, OX, OU, X[: OX + 1], [U], dis.dis[0].scaleX, dis.dis[0].scaleU, P, 0)
            for (X, U) in st]
    nz, Lv = OP.assemblebig(0, nstep, dt, P, big, bigasm, pgr, outs)
    return P, big, outs, nz, Lv


@pytest.mark.parametrize("OX", [2, 1, 0])
def test_directxua_whole_problem(mb, OX):
    OU, nstep, dt, N = 0, 7, 0.1, 9
    model = udof_chain(mb, N, np.random.default_rng(1))
    mb.setscale(model, scale=dict(X=dict(t1=2., t2=2., t3=2.), U=dict(t1=5., t2=5., t3=5.)))
    st0 = mb.initialize(model); dis = st0.dis
    nX, nU = model.getndof("X"), model.getndof("U")
    st = states(mb, nX, nU, nstep)
    P, big, outs, nz, Lv = oracle_all(mb, model, dis, OX, OU, nstep, dt, st)
    eng = mb.directxua.prepare(OX, OU, model, dis, nstep, dt)
    # class-pair patterns and element→nz maps, bit-exact (asm[arrnum(α,β)] of prepare(AssemblyDirect))
    for which, ab in [(0, (1, 2)), (0, (2, 1)), (0, (2, 2)), (1, (1, 3)), (1, (2, 3)), (2, (3, 1)), (2, (3, 2)), (3, (3, 3))]:
        cp, rv = eng.class_pattern(which)
        assert np.array_equal(cp, P["pat"][ab][2]) and np.array_equal(rv, P["pat"][ab][3]), ab
        assert np.array_equal(eng.direct_asm(1, which).T, P["asm"][OP.arrnum(*ab)][0]), ab
    # all-steps structure
    assert eng.ncol == big["n"] and eng.nnzbig == len(big["rowval"])
    cp, rv = eng.big_pattern()
    assert np.array_equal(cp, big["colptr"]) and np.array_equal(rv, big["rowval"])
    for s, (X, U) in enumerate(st):
        eng.set_state(s, X[: OX + 1], U)
    Lvv = np.zeros(eng.nnzbig); Lvec = np.zeros(eng.ncol)
    eng.direct_assemble(Lvv=Lvv, Lv=Lvec)
    for s in (0, 3, nstep - 1):
        o = outs[s]
        scale = max(np.abs(o["L2"][(1, 2)]).max(), 1e-300)
        assert rel(eng.step_block(s, 0), o["L1"][1], scale) <= TOL
        for der in range(OX + 1):
            assert rel(eng.step_block(s, 1, der), o["L2"][(1, 2)][der], scale) <= TOL
            assert rel(eng.step_block(s, 2, der), o["L2"][(2, 1)][der], scale) <= TOL
        assert rel(eng.step_block(s, 3), o["L2"][(1, 3)][0], scale) <= TOL and rel(eng.step_block(s, 4), o["L2"][(3, 1)][0], scale) <= TOL
    assert rel(Lvv, nz) <= TOL, rel(Lvv, nz)
    assert rel(Lvec, Lv, np.abs(nz).max()) <= TOL
    Lvv2 = np.zeros(eng.nnzbig)
    eng.direct_assemble(Lvv=Lvv2)
    assert np.array_equal(Lvv, Lvv2)          # deterministic
    eng.close()


@pytest.mark.parametrize("lo,hi", [(0, 3), (3, 6), (6, 9)])
def test_directxua_time_shard(mb, lo, hi):
    """a handle that owns steps [lo,hi) produces exactly its block columns of the reference's Lvv / rows of Lv"""
    OX, OU, nstep, dt, N = 2, 0, 9, 0.05, 6
    model = udof_chain(mb, N)
    st0 = mb.initialize(model); dis = st0.dis
    nX, nU = model.getndof("X"), model.getndof("U")
    st = states(mb, nX, nU, nstep)
    P, big, outs, nz, Lv = oracle_all(mb, model, dis, OX, OU, nstep, dt, st)
    W = 2 * nX + nU
    c0, c1 = lo * W, hi * W
    p0, p1 = big["colptr"][c0] - 1, big["colptr"][c1] - 1
    eng = mb.directxua.prepare(OX, OU, model, dis, nstep, dt, lo, hi)
    cp, rv = eng.big_pattern()
    assert np.array_equal(cp - 1, big["colptr"][c0:c1 + 1] - 1 - p0) and np.array_equal(rv, big["rowval"][p0:p1])
    for s in range(max(0, lo - 2), min(nstep, hi + 2)):
        eng.set_state(s, st[s][0], st[s][1])
    Lvv = np.zeros(eng.nnzbig); Lvec = np.zeros(eng.ncol)
    eng.direct_assemble(Lvv=Lvv, Lv=Lvec)
    assert rel(Lvv, nz[p0:p1], np.abs(nz).max()) <= TOL
    assert rel(Lvec, Lv[c0:c1], np.abs(nz).max()) <= TOL
    eng.close()


def test_directxua_sliding_window(mb):
    """mb_direct_rebase (BASELINE.json configs[3]: 2000 steps streamed through one window): a handle built for an interior window, moved forward by
    its own length twice and then backward by an overlapping amount, yields the reference's Lvv columns / Lv rows of each window — structure
    bit-exact (rows shifted), values ≤ 1e-12 — while evaluating only the steps that were not stored before the move."""
    OX, OU, nstep, dt, N, L = 2, 0, 26, 0.07, 5, 5
    model = udof_chain(mb, N, np.random.default_rng(3))
    mb.setscale(model, scale=dict(X=dict(t1=2., t2=2., t3=2.), U=dict(t1=5., t2=5., t3=5.)))
    st0 = mb.initialize(model); dis = st0.dis
    nX, nU = model.getndof("X"), model.getndof("U")
    st = states(mb, nX, nU, nstep)
    P, big, outs, nz, Lv = oracle_all(mb, model, dis, OX, OU, nstep, dt, st)
    W = 2 * nX + nU
    eng = mb.directxua.prepare(OX, OU, model, dis, nstep, dt, 4, 4 + L)
    stored = set()

    def check(lo):
        c0, c1 = lo * W, (lo + L) * W
        p0, p1 = big["colptr"][c0] - 1, big["colptr"][c1] - 1
        cp, rv = eng.big_pattern()
        assert np.array_equal(cp - 1, big["colptr"][c0:c1 + 1] - 1 - p0) and np.array_equal(rv, big["rowval"][p0:p1])
        new = [s for s in range(lo - 2, lo + L + 2) if s not in stored]
        for s in new:
            eng.set_state(s, st[s][0], st[s][1])
        runs = []                                              # contiguous runs of new steps
        for s in new:
            if runs and runs[-1][1] == s: runs[-1][1] = s + 1
            else: runs.append([s, s + 1])
        for a, b in runs:
            eng.direct_assemble(eval_range=(a, b), build_big=False)
        Lvv = np.zeros(eng.nnzbig); Lvec = np.zeros(eng.ncol)
        eng.direct_assemble(eval_range=(lo, lo), build_big=True, Lvv=Lvv, Lv=Lvec)
        assert rel(Lvv, nz[p0:p1], np.abs(nz).max()) <= TOL
        assert rel(Lvec, Lv[c0:c1], np.abs(nz).max()) <= TOL
        stored.clear(); stored.update(range(lo - 2, lo + L + 2))
        return len(new)

    assert check(4) == L + 4
    assert eng.rebase(4 + L) == L * W and check(4 + L) == L        # forward by the window length: only the L new steps are evaluated
    assert eng.rebase(4 + 2 * L) == 2 * L * W and check(4 + 2 * L) == L
    assert eng.rebase(11) == 7 * W and check(11) == 3                # backward, overlapping
    assert eng.rebase(nstep - 3 - L) == (nstep - 3 - L - 4) * W and check(nstep - 3 - L) == 7      # up to the last movable position; two stored steps kept
    with pytest.raises(Exception):
        eng.rebase(nstep - 2 - L)                                    # would reach the last step, whose stencil is one-sided
    eng.close()
    edge = mb.directxua.prepare(OX, OU, model, dis, nstep, dt, 0, L)
    with pytest.raises(Exception):
        edge.rebase(L)                                               # a window that contains step 0 cannot be moved
    edge.close()


def test_directxua_curved_constraints(mb):
    """Host-evaluated element types that are NOT linear in X on DirectXUA's second-order branch (L = Λ∘R, src/DirectXUA.jl:152-171, src/Assemble.jl:721-726):
    DofConstraint{:X} (src/BasicElements.jl:425-438) in `positive` mode and in `equal` mode with a quadratic gap (plus one switched `off`) on a Udof beam chain.
    Besides L1[Λ], L1[X], L2[Λ,X], L2[X,Λ] they feed L2[X,X] = Λ·∂²R/∂X² (mb_direct_set_host_xx).  Lvv and Lv against the oracle, constraints restated by hand."""
    OX, OU, nstep, dt, N, t0 = 2, 0, 7, 0.1, 4, 0.3
    model = udof_chain(mb, N, np.random.default_rng(7))
    nod = np.arange(1, N + 2)
    cons = [("positive", 2, "t2", (0.4, 0.3, 0.1, -0.05)), ("equal", 3, "t3", (0.5, -0.2, 0.0, 0.02)), ("off", 1, "t1", (0.0, 1.0, 0.0, 0.0))]

    def gapfun(c):
        a, b, c0, ct = c
        return lambda x, t: (a * x[:, 0] ** 2 + b * x[:, 0] + c0 + ct * t, (2 * a * x + b), np.full((x.shape[0], 1, 1), 2 * a))
    for mode, inod, field, c in cons:
        mb.addelement(model, mb.DofConstraint, [nod[inod]], xinod=(1,), xfield=(field,), λinod=1, λclass="X", λfield="λ" + field, gap=gapfun(c), mode=mode)
    mb.setscale(model, scale=dict(X=dict(t1=2., t2=2., t3=2., λt2=3., λt3=.5), U=dict(t1=5., t2=5., t3=5.)), Λscale=4.)
    st0 = mb.initialize(model); dis = st0.dis
    nX, nU, nA = model.getndof(("X", "U", "A"))
    st = states(mb, nX, nU, nstep)
    Lam = [mb.synthetic.uniform_pm1(700 + s, nX) for s in range(nstep)]
    time = t0 + dt * np.arange(nstep)
    odis = [dict(X=d.X, U=d.U, A=d.A) for d in dis.dis]
    P = OP.prepare_direct(odis, nX, nU, nA, OX, OU, 0)
    big, bigasm, pgr, pgc = OP.preparebig(0, [nstep], P["nL2"], P["pat"])
    A = P["asm"]; sL, sX = dis.scaleΛ, dis.scaleX
    nnzXX = len(P["pat"][(2, 2)][3])
    outs = []
    for s, (X, U) in enumerate(st):
        o = OE.direct_out_zeros(P, OX, OU)
        OE.direct_assemble_step_beams(model.ele[0].eleobj, dis.dis[0].X, dis.dis[0].U, OX, OU, X[: OX + 1], [U], dis.dis[0].scaleX, dis.dis[0].scaleU, P, 0, out=o)
        o["L1"][2] = np.zeros((OX + 1, nX)); hxx = np.zeros(nnzXX)
        for k, (mode, inod, field, (a, b, c0, ct)) in enumerate(cons, start=1):
            ix = dis.dis[k].X[0] - 1
            x, lam = X[0][ix[0]], X[0][ix[1]]
            g, dg, d2g = a * x * x + b * x + c0 + ct * time[s], 2 * a * x + b, 2 * a
            L1, L2 = Lam[s][ix]
            if mode == "positive":
                R = np.array([-dg * lam, -g * lam]); K = np.array([[-d2g * lam, -dg], [-dg * lam, -g]])
                H = np.array([[L2 * (-d2g * lam), L1 * (-d2g) + L2 * (-dg)], [L1 * (-d2g) + L2 * (-dg), 0.]])
            elif mode == "equal":
                R = np.array([-dg * lam, -g]); K = np.array([[-d2g * lam, -dg], [-dg, 0.]])
                H = np.array([[L2 * (-d2g), L1 * (-d2g)], [L1 * (-d2g), 0.]])
            else:
                R = np.array([0., -lam]); K = np.array([[0., 0.], [0., -1.]]); H = np.zeros((2, 2))
            aL, aXv, aLX, aXL, aXX = (A[OP.arrnum(1)][k][:, 0], A[OP.arrnum(2)][k][:, 0], A[OP.arrnum(1, 2)][k][:, 0], A[OP.arrnum(2, 1)][k][:, 0],
                                      A[OP.arrnum(2, 2)][k][:, 0])
            for i in range(2):
                o["L1"][1][aL[i] - 1] += R[i] * sL[ix[i]]
                o["L1"][2][0, aXv[i] - 1] += (K[:, i] @ Lam[s][ix]) * sX[ix[i]]
                for j in range(2):
                    o["L2"][(1, 2)][0, aLX[i + 2 * j] - 1] += K[i, j] * sL[ix[i]] * sX[ix[j]]
                    o["L2"][(2, 1)][0, aXL[j + 2 * i] - 1] += K[i, j] * sL[ix[i]] * sX[ix[j]]
                    hxx[aXX[i + 2 * j] - 1] += H[i, j] * sX[ix[i]] * sX[ix[j]]
        o["L2"][(2, 2)] = {(1, 1): hxx}
        outs.append(o)
    nz, Lv = OP.assemblebig(0, nstep, dt, P, big, bigasm, pgr, outs)
    assert max(np.abs(o["L2"][(2, 2)][(1, 1)]).max() for o in outs) > 1e-3

    eng = mb.directxua.prepare(OX, OU, model, dis, nstep, dt, t0=t0)
    try:
        cp, rv = eng.big_pattern()
        assert np.array_equal(cp, big["colptr"]) and np.array_equal(rv, big["rowval"])
        for s, (X, U) in enumerate(st):
            eng.set_state(s, X, U); eng.set_lambda(s, Lam[s])
            mb.directxua.host_elements(eng, s, X, Lam[s], time[s], model.scaleΛ)
        Lvv = np.zeros(eng.nnzbig); Lvec = np.zeros(eng.ncol)
        eng.direct_assemble(Lvv=Lvv, Lv=Lvec)
        assert rel(Lvv, nz) <= TOL, rel(Lvv, nz)
        assert rel(Lvec, Lv, np.abs(nz).max()) <= TOL
        Lvv2 = np.zeros(eng.nnzbig)
        eng.direct_assemble(Lvv=Lvv2)
        assert np.array_equal(Lvv, Lvv2)          # the X-X entries are added once per assembly, not accumulated
    finally:
        eng.close()


@pytest.mark.parametrize("OX", [2, 0])
def test_directxua_beam_bar_soil(mb, OX):
    """BASELINE.json configs[4] in small: EulerBeam3D{Udof} + Bar3D{Udof} + SoilContact on shared nodes through the DirectXUA first-order path
    (DirectXUA.jl:85-120); every element type on the device; Bar3D reads state.time (weight ramp, BarElement.jl:144)."""
    OU, nstep, dt, N, t0 = 0, 7, 0.1, 10, -7.4
    rng = np.random.default_rng(5)
    model = mb.Model()
    coord = np.cumsum(np.concatenate([[[0., 0., -0.3]], rng.uniform(0.5, 1.0, (N, 3)) * [1, .3, .02]]), axis=0)
    nod = mb.addnode(model, coord)
    unod = mb.addnode(model, np.zeros((N, 0)))
    mesh = np.stack([nod[:-1], nod[1:], unod], axis=1)
    bmat = mb.BeamCrossSection(EA=1e3, EI2=30., EI3=20., GJ=40., mu=2., iota1=.3, w=5., Ca2=3., Ca3=3., Cq2=2., Cq3=2., Cl1=.5)
    mb.addelement(model, mb.EulerBeam3D, mesh[: N // 2], mat=bmat, orient2=(0., 0.2, 1.), Udof=True)
    rmat = mb.AxisymmetricBarCrossSection(EA=500., mu=1.5, w=3., Cat=.2, Clt=.3, Cqt=.4, Can=2., Cln=.6, Cqn=1.2)
    mb.addelement(model, mb.Bar3D, mesh[N // 2:], mat=rmat, Udof=True)
    mb.addelement(model, mb.SoilContact, nod[: N // 2, None], z0=0.0, Kh=30., Kv=200., Ch=3., Cv=7.)
    mb.setscale(model, scale=dict(X=dict(t1=2., t2=2., t3=2.), U=dict(t1=5., t2=5., t3=5.)), Λscale=7.)
    st0 = mb.initialize(model); dis = st0.dis
    nX, nU, nA = model.getndof(("X", "U", "A"))
    st = states(mb, nX, nU, nstep)
    Lam = [mb.synthetic.uniform_pm1(500 + s, nX) for s in range(nstep)]          # SoilContact takes the second-order path: Λ enters L1[X]
    odis = [dict(X=d.X, U=d.U, A=d.A) for d in dis.dis]
    P = OP.prepare_direct(odis, nX, nU, nA, OX, OU, 0)
    big, bigasm, pgr, pgc = OP.preparebig(0, [nstep], P["nL2"], P["pat"])
    bars, soil = model.ele[1].eleobj, model.ele[2].eleobj
    outs = []
    for s, (X, U) in enumerate(st):
        o = OE.direct_out_zeros(P, OX, OU)
        OE.direct_assemble_step_beams(model.ele[0].eleobj, dis.dis[0].X, dis.dis[0].U, OX, OU, X[: OX + 1], [U], dis.dis[0].scaleX, dis.dis[0].scaleU, P, 0, out=o)
        OE.direct_addin_generic(lambda e, xv, sd, uv, usd: OE.bar_residual(bars[e], xv, sd, uv, usd, t=t0 + s * dt)[:2], 6, 3, dis.dis[1].X, dis.dis[1].U,
                                OX, OU, X[: OX + 1], [U], dis.dis[1].scaleX, dis.dis[1].scaleU, P, 1, o)
        OE.direct_addin_soil_second_order(soil, dis.dis[2].X, OX, X[: OX + 1], Lam[s], dis.dis[2].scaleX, model.scaleΛ, P, 2, o)
        outs.append(o)
    nz, Lv = OP.assemblebig(0, nstep, dt, P, big, bigasm, pgr, outs)
    eng = mb.directxua.prepare(OX, OU, model, dis, nstep, dt, t0=t0)
    cp, rv = eng.big_pattern()
    assert np.array_equal(cp, big["colptr"]) and np.array_equal(rv, big["rowval"])
    for ityp in (1, 2, 3):
        assert np.array_equal(eng.direct_asm(ityp, 0).T, P["asm"][OP.arrnum(1, 2)][ityp - 1])
    for s, (X, U) in enumerate(st):
        eng.set_state(s, X[: OX + 1], U)
        eng.set_lambda(s, Lam[s])
    Lvv = np.zeros(eng.nnzbig); Lvec = np.zeros(eng.ncol)
    eng.direct_assemble(Lvv=Lvv, Lv=Lvec)
    W = 2 * nX + nU
    assert np.abs(Lvec.reshape(nstep, W)[:, nX: 2 * nX]).max() > 0            # L1[X] of the soil springs reached Lv
    for s in (0, 4):
        o = outs[s]
        scale = np.abs(o["L2"][(1, 2)]).max()
        assert rel(eng.step_block(s, 0), o["L1"][1], scale) <= TOL
        for der in range(OX + 1):
            assert rel(eng.step_block(s, 1, der), o["L2"][(1, 2)][der], scale) <= TOL
            assert rel(eng.step_block(s, 2, der), o["L2"][(2, 1)][der], scale) <= TOL
        assert rel(eng.step_block(s, 3), o["L2"][(1, 3)][0], scale) <= TOL and rel(eng.step_block(s, 4), o["L2"][(3, 1)][0], scale) <= TOL
    assert rel(Lvv, nz) <= TOL, rel(Lvv, nz)
    assert rel(Lvec, Lv, np.abs(nz).max()) <= TOL
    z = np.array([st[s][0][0][dis.dis[2].X[:, 2] - 1] for s in range(nstep)])
    assert (z < 0).any() and (z >= 0).any()          # both SoilContact branches
    eng.close()
    # the same three device types inside the GENERAL form (mb_xua_add_device_eletyp / mb_xua_eval_device)
    from muscade_b200 import xua
    g = xua.XUAEngine(0)
    try:
        g.prepare(model, dis, OX, OU, 0, [nstep], [dt])
        g.set_time0(1, t0)
        cp, rv = g.big_pattern()
        assert np.array_equal(cp, big["colptr"]) and np.array_equal(rv, big["rowval"])
        sts = [[mb.State(t0 + s * dt, [Lam[s]], st[s][0][: OX + 1], [st[s][1]], st0.A, None, model, dis) for s in range(nstep)]]
        g.assemblebig(sts)
        gLvv, gLv = g.big()
        assert rel(gLvv, nz) <= TOL and rel(gLv, Lv, np.abs(nz).max()) <= TOL
    finally:
        g.close()


def test_directxua_sparser_and_decrementbig(mb):
    """sparser!(cLvv,Lvv,rtol) (SparseTools.jl:172-199) and decrementbig! (DirectXUA.jl:357-383) on the device against their restatements:
    compacted structure bit-exact, values copied; states after the update to 1e-14 (Δt^(1−βder) is an un-pinned Julia `^`), Δ² to 1e-13."""
    OX, OU, nstep, dt, N = 2, 0, 8, 0.1, 6
    model = udof_chain(mb, N, np.random.default_rng(2))
    mb.setscale(model, scale=dict(X=dict(t1=2., t2=2., t3=2.), U=dict(t1=5., t2=5., t3=5.)), Λscale=1e3)
    st0 = mb.initialize(model); dis = st0.dis
    nX, nU = model.getndof("X"), model.getndof("U")
    st = states(mb, nX, nU, nstep)
    eng = mb.directxua.prepare(OX, OU, model, dis, nstep, dt)
    for s, (X, U) in enumerate(st):
        eng.set_state(s, X, U)
    Lvv = np.zeros(eng.nnzbig)
    eng.direct_assemble(Lvv=Lvv)
    cp, rv = eng.big_pattern()
    for rtol in (1e-20, 1e-3):
        nkeep = eng.sparser(rtol)
        c2, r2, v2 = eng.sparse()
        oc, orow, ov = OP.sparser(cp, rv, Lvv, rtol)
        assert nkeep == len(ov) and np.array_equal(c2, oc) and np.array_equal(r2, orow) and np.array_equal(v2, ov)
    assert eng.sparser(1e-20) < 0.75 * eng.nnzbig              # the structurally-zero X-X / U-U blocks are gone
    # decrementbig!
    rng = np.random.default_rng(8)
    W = 2 * nX + nU
    dv = rng.standard_normal(nstep * W) * 1e-2
    Lam = [rng.standard_normal(nX) for _ in range(nstep)]
    sL, sX, sU = dis.scaleΛ, dis.scaleX, dis.scaleU
    assert len(set(sL)) > 1 and len(set(sX)) > 1
    eng.set_dof_scale(sL, sX, sU)
    for s in range(nstep):
        eng.set_lambda(s, Lam[s])
    ost = [dict(L=[Lam[s].copy()], X=[x.copy() for x in st[s][0]], U=[st[s][1].copy()]) for s in range(nstep)]
    d2ref = OP.decrementbig(ost, dv, OX, OU, dt, nstep, nX, nU, sL, sX, sU)
    d2 = eng.decrement(dv)
    assert np.abs(d2 - d2ref).max() <= 1e-13 * d2ref.max()
    for s in range(nstep):
        X, U, L = eng.get_state(s)
        assert np.abs(L - ost[s]["L"][0]).max() <= 1e-14 * np.abs(ost[s]["L"][0]).max()
        assert np.abs(U - ost[s]["U"][0]).max() <= 1e-14 * np.abs(ost[s]["U"][0]).max()
        for d in range(3):
            ref = ost[s]["X"][d]
            assert np.abs(X[d] - ref).max() <= 1e-14 * max(np.abs(ref).max(), np.abs(dv).max() / dt ** d), (s, d)
    eng.close()
