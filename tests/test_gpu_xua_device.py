"""GPU parity: DEVICE element types in the general DirectXUA path (mb_xua_add_device_eletyp / mb_xua_eval_device) and the ElementCost accelerator for strain
gauges on beams (mb_xua_set_gauge_cost) — the reference's TestBeamElementStrainGauge.jl numbers through the C ABI, finite differences, and the oracle's beam."""
import numpy as np
import pytest

import muscade_b200 as mb
from muscade_b200 import xua
from oracle import elements as OE
from oracle import pattern as OP

import xua_models as XM

pytestmark = pytest.mark.gpu

P5 = np.array([[0., .5, 0.], [0., 0, .5], [0., -.5, 0.], [0., 0, -.5], [0., .5, 0.]]).T          # test/TestBeamElementStrainGauge.jl:10-11
D5 = np.array([[1., 0., 0.], [1., 0., 0.], [1., 0., 0.], [1., 0., 0.], [1 / np.sqrt(2), 0, 1 / np.sqrt(2)]]).T
MAT = dict(EA=10., EI2=3., EI3=3., GJ=4., mu=1., iota1=1.0)
SIGMA = 15e-6
measured = lambda t: np.array([np.cos(t), 0., -np.cos(t), 0., np.cos(t) / 2]) * 0.001                # :93


@pytest.fixture
def xeng(mb):
    made = []

    def make():
        e = xua.XUAEngine(0); made.append(e); return e
    yield make
    for e in made:
        e.close()


def costed_beam_model(coords, conn, Udof=False):
    m = mb.Model("gauged")
    nod = mb.addnode(m, np.asarray(coords, float))
    nodes = nod[np.asarray(conn)]
    if Udof:
        un = np.array([mb.addnode(m, np.zeros(3)) for _ in range(len(conn))])
        nodes = np.concatenate([nodes, un[:, None]], axis=1)
    cost = mb.QuadraticGaugeCost(SIGMA, measured)
    mb.addelement(m, mb.ElementCost, nodes, req=("ε",), cost=cost, ElementType=mb.StrainGaugeOnEulerBeam3D,
                  elementkwargs=dict(P=P5, D=D5, elementkwargs=dict(mat=mb.BeamCrossSection(**MAT), orient2=(0., 1., 0.), Udof=Udof)))
    return m


def _eval(eng, m, s0, X0, Lam=None, t=0.):
    st = s0.with_orders(1, 1, 1)
    s = mb.State(t, [np.zeros_like(X0) if Lam is None else np.asarray(Lam, float)], [np.asarray(X0, float)], [u.copy() for u in st.U], st.A, None, m, s0.dis)
    eng.zero()
    eng.assemble_step(1, 1, s)
    return s


def test_reference_strain_gauge_goldens(xeng):
    """test/TestBeamElementStrainGauge.jl: eleres of :36-58 (three states) and ∇L[2][1] of the ElementCost-wrapped beam, :108-111"""
    m = costed_beam_model([[0., 0, 0], [4., 0, 0]], [[0, 1]]); s0 = mb.initialize(m)
    eng = xeng(); eng.prepare(m, s0.dis, 0, 0, 0, [1], [1.])
    G = mb.StrainGaugeOnEulerBeam3D.gauge_matrix(P5, D5)
    for X0, kap, eps in (([0, 0, 0, 0, .1, 0, 0, 0, 0, 0, -.1, 0], np.array([0, 0, 1]) / 20, np.array([0, -1, 0, 1, 0]) / 40),
                         ([0, 0, 0, 0, 0, .1, 0, 0, 0, 0, 0, -.1], np.array([0, -1, 0]) / 20, np.array([1, 0, -1, 0, .5]) / 40),
                         ([0, 0, 0, 0, 0, 0, 0, 0, 0, 1., 0, 0], np.array([.25, 0, 0]), np.array([0, 0, 0, 0, 0.0625]))):
        _eval(eng, m, s0, X0)
        e4, J, c = eng.get_gauge(1)
        assert abs(e4[0, 0]) < 1e-12 and np.allclose(e4[0, 1:], kap, atol=1e-12) and np.allclose(G @ e4[0], eps, atol=1e-12)
        r = G @ e4[0] - measured(0.)
        assert np.isclose(c[0], r @ r / (2 * SIGMA ** 2), rtol=1e-12)
    X = np.array([0, 0, 0, 0, .1, 0, 0, 0, 0, 0, -.1, 0])
    _eval(eng, m, s0, X, Lam=X)
    g, H = eng.get_packet(1)
    ref = np.array([277777.7777783019, 0.0, 1.1102230246251565e-16, 201441.02435832855, 2.7777777927777793e7, -1.2430497627256343e6, -277777.7777783019, 0.0,
                    -1.1102230246251565e-16, -76336.75341947998, -2.7777777927777793e7, 1.2569502372743965e6])
    assert np.allclose(g[0, 12:24], ref, rtol=1e-9, atol=1e-6 * 0 + 1e-9 * np.abs(ref).max())


def test_gauge_jacobian_and_gauss_newton_hessian(xeng):
    rng = np.random.default_rng(2)
    n = 37
    coords = np.cumsum(rng.uniform(0.5, 1.5, (n + 1, 3)), axis=0)
    m = costed_beam_model(coords, [[i, i + 1] for i in range(n)]); s0 = mb.initialize(m); dis = s0.dis
    eng = xeng(); eng.prepare(m, dis, 0, 0, 0, [1], [1.])
    nX = m.getndof("X")
    X0 = rng.normal(0, 0.05, nX); Lam = rng.normal(0, 1., nX)
    _eval(eng, m, s0, X0, Lam)
    e4, J, c = eng.get_gauge(1)
    g, H = eng.get_packet(1)
    idx = dis.dis[0].X
    # J by central differences of the kernel's own requestables
    h = 1e-6
    Jfd = np.zeros_like(J)
    for e in range(0, n, 6):                       # element by element: perturb its own 12 dofs
        for d in range(12):
            Xp, Xm = X0.copy(), X0.copy(); Xp[idx[e, d] - 1] += h; Xm[idx[e, d] - 1] -= h
            _eval(eng, m, s0, Xp); ep = eng.get_gauge(1)[0][e]
            _eval(eng, m, s0, Xm); em_ = eng.get_gauge(1)[0][e]
            Jfd[e, :, d] = (ep - em_) / (2 * h)
        assert np.allclose(J[e], Jfd[e], rtol=1e-6, atol=1e-8), e
    # packet = [R; Λᵀ∂R/∂X + Jᵀ Gᵀ r/σ²], ∇²L[X,X] = Jᵀ GᵀG J/σ² (no Λ·∂²R/∂X²), Λ rows = ∂R/∂X
    G = mb.StrainGaugeOnEulerBeam3D.gauge_matrix(P5, D5)
    r = e4 @ G.T - measured(0.)[None, :]
    dRdX = H[:, :12, 12:24]
    assert np.allclose(H[:, 12:24, :12], dRdX.transpose(0, 2, 1), rtol=0, atol=0)
    gX = np.einsum("ei,eij->ej", Lam[idx - 1], dRdX) + np.einsum("ekd,gk,eg->ed", J, G, r) / SIGMA ** 2
    assert np.allclose(g[:, 12:24], gX, rtol=1e-12, atol=1e-12 * np.abs(gX).max())
    HXX = np.einsum("eki,gk,gl,elj->eij", J, G, G, J) / SIGMA ** 2
    assert np.allclose(H[:, 12:24, 12:24], HXX, rtol=1e-12, atol=1e-12 * np.abs(HXX).max())
    assert not H[:, :12, :12].any()
    assert np.allclose(c, (r * r).sum(1) / (2 * SIGMA ** 2), rtol=1e-12)


@pytest.mark.parametrize("OX,OU,Udof", [(0, 0, False), (2, 0, True), (1, 0, True)])
def test_device_beams_against_oracle(xeng, OX, OU, Udof):
    """plain EulerBeam3D types evaluated by the device kernels inside the general form, with host-evaluated costs beside them: out, Lvv.nzval, Lv against the oracle
    (its C++ beam through DirectXUA's first-order addin!, src/DirectXUA.jl:85-120)"""
    rng = np.random.default_rng(4)
    n, nstep, dt = 11, 6, 0.5
    m = mb.Model("beams")
    nod = mb.addnode(m, np.cumsum(rng.uniform(0.5, 1.5, (n + 1, 3)), axis=0))
    nodes = np.stack([nod[:-1], nod[1:]], axis=1)
    if Udof:
        un = np.array([mb.addnode(m, np.zeros(3)) for _ in range(n)])
        nodes = np.concatenate([nodes, un[:, None]], axis=1)
    mb.addelement(m, mb.EulerBeam3D, nodes, mat=mb.BeamCrossSection(EA=1e3, EI2=30., EI3=20., GJ=40., mu=1.5, iota1=0.7, w=0.3, Ca2=0.1, Cl2=0.2), Udof=Udof)
    mb.addelement(m, mb.SingleDofCost, nod[::3, None], clas="X", field="t2", cost=XM.l1)
    if Udof:
        mb.addelement(m, mb.SingleDofCost, un[:, None], clas="U", field="t1", cost=XM.fu)
    mb.setscale(m, scale=dict(X=dict(t1=2., t2=2., t3=2., r1=0.5, r2=0.5, r3=0.5), U=dict(t1=3., t2=3., t3=3.)))
    s0 = mb.initialize(m); dis = s0.dis
    nX, nU, nA = m.getndof(("X", "U", "A"))
    eng = xeng(); eng.prepare(m, dis, OX, OU, 0, [nstep], [dt])
    st = s0.with_orders(1, OX + 1, OU + 1)
    states = [[mb.State(1. + dt * k, [rng.normal(0, 1., nX)], [rng.normal(0, 0.05, nX) for _ in range(OX + 1)], [rng.normal(0, 0.5, nU) for _ in range(OU + 1)],
                        st.A, None, m, dis) for k in range(nstep)]]
    eng.assemblebig(states)
    P = OP.prepare_direct(XM.dis_lists(dis), nX, nU, nA, OX, OU, 0)
    big, basm, pgr, _ = OP.preparebig(0, [nstep], P["nL2"], P["pat"])
    outs = []
    ed = dis.dis[0]
    for s in states[0]:
        out = OP.out_zeros(P)
        b = OE.direct_assemble_step_beams(m.ele[0].eleobj, ed.X, ed.U if Udof else None, OX, 0, s.X, s.U, ed.scaleX, ed.scaleU if Udof else None, P, 0)
        out["L1"][1][0] += b["L1"][1]
        for d in range(OX + 1):
            out["L2"][(1, 2)][0, d] += b["L2"][(1, 2)][d]; out["L2"][(2, 1)][d, 0] += b["L2"][(2, 1)][d]
        if Udof:
            out["L2"][(1, 3)][0, 0] += b["L2"][(1, 3)][0]; out["L2"][(3, 1)][0, 0] += b["L2"][(3, 1)][0]
        for ityp in range(1, len(m.ele)):
            g, H = xua.packets(m.ele[ityp], dis.dis[ityp], OX, OU, 0, s.Λ[0], s.X, s.U, s.A, s.time)
            OP.lagrangian_addition(P, OX, OU, 0, ityp, g, H, out)
        outs.append(out)
    nz_o, Lv_o = OP.assemblebig_general(0, [nstep], [dt], P, big, basm, pgr, None, [outs])
    nz, Lv = eng.big()
    assert np.abs(nz - nz_o).max() <= 1e-12 * np.abs(nz_o).max() and np.abs(Lv - Lv_o).max() <= 1e-12 * max(np.abs(Lv_o).max(), np.abs(nz_o).max())
    last = outs[-1]
    assert np.abs(eng.get_out(1) - last["L1"][1][0]).max() <= 1e-12 * np.abs(nz_o).max()
    for d in range(OX + 1):
        assert np.abs(eng.get_out(1, 2, 1, d + 1) - last["L2"][(1, 2)][0, d]).max() <= 1e-12 * np.abs(last["L2"][(1, 2)]).max()
        assert np.abs(eng.get_out(2, 1, d + 1, 1) - last["L2"][(2, 1)][d, 0]).max() <= 1e-12 * np.abs(last["L2"][(2, 1)]).max()
    assert np.array_equal(eng.get_out(2, 2, 1, 1), last["L2"][(2, 2)][0, 0])          # host-evaluated costs only: bit-exact


def test_strain_gauge_identification_converges(mb):
    """solve(DirectXUA{0,0,0}) on a cantilever of Udof beams with gauges: the unknown distributed load is identified from synthetic strain measurements.
    Newton with the accelerator's Gauss-Newton Hessian converges, and the gauge strains of the solution reproduce the measurements."""
    n = 8
    coords = np.stack([np.linspace(0., 8., n + 1), np.zeros(n + 1), np.zeros(n + 1)], axis=1)
    m = mb.Model("cantilever")
    nod = mb.addnode(m, coords)
    un = np.array([mb.addnode(m, np.zeros(3)) for _ in range(n)])
    nodes = np.concatenate([np.stack([nod[:-1], nod[1:]], axis=1), un[:, None]], axis=1)
    target = np.zeros((n, 5)); target[:, 1] = -2e-4 * (1 - np.arange(n) / n); target[:, 3] = -target[:, 1]          # bending about axis 3, fading towards the tip
    cost = mb.QuadraticGaugeCost(1e-5, lambda t: target)
    mb.addelement(m, mb.ElementCost, nodes, req=("ε",), cost=cost, ElementType=mb.StrainGaugeOnEulerBeam3D,
                  elementkwargs=dict(P=P5, D=D5, elementkwargs=dict(mat=mb.BeamCrossSection(EA=1e4, EI2=300., EI3=300., GJ=400., mu=1., iota1=1.), orient2=(0., 1., 0.), Udof=True)))
    for f in ("t1", "t2", "t3", "r1", "r2", "r3"):
        mb.addelement(m, HoldD2, [nod[0]], field=f)
    mb.addelement(m, mb.SingleDofCost, un[:, None], clas="U", field="t1", cost=lambda u, t: u ** 2 * 1e-2)
    mb.addelement(m, mb.SingleDofCost, un[:, None], clas="U", field="t2", cost=lambda u, t: u ** 2 * 1e-2)
    mb.addelement(m, mb.SingleDofCost, un[:, None], clas="U", field="t3", cost=lambda u, t: u ** 2 * 1e-6)
    s0 = mb.initialize(m)
    st = xua.solve(0, 0, 0, [s0], [np.array([0., 1.])], maxiter=30, maxΔλ=1e-3, maxΔx=1e-7, maxΔu=1e-5)
    eng = xua.XUAEngine(0)
    try:
        eng.prepare(m, s0.dis, 0, 0, 0, [2], [1.])
        eng.zero(); eng.assemble_step(1, 1, st[0][0])
        e4, J, c = eng.get_gauge(1)
    finally:
        eng.close()
    G = mb.StrainGaugeOnEulerBeam3D.gauge_matrix(P5, D5)
    eps = e4 @ G.T
    assert np.abs(eps[:, [1, 3]] - target[:, [1, 3]]).max() < 0.1 * np.abs(target).max()
    assert np.abs(st[0][0].U[0]).max() > 1e-3


class HoldD2(mb.LagrangianElement):
    """Hold(nod;field) (src/BasicElements.jl:440-448) = DofConstraint with gap = x, mode equal: R = (−λ, −x)"""
    type_parameters = ("field",)

    @classmethod
    def doflist(cls, field, **kw):
        return (1, 1), ("X", "X"), (field, "λ" + field)

    @classmethod
    def construct(cls, coords, field):
        return np.zeros((coords[0].shape[0], 0)), dict(field=field)

    @staticmethod
    def residual(o, extra, X, U, A, t, SP):
        x, lam = X[0][0], X[0][1]
        return [-lam, -x]


@pytest.mark.parametrize("OX,per_element,Udof", [(0, False, True), (2, True, True), (1, False, True), (2, False, False)])
def test_gauge_cost_in_the_windowed_path_equals_general_form(mb, OX, per_element, Udof):
    """ElementCost{StrainGaugeOnEulerBeam3D} on the beam-specialised, windowed path (mb_direct_set_gauge_cost: the costed beam's ∇L[X_d], ∇L[U], Gauss-Newton X₀-X₀ block and
    scale.Λ-scaled Λ rows enter the block-implicit Lvv / Lv) against the general form (mb_xua_*, itself checked against the oracle and the reference's goldens above):
    same structure, Lvv values and Lv within 1e-12, with random Λ, X, X′, X″, U, host-evaluated single-dof costs beside the gauges, non-unit scales."""
    rng = np.random.default_rng(11)
    n, nstep, dt = 9, 7, 0.25
    m = mb.Model("gauged chain")
    nod = mb.addnode(m, np.cumsum(rng.uniform(0.5, 1.5, (n + 1, 3)), axis=0))
    nodes = np.stack([nod[:-1], nod[1:]], axis=1)
    if Udof:
        un = np.array([mb.addnode(m, np.zeros(3)) for _ in range(n)])
        nodes = np.concatenate([nodes, un[:, None]], axis=1)
    tgt = rng.normal(0., 1e-3, (n, 5))
    meas = (lambda t: tgt * np.cos(t)) if per_element else (lambda t: tgt[0] * np.cos(t))
    cost = mb.QuadraticGaugeCost(2e-3, meas)
    mb.addelement(m, mb.ElementCost, nodes, req=("ε",), cost=cost, ElementType=mb.StrainGaugeOnEulerBeam3D,
                  elementkwargs=dict(P=P5, D=D5, elementkwargs=dict(mat=mb.BeamCrossSection(EA=1e3, EI2=30., EI3=20., GJ=40., mu=1.5, iota1=0.7, Ca2=0.1), orient2=(0., 1., 0.), Udof=Udof)))
    mb.addelement(m, mb.SingleDofCost, nod[::3, None], clas="X", field="t2", cost=XM.l1)
    if Udof:
        mb.addelement(m, mb.SingleDofCost, un[:, None], clas="U", field="t1", cost=XM.fu)
    mb.setscale(m, scale=dict(X=dict(t1=2., t2=2., t3=2., r1=0.5, r2=0.5, r3=0.5), U=dict(t1=3., t2=3., t3=3.)), Λscale=7.)
    s0 = mb.initialize(m); dis = s0.dis
    nX, nU = m.getndof("X"), m.getndof("U")
    time = 1. + dt * np.arange(nstep)
    st = s0.with_orders(1, OX + 1, 1)
    states = [[mb.State(time[k], [rng.normal(0, 1., nX)], [rng.normal(0, 0.05, nX) for _ in range(OX + 1)], [rng.normal(0, 0.5, nU)], st.A, None, m, dis) for k in range(nstep)]]
    spec = mb.directxua.prepare(OX, 0, m, dis, nstep, dt, t0=time[0])
    gen = xua.XUAEngine(0)
    try:
        assert len(spec.gauge_costs) == 1
        for k, s in enumerate(states[0]):
            spec.set_state(k, s.X, s.U[0]); spec.set_lambda(k, s.Λ[0])
            spec.set_host_cost(k, *mb.directxua.host_costs(spec, k, s.X[0], s.U[0], time[k])[:4])
        spec.set_gauge_times(time)
        Lvv = np.zeros(spec.nnzbig); Lv = np.zeros(spec.ncol)
        spec.direct_assemble(Lvv=Lvv, Lv=Lv)
        cp, rv = spec.big_pattern()
        nbig, nnz = gen.prepare(m, dis, OX, 0, 0, [nstep], [dt])
        assert nbig == spec.ncol and nnz == spec.nnzbig
        gcp, grv = gen.big_pattern()
        assert np.array_equal(cp, gcp) and np.array_equal(rv, grv)
        gen.assemblebig(states)
        gLvv, gLv = gen.big()
        scale = np.abs(gLvv).max()
        assert np.abs(gLvv - Lvv).max() <= 1e-12 * scale
        assert np.abs(gLv - Lv).max() <= 1e-12 * max(scale, np.abs(gLv).max())
        # the cost is really there: X-X entries of the (step,step) blocks and the U part of Lv are non-zero
        W = 2 * nX + nU
        import scipy.sparse as sp
        A = sp.csc_matrix((Lvv, rv - 1, cp - 1), shape=(nbig, nbig)).tocsr()
        assert abs(A[nX:2 * nX, nX:2 * nX]).max() > 0 and (not Udof or np.abs(Lv[2 * nX:W]).max() > 0)
    finally:
        spec.close(); gen.close()


def test_gauge_cost_windowed_path_errors(mb):
    """argument errors of the two entry points: cost on a non-beam type, after prepare, measurements missing at assembly, per_element flipped"""
    import ctypes as C
    from muscade_b200._lib import ptr
    m = costed_beam_model([[0., 0, 0], [4., 0, 0], [8., 0, 0]], [[0, 1], [1, 2]]); s0 = mb.initialize(m)
    eng = mb.directxua.prepare(0, 0, m, s0.dis, 6, 0.1)
    try:
        G = np.zeros((5, 4))
        assert eng.L.mb_direct_set_gauge_cost(eng.h, 1, 5, ptr(G), 1.) != 0                # already prepared
        with pytest.raises(mb.MuscadeB200Error, match="measurements"):
            eng.direct_assemble()
        eng.set_gauge_measurements(0, 1, np.zeros(5))
        with pytest.raises(mb.MuscadeB200Error, match="per_element"):
            eng.set_gauge_measurements(1, 1, np.zeros((2, 5)))
        with pytest.raises(mb.MuscadeB200Error):
            eng.set_gauge_measurements(99, 1, np.zeros(5))
    finally:
        eng.close()


def test_gauge_identification_through_the_windowed_solver(mb):
    """solve(DirectXUA{0,0,0}) of the gauged cantilever through the beam-specialised path (directxua.solve) and through the general form (xua.solve): same converged states"""
    n, nstep = 8, 6
    coords = np.stack([np.linspace(0., 8., n + 1), np.zeros(n + 1), np.zeros(n + 1)], axis=1)
    m = mb.Model("cantilever")
    nod = mb.addnode(m, coords)
    un = np.array([mb.addnode(m, np.zeros(3)) for _ in range(n)])
    nodes = np.concatenate([np.stack([nod[:-1], nod[1:]], axis=1), un[:, None]], axis=1)
    target = np.zeros((n, 5)); target[:, 1] = -2e-4 * (1 - np.arange(n) / n); target[:, 3] = -target[:, 1]
    cost = mb.QuadraticGaugeCost(1e-5, lambda t: target * (1. + 0.2 * t))
    mb.addelement(m, mb.ElementCost, nodes, req=("ε",), cost=cost, ElementType=mb.StrainGaugeOnEulerBeam3D,
                  elementkwargs=dict(P=P5, D=D5, elementkwargs=dict(mat=mb.BeamCrossSection(EA=1e4, EI2=300., EI3=300., GJ=400., mu=1., iota1=1.), orient2=(0., 1., 0.), Udof=True)))
    for f in ("t1", "t2", "t3", "r1", "r2", "r3"):
        mb.addelement(m, mb.Hold, [nod[0]], field=f)
    mb.addelement(m, mb.SingleDofCost, un[:, None], clas="U", field="t1", cost=lambda u, t: u ** 2 * 1e-2)
    mb.addelement(m, mb.SingleDofCost, un[:, None], clas="U", field="t2", cost=lambda u, t: u ** 2 * 1e-2)
    mb.addelement(m, mb.SingleDofCost, un[:, None], clas="U", field="t3", cost=lambda u, t: u ** 2 * 1e-6)
    s0 = mb.initialize(m)
    time = np.linspace(0., 1., nstep)
    a = mb.directxua.solve(0, 0, s0, time, maxiter=30, maxΔλ=1e-3, maxΔx=1e-7, maxΔu=1e-5)
    b = xua.solve(0, 0, 0, [s0], [time], maxiter=30, maxΔλ=1e-3, maxΔx=1e-7, maxΔu=1e-5)[0]
    umax = max(np.abs(s.U[0]).max() for s in b)
    assert umax > 1e-3
    for sa, sb in zip(a, b):
        assert np.abs(sa.X[0] - sb.X[0]).max() <= 1e-6 * max(np.abs(sb.X[0]).max(), 1e-6)
        assert np.abs(sa.U[0] - sb.U[0]).max() <= 1e-6 * umax
    assert np.abs(a[-1].U[0]).max() > 1.1 * np.abs(a[0].U[0]).max()          # the measurements grow with t, and so does the identified load


def test_gauge_cost_sliding_window(mb):
    """the costed beam type inside mb_direct_rebase's sliding window (how BASELINE.json configs[3]-sized gauged models run: the general form materialises all steps, the
    windowed path does not): each window's Lvv columns / Lv rows equal the full general-form assembly; the per-step L2[X,X], L1[U] and the measurements move with the window."""
    rng = np.random.default_rng(5)
    n, nstep, dt, OX, L = 6, 24, 0.1, 2, 5
    m = mb.Model("gauged chain")
    nod = mb.addnode(m, np.cumsum(rng.uniform(0.5, 1.5, (n + 1, 3)), axis=0))
    un = np.array([mb.addnode(m, np.zeros(3)) for _ in range(n)])
    nodes = np.concatenate([np.stack([nod[:-1], nod[1:]], axis=1), un[:, None]], axis=1)
    tgt = rng.normal(0., 1e-3, 5)
    cost = mb.QuadraticGaugeCost(2e-3, lambda t: tgt * np.cos(t))
    mb.addelement(m, mb.ElementCost, nodes, req=("ε",), cost=cost, ElementType=mb.StrainGaugeOnEulerBeam3D,
                  elementkwargs=dict(P=P5, D=D5, elementkwargs=dict(mat=mb.BeamCrossSection(EA=1e3, EI2=30., EI3=20., GJ=40., mu=1.5, iota1=0.7), orient2=(0., 1., 0.), Udof=True)))
    mb.setscale(m, scale=dict(X=dict(t1=2., t2=2., t3=2.), U=dict(t1=3., t2=3., t3=3.)), Λscale=3.)
    s0 = mb.initialize(m); dis = s0.dis
    nX, nU = m.getndof("X"), m.getndof("U")
    W = 2 * nX + nU
    time = dt * np.arange(nstep)
    st = s0.with_orders(1, OX + 1, 1)
    states = [[mb.State(time[k], [rng.normal(0, 1., nX)], [rng.normal(0, 0.05, nX) for _ in range(OX + 1)], [rng.normal(0, 0.5, nU)], st.A, None, m, dis) for k in range(nstep)]]
    gen = xua.XUAEngine(0)
    spec = mb.directxua.prepare(OX, 0, m, dis, nstep, dt, 4, 4 + L)
    try:
        gen.prepare(m, dis, OX, 0, 0, [nstep], [dt])
        gcp, grv = gen.big_pattern()
        gen.assemblebig(states)
        nz, Lv = gen.big()
        stored = set()

        def check(lo):
            c0, c1 = lo * W, (lo + L) * W
            p0, p1 = gcp[c0] - 1, gcp[c1] - 1
            cp, rv = spec.big_pattern()
            assert np.array_equal(cp - 1, gcp[c0:c1 + 1] - 1 - p0) and np.array_equal(rv, grv[p0:p1])
            for k in range(lo - 2, lo + L + 2):
                if k in stored:
                    continue
                s = states[0][k]
                spec.set_state(k, s.X, s.U[0]); spec.set_lambda(k, s.Λ[0]); spec.set_gauge_measurements(k, 1, cost.measured(time[k]))
                spec.direct_assemble(eval_range=(k, k + 1), build_big=False)
            a = np.zeros(spec.nnzbig); b = np.zeros(spec.ncol)
            spec.direct_assemble(eval_range=(lo, lo), build_big=True, Lvv=a, Lv=b)
            assert np.abs(a - nz[p0:p1]).max() <= 1e-12 * np.abs(nz).max() and np.abs(b - Lv[c0:c1]).max() <= 1e-12 * max(np.abs(nz).max(), np.abs(Lv).max())
            stored.clear(); stored.update(range(lo - 2, lo + L + 2))

        check(4)
        spec.rebase(4 + L); check(4 + L)
        spec.rebase(12); check(12)             # forward, overlapping
        spec.rebase(7); check(7)               # backward
    finally:
        spec.close(); gen.close()


@pytest.mark.parametrize("OX", [0, 2])
def test_sweepx_sees_a_costed_beam_as_its_target(mb, OX):
    """An X-analysis of a model whose beams are wrapped in ElementCost (the SweepX run that provides DirectXUA's initial state): R = ∂L/∂Λ of L = getlagrangian(target) + cost
    is the target's residual (src/BasicElements.jl:117-132), so Lλ and Lλx are those of the plain beams, bit for bit"""
    rng = np.random.default_rng(2)
    n = 7
    coords = np.cumsum(rng.uniform(0.5, 1.5, (n + 1, 3)), axis=0)
    mat = dict(EA=1e3, EI2=30., EI3=20., GJ=40., mu=1.5, iota1=0.7, Ca2=0.1)
    res = []
    for gauged in (False, True):
        m = mb.Model("chain")
        nod = mb.addnode(m, coords)
        nodes = np.stack([nod[:-1], nod[1:]], axis=1)
        if gauged:
            mb.addelement(m, mb.ElementCost, nodes, req=("ε",), cost=mb.QuadraticGaugeCost(SIGMA, measured), ElementType=mb.StrainGaugeOnEulerBeam3D,
                          elementkwargs=dict(P=P5, D=D5, elementkwargs=dict(mat=mb.BeamCrossSection(**mat), orient2=(0., 1., 0.))))
        else:
            mb.addelement(m, mb.EulerBeam3D, nodes, mat=mb.BeamCrossSection(**mat), orient2=(0., 1., 0.))
        state = mb.initialize(m).with_orders(1, OX + 1, 1)
        ndof = m.getndof("X")
        for d in range(OX + 1):
            state.X[d] = mb.synthetic.uniform_pm1(5 + d, ndof) * (0.1 if d == 0 else 0.3)
        out, asm, gr = mb.sweepx.prepare(OX, m, state.dis)
        try:
            out.c = mb.synthetic.newmark_coefficients(OX, 0.3)
            mb.sweepx.assemble("iter", out, asm, state.dis, m, state, 0.3)
            res.append((out.Lλ.copy(), out.Lλx.data.copy(), out.Lλx.indices.copy()))
        finally:
            out.engine.close()
    assert np.array_equal(res[0][2], res[1][2]) and np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert np.abs(res[0][1]).max() > 0
