"""General DirectXUA path, CPU side: the host's dual numbers (adiff2.D2), the element packets (xua.packets) and the oracle's restatement of
DirectXUA_lagrangian_addition! / assemblebig! are pinned to the numbers test/TestDirectXUA.jl holds (out of the last assembled step, :93-108; assembleA!, :67-76)."""
import numpy as np

import muscade_b200 as mb
from muscade_b200 import xua
from muscade_b200.adiff2 import D2, exp10, sqrt, sin
from oracle import pattern as OP

import xua_models as XM


def test_d2_against_closed_forms():
    rng = np.random.default_rng(0)
    x = rng.uniform(0.5, 2., (7, 3))
    a, b, c = D2.variables(x, scale=[1., 2., 0.5])
    f = a * b / c + exp10(a) * sqrt(b) - sin(c) ** 3 + 2. * a - b / 3. + 1.
    X, Y, Z = x[:, 0], x[:, 1], x[:, 2]
    L10 = np.log(10.)
    v = X * Y / Z + 10 ** X * np.sqrt(Y) - np.sin(Z) ** 3 + 2 * X - Y / 3 + 1
    g = np.stack([Y / Z + L10 * 10 ** X * np.sqrt(Y) + 2, X / Z + 10 ** X * .5 / np.sqrt(Y) - 1 / 3, -X * Y / Z ** 2 - 3 * np.sin(Z) ** 2 * np.cos(Z)], 1) * [1., 2., .5]
    H = np.zeros((7, 3, 3))
    H[:, 0, 0] = L10 ** 2 * 10 ** X * np.sqrt(Y); H[:, 0, 1] = 1 / Z + L10 * 10 ** X * .5 / np.sqrt(Y); H[:, 0, 2] = -Y / Z ** 2
    H[:, 1, 1] = -10 ** X * .25 / Y ** 1.5; H[:, 1, 2] = -X / Z ** 2
    H[:, 2, 2] = 2 * X * Y / Z ** 3 - 6 * np.sin(Z) * np.cos(Z) ** 2 + 3 * np.sin(Z) ** 3
    H = H + np.triu(H, 1).transpose(0, 2, 1)
    H = H * np.outer([1., 2., .5], [1., 2., .5])
    assert np.allclose(f.v, v, rtol=1e-14) and np.allclose(f.grad(7, 3), g, rtol=1e-13) and np.allclose(f.hess(7, 3), H, rtol=1e-12, atol=1e-13)


def _assemble_outs(model, dis, P, OX, OU, IA, states):
    """oracle: out of assembleA! and of assemble! at every state"""
    outA = None
    if IA:
        outA = OP.out_zeros(P)
        for ityp, (et, ed) in enumerate(zip(model.ele, dis.dis)):
            if getattr(et.ElType, "acost", False):
                g, H = xua.packets(et, ed, OX, OU, IA, None, None, None, states[0][0].A, states[0][0].time, assembleA=True)
                # an Acost carries A partials only: place them where DirectXUA.jl:70-84 adds them
                na = ed.A.shape[1]
                asmA = P["asm"][OP.arrnum(4)][ityp]; asmAA = P["asm"][OP.arrnum(4, 4)][ityp]
                for e in range(et.nele):
                    for k in range(na):
                        outA["L1"][4][0, asmA[k, e] - 1] += g[e, k]
                    for kb in range(na):
                        for ka in range(na):
                            outA["L2"][(4, 4)][0, 0, asmAA[ka + na * kb, e] - 1] += H[e, ka, kb]
    outs = []
    for exp_states in states:
        row = []
        for s in exp_states:
            out = OP.out_zeros(P)
            for ityp, (et, ed) in enumerate(zip(model.ele, dis.dis)):      # Acost types too: src/Assemble.jl:477 does not match Vector{<:Acost}
                g, H = xua.packets(et, ed, OX, OU, IA, s.Λ[0], s.X, s.U, s.A, s.time)
                OP.lagrangian_addition(P, OX, OU, IA, ityp, g, H, out)
            row.append(out)
        outs.append(row)
    return outA, outs


def test_packets_and_oracle_reproduce_testdirectxua_out():
    """test/TestDirectXUA.jl:67-76 (assembleA!) and :93-108 (out after assemblebig!, i.e. of step 6 at t = 6)"""
    OX, OU, IA, nstep = 2, 0, 1, 6
    m = XM.model_testdirectxua()
    s0 = mb.initialize(m)
    dis = s0.dis
    P = OP.prepare_direct(XM.dis_lists(dis), 2, 4, 6, OX, OU, IA)
    st = s0.with_orders(1, OX + 1, OU + 1)
    states = [[mb.State(1. * i, [v.copy() for v in st.Λ], [v.copy() for v in st.X], [v.copy() for v in st.U], st.A, None, m, dis) for i in range(1, nstep + 1)]]
    outA, outs = _assemble_outs(m, dis, P, OX, OU, IA, states)
    D = lambda colptr, rowval, nz, mm, nn: _dense(colptr, rowval, nz, mm, nn)
    pat = P["pat"]
    # prepareA_out (:67-76)
    AA = D(pat[(4, 4)][2], pat[(4, 4)][3], outA["L2"][(4, 4)][0, 0], 6, 6)
    assert np.allclose(AA, 2e-14 * np.eye(6), rtol=1e-12, atol=0) and not outA["L1"][4].any()
    out = outs[0][-1]
    assert np.allclose(out["L1"][1], [[0., 0.]])                                                                   # :94
    assert np.allclose(out["L1"][2], [[0.055883099639785175, -0.1920340573300732], [0., 0.], [0., 0.]], rtol=1e-13)  # :95
    assert not out["L1"][3].any() and not out["L1"][4].any()                                                       # :96-97
    assert P["nL2"][(1, 1)] == (0, 0)                                                                              # :98
    d = lambda a, b, i, j: D(pat[(a, b)][2], pat[(a, b)][3], out["L2"][(a, b)][i - 1, j - 1], pat[(a, b)][0], pat[(a, b)][1])
    assert np.allclose(d(2, 2, 1, 1), [[2., 0.], [0., 2.]]) and not d(2, 2, 1, 2).any() and not d(2, 2, 1, 3).any()   # :99-101
    assert np.allclose(d(2, 1, 1, 1), [[2.1, -1.1], [-1.1, 1.1]], rtol=1e-13)                                      # :102
    assert np.allclose(d(2, 1, 2, 1), [[.05, 0.], [0., 0.]], rtol=1e-13)                                           # :103
    assert np.allclose(d(2, 1, 3, 1), [[1., 0.], [0., 1.]], rtol=1e-13)                                            # :104
    assert np.allclose(d(3, 3, 1, 1), np.diag([0., 0., 2., 2.]))                                                   # :105
    assert not d(3, 4, 1, 1).any() and d(3, 4, 1, 1).shape == (4, 6)                                               # :106
    # :107 — 2e-14·I after the LAST assemble!: the Acost elements went through it (src/Assemble.jl:477 does not match Vector{<:Acost})
    assert np.allclose(d(4, 4, 1, 1), 2e-14 * np.eye(6), rtol=1e-12, atol=0)


def _dense(colptr, rowval, nz, m, n):
    A = np.zeros((m, n))
    for j in range(n):
        for k in range(colptr[j] - 1, colptr[j + 1] - 1):
            A[rowval[k] - 1, j] = nz[k]
    return A


def test_turbine_and_anchorline_goldens():
    """test/TestAssemble.jl:9-41 — the toy elements of test/SomeElements.jl restated against adiff2.D2 give the reference's R, ∇R and ∇L (diffed_residual / diffed_lagrangian{2})"""
    m = mb.Model("t")
    n1 = mb.addnode(m, [0., 0., -10.]); n2 = mb.addnode(m, [])
    mb.addelement(m, XM.Turbine, [n1, n2], seadrag=2., sea=lambda t, x: (1., 0.), skydrag=3., sky=lambda t, x: (0., 1.))
    s0 = mb.initialize(m)
    s0.X[0][:] = [1., 2.]
    XM.Turbine.no_second_order = True
    try:
        g, H = xua.packets(m.ele[0], s0.dis.dis[0], 0, 0, 1, s0.Λ[0], s0.X, s0.U, s0.A, 0.)
    finally:
        XM.Turbine.no_second_order = False
    assert np.allclose(g[0, :2], [-2., -3.])                                        # R                      :19
    assert np.allclose(H[0, :2, 2:4], 0.) and np.allclose(H[0, :2, 4:6], [[-2., 0.], [0., -3.]])      # ∇R wrt X, A            :20-22
    m = mb.Model("a")
    n1 = mb.addnode(m, [0., 0., 100.]); n3 = mb.addnode(m, [])
    mb.addelement(m, XM.AnchorLine, [n1, n3], Δxₘtop=[0., 2., 0.], xₘbot=[94., 0.], L=170., buoyancy=-1.)
    s0 = mb.initialize(m)
    g, H = xua.packets(m.ele[0], s0.dis.dis[0], 0, 0, 1, np.ones(3), s0.X, s0.U, s0.A, 0.)
    assert np.allclose(g[0, 0:3], [-12.25628901693551, 0.2607721067433087, 24.51257803387102], rtol=1e-10)      # ∇L[1][1]   :37
    assert np.allclose(g[0, 3:6], [-0.91509745608786, 0.14708204066349, 1.3086506986891027], rtol=1e-10)        # ∇L[2][1]   :38
    assert np.allclose(g[0, 6:8], [-156.06324599170992, 12.517061123678818], rtol=1e-10)                        # ∇L[4][1]   :40


def test_gauge_types_do_not_merge_across_different_gauges():
    """ElementCost{StrainGaugeOnEulerBeam3D} types are keyed by their gauge positions / directions and by the cost: a second addelement with other gauges is another type
    (its gauge matrix is type data), the same gauges and cost extend the first type"""
    import muscade_b200 as mb
    P = np.array([[0., .5, 0.], [0., 0, .5]]).T; D = np.array([[1., 0., 0.], [1., 0., 0.]]).T
    cost = mb.QuadraticGaugeCost(1e-5, lambda t: np.zeros(2))
    m = mb.Model("g")
    nod = mb.addnode(m, np.arange(5)[:, None] * np.array([1., 0., 0.])[None, :])
    kw = dict(req=("ε",), cost=cost, ElementType=mb.StrainGaugeOnEulerBeam3D)
    ek = dict(mat=mb.BeamCrossSection(EA=10., EI2=3., EI3=3., GJ=4., mu=1., iota1=1.), orient2=(0., 1., 0.))
    mb.addelement(m, mb.ElementCost, np.stack([nod[:1], nod[1:2]], axis=1), elementkwargs=dict(P=P, D=D, elementkwargs=ek), **kw)
    mb.addelement(m, mb.ElementCost, np.stack([nod[1:2], nod[2:3]], axis=1), elementkwargs=dict(P=P, D=D, elementkwargs=ek), **kw)
    assert len(m.ele) == 1 and m.ele[0].nele == 2
    mb.addelement(m, mb.ElementCost, np.stack([nod[2:3], nod[3:4]], axis=1), elementkwargs=dict(P=2 * P, D=D, elementkwargs=ek), **kw)
    assert len(m.ele) == 2 and not np.array_equal(m.ele[0].extra["G"], m.ele[1].extra["G"])


def _anchor_wrapped(kind):
    """test/TestElementCost.jl:7-13, 63-68: AnchorLine wrapped in ElementCost (cost = Fh²) or ElementConstraint (gap = Fh², λ on node 1, mode equal)"""
    import muscade_b200 as mb
    m = mb.Model("TestModel")
    n1 = mb.addnode(m, [0., 0., 100.]); n3 = mb.addnode(m, [])
    ek = dict(Δxₘtop=[5., 0, 0], xₘbot=[250., 0], L=290., buoyancy=-5e3)
    sq = lambda eleres, t: eleres.Fh * eleres.Fh
    if kind == "cost":
        mb.addelement(m, mb.ElementCost, [[n1, n3]], req=("Fh",), cost=sq, ElementType=XM.AnchorLine, elementkwargs=ek)
    else:
        mb.addelement(m, mb.ElementConstraint, [[n1, n3]], λinod=1, λfield="λ", req=("Fh",), gap=sq, mode=mb.equal, ElementType=XM.AnchorLine, elementkwargs=ek)
    return m


def test_elementcost_and_elementconstraint_reference_goldens():
    """test/TestElementCost.jl:29-33 and :81-85: dof lists, L and ∇L (order Λ, X, U, A) of an AnchorLine wrapped in ElementCost / ElementConstraint at Λ = 0, X = 1, U = 1, A = 0 —
    through the packets the device assembles (xua.packets: the generic second-order path, adiff2.D2 in the place of the reference's nested duals)"""
    import muscade_b200 as mb
    from muscade_b200 import xua
    gold = {"cost": (1.926851845351649e11, [-438861.1307445675, 9278.602091074139, 1.8715107899328927e6, -2.322235123921358e10, 4.9097753633879846e8, 9.903105530914653e10,
                                           -6.735986859485705e12, 3.853703690703298e11]),
            "constraint": (-1.926851845351649e11, [-438861.1307445675, 9278.602091074139, 1.8715107899328927e6, 2.322235123921358e10, -4.9097753633879846e8, -9.903105530914653e10,
                                                  -1.926851845351649e11, 6.735986859485705e12, -3.853703690703298e11])}
    for kind in ("cost", "constraint"):
        m = _anchor_wrapped(kind)
        et = m.ele[0]
        inod, clas, field = et.inod, et.clas, et.field
        if kind == "cost":
            assert (tuple(inod), tuple(clas), tuple(field)) == ((1, 1, 1, 2, 2), ("X", "X", "X", "A", "A"), ("tx1", "tx2", "rx3", "ΔL", "Δbuoyancy"))
        else:
            assert (tuple(inod), tuple(clas), tuple(field)) == ((1, 1, 1, 2, 2, 1), ("X", "X", "X", "A", "A", "U"), ("tx1", "tx2", "rx3", "ΔL", "Δbuoyancy", "λ"))
        s0 = mb.initialize(m); dis = s0.dis
        nX, nU, nA = m.getndof(("X", "U", "A"))
        assert (nX, nU, nA) == (3, 0 if kind == "cost" else 1, 2)
        Λ, X, U, A = np.zeros(nX), [np.ones(nX)], [np.ones(nU)], np.zeros(nA)
        g, H = xua.packets(et, dis.dis[0], 0, 0, 1, Λ, X, U, A, 0.)
        Lval, grad = gold[kind]
        assert np.allclose(g[0], grad, rtol=1e-10)
        ed = dis.dis[0]
        L = et.ElType.lagrangian(et.eleobj, et.extra, list(Λ[ed.X[0] - 1][:, None] * np.ones(1)), [list(X[0][ed.X[0] - 1][:, None] * np.ones(1))],
                                 [list(U[0][ed.U[0] - 1][:, None] * np.ones(1))] if nU else [[]], list(A[ed.A[0] - 1][:, None] * np.ones(1)), 0., None)
        assert np.allclose(D2.lift(L).v, Lval, rtol=1e-12)
        assert np.abs(H[0] - H[0].T).max() <= 1e-9 * np.abs(H[0]).max()          # equal mode: a true potential, symmetric Hessian


def test_kkt_pseudo_potential_derivatives():
    """KKT(λ,g,γ) (src/BasicElements.jl:289-307): value 0, gradient λ·∇g + (gλ − γ)·∇λ, second derivative = derivative of that expression (not symmetric)"""
    from muscade_b200.adiff2 import D2, KKT
    x = D2.variables(np.array([[0.7, 1.3, -0.4]]))
    lam, g = x[0], x[1] * x[1] + x[2]                  # λ = z₀, g = z₁² + z₂
    k = KKT(lam, g, 0.25)
    gv, lv = 1.3 ** 2 - 0.4, 0.7
    S = gv * lv - 0.25
    assert k.v[0] == 0.
    assert np.allclose(k.g[0], [S, lv * 2 * 1.3, lv])
    # ∂ⱼ of gradient entry i, by hand: grad = (gλ − γ, 2λz₁, λ)
    assert np.allclose(k.H[0], [[gv, lv * 2 * 1.3, lv], [2 * 1.3, 2 * lv, 0.], [1., 0., 0.]])


def test_dofconstraint_reference_goldens():
    """test/TestDofConstraints.jl:27-86 — DofConstraint{:X} residual and ∂R/∂X in equal / off / positive (γ = 0) mode, in contact and with a gap (toolbox.DofConstraint's
    closed forms); :96-157 — the U-class constraint's lagrangian −gap·λ / −λ²/2 / −KKT(λ,gap,γ) with γ = 1 and γ = 0 through adiff2.D2 and adiff2.KKT:
    ∇L and its (non-symmetric) derivative"""
    import muscade_b200 as mb
    from muscade_b200.adiff2 import KKT
    gap = lambda x, t: (.3 * x[:, 0] + .4 * x[:, 1], np.array([.3, .4]))
    ctc, gp = np.array([[4., -3., 10.]]), np.array([[4., 3., 10.]])
    E = np.array([[0, 0, -.3], [0, 0, -.4], [-.3, -.4, 0]])
    P0 = lambda gv: np.array([[0, 0, -.3], [0, 0, -.4], [-3., -4., -gv]])
    for mode, X, R, K in (("equal", ctc, [-3, -4, 0], E), ("equal", gp, [-3, -4, -2.4], E), ("off", ctc, [0, 0, -10], np.diag([0, 0, -1.])),
                          ("positive", ctc, [-3, -4, 0], P0(0.)), ("positive", gp, [-3, -4, -24.], P0(2.4))):
        r, k0, _, _ = mb.DofConstraint.residual(dict(gap=gap, gargs=(), mode=mode, Nx=2), [X], 0.)
        assert np.allclose(r[0], R) and np.allclose(k0[0], K), mode
    for γ, U, R, H in ((1., ctc, [-3., -4., 1.], P0(0.)), (1., gp, [-3., -4., -23.], P0(2.4)), (0., ctc, [-3, -4, 0], P0(0.)), (0., gp, [-3., -4., -24.], P0(2.4))):
        u = D2.variables(U)
        L = -KKT(u[2], .3 * u[0] + .4 * u[1], γ)
        assert np.allclose(L.g[0], R) and np.allclose(L.H[0], H), γ
    u = D2.variables(ctc)
    L = -(.3 * u[0] + .4 * u[1]) * u[2]                                   # equal
    assert np.allclose(L.g[0], [-3, -4, 0]) and np.allclose(L.H[0], E)
    L = -0.5 * (u[2] * u[2])                                             # off
    assert np.allclose(L.g[0], [0, 0, -10]) and np.allclose(L.H[0], np.diag([0, 0, -1.]))
