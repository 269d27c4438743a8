"""Host logic (no GPU): model description, dof numbering, Disassembler — against the reference's own integer pins
(test/TestModelDescription.jl, test/TestAssemble.jl:53-67) and against the closed-form chain numbering used by the bench."""
import numpy as np
import pytest


class _Gen:
    """Stand-in element type defined only by its doflist (enough for every integer structure)."""
    kind = "host"
    LIST = None
    NAME = None

    @classmethod
    def doflist(cls, **kw): return cls.LIST
    @classmethod
    def typekey(cls, **kw): return (cls.NAME,)
    @classmethod
    def construct(cls, coords, **kw): return np.zeros((coords[0].shape[0], 0))


class Turbine(_Gen):      # test/SomeElements.jl:54-56
    NAME = "Turbine"; LIST = ((1, 1, 2, 2), ("X", "X", "A", "A"), ("tx1", "tx2", "Δseadrag", "Δskydrag"))


class AnchorLine(_Gen):   # test/SomeElements.jl:74-76
    NAME = "AnchorLine"; LIST = ((1, 1, 1, 2, 2), ("X", "X", "X", "A", "A"), ("tx1", "tx2", "rx3", "ΔL", "Δbuoyancy"))


def test_disassembler_pins(mb):
    """test/TestAssemble.jl:41-67"""
    model = mb.Model("TestModel")
    n1 = mb.addnode(model, [0, 0, 100.]); n2 = mb.addnode(model, []); n3 = mb.addnode(model, [])
    e1 = mb.addelement(model, Turbine, [n1, n2]); e2 = mb.addelement(model, AnchorLine, [n1, n3])
    assert e1 == (1, 1) and e2 == (2, 1)
    dis = mb.Disassembler(model)
    assert dis.dis[0].X.tolist() == [[1, 2]] and dis.dis[0].U.shape == (1, 0) and dis.dis[0].A.tolist() == [[1, 2]]
    assert dis.dis[1].X.tolist() == [[1, 2, 3]] and dis.dis[1].A.tolist() == [[3, 4]]
    assert np.allclose(dis.scaleX, 1) and np.allclose(dis.scaleΛ, 1) and dis.scaleU.size == 0 and np.allclose(dis.scaleA, [1, 1, 1, 1])
    assert dis.fieldX == ["tx1", "tx2", "rx3"] and dis.fieldU == [] and dis.fieldA == ["Δseadrag", "Δskydrag", "ΔL", "Δbuoyancy"]


def test_model_description_pins(mb):
    """test/TestModelDescription.jl: node/element/dof numbering after addelement!"""
    model = mb.Model("TestModel")
    n1 = mb.addnode(model, [0, 0, 100.]); n2 = mb.addnode(model, []); n3 = mb.addnode(model, [])
    mb.addelement(model, Turbine, [n1, n2]); mb.addelement(model, AnchorLine, [n1, n3])
    assert (n1, n2, n3) == (1, 2, 3)
    assert model.getndof() == 7 and model.getndof(("X", "U", "A")) == (3, 0, 4)
    # model.ele[·][1].dofID  (TestModelDescription.jl:47-58): Turbine X1 X2 A1 A2 ; AnchorLine X1 X2 X3 A3 A4
    assert model.ele[0].dofID.tolist() == [[1, 2, 1, 2]] and model.ele[0].clas == ("X", "X", "A", "A")
    assert model.ele[1].dofID.tolist() == [[1, 2, 3, 3, 4]] and model.ele[1].nodID.tolist() == [[1, 3]]
    # model.dof[DofID].nodID / idoftyp  (:60-80)
    nodX = np.concatenate(model.dof_nod["X"]); typX = np.concatenate(model.dof_typ["X"])
    nodA = np.concatenate(model.dof_nod["A"]); typA = np.concatenate(model.dof_typ["A"])
    assert nodX.tolist() == [1, 1, 1] and typX.tolist() == [1, 2, 5]
    assert nodA.tolist() == [2, 2, 3, 3] and typA.tolist() == [3, 4, 6, 7]
    # model.doftyp[·].dofID (:83-99)
    assert np.concatenate(model.doftyp[0].dofID).tolist() == [1] and np.concatenate(model.doftyp[2].dofID).tolist() == [1]
    assert np.concatenate(model.doftyp[4].dofID).tolist() == [3]
    assert model.getnele() == 2 and model.getneletyp() == 2
    assert [(d.clas, d.field) for d in model.doftyp] == [("X", "tx1"), ("X", "tx2"), ("A", "Δseadrag"), ("A", "Δskydrag"), ("X", "rx3"), ("A", "ΔL"), ("A", "Δbuoyancy")]
    mb.setscale(model, scale=dict(X=dict(tx1=10., rx3=2.)), Λscale=1e3)
    dis = mb.Disassembler(model)
    assert dis.scaleX.tolist() == [10., 1., 2.] and dis.scaleΛ.tolist() == [1e4, 1e3, 2e3]
    st = mb.initialize(model)
    with pytest.raises(mb.MuscadeB200Error):
        mb.addnode(model, [0.])        # model is initialized and can no longer be edited
    assert st.X[0].shape == (3,) and st.A.shape == (4,) and st.time == -np.inf


def test_chain_numbering_matches_generic_addelement(mb):
    """the bench's closed-form chain (synthetic.chain) is what addnode!/addelement! produce"""
    N = 37
    eleobj, idx, ndof = mb.synthetic.chain(N)
    model = mb.Model()
    k = np.arange(N + 1)[:, None] * np.array([0.8, 0.6, 0.0])[None, :]
    nod = mb.addnode(model, k)
    mesh = np.stack([nod[:-1], nod[1:]], axis=1)
    mat = mb.BeamCrossSection(EA=10., EI2=3., EI3=3., GJ=4., mu=1., iota1=1.)
    ityp, iele = mb.addelement(model, mb.EulerBeam3D, mesh, mat=mat, orient2=(0., 1., 0.))
    dis = mb.Disassembler(model)
    assert ityp == 1 and iele.tolist() == list(range(1, N + 1))
    assert model.getndof("X") == ndof and np.array_equal(dis.dis[0].X, idx)
    assert np.array_equal(model.ele[0].eleobj, eleobj)
    assert dis.fieldX[:6] == ["t1", "t2", "t3", "r1", "r2", "r3"]


def test_shared_dofs_and_multiple_types(mb):
    """StaticBeamAnalysis-like model: beams + 6 Hold types + DofLoad ⇒ 8 element types, λ dofs appended after beam dofs"""
    nel = 8
    model = mb.Model()
    th = 3 * np.pi / 2 + np.arange(nel + 1) / nel * np.pi / 4
    nod = mb.addnode(model, np.stack([100 * np.cos(th), np.zeros(nel + 1), 100 + 100 * np.sin(th)], axis=1))
    mat = mb.BeamCrossSection(EA=1e9, EI2=833.33e3, EI3=833.33e3, GJ=705e3, mu=1., iota1=1.)
    mb.addelement(model, mb.EulerBeam3D, np.stack([nod[:-1], nod[1:]], axis=1), mat=mat)
    for f in ["t1", "t2", "t3", "r1", "r2", "r3"]:
        mb.addelement(model, mb.Hold, [nod[0]], field=f)
    mb.addelement(model, mb.DofLoad, [nod[-1]], field="t2", value=lambda t: 300. * t)
    assert model.getneletyp() == 8 and model.getndof("X") == 6 * (nel + 1) + 6
    dis = mb.Disassembler(model)
    assert dis.dis[1].X.tolist() == [[1, 55]] and dis.dis[6].X.tolist() == [[6, 60]] and dis.dis[7].X.tolist() == [[50]]
    assert dis.fieldX[54:] == ["λt1", "λt2", "λt3", "λr1", "λr2", "λr3"]


def test_dofconstraint_hessian_matches_finite_differences():
    """host-evaluated DofConstraint{:X} on DirectXUA's second-order branch (L = Λ∘R, src/DirectXUA.jl:152-171): the Λ-contracted second derivative handed to
    mb_direct_set_host_xx equals the derivative of (∂R/∂X)ᵀΛ, for every mode and a quadratic two-dof gap"""
    import muscade_b200 as mb
    A = np.array([[0.4, 0.1], [0.1, -0.3]]); bvec = np.array([0.3, -0.2])

    def gap(x, t):
        g = 0.5 * np.einsum("ei,ij,ej->e", x, A, x) + x @ bvec + 0.1 - 0.05 * t
        return g, x @ A + bvec, np.broadcast_to(A, (x.shape[0], 2, 2))
    X = [np.array([[0.3, -0.4, 0.7], [0.1, 0.2, -0.6]])]; Lam = np.array([[0.2, -0.5, 0.9], [-0.3, 0.4, 0.1]])
    for mode in ("equal", "positive", "off"):
        ex = dict(gap=gap, gargs=(), mode=mode, Nx=2)
        H = mb.DofConstraint.hessian(ex, X, Lam, 0.2)
        h = 1e-6
        Hfd = np.zeros((2, 3, 3))
        for j in range(3):
            Xp = [X[0].copy()]; Xp[0][:, j] += h; Xm = [X[0].copy()]; Xm[0][:, j] -= h
            Kp = mb.DofConstraint.residual(ex, Xp, 0.2)[1]; Km = mb.DofConstraint.residual(ex, Xm, 0.2)[1]
            Hfd[:, :, j] = (np.einsum("eki,ek->ei", Kp, Lam) - np.einsum("eki,ek->ei", Km, Lam)) / (2 * h)
        if H is None:
            assert np.abs(Hfd).max() < 1e-9 and mode == "off"
        else:
            assert np.abs(H - Hfd).max() < 1e-8 and np.abs(H - H.transpose(0, 2, 1)).max() == 0.
    lin = dict(gap=lambda x, t: (x[:, 0] - t, np.ones_like(x)), gargs=(), mode="equal", Nx=1)
    assert mb.DofConstraint.hessian(lin, [np.array([[0.3, 0.7]])], np.array([[0.2, -0.5]]), 0.) is None      # affine gap, equal mode: R linear in (x,λ)


def test_host_element_arguments_are_kept_per_element():
    """the reference stores args / costargs / gargs in every element object (src/BasicElements.jl:198-208,275-284): two addelement! calls with the same
    closure and different arguments make ONE element type whose elements differ — the later arguments must not be dropped"""
    import muscade_b200 as mb
    model = mb.Model()
    nod = mb.addnode(model, np.array([[0., 0, 0], [1., 0, 0], [2., 0, 0]]))
    cost = lambda x, t, k: 0.5 * k * x * x
    t1, _ = mb.addelement(model, mb.SingleDofCost, [nod[0]], clas="X", field="t1", cost=cost, costargs=(1.,))
    t2, _ = mb.addelement(model, mb.SingleDofCost, [nod[1]], clas="X", field="t1", cost=cost, costargs=(5.,))
    assert t1 == t2 and model.ele[t1 - 1].nele == 2
    c, c1, c2 = model.ele[t1 - 1].cost_derivs(np.array([2., 2.]), 0.)
    assert c.tolist() == [2., 10.] and c1.tolist() == [2., 10.] and c2.tolist() == [1., 5.]
    load = lambda t, a: a * t
    l1, _ = mb.addelement(model, mb.DofLoad, [nod[0]], field="t2", value=load, args=(1.,))
    l2, _ = mb.addelement(model, mb.DofLoad, np.array([[nod[1]], [nod[2]]]), field="t2", value=load, args=(-3.,))
    assert l1 == l2
    R, K0, K1, K2 = model.ele[l1 - 1].residual([np.zeros((3, 1))], 2.)
    assert R[:, 0].tolist() == [-2., 6., 6.] and K0.shape == (3, 1, 1) and K1 is None


def test_vectorised_constructors_equal_the_oracle_and_the_reference_golden():
    """toolbox.eulerbeam3d_structs / bar3d_structs (the host side of mb_add_eulerbeam3d / mb_add_bar3d: 69 / 38 doubles per element, the reference's struct
    layout) against the oracle's literal constructors — bit for bit over random elements — and against test/TestBeamElement.jl:15-31"""
    import muscade_b200 as mb
    from oracle import elements as OE
    rng = np.random.default_rng(12)
    n = 50
    c1 = rng.uniform(-5, 5, (n, 3)); c2 = c1 + rng.uniform(0.3, 3., (n, 3)) * rng.choice([-1., 1.], (n, 3))
    o2 = rng.uniform(-1, 1, (n, 3))
    mat = mb.BeamCrossSection(EA=10., EI2=3., EI3=2.5, GJ=4., mu=1., iota1=1.3, w=.7, Ca2=1., Cq3=2.)
    for e in range(n):
        got = mb.toolbox.eulerbeam3d_structs(c1[e], c2[e], mat, orient2=o2[e])[0]
        ref = OE.beam_ctor(c1[e], c2[e], np.asarray(mat, float), orient2=o2[e])
        assert np.array_equal(got, ref), e
    got = mb.toolbox.eulerbeam3d_structs(c1, c2, mat)                     # many at once = one at a time
    assert np.array_equal(got, np.stack([OE.beam_ctor(c1[e], c2[e], np.asarray(mat, float)) for e in range(n)]))
    bmat = mb.AxisymmetricBarCrossSection(EA=10., mu=1.5, w=.3, Cat=2., Clt=.2, Cqt=4., Can=1., Cln=.1, Cqn=3.)
    gotb = mb.toolbox.bar3d_structs(c1, c2, bmat)
    assert np.array_equal(gotb, np.stack([OE.bar_ctor(c1[e], c2[e], np.asarray(bmat, float)) for e in range(n)]))
    # the reference's own constructor golden (test/TestBeamElement.jl:8-31), through the product's constructor
    b = mb.toolbox.eulerbeam3d_structs([0, 0, 0], [4, 3, 0], mb.BeamCrossSection(EA=10., EI2=3., EI3=3., GJ=4., mu=1., iota1=1.))[0]
    L = 5.
    assert np.allclose(b[0:3], [2., 1.5, 0.]) and np.allclose(b[3:12].reshape(3, 3).T, [[.8, -.6, 0.], [.6, .8, 0.], [0., 0., 1.]])
    assert np.allclose(b[12:16], [-0.4305681557970263, -0.16999052179242816, 0.16999052179242816, 0.4305681557970263]) and np.allclose(b[16:18], [-.5, .5])
    assert np.allclose(b[18:21], [4., 3., 0.]) and np.allclose(b[21:24], [5., 0., 0.])
    assert np.allclose(b[24:28], [-0.8611363115940526, -0.3399810435848563, 0.3399810435848563, 0.8611363115940526])
    assert np.allclose(b[28:32], [-0.972414176921822, -0.49032285223640754, 0.49032285223640754, 0.972414176921822])
    assert np.allclose(b[32:36], np.array([-0.0646110632135477, -0.221103222500738, -0.221103222500738, -0.0646110632135477]) * L)
    assert np.allclose(b[36:40], 2. / L) and np.allclose(b[40:44], np.array([10.333635739128631, 4.079772523018276, -4.079772523018276, -10.333635739128631]) / L ** 2)
    assert np.allclose(b[44:48], 2. / L) and b[48] == L
    assert np.allclose(b[49:53], np.array([0.17392742256872692, 0.3260725774312731, 0.3260725774312731, 0.17392742256872692]) * L)
