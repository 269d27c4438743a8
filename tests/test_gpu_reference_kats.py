"""End-to-end known answers HELD BY THE REFERENCE'S OWN TESTS, run through the product (host mirror of the SweepX driver + device scatter / state
update through the C ABI): the converged states of test/TestSweepX1.jl:23-26 and test/TestSweepX2.jl:20-24 (SomeElements.jl toy oscillators).
These pin the Newmark-β driver (src/SweepX.jl:3-15, 98-132, 179-226), the :step / :iter missions of addin! (src/SweepX.jl:45-96) for host-evaluated
element types, the device reduction into Lλ / Lλx and mb_sweepx_newmark_decrement to numbers produced by Muscade itself — not by the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def oscillator_types(mb):
    class SdofOscillator(mb.toolbox.ElementType):
        """test/SomeElements.jl:193-207: R = −u + K₁x + K₂x² + C₁x′ + C₂x′² + M₁x″ + M₂x″²; dofs (X tx1, U tu1)"""
        takes_UA = True

        @classmethod
        def doflist(cls, **kw):
            return (1, 1), ("X", "U"), ("tx1", "tu1")

        @classmethod
        def typekey(cls, **kw):
            return ("SdofOscillator",)

        @classmethod
        def construct(cls, coords, K1=0., K2=0., C1=0., C2=0., M1=0., M2=0.):
            return np.zeros((coords[0].shape[0], 0)), dict(p=(K1, K2, C1, C2, M1, M2))

        @staticmethod
        def residual(extra, X, t, U=None, A=None):
            K1, K2, C1, C2, M1, M2 = extra["p"]
            x = X[0][:, 0]
            u = U[:, 0] if U is not None else 0.
            x1 = X[1][:, 0] if len(X) > 1 else np.zeros_like(x)
            x2 = X[2][:, 0] if len(X) > 2 else np.zeros_like(x)
            R = (-u + K1 * x + K2 * x ** 2 + C1 * x1 + C2 * x1 ** 2 + M1 * x2 + M2 * x2 ** 2)[:, None]
            k = lambda a: np.asarray(a, float)[:, None, None]
            return R, k(K1 + 2 * K2 * x), k(C1 + 2 * C2 * x1), k(M1 + 2 * M2 * x2)
    return SdofOscillator


@pytest.mark.parametrize("device_state", [False, True])
def test_sweepx2_sdof_oscillator_reference_states(mb, device_state):
    """test/TestSweepX2.jl:8-24"""
    Osc = oscillator_types(mb)
    model = mb.Model()
    node = mb.addnode(model, np.zeros((1, 0)))
    mb.addelement(model, Osc, [node[0]], K1=1., K2=.3, C1=1., C2=2., M1=3.)
    st = mb.initialize(model, time=0.).with_orders(1, 3, 1)
    st.X[1][0] = 1.                                                  # setdof!(initialstate,[1.];field=:tx1,order=1)
    T = 0.4 * np.arange(1, 101)
    states = mb.sweepx.solve(2, st, T, device_state=device_state)
    X = np.array([s.X[0][0] for s in states]); X1 = np.array([s.X[1][0] for s in states]); X2 = np.array([s.X[2][0] for s in states])
    ref0 = [0.3653456491624315, 0.039495394592936224, -0.9800856952523974, 0.0204208015740059, 0.10202734361080085, -0.08876358375431506,
            0.020279004568508258, 0.009410960025011333, -0.01164572216979256, 0.0044937076208295505]
    ref1 = [0.8267282458121575, -0.5543584424772612, 0.1761431693326722, 0.17142199935793515, -0.07990836783415763, 0.009615311853025504,
            0.01797222270368847, -0.012951778973228531, 0.0032459856194525815, 0.0015257230558188115]
    ref2 = [-0.8663587709392137, -0.033410494488486674, 0.1512397675677002, -0.08357963580025408, -0.012670847567464316, 0.025533223926927903,
            -0.0130068667401185, 0.0010595839803349055, 0.0027793256172318247, -0.002010048120257385]
    # Julia's ≈ is rtol = √eps on the vector norm; the Newton iterations stop at |Δx| ≤ 1e-5, so agreement is at the level of that tolerance
    for got, ref in ((X[::10], ref0), (X1[::10], ref1), (X2[::10], ref2)):
        assert np.linalg.norm(got - ref) <= 1e-6 * np.linalg.norm(ref), (got, ref)


def test_sweepx1_exponential_decay_reference_states(mb):
    """test/TestSweepX1.jl:9-26 — AdjustableSdofOscillator with A = 0 (C·10⁰): R = K x + C x′, SweepX{1}, 20 steps of 0.1"""
    Osc = oscillator_types(mb)
    model = mb.Model()
    node = mb.addnode(model, np.zeros((1, 0)))
    mb.addelement(model, Osc, [node[0]], K1=1., C1=.3)
    st = mb.initialize(model, time=0.).with_orders(1, 2, 1)
    st.X[0][0] = 1.
    t = np.arange(1, 21) * 0.1
    states = mb.sweepx.solve(1, st, t, maxiter=5)
    x = np.array([s.X[0][0] for s in states]); x1 = np.array([s.X[1][0] for s in states])
    rx = [0.857143, 0.612245, 0.437318, 0.31237, 0.223121, 0.159372, 0.113837, 0.0813124, 0.0580803, 0.0414859, 0.0296328, 0.0211663, 0.0151188,
          0.0107991, 0.00771366, 0.00550976, 0.00393554, 0.0028111, 0.00200793, 0.00143424]
    rx1 = [-2.85714, -2.04082, -1.45773, -1.04123, -0.743738, -0.531241, -0.379458, -0.271041, -0.193601, -0.138286, -0.098776, -0.0705543,
           -0.0503959, -0.0359971, -0.0257122, -0.0183659, -0.0131185, -0.00937034, -0.0066931, -0.00478079]
    assert np.linalg.norm(x - rx) <= 1e-5 * np.linalg.norm(rx) and np.linalg.norm(x1 - rx1) <= 1e-5 * np.linalg.norm(rx1)      # rtol of the reference test


def test_assemble_turbine_anchorline_reference_values(mb):
    """test/TestAssemble.jl:43-51,152-158 — assemble!{:iter}(out::AssemblySweepX{0},…) of a Turbine and an AnchorLine (toy elements of test/SomeElements.jl, written against
    adiff2.D2, with A-dofs as plain values; the AnchorLine is given by its `lagrangian` only) at the zero state: the device scatter gives the reference's Lλ and Lλx"""
    import xua_models as XM
    m = mb.Model("TestModel")
    n1 = mb.addnode(m, [0., 0., 100.]); n2 = mb.addnode(m, []); n3 = mb.addnode(m, [])
    mb.addelement(m, XM.Turbine, [n1, n2], seadrag=2., sea=lambda t, x: (1., 0.), skydrag=3., sky=lambda t, x: (0., 1.))
    mb.addelement(m, XM.AnchorLine, [n1, n3], Δxₘtop=[5., 0., 0.], xₘbot=[150., 0.], L=180., buoyancy=-1e3)
    st = mb.initialize(m)
    out, asm, dofgr = mb.sweepx.prepare(0, m, st.dis)
    try:
        assert asm[1, 1].tolist() == [[1], [2]] and asm[1, 2].tolist() == [[1], [2], [3]]                       # :121-122
        assert asm[2, 1].tolist() == [[1], [2], [4], [5]] and asm[2, 2].ravel().tolist() == list(range(1, 10))  # :123-124
        out.c = mb.sweepx.newmark_coefficients(0, 0.)
        mb.sweepx.assemble("iter", out, asm, st.dis, m, st, 0.)
        assert np.allclose(out.Lλ, [-152130.71199858442, -3.0, 0.0], rtol=1e-10, atol=1e-9)                    # :156
        K = np.array([[10323.069597975566, 0., 0.], [0., 1049.1635310247202, 5245.8176551236], [0., 5245.817655123601, 786872.6482685402]])
        assert np.allclose(out.Lλx.toarray(), K, rtol=1e-10, atol=1e-7)                                         # :157
    finally:
        out.engine.close()


def test_sweepx0_turbine_moorings_reference_state(mb):
    """test/TestSweepX0.jl:9-29 — solve(SweepX{0};time=[0.,1.]) of a turbine on three anchor lines: the converged X of step 1 is the reference's"""
    import xua_models as XM
    m = XM.model_testsweepx0()
    st = mb.initialize(m)
    states = mb.sweepx.solve(0, st, [0., 1.])
    s = states[0]
    assert np.allclose(s.X[0], [-5.332268523655259, 21.09778288272267, 0.011304253608808651], rtol=1e-7)       # :24 (Julia's ≈: rtol = √eps)
    assert not s.Λ[0].any() and not s.A.any() and s.time == 0. and m.locked


def test_sweepx0_with_wrapped_anchor_lines_same_state(mb):
    """the same analysis with two of the anchor lines wrapped in ElementCost / ElementConstraint: in an X-analysis R = ∂L/∂Λ is the target's residual (the cost and the
    multiplier do not enter), so the converged state is the reference's of test/TestSweepX0.jl:24"""
    import xua_models as XM
    m = XM.model_mooring_wrapped()
    st = mb.initialize(m)
    s = mb.sweepx.solve(0, st, [0., 1.])[0]
    assert np.allclose(s.X[0], [-5.332268523655259, 21.09778288272267, 0.011304253608808651], rtol=1e-7)
