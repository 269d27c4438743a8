"""EigX over the device assembly (muscade_b200.eigx): the reference's own goldens (test/TestEigX.jl) and beam models against SweepX / the oracle / beam theory"""
import numpy as np
import pytest

import muscade_b200 as mb
from muscade_b200 import eigx
from oracle import elements as OE
from oracle import pattern as OP

import xua_models as XM

pytestmark = pytest.mark.gpu


def test_reference_eigx_real_goldens(mb):
    """test/TestEigX.jl:26-39: ω of the five lowest modes and the snapshot of modes 1, 2 with amplitudes 1, 4+5i (general form: Spring{1} and SdofOscillator through adiff2.D2)"""
    m = XM.model_testeigx(); s0 = mb.initialize(m)
    inc = eigx.solve(s0, nmod=5)
    assert np.allclose(inc.ω, [0.143146493914704, 0.42677293340604844, 0.7024494006334382, 0.9650405646123343, 1.209654848799371], rtol=1e-8)
    st = eigx.increment(s0, inc, [1, 2], [1, 4 + 5j])
    assert np.allclose(st.X[0][0::2], [1.7338129265935287, 4.177182755457795, 4.192402275285353, 1.9005876862087123, -1.1387295193620908, -2.9254875682906434], rtol=1e-6)
    assert np.allclose(st.X[1][0::2], [-0.8521204984886861, -2.015355547777684, -1.8990540291889357, -0.5770550457522317, 1.1113125586121952, 2.0941148057104684], rtol=1e-6)
    assert np.allclose(st.X[2][0::2], [-0.2937262712953869, -0.6962620392508903, -0.6613336756803813, -0.2137967604399907, 0.36006958839215825, 0.694478296097603], rtol=1e-6)


def test_reference_eigx_complex_goldens(mb):
    """test/TestEigX.jl:41-53: |p| of the five lowest complex modes and the snapshot (atol 1e-6 as the reference's test)"""
    m = XM.model_testeigx(); s0 = mb.initialize(m)
    inc = eigx.solve(s0, nmod=5, complex_modes=True)
    gold = np.abs([0.10517722364770647 - 4.762644452370236e-17j, 0.1948227763522931 + 9.598886497048079e-17j, 0.14999999999999958 - 0.3995436605528897j,
                   0.1499999999999998 + 0.39954366055289037j, 0.14999999999999933 + 0.6862471569706331j])
    assert np.allclose(np.sort(np.abs(inc.p)), np.sort(gold), rtol=1e-7)
    i1 = int(np.argmin(np.abs(inc.p - gold[0]))) + 1; i2 = int(np.argmin(np.abs(inc.p - gold[1]))) + 1
    st = eigx.increment(s0, inc, [i1, i2], [1, 4 + 5j])
    assert np.allclose(st.X[0][0::2], [0.6824241336467091, 1.9966601248302318, 3.162812968667382, 4.094394443525039, 4.722313413114684, 5.0], atol=1e-6)
    assert np.allclose(st.X[1][0::2], [0.12071650663891864, 0.35319652886061903, 0.5594815803033334, 0.7242724423928726, 0.8353473307560867, 0.8844683290568793], atol=1e-6)
    assert np.allclose(st.X[2][0::2], [0.02223145453650511, 0.06504555832867998, 0.10303553062324605, 0.13338382896767445, 0.15383965890798573, 0.1628859051167022], atol=1e-6)


def beam_model(n, coords=None, mat=None, holds=True):
    m = mb.Model("beam")
    nod = mb.addnode(m, coords if coords is not None else np.stack([np.linspace(0., 10., n + 1), np.zeros(n + 1), np.zeros(n + 1)], axis=1))
    mb.addelement(m, mb.EulerBeam3D, np.stack([nod[:-1], nod[1:]], axis=1), mat=mb.BeamCrossSection(**(mat or dict(EA=1e6, EI2=1e3, EI3=1e3, GJ=1e3, mu=2., iota1=0.01))),
                  orient2=(0., 1., 0.))
    if holds:
        for f in ("t1", "t2", "t3", "r1", "r2", "r3"):
            mb.addelement(m, mb.Hold, [nod[0]], field=f)
    return m


def test_beam_matrices_against_sweepx_and_oracle(mb):
    """K, C, M of a beam chain at a deformed, moving state (specialised path): K = the SweepX{0} tangent Lλx at the same X₀ up to the static/dynamic terms — checked on the
    oracle's DirectXUA first-order blocks (src/DirectXUA.jl:85-120) for all three derivative orders, ≤ 1e-12"""
    rng = np.random.default_rng(3)
    n = 9
    coords = np.cumsum(rng.uniform(0.5, 1.5, (n + 1, 3)), axis=0)
    m = beam_model(n, coords, dict(EA=1e3, EI2=30., EI3=20., GJ=40., mu=1.5, iota1=0.7, Ca2=0.1, Cq2=0.3), holds=False)
    mb.setscale(m, scale=dict(X=dict(t1=2., t2=2., t3=2., r1=0.5, r2=0.5, r3=0.5)))
    s0 = mb.initialize(m).with_orders(1, 3, 1)
    nX = m.getndof("X")
    s0.X = [rng.normal(0, 0.05, nX), rng.normal(0, 0.1, nX), rng.normal(0, 0.1, nX)]
    s0.time = 0.
    K, C, M, L1 = eigx.assemble_matrices(s0)
    dis = s0.dis; ed = dis.dis[0]
    P = OP.prepare_direct(XM.dis_lists(dis), nX, 0, 0, 2, 0, 0)
    b = OE.direct_assemble_step_beams(m.ele[0].eleobj, ed.X, None, 2, 0, s0.X, s0.U, ed.scaleX, None, P, 0)
    ref = np.abs(b["L2"][(1, 2)][0]).max()
    for d, A in enumerate((K, C, M)):
        assert np.abs(A.data - b["L2"][(1, 2)][d]).max() <= 1e-12 * ref
    assert np.abs(L1 - b["L1"][1]).max() <= 1e-12 * np.abs(b["L1"][1]).max()
    assert np.abs(C.data).max() > 0 and np.abs(M.data).max() > 0


def test_cantilever_frequencies_against_beam_theory(mb):
    """clamped-free Euler beam, 40 elements, at rest: the lowest bending frequencies (two planes, equal EI) ω = (βL)²·√(EI/(μL⁴)), βL = 1.8751, 4.6941"""
    n, L, EI, mu = 40, 10., 1e3, 2.
    s0 = mb.initialize(beam_model(n))
    inc = eigx.solve(s0, nmod=4)
    w = np.sort(inc.ω)
    th = np.array([1.8751 ** 2, 1.8751 ** 2, 4.6941 ** 2, 4.6941 ** 2]) * np.sqrt(EI / (mu * L ** 4))
    assert np.allclose(w, th, rtol=2e-2)
    st = eigx.increment(s0, inc, [1], [0.1])
    assert abs(st.X[0]).max() == pytest.approx(0.1, rel=1e-9)          # normalised mode shapes, unit scales
