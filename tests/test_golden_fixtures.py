"""Committed fixtures (tests/golden/*.npz, written by tests/golden/make_fixtures.py from the oracle):
  not gpu: the oracle reproduces them (the checker did not drift);
  gpu    : the CUDA engine, through the C ABI, against the committed vectors — no oracle call on this path."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-12
CASES = [(0, "iter"), (1, "iter"), (1, "step"), (2, "iter"), (2, "step")]


def rel(a, b, floor=0.):
    return np.abs(a - b).max() / max(np.abs(b).max(), floor, 1e-300)


def test_oracle_reproduces_fixtures():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_fixtures", os.path.join(GOLD, "make_fixtures.py"))
    mf = importlib.util.module_from_spec(spec); spec.loader.exec_module(mf)
    for name, fresh in (("sweepx_chain8.npz", mf.sweepx_chain()), ("directxua_chain4x6.npz", mf.direct_chain())):
        gold = np.load(os.path.join(GOLD, name))
        assert set(gold.files) == set(fresh)
        for k in gold.files:      # integers exactly; floats to 1e-13 (libm may pick another sincos variant on another CPU)
            if gold[k].dtype.kind == "i": assert np.array_equal(gold[k], fresh[k]), (name, k)
            else: assert rel(fresh[k], gold[k], 1e-3) <= 1e-13, (name, k)


@pytest.mark.gpu
@pytest.mark.parametrize("OX,mission", CASES)
@pytest.mark.parametrize("zero", [False, True])
def test_gpu_sweepx_against_committed_vectors(mb, engine_factory, OX, mission, zero):
    gold = np.load(os.path.join(GOLD, "sweepx_chain8.npz"))
    eleobj, idx, ndof = mb.synthetic.chain(8, dynamic=OX > 0)
    eng = engine_factory()
    eng.add_eulerbeam3d(eleobj, idx, np.ones(12))
    assert eng.sweepx_prepare(ndof) == len(gold["rowval"]) == 900          # nnz = 108N+36 for N=8
    cp, rv = eng.sweepx_pattern()
    assert np.array_equal(cp, gold["colptr"]) and np.array_equal(rv, gold["rowval"])
    X = mb.synthetic.state(ndof, nder=OX + 1, zero=zero)
    L, nz = eng.sweepx_assemble(OX, mission, X, mb.synthetic.newmark_coefficients(OX, 0.3))
    key = "OX%d_%s_%s" % (OX, mission, "zero" if zero else "rand")
    assert rel(nz, gold[key + "_nz"]) <= TOL
    assert rel(L, gold[key + "_L"], np.abs(gold[key + "_nz"]).max()) <= TOL


@pytest.mark.gpu
def test_gpu_directxua_against_committed_vectors(mb):
    gold = np.load(os.path.join(GOLD, "directxua_chain4x6.npz"))
    N, nstep, dt = 4, 6, 0.1
    model = mb.Model()
    nod = mb.addnode(model, np.arange(N + 1)[:, None] * np.array([.8, .6, 0.])[None, :])
    unod = mb.addnode(model, np.zeros((N, 0)))
    mat = mb.BeamCrossSection(EA=10., EI2=3., EI3=2.5, GJ=4., mu=1., iota1=1.2, w=.3, Ca2=2., Ca3=1.5, Cq2=1., Cq3=.7, Cl1=.2)
    mb.addelement(model, mb.EulerBeam3D, np.stack([nod[:-1], nod[1:], unod], axis=1), mat=mat, Udof=True)
    st0 = mb.initialize(model)
    nX, nU = model.getndof("X"), model.getndof("U")
    eng = mb.directxua.prepare(2, 0, model, st0.dis, nstep, dt)
    cp, rv = eng.big_pattern()
    assert np.array_equal(cp, gold["colptr"]) and np.array_equal(rv, gold["rowval"])
    for s in range(nstep):
        eng.set_state(s, [mb.synthetic.uniform_pm1(10 + 3 * s + d, nX) * (0.1 if d == 0 else 0.3) for d in range(3)], mb.synthetic.uniform_pm1(99 + s, nU))
    Lvv = np.zeros(eng.nnzbig); Lv = np.zeros(eng.ncol)
    eng.direct_assemble(Lvv=Lvv, Lv=Lv)
    assert rel(Lvv, gold["nzval"]) <= TOL
    assert rel(Lv, gold["Lv"], np.abs(gold["nzval"]).max()) <= TOL
    eng.close()
