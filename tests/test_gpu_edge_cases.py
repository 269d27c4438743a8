"""Edge cases of the assembly path on the GPU (through the C ABI): empty element types, one element, free dofs (empty CSC columns), nodes shared by
many elements (long contributor lists), error behaviour of the ABI, NaN location in DirectXUA.  Integer structures bit-exact, floats ≤ 1e-12."""
import numpy as np
import pytest

from oracle import elements as OE
from oracle import pattern as OP

pytestmark = pytest.mark.gpu
TOL = 1e-12


def rel(a, b, floor=0.):
    return np.abs(a - b).max() / max(np.abs(b).max(), floor, 1e-300)


def oracle_beams(eleobj, idx, ndof, OX, mission, X, scale, nm):
    dis = [dict(X=idx, U=np.zeros((idx.shape[0], 0), np.int64), A=np.zeros((idx.shape[0], 0), np.int64))]
    asm1, asm2, colptr, rowval = OP.prepare_sweepx(dis, ndof, 0, 0)
    L = np.zeros(ndof); nz = np.zeros(len(rowval))
    OE.sweepx_assemble_beams(eleobj, idx, asm1[0].T, asm2[0].T, OX, mission, X, scale, nm, L, nz)
    return colptr, rowval, L, nz


def test_single_element_and_free_dofs(mb, engine_factory):
    """one element in a model whose dof vector is larger than what the element touches: columns without entries, Lλ zero there"""
    eleobj, idx, _ = mb.synthetic.chain(1, dynamic=True)
    idx = idx + 5                                   # dofs 1..5 and 18..25 are free
    ndof = 25
    X = mb.synthetic.state(ndof, nder=3)
    nm = mb.synthetic.newmark_coefficients(2, 0.3)
    eng = engine_factory()
    eng.add_eulerbeam3d(eleobj, idx, np.ones(12))
    assert eng.sweepx_prepare(ndof) == 144
    colptr, rowval, Lref, nzref = oracle_beams(eleobj, idx, ndof, 2, "step", X, np.ones(12), nm)
    cp, rv = eng.sweepx_pattern()
    assert np.array_equal(cp, colptr) and np.array_equal(rv, rowval)
    assert (np.diff(cp)[:5] == 0).all() and (np.diff(cp)[17:] == 0).all()
    L, nz = eng.sweepx_assemble(2, "step", X, nm)
    assert rel(nz, nzref) <= TOL and rel(L, Lref, np.abs(nzref).max()) <= TOL
    assert (L[:5] == 0).all() and (L[17:] == 0).all()


def test_empty_element_type_between_others(mb, engine_factory):
    """an element type with zero elements keeps its ieletyp slot and contributes nothing"""
    N = 40
    eleobj, idx, ndof = mb.synthetic.chain(N)
    X = mb.synthetic.state(ndof)
    nm = mb.synthetic.newmark_coefficients(0, 0.)
    eng = engine_factory()
    t1 = eng.add_eulerbeam3d(eleobj[:25], idx[:25], np.ones(12))
    t2 = eng.add_eulerbeam3d(eleobj[:0], idx[:0], np.ones(12))
    t3 = eng.add_eulerbeam3d(eleobj[25:], idx[25:], np.ones(12))
    assert (t1, t2, t3) == (1, 2, 3)
    eng.sweepx_prepare(ndof)
    colptr, rowval, Lref, nzref = oracle_beams(eleobj, idx, ndof, 0, "iter", X, np.ones(12), nm)
    cp, rv = eng.sweepx_pattern()
    assert np.array_equal(cp, colptr) and np.array_equal(rv, rowval)
    L, nz = eng.sweepx_assemble(0, "iter", X, nm)
    assert rel(nz, nzref) <= TOL and rel(L, Lref, np.abs(nzref).max()) <= TOL
    a1, a2 = eng.sweepx_asm(t2)
    assert a1.shape[0] == 0 and a2.shape[0] == 0


@pytest.mark.parametrize("OX,mission", [(0, "iter"), (2, "iter")])
def test_star_topology_many_contributors(mb, engine_factory, OX, mission):
    """24 beams meeting in one hub node: the hub's 36 non-zeros collect 24 contributions each, summed in element order"""
    M = 24
    rng = np.random.default_rng(9)
    hub = np.zeros(3)
    tips = rng.standard_normal((M, 3)); tips /= np.linalg.norm(tips, axis=1)[:, None]
    mat = OE.beam_cross_section(EA=10., EI2=3., EI3=2.5, GJ=4., mu=1., iota1=1.2, w=.3, Ca2=2., Ca3=1.5, Cq2=1., Cq3=.7)
    eleobj = np.stack([OE.beam_ctor(hub, tips[k] * (1 + .1 * k), mat, orient2=(0.3, 1., 0.2)) for k in range(M)])
    idx = np.stack([np.concatenate([np.arange(1, 7), 6 * (k + 1) + np.arange(1, 7)]) for k in range(M)]).astype(np.int64)
    idx = idx[rng.permutation(M)] if False else idx
    ndof = 6 * (M + 1)
    X = mb.synthetic.state(ndof, nder=OX + 1)
    scale = np.array([2., 2, 2, .5, .5, .5] * 2)
    nm = mb.synthetic.newmark_coefficients(OX, 0.3)
    eng = engine_factory()
    eng.add_eulerbeam3d(eleobj, idx, scale); eng.sweepx_prepare(ndof)
    colptr, rowval, Lref, nzref = oracle_beams(eleobj, idx, ndof, OX, mission, X, scale, nm)
    cp, rv = eng.sweepx_pattern()
    assert np.array_equal(cp, colptr) and np.array_equal(rv, rowval)
    L, nz = eng.sweepx_assemble(OX, mission, X, nm)
    assert rel(nz, nzref) <= TOL and rel(L, Lref, np.abs(nzref).max()) <= TOL
    L2, nz2 = eng.sweepx_assemble(OX, mission, X, nm)
    assert np.array_equal(nz, nz2) and np.array_equal(L, L2)


def test_abi_error_behaviour(mb):
    """status codes and messages instead of exceptions across the ABI: wrong call order, bad arguments, bad indices"""
    eleobj, idx, ndof = mb.synthetic.chain(4)
    eng = mb.Engine(0)
    try:
        with pytest.raises(mb.MuscadeB200Error, match="prepare"):
            eng.sweepx_assemble(0, "iter", mb.synthetic.state(ndof), mb.synthetic.newmark_coefficients(0, 0.))
        bad = idx.copy(); bad[2, 5] = 0
        with pytest.raises(mb.MuscadeB200Error, match="1-based"):
            eng.add_eulerbeam3d(eleobj, bad, np.ones(12))
        eng.add_eulerbeam3d(eleobj, idx, np.ones(12))
        eng.sweepx_prepare(ndof)
        with pytest.raises(mb.MuscadeB200Error, match="before prepare"):
            eng.add_eulerbeam3d(eleobj, idx, np.ones(12))
        with pytest.raises(mb.MuscadeB200Error, match="state vectors missing"):
            eng.L.mb_sweepx_assemble.restype = int
            X = mb.synthetic.state(ndof)
            rc = eng.L.mb_sweepx_assemble(eng.h, 2, 1, X[0].ctypes.data, None, None, None, 0., mb.synthetic.newmark_coefficients(2, .3), None, None, None)
            mb._lib.check(eng.h, rc)
        # the handle is still usable after errors
        L, nz = eng.sweepx_assemble(0, "iter", mb.synthetic.state(ndof), mb.synthetic.newmark_coefficients(0, 0.))
        assert np.isfinite(nz).all()
    finally:
        eng.close()


def test_directxua_nan_reports_step_and_element(mb):
    """NaN guard (src/Assemble.jl:630) in the all-steps assembly: first offending (step, element type, element)"""
    from test_gpu_directxua import udof_chain, states
    OX, nstep, dt, N = 2, 7, 0.1, 9
    model = udof_chain(mb, N)
    st0 = mb.initialize(model); dis = st0.dis
    nX, nU = model.getndof("X"), model.getndof("U")
    st = states(mb, nX, nU, nstep)
    st[4][0][1][dis.dis[0].X[6, 2] - 1] = np.nan          # velocity of a dof of element 7 (and its neighbour 6), step 5
    eng = mb.directxua.prepare(OX, 0, model, dis, nstep, dt)
    try:
        for s, (X, U) in enumerate(st):
            eng.set_state(s, X, U)
        with pytest.raises(mb.MuscadeB200Error) as ei:
            eng.direct_assemble(Lv=np.zeros(eng.ncol))
        d = ei.value.dbg
        assert d["step"] == 5 and d["ieletyp"] == 1 and d["iele"] == min(e for e in range(N) if (dis.dis[0].X[e] == dis.dis[0].X[6, 2]).any()) + 1
    finally:
        eng.close()


def test_dof_index_beyond_model_size_is_rejected(mb):
    """prepare checks every group's dof numbers against ndofX / ndofU (they are not known when the groups are added): an index beyond the model would
    scatter outside colptr and read outside the state vectors — MB_ERR_ARG instead"""
    eleobj, idx, ndof = mb.synthetic.chain(5)
    e = mb.Engine(0)
    e.add_eulerbeam3d(eleobj, idx, np.ones(12))
    with pytest.raises(mb.MuscadeB200Error, match="exceeds ndofX"):
        e.sweepx_prepare(ndof - 1)
    e.close()
    e = mb.Engine(0)
    e.add_eulerbeam3d(eleobj, idx, np.ones(12))
    assert e.sweepx_prepare(ndof) == 108 * 5 + 36
    e.close()
