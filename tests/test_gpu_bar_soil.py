"""GPU parity for Bar3D (toolbox/BarElement.jl) and SoilContact (toolbox/SoilContact.jl) kernels, alone and mixed with beams
in one model (the SCR-riser style of examples/DynamicBeamAnalysis.jl: beams + soil springs on the same nodes)."""
import numpy as np
import pytest

from oracle import elements as OE
from oracle import pattern as OP

pytestmark = pytest.mark.gpu
TOL = 1e-12


def rel(a, b, floor=0.):
    return np.abs(a - b).max() / max(np.abs(b).max(), floor, 1e-300)


def mixed_model(mb, N=40):
    rng = np.random.default_rng(4)
    model = mb.Model()
    coord = np.cumsum(np.concatenate([[[0., 0., -0.3]], rng.uniform(0.5, 1.0, (N, 3)) * [1, .3, .02]]), axis=0)
    nod = mb.addnode(model, coord)
    mesh = np.stack([nod[:-1], nod[1:]], axis=1)
    bmat = mb.BeamCrossSection(EA=1e3, EI2=30., EI3=20., GJ=40., mu=2., iota1=.3, w=5., Ca2=3., Ca3=3., Cq2=2., Cq3=2., Cl1=.5)
    mb.addelement(model, mb.EulerBeam3D, mesh[: N // 2], mat=bmat, orient2=(0., 0.2, 1.))
    rmat = mb.AxisymmetricBarCrossSection(EA=500., mu=1.5, w=3., Cat=.2, Clt=.3, Cqt=.4, Can=2., Cln=.6, Cqn=1.2)
    mb.addelement(model, mb.Bar3D, mesh[N // 2:], mat=rmat)
    mb.addelement(model, mb.SoilContact, nod[: N // 2, None], z0=0.0, Kh=30., Kv=200., Ch=3., Cv=7.)
    return model


@pytest.mark.parametrize("pipe", [False, True])
@pytest.mark.parametrize("OX,mission,N", [(0, "iter", 40), (1, "step", 40), (2, "iter", 40), (2, "step", 40), (0, "iter", 730), (2, "step", 730)])
def test_mixed_beam_bar_soil(mb, OX, mission, N, pipe, monkeypatch):
    """N = 730: 365 bars / 365 soil springs = several CTAs of full warps plus a partial one (the warp-staged stores of bar_kernel / soil_kernel)"""
    if pipe:       # force the chunked host-buffer pipeline of mb_sweepx_assemble (normally only for ≥ 4M non-zeros)
        monkeypatch.setenv("MB_E2E_MIN_NNZ", "0"); monkeypatch.setenv("MB_E2E_CHUNKS", "4")
    model = mixed_model(mb, N)
    mb.setscale(model, scale=dict(X=dict(t1=3., t2=3., t3=3., r1=1., r2=1., r3=1.)))
    state = mb.initialize(model)
    dis = state.dis
    ndof = model.getndof("X")
    state = state.with_orders(1, OX + 1, 1)
    state.X[0] = mb.synthetic.uniform_pm1(5, ndof) * 0.1
    for d in range(1, OX + 1):
        state.X[d] = mb.synthetic.uniform_pm1(5 + d, ndof) * 0.3
    state.time = -7.3          # inside the weight ramp of Bar3D (BarElement.jl:144)
    out, asm, gr = mb.sweepx.prepare(OX, model, dis)
    out.c = mb.synthetic.newmark_coefficients(OX, 0.3)
    mb.sweepx.assemble(mission, out, asm, dis, model, state, 0.3)
    # oracle
    odis = [dict(X=d.X, U=np.zeros((d.X.shape[0], 0), np.int64), A=np.zeros((d.X.shape[0], 0), np.int64)) for d in dis.dis]
    asm1, asm2, colptr, rowval = OP.prepare_sweepx(odis, ndof, 0, 0)
    assert np.array_equal(out.Lλx.indptr + 1, colptr) and np.array_equal(out.Lλx.indices + 1, rowval)
    L = np.zeros(ndof); nz = np.zeros(len(rowval))
    X = state.X[: OX + 1]
    OE.sweepx_assemble_beams(model.ele[0].eleobj, dis.dis[0].X, asm1[0].T, asm2[0].T, OX, mission, X, dis.dis[0].scaleX, out.c, L, nz)
    bars = model.ele[1].eleobj
    OE.sweepx_addin_generic(lambda e, xv, sd: OE.bar_residual(bars[e], xv, sd, t=state.time)[:2], 6, dis.dis[1].X, asm1[1].T, asm2[1].T,
                            OX, mission, X, dis.dis[1].scaleX, out.c, L, nz)
    soil = model.ele[2].eleobj
    OE.sweepx_addin_generic(lambda e, xv, sd: OE.soil_residual(soil[e], xv, sd)[:2], 3, dis.dis[2].X, asm1[2].T, asm2[2].T,
                            OX, mission, X, dis.dis[2].scaleX, out.c, L, nz)
    for k in range(3):
        assert np.array_equal(asm[1, k + 1], asm1[k]) and np.array_equal(asm[2, k + 1], asm2[k])
    assert rel(out.Lλx.data, nz) <= TOL, rel(out.Lλx.data, nz)
    assert rel(out.Lλ, L, np.abs(nz).max()) <= TOL
    assert (state.X[0][dis.dis[2].X[:, 2] - 1] < 0).any() and (state.X[0][dis.dis[2].X[:, 2] - 1] >= 0).any()   # both soil branches exercised
    out.engine.close()


@pytest.mark.parametrize("OX,mission", [(0, "iter"), (2, "iter"), (2, "step")])
def test_odd_sized_type_before_beams(mb, OX, mission):
    """an element type with an odd number of tangent entries (21 SoilContact × 9) stored BEFORE the beams: the beams' element tangents then start at an
    8-byte (not 16-byte) aligned address — the kernels' vector stores must cope"""
    rng = np.random.default_rng(9)
    N = 30
    model = mb.Model()
    coord = np.cumsum(np.concatenate([[[0., 0., -0.3]], rng.uniform(0.5, 1.0, (N, 3)) * [1, .3, .02]]), axis=0)
    nod = mb.addnode(model, coord)
    mesh = np.stack([nod[:-1], nod[1:]], axis=1)
    mb.addelement(model, mb.SoilContact, nod[:21, None], z0=0.0, Kh=30., Kv=200., Ch=3., Cv=7.)
    bmat = mb.BeamCrossSection(EA=1e3, EI2=30., EI3=20., GJ=40., mu=2., iota1=.3, w=5., Ca2=3., Ca3=3., Cq2=2., Cq3=2., Cl1=.5)
    mb.addelement(model, mb.EulerBeam3D, mesh, mat=bmat, orient2=(0., 0.2, 1.))
    state = mb.initialize(model)
    dis = state.dis
    ndof = model.getndof("X")
    state = state.with_orders(1, OX + 1, 1)
    state.X[0] = mb.synthetic.uniform_pm1(5, ndof) * 0.1
    for d in range(1, OX + 1):
        state.X[d] = mb.synthetic.uniform_pm1(5 + d, ndof) * 0.3
    out, asm, gr = mb.sweepx.prepare(OX, model, dis)
    out.c = mb.synthetic.newmark_coefficients(OX, 0.3)
    mb.sweepx.assemble(mission, out, asm, dis, model, state, 0.3)
    odis = [dict(X=d.X, U=np.zeros((d.X.shape[0], 0), np.int64), A=np.zeros((d.X.shape[0], 0), np.int64)) for d in dis.dis]
    asm1, asm2, colptr, rowval = OP.prepare_sweepx(odis, ndof, 0, 0)
    L = np.zeros(ndof); nz = np.zeros(len(rowval))
    X = state.X[: OX + 1]
    soil = model.ele[0].eleobj
    OE.sweepx_addin_generic(lambda e, xv, sd: OE.soil_residual(soil[e], xv, sd)[:2], 3, dis.dis[0].X, asm1[0].T, asm2[0].T,
                            OX, mission, X, dis.dis[0].scaleX, out.c, L, nz)
    OE.sweepx_assemble_beams(model.ele[1].eleobj, dis.dis[1].X, asm1[1].T, asm2[1].T, OX, mission, X, dis.dis[1].scaleX, out.c, L, nz)
    assert rel(out.Lλx.data, nz) <= TOL, rel(out.Lλx.data, nz)
    assert rel(out.Lλ, L, np.abs(nz).max()) <= TOL
    out.engine.close()
