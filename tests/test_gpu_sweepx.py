"""GPU parity, SweepX path: CUDA engine (through the C ABI) vs the oracle on the same seeded inputs.

Bars: integer structures bit-exact (colptr, rowval, asm maps); floats max|Δ| ≤ 1e-12·max|ref| per assembled array
(SURVEY.md §8d "Parity measure"; north_star: 1e-12 relative in FP64)."""
import numpy as np
import pytest

from oracle import elements as OE
from oracle import pattern as OP

pytestmark = pytest.mark.gpu
TOL = 1e-12


def rel(a, b, floor=0.):
    """max|a-b| / max(max|b|, floor).  For the gradient Lλ the floor is max|nzval|: at an equilibrium state (e.g. all-zero)
    Lλ is pure round-off of terms of the tangent's magnitude, and a ratio of two noises says nothing."""
    return np.abs(a - b).max() / max(np.abs(b).max(), floor, 1e-300)


def oracle_assemble(eleobj, idx, ndof, OX, mission, X, scale, nm):
    dis = [dict(X=idx, U=np.zeros((idx.shape[0], 0), np.int64), A=np.zeros((idx.shape[0], 0), np.int64))]
    asm1, asm2, colptr, rowval = OP.prepare_sweepx(dis, ndof, 0, 0)
    L = np.zeros(ndof); nz = np.zeros(len(rowval))
    OE.sweepx_assemble_beams(eleobj, idx, asm1[0].T, asm2[0].T, OX, mission, X, scale, nm, L, nz)
    return asm1[0].T, asm2[0].T, colptr, rowval, L, nz


@pytest.mark.parametrize("OX,mission", [(0, "iter"), (1, "iter"), (1, "step"), (2, "iter"), (2, "step")])
@pytest.mark.parametrize("zero", [False, True])
def test_chain_parity(mb, engine_factory, OX, mission, zero):
    N = 500
    eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=OX > 0)
    X = mb.synthetic.state(ndof, nder=OX + 1, zero=zero)
    scale = np.ones(12)
    nm = mb.synthetic.newmark_coefficients(OX, 0.3)
    eng = engine_factory()
    ityp = eng.add_eulerbeam3d(eleobj, idx, scale)
    nnz = eng.sweepx_prepare(ndof)
    a1, a2, colptr, rowval, Lref, nzref = oracle_assemble(eleobj, idx, ndof, OX, mission, X, scale, nm)
    assert nnz == len(rowval) == 108 * N + 36
    cp, rv = eng.sweepx_pattern()
    assert np.array_equal(cp, colptr) and np.array_equal(rv, rowval)
    g1, g2 = eng.sweepx_asm(ityp)
    assert np.array_equal(g1, a1) and np.array_equal(g2, a2)
    L, nz = eng.sweepx_assemble(OX, mission, X, nm)
    assert rel(L, Lref, np.abs(nzref).max()) <= TOL, rel(L, Lref)
    assert rel(nz, nzref) <= TOL, rel(nz, nzref)
    # determinism: bit-identical on repetition (no atomics in the reduction)
    L2, nz2 = eng.sweepx_assemble(OX, mission, X, nm)
    assert np.array_equal(L, L2) and np.array_equal(nz, nz2)


def test_scaled_and_shuffled(mb, engine_factory):
    """non-unit scales (X:(t=10,r=1)) and a non-chain dof numbering / element order"""
    N = 300
    eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=True)
    rng = np.random.default_rng(3)
    perm = rng.permutation(ndof)
    idx = (perm[idx - 1] + 1).astype(np.int64)
    order = rng.permutation(N)
    eleobj, idx = eleobj[order], idx[order]
    X = [x[np.argsort(perm)] * 0 + x for x in mb.synthetic.state(ndof, nder=3)]
    scale = np.array([10., 10., 10., 1., 1., 1.] * 2)
    nm = mb.synthetic.newmark_coefficients(2, 0.3)
    eng = engine_factory()
    ityp = eng.add_eulerbeam3d(eleobj, idx, scale)
    eng.sweepx_prepare(ndof)
    a1, a2, colptr, rowval, Lref, nzref = oracle_assemble(eleobj, idx, ndof, 2, "step", X, scale, nm)
    cp, rv = eng.sweepx_pattern()
    assert np.array_equal(cp, colptr) and np.array_equal(rv, rowval)
    g1, g2 = eng.sweepx_asm(ityp)
    assert np.array_equal(g1, a1) and np.array_equal(g2, a2)
    L, nz = eng.sweepx_assemble(2, "step", X, nm)
    assert rel(L, Lref, np.abs(nzref).max()) <= TOL and rel(nz, nzref) <= TOL


def test_nan_reported_with_location(mb, engine_factory):
    """NaN guard of getresidual (src/Assemble.jl:630): first offending element, 1-based, deterministic"""
    N = 64
    eleobj, idx, ndof = mb.synthetic.chain(N)
    X = mb.synthetic.state(ndof)
    X[0][6 * 40 + 1] = np.nan
    eng = engine_factory()
    eng.add_eulerbeam3d(eleobj, idx, np.ones(12))
    eng.sweepx_prepare(ndof)
    with pytest.raises(mb.MuscadeB200Error) as ei:
        eng.sweepx_assemble(0, "iter", X, mb.synthetic.newmark_coefficients(0, 0.))
    assert ei.value.dbg["ieletyp"] == 1 and ei.value.dbg["iele"] == 40


@pytest.mark.parametrize("OX,mission", [(0, "iter"), (2, "step")])
@pytest.mark.parametrize("shuffle", [False, True])
def test_pipelined_host_path_is_bit_identical(mb, engine_factory, monkeypatch, OX, mission, shuffle):
    """mb_sweepx_assemble with host buffers evaluates element chunks and ships completed prefixes of nzval / Lλ while the next chunk
    computes (large models only; forced here).  Same kernels, same summation order ⇒ bit-identical to the one-shot path, also when the
    element order is random (no prefix completes early) and when the chunk count does not divide the element count."""
    N = 1013
    eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=OX > 0)
    if shuffle:
        perm = np.random.default_rng(3).permutation(N)
        eleobj, idx = eleobj[perm], idx[perm]
    X = mb.synthetic.state(ndof, nder=OX + 1)
    nm = mb.synthetic.newmark_coefficients(OX, 0.3)
    ref = engine_factory()
    ref.add_eulerbeam3d(eleobj, idx, np.ones(12)); ref.sweepx_prepare(ndof)
    L0, nz0 = ref.sweepx_assemble(OX, mission, X, nm)
    monkeypatch.setenv("MB_E2E_MIN_NNZ", "0"); monkeypatch.setenv("MB_E2E_CHUNKS", "7")
    eng = engine_factory()
    eng.add_eulerbeam3d(eleobj, idx, np.ones(12)); eng.sweepx_prepare(ndof)
    for _ in range(2):
        L1, nz1 = eng.sweepx_assemble(OX, mission, X, nm)
        assert np.array_equal(L0, L1) and np.array_equal(nz0, nz1)
    # NaN location is still the first offending element in element order
    Xbad = [x.copy() for x in X]
    Xbad[0][idx[700, 3] - 1] = np.nan
    with pytest.raises(mb.MuscadeB200Error) as ei:
        eng.sweepx_assemble(OX, mission, Xbad, nm)
    assert ei.value.dbg["iele"] == min(i for i in range(N) if (idx[i] == idx[700, 3]).any()) + 1


@pytest.mark.parametrize("OX", [0, 1, 2])
def test_device_newmark_decrement_is_bit_identical(mb, engine_factory, OX):
    """mb_sweepx_newmark_decrement ≡ Newmarkβdecrement!{OX} (SweepX.jl:98-132) with getdof!/decrement! (Assemble.jl:206-233):
    same rounding sequence as the host restatement (no FMA contraction) ⇒ bit-identical states, first and later iterations."""
    N = 300
    eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=True)
    eng = engine_factory()
    eng.add_eulerbeam3d(eleobj, idx, np.ones(12)); eng.sweepx_prepare(ndof)
    rng = np.random.default_rng(11)
    scale = rng.uniform(0.1, 10., ndof)
    X = [rng.standard_normal(ndof) for _ in range(OX + 1)]
    nm = mb.synthetic.newmark_coefficients(OX, 0.37)
    eng.set_dof_scale(scale); eng.set_state(X)

    class S:                       # the slice of State / DofGroup the host restatement touches
        pass
    st = S(); st.X = [x.copy() for x in X]; st.Λ = []; st.U = []; st.A = np.zeros(0)
    gr = S(); gr.iX = gr.jX = np.arange(1, ndof + 1); gr.scaleX = scale
    gr.iΛ = gr.iU = gr.iA = gr.jΛ = gr.jU = gr.jA = np.zeros(0, np.int64)
    buf = (np.zeros(ndof), np.zeros(ndof))
    for it in range(3):
        dx = rng.standard_normal(ndof) * 1e-2
        mb.sweepx.newmark_decrement(OX, st, dx, gr, nm, it == 0, buf)
        dx2, _ = eng.newmark_decrement(OX, it == 0, dx, nm)
        got = eng.get_state(OX)
        for d in range(OX + 1):
            assert np.array_equal(got[d], st.X[d]), (it, d, np.abs(got[d] - st.X[d]).max())
        assert abs(dx2 - dx @ dx) <= 1e-13 * (dx @ dx)
    # assembling at the resident state == assembling at the same state passed from the host
    L0, nz0 = eng.sweepx_assemble_resident(OX, "iter", nm)
    L1, nz1 = eng.sweepx_assemble(OX, "iter", st.X, nm)
    assert np.array_equal(L0, L1) and np.array_equal(nz0, nz1)


@pytest.mark.parametrize("OX", [0, 2])
def test_beam_requestables_batched(mb, engine_factory, OX):
    """getresult(state,req,els) for all EulerBeam3D elements at once (src/Output.jl:131-181; requestables of toolbox/BeamElement.jl:28-64,
    151-174) vs the oracle's per-element capture; from host state and from the device-resident state."""
    N = 257
    eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=True)
    X = mb.synthetic.state(ndof, nder=OX + 1)
    eng = engine_factory()
    ityp = eng.add_eulerbeam3d(eleobj, idx, np.ones(12)); eng.sweepx_prepare(ndof)
    res = eng.beam_results(ityp, OX, X)
    ref = np.stack([OE.beam_results(eleobj[e], np.stack([X[d][idx[e] - 1] for d in range(OX + 1)])) for e in range(N)])
    assert res["raw"].shape == ref.shape == (N, 77)
    for lo, hi in [(0, 1), (1, 10), (10, 13)] + [(13 + 16 * g + a, 13 + 16 * g + b) for g in range(4) for a, b in [(0, 3), (3, 6), (6, 7), (7, 10), (10, 13), (13, 16)]]:
        scale = max(np.abs(ref[:, lo:hi]).max(), 1e-3)
        assert np.abs(res["raw"][:, lo:hi] - ref[:, lo:hi]).max() <= 1e-12 * max(scale, 1.), (lo, hi)
    assert np.array_equal(res["fi"][:, 2], res["raw"][:, 13 + 32 + 6]) and res["rsm"].shape == (N, 3, 3)
    eng.set_state(X)
    res2 = eng.beam_results(ityp, OX)
    assert np.array_equal(res2["raw"], res["raw"])


@pytest.mark.parametrize("topology", ["chain", "shuffled", "star", "mixed", "chain_pipelined"])
def test_fused_epilogue_bit_identical(mb, engine_factory, monkeypatch, topology):
    """MB_FUSE=1: the static kernel sums the non-zeros a warp of five elements holds all contributors of in its shared-memory tile (beam_kernel.cuh, "fused
    epilogue") — same element-order sums, so Lλ and nzval are the BITS of the default path, whatever the mesh (pattern dedupe, partial last warp, nodes with
    many elements, another element type on the same dofs)."""
    N = {"star": 40, "chain_pipelined": 50003}.get(topology, 1003)          # 50 003 elements: nnz beyond the threshold of the chunked host-buffer pipeline
    rng = np.random.default_rng(11)
    eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=False)
    bars = None
    if topology == "shuffled":
        perm = rng.permutation(ndof)
        idx = (perm[idx - 1] + 1).astype(np.int64)
        order = rng.permutation(N); eleobj, idx = eleobj[order], idx[order]
    elif topology == "star":                                  # every element's first node is node 1
        idx = idx.copy(); idx[:, :6] = np.arange(1, 7)
    X = mb.synthetic.state(ndof, nder=1)
    nm = mb.synthetic.newmark_coefficients(0, 0.)
    out = []
    for fuse in ("0", "1"):
        monkeypatch.setenv("MB_FUSE", fuse)
        eng = engine_factory()
        eng.add_eulerbeam3d(eleobj, idx, np.ones(12))
        if topology == "mixed":                               # a second beam type sharing the first type's nodes: contributors from two groups
            eng.add_eulerbeam3d(eleobj[:200], idx[100:300], np.ones(12))
        eng.sweepx_prepare(ndof)
        L, nz = eng.sweepx_assemble(0, "iter", X, nm)
        L2, nz2 = eng.sweepx_assemble(0, "iter", X, nm)
        assert np.array_equal(L, L2) and np.array_equal(nz, nz2)
        out.append((L, nz))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


@pytest.mark.parametrize("shuffled", [False, True])
@pytest.mark.parametrize("OX", [0, 2])
def test_host_state_pipeline_bit_identical(mb, engine_factory, monkeypatch, OX, shuffled):
    """mb_sweepx_assemble with host state in and host Lλ out (nzval left in HBM): the state is copied in pieces and element chunk j starts when the dofs it reads have
    arrived; Lλ leaves range by range.  Same kernels on the same data ⇒ the bits of the one-shot call (MB_E2E_CHUNKS=1), also when the numbering gives the pipeline
    nothing to overlap (shuffled: the first chunk already needs the whole state)."""
    N = 60011
    eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=OX > 0)
    if shuffled:
        rng = np.random.default_rng(7)
        perm = rng.permutation(ndof); idx = (perm[idx - 1] + 1).astype(np.int64)
    X = mb.synthetic.state(ndof, nder=OX + 1)
    nm = mb.synthetic.newmark_coefficients(OX, 0.3)
    out = []
    for chunks in ("1", "8"):
        monkeypatch.setenv("MB_E2E_CHUNKS", chunks)
        eng = engine_factory()
        eng.add_eulerbeam3d(eleobj, idx, np.ones(12)); eng.sweepx_prepare(ndof)
        L, _ = eng.sweepx_assemble(OX, "iter", X, nm, nzval_on_device=True)
        L2, nz = eng.sweepx_assemble(OX, "iter", X, nm)                      # and with the CSC values copied out as well
        assert np.array_equal(L, L2)
        out.append((L, nz))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
