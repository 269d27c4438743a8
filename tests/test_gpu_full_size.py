"""BASELINE.json's full size (10 M EulerBeam3D elements, configs[2]) on the GPU.  The oracle cannot assemble 10 M elements in seconds, but the
tangent and gradient entries of a node depend on its two elements only: windows of P nodes at the start, middle and end of the model are
re-assembled by the oracle from the very same element structs and states and compared to 1e-12.  Between the windows a size-independent
property is used: the state repeats every P nodes and the mesh is uniform up to the rounding of its coordinates (k·h at k ~ 10⁷ carries
≈1e-9 relative error into the element length), so the whole value array is P-periodic to 1e-6.  The run is repeated for bit-reproducibility.
(The non-zero layout of the chain is regular: node k ≥ 1 owns 108 values from 72 + 108(k−1))"""
import numpy as np
import pytest

from oracle import elements as OE
from oracle import pattern as OP

pytestmark = pytest.mark.gpu
TOL = 1e-12
N, P = 10_000_000, 50


def periodic_state(mb, nnode, nder):
    base = mb.synthetic.state(6 * P, nder=nder)
    reps = -(-nnode // P)
    return [np.tile(b.reshape(P, 6), (reps, 1))[:nnode].ravel().copy() for b in base]


@pytest.mark.parametrize("OX,mission", [(0, "iter"), (2, "iter")])
def test_ten_million_elements_periodic_windows(mb, OX, mission):
    nm = mb.synthetic.newmark_coefficients(OX, 0.3)

    def oracle_window(eleobj, X, k0):
        """nodes k0 .. k0+P-1 from elements k0-1 .. k0+P-1, local numbering"""
        ne = P + 1
        eo = eleobj[k0 - 1: k0 + P]
        ix = (np.arange(ne)[:, None] * 6 + np.arange(1, 13)[None, :]).astype(np.int64)
        nd = 6 * (ne + 1)
        Xl = [x[6 * (k0 - 1): 6 * (k0 + P + 1)].copy() for x in X]
        dis = [dict(X=ix, U=np.zeros((ne, 0), np.int64), A=np.zeros((ne, 0), np.int64))]
        asm1, asm2, colptr, rowval = OP.prepare_sweepx(dis, nd, 0, 0)
        assert colptr[6] - 1 == 72 and colptr[12] - colptr[6] == 108          # the regular layout used below
        Lr = np.zeros(nd); nzr = np.zeros(len(rowval))
        OE.sweepx_assemble_beams(eo, ix, asm1[0].T, asm2[0].T, OX, mission, Xl, np.ones(12), nm, Lr, nzr)
        return nzr[72: 72 + 108 * P].reshape(P, 108), Lr[6: 6 * (P + 1)].reshape(P, 6)       # local nodes 1 .. P
    # the full model
    eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=True)
    X = periodic_state(mb, N + 1, OX + 1)
    eng = mb.Engine(0)
    try:
        eng.add_eulerbeam3d(eleobj, idx, np.ones(12))
        assert eng.sweepx_prepare(ndof) == 108 * N + 36
        del idx
        L, nz = eng.sweepx_assemble(OX, mission, X, nm)
        for k0 in (P, (N // 2 // P) * P, ((N - P) // P - 1) * P):                  # node windows aligned with the period
            ref_nz, ref_L = oracle_window(eleobj, X, k0)
            scale = np.abs(ref_nz).max()
            got = nz[72 + 108 * (k0 - 1): 72 + 108 * (k0 + P - 1)].reshape(P, 108)
            assert np.abs(got - ref_nz).max() <= TOL * scale, k0
            assert np.abs(L[6 * k0: 6 * (k0 + P)].reshape(P, 6) - ref_L).max() <= TOL * scale, k0
        # global periodicity of everything in between (cheap full-array property) and bit-reproducibility
        body = nz[72 + 108 * (P - 1): 72 + 108 * (P - 1) + 108 * P * ((N - 2 * P) // P)].reshape(-1, P * 108)
        assert np.abs(body - body[0]).max() <= 1e-6 * scale
        del body
        chk = (float(nz.sum()), float(np.abs(nz).max()), float(L.sum()))
        L2, nz2 = eng.sweepx_assemble(OX, mission, X, nm)
        assert (float(nz2.sum()), float(np.abs(nz2).max()), float(L2.sum())) == chk and np.array_equal(nz[:10**6], nz2[:10**6])
    finally:
        eng.close()


def test_general_form_equals_specialised_path_at_size(mb):
    """A size the oracle's loops cannot do in seconds: 20 000 EulerBeam3D{Udof} × 6 steps, DirectXUA{2,0,0}.  The general form (mb_xua_*: packets from the device kernels, segmented
    reductions, tabulated weighted adds through Lvvasm) and the beam-specialised path (mb_direct_*: block-implicit Lvv) implement the same assemblebig! by different means:
    identical structures, Lvv values and Lv within 1e-13 of each other (the specialised path is the one checked against the oracle at small sizes)."""
    from muscade_b200 import xua
    Ne, OX, OU, nstep, dt = 20000, 2, 0, 6, 0.1
    m = mb.Model("udof chain")
    h = np.array([0.8, 0.6, 0.])
    nod = mb.addnode(m, np.arange(Ne + 1, dtype=float)[:, None] * h[None, :])
    un = mb.addnode(m, np.zeros((Ne, 0)))
    mb.addelement(m, mb.EulerBeam3D, np.stack([nod[:-1], nod[1:], un], axis=1),
                  mat=mb.BeamCrossSection(EA=10., EI2=3., EI3=3., GJ=4., mu=1., iota1=1., Ca2=169.6, Ca3=169.6, Cq2=235.2, Cq3=235.2), Udof=True)
    mb.setscale(m, scale=dict(X=dict(t1=2., t2=2., t3=2.), U=dict(t1=5., t2=5., t3=5.)))
    s0 = mb.initialize(m); dis = s0.dis
    nX, nU = m.getndof("X"), m.getndof("U")
    st = [([mb.synthetic.uniform_pm1(10 + 3 * s + d, nX) * (0.05 if d == 0 else 0.1) for d in range(3)], 0.5 * mb.synthetic.uniform_pm1(99 + s, nU)) for s in range(nstep)]
    spec = mb.directxua.prepare(OX, OU, m, dis, nstep, dt)
    gen = xua.XUAEngine(0)
    try:
        for s, (X, U) in enumerate(st):
            spec.set_state(s, X, U)
        Lvv = np.zeros(spec.nnzbig); Lv = np.zeros(spec.ncol)
        spec.direct_assemble(Lvv=Lvv, Lv=Lv)
        cp, rv = spec.big_pattern()
        nbig, nnz = gen.prepare(m, dis, OX, OU, 0, [nstep], [dt])
        assert nbig == spec.ncol and nnz == spec.nnzbig
        gcp, grv = gen.big_pattern()
        assert np.array_equal(cp, gcp) and np.array_equal(rv, grv)
        states = [[mb.State(dt * s, [np.zeros(nX)], st[s][0], [st[s][1]], s0.A, None, m, dis) for s in range(nstep)]]
        gen.assemblebig(states)
        gLvv, gLv = gen.big()
        scale = np.abs(Lvv).max()
        assert np.abs(gLvv - Lvv).max() <= 1e-13 * scale and np.abs(gLv - Lv).max() <= 1e-13 * max(scale, np.abs(Lv).max())
        assert np.count_nonzero(Lvv) > 0.2 * Lvv.size
    finally:
        spec.close(); gen.close()
