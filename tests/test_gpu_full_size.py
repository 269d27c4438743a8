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
