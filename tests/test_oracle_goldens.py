"""Oracle pinning (no GPU): the CPU restatement against every golden vector the reference's own tests hold for this path.

Each test names the reference test it ports.  `≈` in Julia is rtol=√eps on the norm for arrays; the same is used here."""
import numpy as np
import pytest

from oracle import elements as OE
from oracle import pattern as OP

RT = 1.5e-8


def approx(a, b, rtol=RT, atol=0.):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return np.linalg.norm(a - b) <= max(atol, rtol * max(np.linalg.norm(a), np.linalg.norm(b)))


# ------------------------------------------------------------------------------------------------ test/TestRotations.jl
def test_sinc1_family():
    th = [0, 1e-8, 1e-6, 1e-2, .1, 1, np.pi]
    gold = [[1.0, 1.0, 0.9999999999998334, 0.9999833334166666, 0.9983341664682817, 0.8414709848078965, 0.0],
            [0.0, -3.3333333333333334e-9, -3.3333333333329995e-7, -0.003333300000107897, -0.03330001190255594, -0.30116867893975674, -0.3183098861837907],
            [-0.3333333333333333, -0.3333333333333333, -0.33333333333323334, -0.33332333339285697, -0.33233392841710757, -0.23913362692838303, 0.20264236728467552],
            [0.0, 2.0e-9, 1.999999999999762e-7, 0.001999976190568783, 0.019976199733646196, 0.1770985749170091, 0.12480067958459377],
            [0.2, 0.2, 0.1999999999999286, 0.19999285718915333, 0.19928617712243368, 0.13307670326700266, -0.14309499132952147]]
    for k in range(5):
        assert approx([OE.sinc1k(k, t) for t in th], gold[k])


def test_scac_and_derivatives():
    x = [-1 + 1e-11, -1 + 1e-9, 0, 1 - 1e-9, 1 - 1e-11, 1]
    gold = [[1.4235271721825872e-6, 1.4235453308738641e-5, 0.6366197723675814, 0.9999999996666666, 0.9999999999966667, 1.0],
            [71176.45403923128, 7117.8281743984935, 0.40528473456935105, 0.33333333337777776, 0.33333333333377774, 0.3333333333333333],
            [-3.55881227537807e15, -3.5588128658975117e12, -0.12059522143638957, -0.04444444447619157, -0.04444444444476192, -0.044444444444444446],
            [5.338217971349224e26, 5.338219446646726e21, 0.17496482731099405, 0.03174712798947714, 0.031747127950916276, 0.031747127950526775],
            [-1.334554382408856e38, -1.33455489871291e31, 0., -0.038950361148861766, -0.03895036108203678, -0.038950361081361774]]
    for n in range(5):
        assert approx([OE.scac_d(n, t) for t in x], gold[n])
    # entries that the norm-based ≈ does not resolve next to 1e38: checked one by one (orders 0-3, all abscissae)
    for n in range(4):
        for t, g in zip(x, gold[n]):
            assert abs(OE.scac_d(n, t) - g) <= 1e-7 * abs(g)


def test_rodrigues_roundtrip():
    for v in ([.1, .2, .3], [1e-7, 2e-7, 1e-8], [0., 0., 0.]):
        M, w, dw = OE.rodrigues_roundtrip(v)
        assert approx(w, v, atol=1e-20) and approx(dw, np.eye(3)) and approx(M @ M.T, np.eye(3))


# ------------------------------------------------------------------------------------------------ test/TestBeamElement.jl
def test_beam_constructor():
    L = 5
    b = OE.beam_ctor([0, 0, 0], [4, 3, 0], OE.beam_cross_section(EA=10., EI2=3., EI3=3., GJ=4., mu=1., iota1=1.0))
    f = lambda n: OE.beam_field(b, n)
    assert approx(f("cm"), [2.0, 1.5, 0.0]) and approx(f("rm"), [[0.8, -0.6, 0.0], [0.6, 0.8, -0.0], [0.0, 0.0, 1.0]])
    assert approx(f("zgp"), [-0.4305681557970263, -0.16999052179242816, 0.16999052179242816, 0.4305681557970263])
    assert approx(f("znod"), [-0.5, 0.5]) and approx(f("tgm"), [4.0, 3.0, 0.0]) and approx(f("tge"), [5.0, 0.0, 0.0])
    assert approx(f("ya"), [-0.8611363115940526, -0.3399810435848563, 0.3399810435848563, 0.8611363115940526])
    assert approx(f("yu"), [-0.972414176921822, -0.49032285223640754, 0.49032285223640754, 0.972414176921822])
    assert approx(f("yv"), np.array([-0.0646110632135477, -0.221103222500738, -0.221103222500738, -0.0646110632135477]) * L)
    assert approx(f("ka"), np.array([2.0, 2.0, 2.0, 2.0]) / L) and approx(f("kv"), np.array([2.0, 2.0, 2.0, 2.0]) / L)
    assert approx(f("ku"), np.array([10.333635739128631, 4.079772523018276, -4.079772523018276, -10.333635739128631]) / L ** 2)
    assert approx(f("dL"), np.array([0.17392742256872692, 0.3260725774312731, 0.3260725774312731, 0.17392742256872692]) * L)


EA, EI2, EI3, GJ, L, mu, i1 = 10., 3., 2., 4., 2., 1., 5.


def _beam(**kw):
    return OE.beam_ctor([0, 0, 0], [L, 0, 0], OE.beam_cross_section(EA=EA, EI2=EI2, EI3=EI3, GJ=GJ, mu=mu, iota1=i1, **kw))


def _R(b, *X):
    return OE.beam_residual(b, np.array(X))[0]


def test_beam_residuals():
    b = _beam()
    x = np.zeros(12); x[6] = 0.1
    assert approx(_R(b, x), [-EA / L * .1, 0, 0, 0, 0, 0, EA / L * .1, 0, 0, 0, 0, 0])
    x = np.zeros(12); x[7] = 0.01
    assert approx(_R(b, x), [0, -12 * EI3 / L ** 3 * .01, 0, 0, 0, -6 * EI3 / L ** 2 * .01, 0, 12 * EI3 / L ** 3 * .01, 0, 0, 0, -6 * EI3 / L ** 2 * .01], atol=1e-2)
    x = np.zeros(12); x[11] = 0.01
    assert approx(_R(b, x), [0, 6 * EI3 / L ** 2 * .01, 0, 0, 0, 2 * EI3 / L * .01, 0, -6 * EI3 / L ** 2 * .01, 0, 0, 0, 4 * EI3 / L * .01], atol=1e-2)
    x = np.zeros(12); x[9] = 0.1
    assert approx(_R(b, x), [0, 0, 0, -GJ / L * .1, 0, 0, 0, 0, 0, GJ / L * .1, 0, 0])


def test_beam_K_and_M_closed_form():
    """diffed_residual at X=(0,0,0): K and M against the closed-form Euler-beam matrices (TestBeamElement.jl:84-239)"""
    R, (K, C, M), _ = OE.diffed_residual(_beam(), np.zeros((3, 12)))
    k = lambda i, j: K[i - 1, j - 1]; m = lambda i, j: M[i - 1, j - 1]
    ap = lambda a, b: abs(a - b) <= RT * max(abs(a), abs(b))
    assert ap(k(1, 1), EA / L) and ap(k(7, 7), EA / L) and ap(k(1, 7), -EA / L) and ap(k(7, 1), -EA / L)
    for (i, j, v) in [(2, 2, 12 * EI3 / L ** 3), (8, 8, 12 * EI3 / L ** 3), (3, 3, 12 * EI2 / L ** 3), (9, 9, 12 * EI2 / L ** 3), (3, 9, -12 * EI2 / L ** 3),
                      (9, 3, -12 * EI2 / L ** 3), (2, 8, -12 * EI3 / L ** 3), (8, 2, -12 * EI3 / L ** 3), (8, 6, -6 * EI3 / L ** 2), (2, 12, 6 * EI3 / L ** 2),
                      (9, 5, 6 * EI2 / L ** 2), (3, 11, -6 * EI2 / L ** 2), (5, 5, 4 * EI2 / L), (11, 11, 4 * EI2 / L), (6, 6, 4 * EI3 / L), (12, 12, 4 * EI3 / L),
                      (5, 11, 2 * EI2 / L), (11, 5, 2 * EI2 / L), (6, 12, 2 * EI3 / L), (12, 6, 2 * EI3 / L), (5, 9, 6 * EI2 / L ** 2), (11, 3, -6 * EI2 / L ** 2),
                      (6, 8, -6 * EI3 / L ** 2), (12, 2, 6 * EI3 / L ** 2), (4, 4, GJ / L), (10, 10, GJ / L), (4, 10, -GJ / L), (10, 4, -GJ / L)]:
        assert ap(k(i, j), v), (i, j, k(i, j), v)
    for (i, j, v) in [(1, 1, mu * L / 3), (7, 7, mu * L / 3), (1, 7, mu * L / 6), (7, 1, mu * L / 6), (8, 8, 156 * mu * L / 420), (2, 2, 156 * mu * L / 420),
                      (3, 3, 156 * mu * L / 420), (9, 9, 156 * mu * L / 420), (3, 9, 54 * mu * L / 420), (9, 3, 54 * mu * L / 420), (2, 8, 54 * mu * L / 420),
                      (8, 2, 54 * mu * L / 420), (8, 6, 13 * mu * L ** 2 / 420), (2, 12, -13 * mu * L ** 2 / 420), (9, 5, -13 * mu * L ** 2 / 420),
                      (3, 11, 13 * mu * L ** 2 / 420), (5, 5, 4 * mu * L ** 3 / 420), (11, 11, 4 * mu * L ** 3 / 420), (6, 6, 4 * mu * L ** 3 / 420),
                      (12, 12, 4 * mu * L ** 3 / 420), (5, 11, -3 * mu * L ** 3 / 420), (11, 5, -3 * mu * L ** 3 / 420), (6, 12, -3 * mu * L ** 3 / 420),
                      (12, 6, -3 * mu * L ** 3 / 420), (5, 9, -13 * mu * L ** 2 / 420), (11, 3, 13 * mu * L ** 2 / 420), (6, 8, 13 * mu * L ** 2 / 420),
                      (12, 2, -13 * mu * L ** 2 / 420), (4, 4, i1 * L / 4), (10, 10, i1 * L / 4), (4, 10, i1 * L / 4), (10, 4, i1 * L / 4)]:
        assert ap(m(i, j), v), (i, j, m(i, j), v)
    # spurious stiffness / inertia (TestBeamElement.jl:155-172,223-239)
    for A in (K, M):
        for row, cols in [(1, [2, 3, 4, 5, 6, 8, 9, 10, 11, 12]), (7, [2, 3, 4, 5, 6, 8, 9, 10, 11, 12]), (4, [1, 2, 3, 5, 6, 7, 8, 9, 11, 12]),
                          (10, [1, 2, 3, 5, 6, 7, 8, 9, 11, 12]), (2, [1, 3, 4, 5, 7, 9, 10, 11]), (3, [1, 2, 4, 6, 7, 8, 10, 12]), (8, [1, 3, 4, 5, 7, 9, 10, 11]),
                          (9, [1, 2, 4, 6, 7, 8, 10, 12]), (5, [1, 2, 4, 6, 7, 8, 10, 12]), (6, [1, 3, 4, 5, 7, 9, 10, 11]), (11, [1, 2, 4, 6, 7, 8, 10, 12]),
                          (12, [1, 3, 4, 5, 7, 9, 10, 11])]:
            assert np.linalg.norm(A[row - 1, np.array(cols) - 1]) <= 1e-12
    assert np.abs(R).max() == 0.


def test_beam_weight_addedmass_damping():
    w = 10
    b = _beam(w=w)
    x = np.zeros(12); x[4] = w * L ** 3 / (24 * EI2); x[10] = -w * L ** 3 / (24 * EI2)
    assert approx(_R(b, x), [0, 0, w * L / 2, 0, 0, 0, 0, 0, w * L / 2, 0, 0, 0])
    Ca1, Ca2, Ca3, a1, a2, a3 = 1., 2., 3., 4., 3., 2.
    b = _beam(Ca1=Ca1, Ca2=Ca2, Ca3=Ca3)
    z = np.zeros(12)
    acc = z.copy(); acc[[0, 6]] = a1
    assert approx(_R(b, z, z, acc), [(mu + Ca1) * a1 * L / 2, 0, 0, 0, 0, 0, (mu + Ca1) * a1 * L / 2, 0, 0, 0, 0, 0])
    d = z.copy(); d[5] = -(mu + Ca2) * a2 * L ** 3 / (24 * EI3); d[11] = -d[5]; acc = z.copy(); acc[[1, 7]] = a2
    assert approx(_R(b, d, z, acc), [0, (mu + Ca2) * a2 * L / 2, 0, 0, 0, 0, 0, (mu + Ca2) * a2 * L / 2, 0, 0, 0, 0])
    d = z.copy(); d[4] = (mu + Ca3) * a3 * L ** 3 / (24 * EI2); d[10] = -d[4]; acc = z.copy(); acc[[2, 8]] = a3
    assert approx(_R(b, d, z, acc), [0, 0, (mu + Ca3) * a3 * L / 2, 0, 0, 0, 0, 0, (mu + Ca3) * a3 * L / 2, 0, 0, 0])
    Cl, Cq, v = (1., 2., 3.), (.1, .2, .3), (1.0, 1.1, 0.1)
    b = _beam(Cl1=Cl[0], Cl2=Cl[1], Cl3=Cl[2], Cq1=Cq[0], Cq2=Cq[1], Cq3=Cq[2])
    f = [(Cl[i] + Cq[i] * abs(v[i])) * v[i] for i in range(3)]
    vel = z.copy(); vel[[0, 6]] = v[0]
    assert approx(_R(b, z, vel, z), [f[0] * L / 2, 0, 0, 0, 0, 0, f[0] * L / 2, 0, 0, 0, 0, 0])
    d = z.copy(); d[5] = -f[1] * L ** 3 / (24 * EI3); d[11] = -d[5]; vel = z.copy(); vel[[1, 7]] = v[1]
    assert approx(_R(b, d, vel, z), [0, f[1] * L / 2, 0, 0, 0, 0, 0, f[1] * L / 2, 0, 0, 0, 0])
    d = z.copy(); d[4] = f[2] * L ** 3 / (24 * EI2); d[10] = -d[4]; vel = z.copy(); vel[[2, 8]] = v[2]
    assert approx(_R(b, d, vel, z), [0, 0, f[2] * L / 2, 0, 0, 0, 0, 0, f[2] * L / 2, 0, 0, 0])


# ------------------------------------------------------------------------------------------------ test/TestAssemble.jl (integer pins)
def _dis_assemble():
    # Turbine: X (1,2) A (1,2); AnchorLine: X (1,2,3) A (3,4)    (TestAssemble.jl:53-60)
    e0 = np.zeros((1, 0), np.int64)
    return [dict(X=np.array([[1, 2]]), U=e0, A=np.array([[1, 2]])), dict(X=np.array([[1, 2, 3]]), U=e0, A=np.array([[3, 4]]))]


def test_asmvec_and_prepare_sweepx_maps():
    dis = _dis_assemble()
    gr = OP.allXdofs(3, 0, 4)
    assert gr["jX"].tolist() == [1, 2, 3] and OP.indexedstate(gr)[1].tolist() == [1, 2, 3] and OP.indexedstate(gr)[0].tolist() == [0, 0, 0]
    a = OP.asmvec(gr, dis)
    assert a[0].tolist() == [[1], [2]] and a[1].tolist() == [[1], [2], [3]]                       # :112-117
    asm1, asm2, colptr, rowval = OP.prepare_sweepx(dis, 3, 0, 4)
    assert asm2[0][:, 0].tolist() == [1, 2, 4, 5] and asm2[1][:, 0].tolist() == [1, 2, 3, 4, 5, 6, 7, 8, 9]   # :125-126
    assert colptr.tolist() == [1, 4, 7, 10] and rowval.tolist() == [1, 2, 3, 1, 2, 3, 1, 2, 3]   # the 9-nnz pattern implied by asm[2,2]


# ------------------------------------------------------------------------------------------------ test/TestSparseTools.jl
def test_sparsetools_prepare_and_addin():
    # block = sparse([1,1,1,2,3,4],[1,2,4,2,3,4],ones(6)); pattern = 3 diagonal blocks   (:29-42)
    colptr = np.array([1, 2, 4, 5, 7]); rowval = np.array([1, 1, 2, 3, 1, 4])
    blk = (4, 4, colptr, rowval)
    big, asm, pgr, pgc = OP.sparsetools_prepare(3, 3, [(1, 1, blk), (2, 2, blk), (3, 3, blk)])
    assert asm["colptr"].tolist() == [1, 2, 3, 4] and asm["rowval"].tolist() == [1, 2, 3]
    assert [v.tolist() for v in asm["nzval"]] == [[1, 2, 3, 4, 5, 6], [7, 8, 9, 10, 11, 12], [13, 14, 15, 16, 17, 18]]
    assert pgr.tolist() == [1, 5, 9, 13] and pgc.tolist() == [1, 5, 9, 13]
    nz = np.zeros(18)
    for k in (1, 2, 3):
        OP.addin_block(asm, nz, np.ones(6), k, k)
    assert nz.tolist() == [1.] * 18
    # first example (:8-27): 3×2 pattern with a hole at (1,1); blocks land at the right place
    cp = np.array([1, 2, 5, 6]); rv = np.array([1, 1, 2, 3, 3]); vals = np.arange(1., 6.)
    b3 = (3, 3, cp, rv)
    blocks = [(2, 1, b3), (3, 1, b3), (1, 2, b3), (2, 2, b3), (3, 2, b3)]
    big, asm, pgr, pgc = OP.sparsetools_prepare(3, 2, blocks)
    nz = np.zeros(len(big["rowval"]))
    for (r, c, _) in blocks:
        OP.addin_block(asm, nz, vals, r, c)
    import scipy.sparse as sp
    B = sp.csc_matrix((nz, big["rowval"] - 1, big["colptr"] - 1), shape=(big["m"], big["n"])).toarray()
    blockd = sp.csc_matrix((vals, rv - 1, cp - 1), shape=(3, 3)).toarray()
    assert np.array_equal(B[3:6, 0:3], blockd) and np.array_equal(B[3:6, 3:6], blockd) and np.all(B[0:3, 0:3] == 0)
    with pytest.raises(KeyError):
        OP.find_block(asm, 1, 1)


# ------------------------------------------------------------------------------------------------ test/TestFiniteDifferences.jl
def test_finitediff():
    n, dt = 10, 0.1
    t = np.arange(n) * dt
    x = 1 + t + 0.5 * t ** 2 + 1 / 3 * t ** 3
    x1 = [np.zeros(n) for _ in range(3)]
    for order in range(3):
        for s in range(1, n + 1):
            for (ds, w) in OP.finitediff(order, n, s):
                x1[order][s - 1] += x[s + ds - 1] * w / dt ** order
    assert approx(x1[0], x)
    assert approx(300 * x1[1], [316, 334, 373, 418, 469, 526, 589, 658, 733, 772])
    assert approx(10 * x1[2], [12, 12, 14, 16, 18, 20, 22, 24, 26, 26])
    with pytest.raises(ValueError):
        OP.finitediff(1, 5, 1)


# ------------------------------------------------------------------------------------------------ test/TestDirectXUA.jl (integer pins)
def _dis_directxua():
    """dof structure of the TestDirectXUA.jl model (:30-51): El1×2, Spring{1}, 6 SingleAcost types… only doflists matter here.
    X: n1.tx1=1, n2.tx1=2 ; U: n1.u=1, n2.u=2, n1.utx1=3, n2.utx1=4 ; A: n1.ΞC=1, n1.ΞM=2, n2.ΞC=3, n2.ΞM=4, n3.ΞL₀=5, n3.ΞEI=6"""
    I = lambda *r: np.array(r, np.int64).reshape(len(r), -1)
    e = lambda n: np.zeros((n, 0), np.int64)
    return [dict(X=I([1], [2]), U=I([1], [2]), A=I([1, 2], [3, 4])),      # El1 (two elements)
            dict(X=I([1, 2]), U=e(1), A=I([5, 6])),                        # Spring{1}
            dict(X=e(1), U=e(1), A=I([5])), dict(X=e(1), U=e(1), A=I([6])),   # SingleAcost ΞL₀, ΞEI (n3)
            dict(X=e(2), U=e(2), A=I([1], [3])), dict(X=e(2), U=e(2), A=I([2], [4])),   # SingleAcost ΞC (n1,n2), ΞM (n1,n2)
            dict(X=I([1], [2]), U=I([3], [4]), A=e(2)),                    # SingleUdof (tx1, utx1) ×2
            dict(X=I([1]), U=e(1), A=e(1)), dict(X=I([2]), U=e(1), A=e(1))]   # SingleDofCost l1, l2 (distinct functor types)


def test_directxua_prepare_and_big_pattern():
    dis = _dis_directxua()
    P = OP.prepare_direct(dis, 2, 4, 6, OX=2, OU=0, IA=1)
    asm = P["asm"]
    T = lambda a: a.tolist()
    assert T(asm[1][0]) == [[1, 2]] and T(asm[1][1]) == [[1], [2]]                 # TestDirectXUA.jl:110-111
    assert T(asm[2][0]) == [[1, 2]] and T(asm[2][1]) == [[1], [2]]                 # :112-113
    assert T(asm[3][0]) == [[1, 2]] and asm[3][1].shape == (0, 1)                  # :114-115
    assert T(asm[4][0]) == [[1, 3], [2, 4]] and T(asm[4][1]) == [[5], [6]]         # :116-117
    assert T(asm[5][0]) == [[1, 4]]                                                # :118
    assert T(asm[20][0]) == [[1, 5], [2, 6], [3, 7], [4, 8]]                       # :119
    big, bigasm, pgr, pgc = OP.preparebig(1, [6], P["nL2"], P["pat"])
    assert big["m"] == 54 and big["n"] == 54                                       # :123-124
    assert big["colptr"][:50].tolist() == [1, 13, 25, 43, 61, 68, 75, 80, 85, 97, 109, 133, 157, 164, 171, 176, 181, 193, 205, 235, 265, 272, 279, 284, 289,
                                            301, 313, 343, 373, 380, 387, 392, 397, 409, 421, 445, 469, 476, 483, 488, 493, 505, 517, 535, 553, 560, 567,
                                            572, 577, 597]                         # :125
    assert big["rowval"][:60].tolist() == [3, 4, 5, 7, 11, 12, 19, 20, 49, 50, 53, 54, 3, 4, 6, 8, 11, 12, 19, 20, 51, 52, 53, 54, 1, 2, 3, 4, 5, 7, 9, 10, 11,
                                            12, 13, 15, 19, 20, 49, 50, 53, 54, 1, 2, 3, 4, 6, 8, 9, 10, 11, 12, 14, 16, 19, 20, 51, 52, 53, 54]   # :126
    assert pgr.tolist() == [1, 3, 5, 9, 11, 13, 17, 19, 21, 25, 27, 29, 33, 35, 37, 41, 43, 45, 49, 55]          # :130-131
    assert bigasm["colptr"].tolist() == [1, 6, 14, 20, 25, 36, 42, 47, 61, 67, 72, 86, 92, 97, 108, 114, 119, 127, 133, 152]   # :132
    assert bigasm["nzval"][0].tolist() == [1, 2, 13, 14] and bigasm["nzval"][3].tolist() == [7, 8, 19, 20]       # :133-134
    assert bigasm["nzval"][150].tolist() == [595, 596, 615, 616, 635, 636, 655, 656, 681, 682, 707, 708]         # :135


def test_big_pattern_repeats_over_interior_windows():
    """The property mb_direct_rebase rests on, checked on the restatement of makepattern / SparseTools.prepare (src/DirectXUA.jl:245-307,
    src/SparseTools.jl:32-94): away from the first and the last step the finite-difference stencils are central (src/FiniteDifferences.jl:8-31), so the
    Lvv columns of the steps [lo,lo+L), 3 ≤ lo, lo+L ≤ nstep−3 (0-based) have the same structure for every lo, up to a shift of the row numbers by W per step."""
    dis = _dis_directxua()
    nX, nU, nstep, L = 2, 4, 14, 3
    P = OP.prepare_direct(dis, nX, nU, 6, OX=2, OU=0, IA=0)
    big, bigasm, pgr, pgc = OP.preparebig(0, [nstep], P["nL2"], P["pat"])
    W = 2 * nX + nU
    assert big["n"] == nstep * W
    cp, rv = big["colptr"], big["rowval"]

    def window(lo):
        c0, c1 = lo * W, (lo + L) * W
        p0, p1 = cp[c0] - 1, cp[c1] - 1
        return cp[c0:c1 + 1] - cp[c0], rv[p0:p1] - lo * W
    ref = window(3)
    for lo in range(4, nstep - 3 - L + 1):
        w = window(lo)
        assert np.array_equal(w[0], ref[0]) and np.array_equal(w[1], ref[1]), lo
    for lo in (0, 2, nstep - 2 - L, nstep - L):            # windows that reach the one-sided stencils at either end differ
        w = window(lo)
        assert not (np.array_equal(w[0], ref[0]) and np.array_equal(w[1], ref[1])), lo


# ------------------------------------------------------------------------------------------------ test/TestBarElement.jl
def test_bar_element_goldens():
    EAb, L0, mub = 10., 2., 1.
    b = OE.bar_ctor([0, 0, 0], [L0, 0, 0], OE.bar_cross_section(EA=EAb, mu=mub))
    assert approx(b[0:3], [L0 / 2, 0, 0]) and approx(b[3:6], [L0, 0, 0]) and approx(b[9], L0)
    assert approx(b[20:24], [0.34785484513745385, 0.6521451548625462, 0.6521451548625462, 0.34785484513745385])
    assert approx(b[24:28], [-0.4305681557970263, -0.16999052179242816, 0.16999052179242816, 0.4305681557970263])
    assert approx(b[30:34], [0.9305681557970262, 0.6699905217924281, 0.33000947820757187, 0.06943184420297371])
    assert approx(b[34:38], [0.06943184420297371, 0.33000947820757187, 0.6699905217924281, 0.9305681557970262])
    dx = .1
    x = np.array([0, 0, 0, dx, 0, 0.])
    assert approx(OE.bar_residual(b, [x])[0], [-EAb / L0 * dx, 0, 0, EAb / L0 * dx, 0, 0])
    X = np.zeros((3, 6)); X[0] = x
    seed = np.zeros((3, 6, 18))
    for d in range(3):
        seed[d, np.arange(6), 6 * d + np.arange(6)] = 1
    R, dR, rc = OE.bar_residual(b, X, seed)
    K, M = dR[:, :6], dR[:, 12:]
    kt = (EAb / L0) * dx / (L0 + dx)
    ap = lambda a, c: abs(a - c) <= RT * max(abs(a), abs(c))
    assert ap(K[0, 0], EAb / L0) and ap(K[3, 3], EAb / L0) and ap(K[0, 3], -EAb / L0) and ap(K[3, 0], -EAb / L0)
    for (i, j, v) in [(1, 1, kt), (2, 2, kt), (4, 4, kt), (5, 5, kt), (1, 4, -kt), (4, 1, -kt), (2, 5, -kt), (5, 2, -kt)]:
        assert ap(K[i, j], v)
    assert np.linalg.norm(K[np.ix_([0, 3], [1, 2, 4, 5])]) < 1e-12 and np.linalg.norm(K[np.ix_([1, 4], [0, 2, 3, 5])]) < 1e-12
    for i in range(6):
        assert ap(M[i, i], mub * L0 / 3) and ap(M[i, (i + 3) % 6], mub * L0 / 6)
    w = 10
    bw = OE.bar_ctor([0, 0, 0], [L0, 0, 0], OE.bar_cross_section(EA=EAb, mu=mub, w=w))
    assert approx(OE.bar_residual(bw, [np.zeros(6)], t=0.)[0], [0, 0, w * L0 / 2, 0, 0, w * L0 / 2])


def test_soil_contact_branches():
    """toolbox/SoilContact.jl:14-18: springs/dampers only below z₀; above, an integer-zero residual without partials"""
    p = np.array([0.5, 30., 200., 3., 7.])
    X = np.array([[.1, .2, .3], [1., 2., 3.]]); seed = np.zeros((2, 3, 3)); seed[0, np.arange(3), np.arange(3)] = 1.; seed[1] = 2.5 * seed[0]
    R, dR, c = OE.soil_residual(p, X, seed)
    assert c == 1 and approx(R, [30 * .1 + 3 * 1, 30 * .2 + 3 * 2, 200 * (.3 - .5) + 7 * 3]) and approx(dR, np.diag([30 + 7.5, 30 + 7.5, 200 + 17.5]))
    X[0, 2] = 0.5
    R, dR, c = OE.soil_residual(p, X, seed)
    assert c == 0 and np.all(R == 0) and np.all(dR == 0)


# ------------------------------------------------------------------------------------------------ test/TestSparseTools.jl:58-76
def test_sparser_golden():
    import scipy.sparse as sp
    i = np.array([3, 7, 2, 3, 6, 2, 7, 9, 2, 6, 4, 5, 9, 3, 9, 1, 7, 8, 10, 4, 9, 7])
    j = np.array([2, 2, 3, 3, 3, 4, 4, 4, 6, 6, 7, 7, 7, 8, 8, 9, 9, 9, 9, 10, 10, 11])
    v = np.array([0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0])
    s = sp.csc_matrix((v + 10., (i - 1, j - 1)), shape=(10, 11))        # +10: keep the explicit zeros of Julia's sparse(i,j,v) stored
    s.sort_indices()
    nz = s.data - 10.
    colptr, rowval, nzval = OP.sparser(s.indptr.astype(np.int64) + 1, s.indices.astype(np.int64) + 1, nz, rtol=0.6)   # keep ⇔ v > 0.5 for v ∈ {0,1}
    assert colptr.tolist() == [1, 1, 2, 3, 5, 5, 5, 6, 6, 7, 8, 8]
    assert rowval.tolist() == [7, 6, 2, 7, 4, 1, 9]
    assert np.array_equal(nzval, np.ones(7))


# ------------------------------------------------------------------------------------------------ test/TestBeamElementStrainGauge.jl:30-56
def test_beam_requestables_goldens():
    """♢κ and the axial strain of the beam as read by StrainGaugeOnEulerBeam3D (eleres.κ, eleres.εₐₓ; atol = 1e-12 as in the reference)"""
    b = OE.beam_ctor([0, 0, 0], [4, 0, 0], OE.beam_cross_section(EA=10., EI2=3., EI3=3., GJ=4., mu=1., iota1=1.0))
    for X, kap in [([0, 0, 0, 0, .1, 0, 0, 0, 0, 0, -.1, 0], [0, 0, 1 / 20]),
                   ([0, 0, 0, 0, 0, .1, 0, 0, 0, 0, 0, -.1], [0, -1 / 20, 0]),
                   ([0, 0, 0, 0, 0, 0, 0, 0, 0, 1., 0, 0], [.25, 0, 0])]:
        r = OE.beam_results(b, np.array([X], float))
        assert abs(r[0]) <= 1e-12                                   # εₐₓ
        assert np.abs(r[10:13] - np.array(kap)).max() <= 1e-12       # ♢κ
        # Gauss-point curvature κgp[1] (torsion rate) is uniform along the element and equals ♢κ[1]
        assert np.abs(r[13 + 16 * np.arange(4) + 3] - kap[0]).max() <= 1e-12


# ---------------------------------------------------------------------------------------------------------------------------------------------
# adiff.hpp / Taylor restatements pinned DIRECTLY on the reference's unit-test expressions (not only through the beam matrices they produce)
def _kat(name, n):
    import ctypes as C
    L = OE.lib()
    fn = getattr(L, name)
    fn.argtypes = [np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")]
    fn.restype = C.c_int
    out = np.full(n, np.nan)
    assert fn(out) == n
    return out


def test_adiff_operations_kat():
    """test/TestAdiff.jl:40-52 and :107-121 — nested duals of different precedence: od = oa+oc, oj = od^2 === od*od, ok = oc*oa/oc, the `x^0 → zero` rule
    (Adiff.jl:230) and norm(variate{1,3}(ox)) === sqrt(sum(oX.^2))"""
    o = _kat("orc_kat_adiff", 22)
    od, oj, odod, ok, z0, nrm = o[0:4], o[4:8], o[8:12], o[12:16], o[16:18], o[18:22]
    assert od.tolist() == [4., 1., 1., 0.]                 # ∂ℝ{3,1,∂𝕣11}(∂𝕣11(4.0,[1.0]), [∂𝕣11(1.0,[0.0])])        TestAdiff.jl:110
    assert oj.tolist() == odod.tolist() == [16., 8., 8., 2.]   # oj === od*od                                        TestAdiff.jl:113
    assert ok.tolist() == [1., 1., 0., 0.]                 # ∂ℝ{3,1,∂𝕣11}(∂𝕣11(1.0,[1.0]), [∂𝕣11(0.0,[0.0])])        TestAdiff.jl:115
    assert z0.tolist() == [0., 0.]                         # variate{1}(0.)^0 === ∂ℝ{1,1,𝕣}(0.,[0.])                  TestAdiff.jl:117
    x = np.array([1., 2., 3.]); s = np.sqrt(14.)
    assert nrm[0] == s and np.array_equal(nrm[1:], (2. * x) / 2. / s)                                                # TestAdiff.jl:121


def test_taylor_motion_roundtrip_kat():
    """test/TestTaylor.jl:5-66 — motion{P}(X) packs (x,x′,x″) into nested one-partial duals, motion⁻¹ unpacks value / velocity / acceleration
    (zero for the orders the tuple does not hold), partials of the solver's seeding carried along"""
    o = _kat("orc_kat_motion", 108).reshape(3, 3, 3, 4)              # [ND-1][component][ider][value + 3 partials]
    for nd in (1, 2, 3):
        for c in range(3):
            for d in range(3):
                exp = np.zeros(4)
                if d < nd:
                    exp[0] = 3 * d + c + 1; exp[1 + c] = 1.        # variate{1,3}(SVector(1,2,3)) etc.
                assert np.array_equal(o[nd - 1, c, d], exp), (nd, c, d)


def test_taylor_revariate_and_chainrule_kat():
    """test/TestTaylor.jl:69-74 (revariate{2}) and :119-131 ("chainrule NamedTuple": q == q2) on the oracle's revariate2 / McLaurin"""
    o = _kat("orc_kat_chainrule", 51)
    assert o[:9].tolist() == [3., 1., 0., 1., 0., 0., 0., 0., 0.]   # ∂ℝ{2,2,∂ℝ{1,2}}(∂ℝ{1,2}(3,[1,0]), [∂ℝ{1,2}(1,[0,0]), ∂ℝ{1,2}(0,[0,0])])
    q, q2 = o[9:30], o[30:51]
    assert np.array_equal(q, q2)                                     # TestTaylor.jl:130
    assert q2[0] == 1. + 4. + 6.25 + 9. + 4. and q2[1:5].tolist() == [2., 4., 5., 6.] and not q2[5:].any()


def test_static_dual_build_is_bit_identical():
    """liboracle_np12.so (the -O3 build with 12 compile-time partials that bench.py's CPU arm times) gives the very bits of the checker build"""
    import muscade_b200 as mb
    n = 300
    for OX in (0, 2):
        eleobj, idx, ndof = mb.synthetic.chain(n, dynamic=OX > 0)
        X = mb.synthetic.state(ndof, nder=OX + 1); nm = mb.synthetic.newmark_coefficients(OX, 0.3)
        dis = [dict(X=idx, U=np.zeros((n, 0), np.int64), A=np.zeros((n, 0), np.int64))]
        a1, a2, cp, rv = OP.prepare_sweepx(dis, ndof, 0, 0)
        a1t, a2t = np.ascontiguousarray(a1[0].T), np.ascontiguousarray(a2[0].T)
        res = []
        for fast in (False, True):
            L = np.zeros(ndof); nz = np.zeros(len(rv))
            OE.sweepx_assemble_beams_mt(eleobj, idx, a1t, a2t, OX, X, np.ones(12), nm, L, nz, 2, static_duals=fast)
            res.append((L, nz))
        assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
