"""The PRODUCT's kernel formulation (muscade.jl_b200/csrc/beam_math.cuh compiled for the host by tests/csrc) against the
oracle's literal nested-dual algorithm — no GPU needed.  Tolerance 1e-12 (north_star), observed ~1e-15."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import elements as OE

HERE = os.path.dirname(os.path.abspath(__file__))
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def H():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "csrc"), "-s"])
    L = C.CDLL(os.path.join(HERE, "csrc", "libhostmath.so"))
    L.mbh_beam_residual.argtypes = [f64p, f64p, C.c_int, C.c_int, C.c_int, f64p, f64p, C.c_int, f64p, f64p, f64p, f64p]
    L.mbh_beam_iter_sd.argtypes = [f64p, f64p, C.c_int, f64p, f64p, C.c_double, C.c_double, C.c_int, f64p, f64p, f64p]
    return L


def geo16(e):
    return np.ascontiguousarray(np.concatenate([e[0:3], e[3:12], e[18:21], e[48:49]]))


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


MAT = dict(EA=10, EI2=3, EI3=2.5, GJ=4, mu=1, iota1=1.3, w=0.7, Ca1=.3, Ca2=169.6, Ca3=150., Cl1=.2, Cl2=.5, Cl3=.7, Cq1=.1, Cq2=235.2, Cq3=200.)


@pytest.mark.parametrize("nd,w", [(1, 1), (1, 4), (1, 12), (2, 1), (3, 1), (3, 3)])
@pytest.mark.parametrize("amp", [0.0, 1.0, 8.0])
def test_dense_dual_path_vs_oracle(H, nd, w, amp):
    """diffed_residual-style seeding (all of X₀,X₁,X₂,U₀), dense Dual<W> lanes"""
    rng = np.random.default_rng(11 * nd + w)
    e = OE.beam_ctor([0, 0, 0], [.8, .6, 0.1], OE.beam_cross_section(**MAT), orient2=(0, 1, .2))
    X = np.zeros((nd, 12))
    X[0] = rng.uniform(-1, 1, 12) * np.array([.05] * 3 + [.1] * 3 + [.05] * 3 + [.1] * 3) * amp
    for d in range(1, nd):
        X[d] = rng.uniform(-1, 1, 12) * 0.1 * (amp > 0)
    npx = 12 * nd + 3
    Xs = np.zeros((nd, 12, npx))
    for d in range(nd):
        Xs[d, np.arange(12), 12 * d + np.arange(12)] = 1.0
    U = rng.uniform(-1, 1, 3); Us = np.zeros((3, npx)); Us[np.arange(3), 12 * nd + np.arange(3)] = 1
    R0, dR0, rc = OE.beam_residual(e, X, Xs, U, Us)
    R1 = np.zeros(12); dR1 = np.zeros((12, npx))
    assert H.mbh_beam_residual(geo16(e), np.ascontiguousarray(e[53:69]), nd, w, npx, X, Xs, 1, U, Us, R1, dR1) == 0
    assert rc == 0 and rel(R1, R0) <= 1e-12 and rel(dR1, dR0) <= 1e-12


@pytest.mark.parametrize("nd", [1, 2, 3])
@pytest.mark.parametrize("amp", [0.0, 1.0, 3.0])
def test_sparse_dual_lanes_vs_oracle(H, nd, amp):
    """the kernel's 6-lane statically sparse path with the SweepX :iter seeding X₀+δX, X₁+a₁δX, X₂+b₁δX and non-unit scales"""
    rng = np.random.default_rng(5 + nd)
    e = OE.beam_ctor([0, 0, 0], [.8, .6, 0.1], OE.beam_cross_section(**MAT), orient2=(0, 1, .2))
    scale = np.array([10., 10, 10, 1, 1, 1] * 2); a1, b1 = 3.3, 7.1
    X = np.zeros((3, 12)); X[0] = rng.uniform(-1, 1, 12) * 0.3 * amp
    for d in range(1, nd):
        X[d] = rng.uniform(-1, 1, 12) * 0.1 * (amp > 0)
    U = rng.uniform(-1, 1, 3)
    Xs = np.zeros((nd, 12, 12))
    for d in range(nd):
        Xs[d, np.arange(12), np.arange(12)] = scale * [1, a1, b1][d]
    R0, dR0, rc = OE.beam_residual(e, X[:nd], Xs, U, np.zeros((3, 12)))
    R1 = np.zeros(12); K1 = np.zeros((12, 12))
    H.mbh_beam_iter_sd(geo16(e), np.ascontiguousarray(e[53:69]), nd, np.ascontiguousarray(X), scale, a1, b1, 1, U, R1, K1)
    assert rel(R1, R0) <= 1e-12 and rel(K1, dR0) <= 1e-12


@pytest.mark.parametrize("nd", [2, 3])
@pytest.mark.parametrize("amp", [0.0, 1.0, 3.0, 8.0])
def test_two_phase_active_passive_lanes_vs_oracle(H, nd, amp):
    """the Newmark / DirectXUA lanes as beam_cot_kernel + beam_kernel_sd<ND,split> run them: phase A (time-jets → cotangents) and phase B (order-0
    forward + reverse), both in the active/passive formulation — the lane's rotation and translation dof belong to one node, the other node's rotation
    is carried as plain numbers (and plain jets)"""
    H.mbh_beam_iter_split.argtypes = [f64p, f64p, C.c_int, f64p, f64p, C.c_double, C.c_double, C.c_int, f64p, f64p, f64p]
    rng = np.random.default_rng(50 + nd + int(amp))
    e = OE.beam_ctor([0, 0, 0], [.8, .6, 0.1], OE.beam_cross_section(**MAT), orient2=(0, 1, .2))
    scale = np.array([10., 10, 10, 1, 1, 1] * 2); a1, b1 = 3.3, 7.1
    X = np.zeros((3, 12)); X[0] = rng.uniform(-1, 1, 12) * 0.3 * amp
    for d in range(1, nd):
        X[d] = rng.uniform(-1, 1, 12) * 0.1 * (amp > 0)
    U = rng.uniform(-1, 1, 3)
    Xs = np.zeros((nd, 12, 12))
    for d in range(nd):
        Xs[d, np.arange(12), np.arange(12)] = scale * [1, a1, b1][d]
    R0, dR0, rc = OE.beam_residual(e, X[:nd], Xs, U, np.zeros((3, 12)))
    R1 = np.zeros(12); K1 = np.zeros((12, 12))
    assert H.mbh_beam_iter_split(geo16(e), np.ascontiguousarray(e[53:69]), nd, np.ascontiguousarray(X), scale, a1, b1, 1, U, R1, K1) == 0
    assert rel(R1, R0) <= 1e-12 and rel(K1, dR0) <= 1e-12, (rel(R1, R0), rel(K1, dR0))


@pytest.mark.parametrize("amp", [0.0, 1.0, 3.0, 8.0])
@pytest.mark.parametrize("udof", [0, 1])
def test_static_symmetric_tangent_vs_oracle(H, amp, udof):
    """statics kernel (beam_static_sym): rotation columns by forward-over-reverse, their transposes for the rotation rows of the translation
    columns (the static residual is a gradient ⇒ symmetric tangent), translation×translation block in closed form; non-unit scales"""
    H.mbh_beam_static_sym.argtypes = [f64p, f64p, f64p, f64p, C.c_int, f64p, f64p, f64p]
    rng = np.random.default_rng(77 + int(amp))
    e = OE.beam_ctor([0, 0, 0], [.8, .6, 0.1], OE.beam_cross_section(**MAT), orient2=(0, 1, .2))
    scale = np.array([10., 20, 5, 1, 2, 3, 7, 10, 4, .5, 1, 2])
    X = np.zeros((1, 12)); X[0] = rng.uniform(-1, 1, 12) * 0.3 * amp
    U = rng.uniform(-1, 1, 3) * udof
    Xs = np.zeros((1, 12, 12)); Xs[0, np.arange(12), np.arange(12)] = scale
    R0, dR0, rc = OE.beam_residual(e, X, Xs, U, np.zeros((3, 12)))
    K0 = dR0 * scale[:, None]
    R1 = np.zeros(12); K1 = np.zeros((12, 12))
    assert H.mbh_beam_static_sym(geo16(e), np.ascontiguousarray(e[53:69]), np.ascontiguousarray(X[0]), scale, udof, U, R1, K1) == 0
    assert rel(R1, R0 * scale) <= 1e-12 and rel(K1, K0) <= 1e-12
    assert np.abs(K0 - K0.T).max() <= 1e-12 * np.abs(K0).max()          # the property the kernel relies on, checked on the ORACLE's tangent


@pytest.mark.parametrize("amp", [0.0, 1.0, 3.0, 8.0])
@pytest.mark.parametrize("udof", [0, 1])
def test_static_active_passive_vs_oracle(H, amp, udof):
    """statics kernel, active/passive formulation (beam_static_ap): each lane's seeded rotation dof belongs to its ACTIVE node; the corotated frame is
    rebuilt from N = r_a r_p^T so that everything depending on the passive node only is plain values. Same checks as the symmetric kernel."""
    H.mbh_beam_static_ap.argtypes = [f64p, f64p, f64p, f64p, C.c_int, f64p, f64p, f64p]
    rng = np.random.default_rng(77 + int(amp))
    e = OE.beam_ctor([0, 0, 0], [.8, .6, 0.1], OE.beam_cross_section(**MAT), orient2=(0, 1, .2))
    scale = np.array([10., 20, 5, 1, 2, 3, 7, 10, 4, .5, 1, 2])
    X = np.zeros((1, 12)); X[0] = rng.uniform(-1, 1, 12) * 0.3 * amp
    U = rng.uniform(-1, 1, 3) * udof
    Xs = np.zeros((1, 12, 12)); Xs[0, np.arange(12), np.arange(12)] = scale
    R0, dR0, rc = OE.beam_residual(e, X, Xs, U, np.zeros((3, 12)))
    K0 = dR0 * scale[:, None]
    R1 = np.zeros(12); K1 = np.zeros((12, 12))
    assert H.mbh_beam_static_ap(geo16(e), np.ascontiguousarray(e[53:69]), np.ascontiguousarray(X[0]), scale, udof, U, R1, K1) == 0
    assert rel(R1, R0 * scale) <= 1e-12 and rel(K1, K0) <= 1e-12


def test_reference_pow_quirk_is_reproduced(H):
    """src/Adiff.jl:230 (x^0 → zero) drops 2·ċ² from d²/dt² of sinc1(θ/2)² in Rodrigues for 3-deep duals (SweepX{2}, DirectXUA OX=2).
    The oracle has it by construction; the product reproduces it (sqr_ref). A mathematically 'correct' square differs at 1e-6."""
    rng = np.random.default_rng(2)
    e = OE.beam_ctor([0, 0, 0], [.8, .6, 0.1], OE.beam_cross_section(EA=10, EI2=3, EI3=2.5, GJ=4, mu=1, iota1=0), orient2=(0, 1, .2))
    X = np.zeros((3, 12)); X[0] = rng.uniform(-1, 1, 12) * 0.3; X[1] = rng.uniform(-1, 1, 12) * 0.5
    R0, _, _ = OE.beam_residual(e, X, np.zeros((3, 12, 0)))
    R1 = np.zeros(12); d = np.zeros((12, 1))
    H.mbh_beam_residual(geo16(e), np.ascontiguousarray(e[53:69]), 3, 1, 0, X, np.zeros((3, 12, 0)), 0, np.zeros(3), np.zeros((3, 0)), R1, d)
    assert rel(R1, R0) <= 1e-12


@pytest.mark.parametrize("nd", [1, 2, 3])
@pytest.mark.parametrize("amp", [0.0, 1.0, 3.0])
def test_requestables_vs_oracle(H, nd, amp):
    """getresult values (ε, rₛₘ, ♢κ and per Gauss point x, κgp, fᵢ, mᵢ, fₑ, mₑ): product math vs the oracle's espy capture"""
    H.mbh_beam_results.argtypes = [f64p, f64p, C.c_int, f64p, f64p]
    rng = np.random.default_rng(21 + nd)
    e = OE.beam_ctor([0, 0, 0], [.8, .6, 0.1], OE.beam_cross_section(**MAT), orient2=(0, 1, .2))
    X = np.zeros((nd, 12)); X[0] = rng.uniform(-1, 1, 12) * 0.3 * amp
    for d in range(1, nd):
        X[d] = rng.uniform(-1, 1, 12) * 0.1 * (amp > 0)
    ref = OE.beam_results(e, X)
    out = np.zeros(77)
    assert H.mbh_beam_results(geo16(e), np.ascontiguousarray(e[53:69]), nd, np.ascontiguousarray(X), out) == 0
    scale = max(np.abs(ref).max(), 1.)
    assert np.abs(out - ref).max() <= 1e-12 * scale, np.abs(out - ref).argmax()
