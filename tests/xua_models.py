"""Test elements and models of the general DirectXUA path: test/TestDirectXUA.jl:10-51 and test/SomeElements.jl:164-188 restated against adiff2.D2."""
import numpy as np

import muscade_b200 as mb
from muscade_b200.adiff2 import exp10, sqrt, sin, cos


class El1(mb.LagrangianElement):
    """El1 (test/TestDirectXUA.jl:10-21): r = −u + K·x + C·10^ΞC·x′ + M·10^ΞM·x″; no_second_order is the default Val(false)"""
    type_parameters = ()

    @classmethod
    def doflist(cls, **kw):
        return (1, 1, 1, 1), ("X", "U", "A", "A"), ("tx1", "u", "ΞC", "ΞM")

    @classmethod
    def construct(cls, coords, K, C, M):
        n = coords[0].shape[0]
        return np.tile(np.array([[K, C, M]], float), (n, 1))

    @staticmethod
    def residual(o, extra, X, U, A, t, SP):
        x, x1, x2, u, XC, XM = X[0][0], X[1][0], X[2][0], U[0][0], A[0], A[1]
        return [-u + o[:, 0] * x + o[:, 1] * exp10(XC) * x1 + o[:, 2] * exp10(XM) * x2]


class Spring1(mb.LagrangianElement):
    """Spring{1} (test/SomeElements.jl:164-188), no_second_order = Val(false)"""
    type_parameters = ()

    @classmethod
    def doflist(cls, **kw):
        return (1, 2, 3, 3), ("X", "X", "A", "A"), ("tx1", "tx1", "ΞL₀", "ΞEI")

    @classmethod
    def construct(cls, coords, EA):
        x1, x2 = coords[0][:, 0], coords[1][:, 0]
        return np.stack([x1, x2, np.full_like(x1, EA), np.abs(x1 - x2)], axis=1)

    @staticmethod
    def residual(o, extra, X, U, A, t, SP):
        L0 = o[:, 3] * exp10(A[0]); EA = o[:, 2] * exp10(A[1])
        dx = (X[0][0] + o[:, 0]) - (X[0][1] + o[:, 1])
        L = sqrt(dx * dx)
        T = EA * (L - L0) / L0
        F1 = dx / L * T
        return [F1, -F1]


class Spring2(mb.LagrangianElement):
    """Spring{2} (test/SomeElements.jl:164-188): two translations per node, A-dofs ΞL₀, ΞEI on a third node; no_second_order = Val(false)"""
    type_parameters = ()

    @classmethod
    def doflist(cls, **kw):
        return (1, 1, 2, 2, 3, 3), ("X", "X", "X", "X", "A", "A"), ("tx1", "tx2", "tx1", "tx2", "ΞL₀", "ΞEI")

    @classmethod
    def construct(cls, coords, EA):
        x1, x2 = coords[0], coords[1]
        return np.concatenate([x1, x2, np.full((x1.shape[0], 1), float(EA)), np.sqrt(((x1 - x2) ** 2).sum(1))[:, None]], axis=1)

    @staticmethod
    def residual(o, extra, X, U, A, t, SP):
        L0 = o[:, 5] * exp10(A[0]); EA = o[:, 4] * exp10(A[1])
        d1 = (X[0][0] + o[:, 0]) - (X[0][2] + o[:, 2]); d2 = (X[0][1] + o[:, 1]) - (X[0][3] + o[:, 3])
        L = sqrt(d1 * d1 + d2 * d2)
        T = EA * (L - L0) / L0
        F1, F2 = d1 / L * T, d2 / L * T
        return [F1, F2, -F1, -F2]


def model_testdirectxua001():
    """test/TestDirectXUA001.jl:7-30"""
    m = mb.Model("TestModel")
    n1 = mb.addnode(m, [0., 0.]); n2 = mb.addnode(m, [10., 0.]); n3 = mb.addnode(m, [0., 10.]); n4 = mb.addnode(m, [])
    mb.addelement(m, Spring2, [n1, n2, n4], EA=10)
    mb.addelement(m, Spring2, [n1, n3, n4], EA=10)
    load = lambda t: 0.1 * t
    mb.addelement(m, mb.DofLoad, [n1], field="tx1", value=load)
    mb.addelement(m, mb.DofLoad, [n1], field="tx2", value=load)
    for n in (n2, n3):
        for f in ("tx1", "tx2"):
            mb.addelement(m, mb.Hold, [n], field=f)
    meas = lambda x, t: 0.5 * ((x - 0.12 * t) / 0.01) ** 2
    acost = lambda a: 0.5 * (a / .1) ** 2
    mb.addelement(m, mb.SingleDofCost, [n1], clas="X", field="tx1", cost=meas)
    mb.addelement(m, mb.SingleDofCost, [n1], clas="X", field="tx2", cost=meas)
    mb.addelement(m, mb.SingleAcost, [n4], field="ΞL₀", cost=acost)
    mb.addelement(m, mb.SingleAcost, [n4], field="ΞEI", cost=acost)
    return m


def horner(p, x):
    y = 0.
    for c in reversed(p):
        y = c + x * y
    return y


class Turbine(mb.LagrangianElement):
    """Turbine (test/SomeElements.jl:20-31): R = −sea(t,x)·seadrag·(1+A₁) − sky(t,x)·skydrag·(1+A₂); sea, sky return two components"""
    type_parameters = ("sea", "sky")

    @classmethod
    def doflist(cls, **kw):
        return (1, 1, 2, 2), ("X", "X", "A", "A"), ("tx1", "tx2", "Δseadrag", "Δskydrag")

    @classmethod
    def construct(cls, coords, seadrag, sea, skydrag, sky):
        n = coords[0].shape[0]
        return np.concatenate([coords[0][:, :3], np.tile([[seadrag, skydrag]], (n, 1))], axis=1), dict(sea=sea, sky=sky)

    @staticmethod
    def residual(o, extra, X, U, A, t, SP):
        x = [X[0][0] + o[:, 0], X[0][1] + o[:, 1]]
        sea, sky = extra["sea"](t, x), extra["sky"](t, x)
        return [-sea[i] * o[:, 3] * (1 + A[0]) - sky[i] * o[:, 4] * (1 + A[1]) for i in range(2)]


ANCHOR_P = [2.82040487827, -24.86027164695, 153.69500343165, -729.52107422849, 2458.11921356871, -5856.85610233072, 9769.49700812681, -11141.12651712473,
            8260.66447746395, -3582.36704093187, 687.83550335374]


class AnchorLine(mb.LagrangianElement):
    """AnchorLine (test/SomeElements.jl:62-103): a catenary mooring line given by its `lagrangian` only — L = Λ₁₂·Fd + Λ₃·m₃"""
    type_parameters = ()

    @classmethod
    def doflist(cls, **kw):
        return (1, 1, 1, 2, 2), ("X", "X", "X", "A", "A"), ("tx1", "tx2", "rx3", "ΔL", "Δbuoyancy")

    @classmethod
    def construct(cls, coords, Δxₘtop, xₘbot, L, buoyancy):
        n = coords[0].shape[0]
        return np.concatenate([coords[0][:, :3], np.tile([list(Δxₘtop) + list(xₘbot) + [L, buoyancy]], (n, 1))], axis=1)

    @staticmethod
    def lagrangian(o, extra, Λ, X, U, A, t, SP):
        return AnchorLine.lagrangian_req(o, extra, Λ, X, U, A, t, SP)[0]

    @staticmethod
    def lagrangian_req(o, extra, Λ, X, U, A, t, SP):
        """→ L and the requestables (☼ of test/SomeElements.jl:80-88) for ElementCost / ElementConstraint"""
        L, buoy = o[:, 8] * (1 + A[0]), o[:, 9] * (1 + A[1])
        x = X[0]
        Xtop = [x[0] + o[:, 0], x[1] + o[:, 1], o[:, 2]]
        c, s = cos(x[2]), sin(x[2])
        dX = [c * o[:, 3] - s * o[:, 4], s * o[:, 3] + c * o[:, 4]]                  # arm of the fairlead
        ch = [Xtop[0] + dX[0] - o[:, 6], Xtop[1] + dX[1] - o[:, 7]]                 # from anchor to fairlead
        xaf = sqrt(ch[0] * ch[0] + ch[1] * ch[1])
        cr = exp10(horner(ANCHOR_P, (L - xaf) / Xtop[2])) * Xtop[2]
        Fh = -cr * buoy
        Fd = [ch[0] / xaf * Fh, ch[1] / xaf * Fh]
        m3 = dX[0] * Fd[1] - dX[1] * Fd[0]
        return Λ[0] * Fd[0] + Λ[1] * Fd[1] + Λ[2] * m3, dict(Fh=Fh, cr=cr, xaf=xaf)


def model_testsweepx0():
    """test/TestSweepX0.jl:9-18"""
    m = mb.Model("TestModel")
    n1 = mb.addnode(m, [0., 0., 100.]); n2 = mb.addnode(m, []); n3 = mb.addnode(m, [])
    sea = lambda t, x: (1. * t, 0. * t)
    sky = lambda t, x: (0., 10.)
    mb.addelement(m, Turbine, [n1, n2], seadrag=1e6, sea=sea, skydrag=1e5, sky=sky)
    for i in range(3):
        al = np.array([np.cos(i * 2 * np.pi / 3), np.sin(i * 2 * np.pi / 3)])
        mb.addelement(m, AnchorLine, [n1, n3], Δxₘtop=list(5 * al) + [0.], xₘbot=list(250 * al), L=290., buoyancy=-5e3)
    return m


def fa(a): return a ** 2 * 1e-14
def fu(u, t): return u ** 2
def l1(x, t): return (x - 0.1 * np.sin(t)) ** 2
def l2(x, t): return (x - 0.1 * np.cos(t)) ** 2


def fa_stiff(a): return (a - 0.02) ** 2 * 0.5        # pulls every A-dof towards 0.02: the identified A is not zero


def model_testdirectxua(fa=fa):
    """test/TestDirectXUA.jl:26-51 (fa: the A-cost; the reference's 1e-14·a² leaves the all-steps system with a condition number of 8e15)"""
    m = mb.Model("TrueModel")
    n1 = mb.addnode(m, [0.]); n2 = mb.addnode(m, [1.]); n3 = mb.addnode(m, [])
    mb.addelement(m, El1, [n1], K=1., C=0.05, M=1.)
    mb.addelement(m, El1, [n2], K=0., C=0.0, M=1.)
    mb.addelement(m, Spring1, [n1, n2, n3], EA=1.1)
    mb.addelement(m, mb.SingleAcost, [n3], field="ΞL₀", cost=fa)
    mb.addelement(m, mb.SingleAcost, [n3], field="ΞEI", cost=fa)
    mb.addelement(m, mb.SingleAcost, [n1], field="ΞC", cost=fa)
    mb.addelement(m, mb.SingleAcost, [n1], field="ΞM", cost=fa)
    mb.addelement(m, mb.SingleAcost, [n2], field="ΞC", cost=fa)
    mb.addelement(m, mb.SingleAcost, [n2], field="ΞM", cost=fa)
    mb.addelement(m, mb.SingleUdof, [n1], Xfield="tx1", Ufield="utx1", cost=fu)
    mb.addelement(m, mb.SingleUdof, [n2], Xfield="tx1", Ufield="utx1", cost=fu)
    mb.addelement(m, mb.SingleDofCost, [n1], clas="X", field="tx1", cost=l1)
    mb.addelement(m, mb.SingleDofCost, [n2], clas="X", field="tx1", cost=l2)
    return m


def dis_lists(dis):
    """Disassembler → the oracle's `dis` (list of dicts of 1-based index arrays)"""
    return [dict(X=d.X, U=d.U, A=d.A) for d in dis.dis]


class El0(mb.LagrangianElement):
    """El1 without the derivatives an analysis of lower order does not carry: r = −u + K·x (+ C·10^ΞC·x′ when ox ≥ 1); no_second_order = Val(true) to cover that branch"""
    type_parameters = ("ox",)
    no_second_order = True

    @classmethod
    def doflist(cls, **kw):
        return (1, 1, 1, 1), ("X", "U", "A", "A"), ("tx1", "u", "ΞC", "ΞM")

    @classmethod
    def construct(cls, coords, K, C, ox):
        n = coords[0].shape[0]
        return np.tile(np.array([[K, C, ox]], float), (n, 1))

    @staticmethod
    def residual(o, extra, X, U, A, t, SP):
        r = -U[0][0] + o[:, 0] * X[0][0] * exp10(A[1])
        if len(X) > 1:
            r = r + o[:, 1] * exp10(A[0]) * X[1][0]
        return [r]


def model_chain(n, rng, ox=2):
    """a longer model for the structures at size: n El1 oscillators on a line, springs between neighbours (each with its own A node), U-loads and costs"""
    m = mb.Model("chain")
    nod = mb.addnode(m, np.arange(n, dtype=float)[:, None])
    anod = [mb.addnode(m, []) for _ in range(n - 1)]
    if ox == 2: mb.addelement(m, El1, nod[:, None], K=1.3, C=0.07, M=0.9)
    else: mb.addelement(m, El0, nod[:, None], K=1.3, C=0.07, ox=ox)
    mb.addelement(m, Spring1, np.stack([nod[:-1], nod[1:], np.asarray(anod)], axis=1), EA=2.1)
    for a in anod[::3]:
        mb.addelement(m, mb.SingleAcost, [a], field="ΞL₀", cost=fa)
    mb.addelement(m, mb.SingleUdof, nod[:, None], Xfield="tx1", Ufield="utx1", cost=fu)
    mb.addelement(m, mb.SingleDofCost, nod[::2, None], clas="X", field="tx1", cost=l1)
    return m


class SdofOscillator(mb.LagrangianElement):
    """SdofOscillator (test/SomeElements.jl:191-207): R = −u + K₁x + K₂x² + C₁x′ + C₂x′² + M₁x″ + M₂x″²; dofs (X tx1, U tu1); the parameters are element data"""
    type_parameters = ()

    @classmethod
    def doflist(cls, **kw):
        return (1, 1), ("X", "U"), ("tx1", "tu1")

    @classmethod
    def construct(cls, coords, K1=0., K2=0., C1=0., C2=0., M1=0., M2=0.):
        return np.tile(np.array([[K1, K2, C1, C2, M1, M2]], float), (coords[0].shape[0], 1))

    @staticmethod
    def residual(o, extra, X, U, A, t, SP):
        x = X[0][0]
        r = -U[0][0] + o[:, 0] * x + o[:, 1] * x * x
        if len(X) > 1:
            r = r + o[:, 2] * X[1][0] + o[:, 3] * X[1][0] * X[1][0]
        if len(X) > 2:
            r = r + o[:, 4] * X[2][0] + o[:, 5] * X[2][0] * X[2][0]
        return [r]


def model_testeigx():
    """test/TestEigX.jl:6-24: 10 Spring{1} between 11 nodes on a line (A-dofs on a node of their own), a mass-damper on every node, a unit spring to ground at node 1"""
    nel, L, EA = 10, 10., 1.
    nnod = nel + 1
    M, Cd = 10. / nnod, 3. / nnod
    m = mb.Model("TestEigX")
    xnod = mb.addnode(m, np.linspace(0., L, nnod)[:, None])
    anod = mb.addnode(m, np.zeros((1, 0)))[0]
    mb.addelement(m, Spring1, np.stack([xnod[:-1], xnod[1:], np.full(nel, anod)], axis=1), EA=EA)
    mb.addelement(m, SdofOscillator, xnod[:, None], M1=M, C1=Cd)
    mb.addelement(m, SdofOscillator, xnod[:1, None], K1=1.)
    return m


def model_mooring_wrapped():
    """the moored turbine of test/TestSweepX0.jl:9-18 with its three AnchorLines wrapped as in test/TestElementCost.jl: one plain, one in ElementCost (cost on the horizontal
    force), one in ElementConstraint (gap on the horizontal force, multiplier λ as a U-dof on the turbine node)"""
    m = mb.Model("MooredTurbine")
    n1 = mb.addnode(m, [0., 0., 100.]); n2 = mb.addnode(m, []); n3 = mb.addnode(m, [])
    sea = lambda t, x: (1. * t, 0. * t)
    sky = lambda t, x: (0., 10.)
    mb.addelement(m, Turbine, [n1, n2], seadrag=1e6, sea=sea, skydrag=1e5, sky=sky)
    ek = lambda i: dict(Δxₘtop=list(5 * np.array([np.cos(i * 2 * np.pi / 3), np.sin(i * 2 * np.pi / 3)])) + [0.],
                        xₘbot=list(250 * np.array([np.cos(i * 2 * np.pi / 3), np.sin(i * 2 * np.pi / 3)])), L=290., buoyancy=-5e3)
    mb.addelement(m, AnchorLine, [n1, n3], **ek(0))
    mb.addelement(m, mb.ElementCost, [n1, n3], req=("Fh",), cost=lambda eleres, t: 1e-8 * (eleres.Fh - 4e5) * (eleres.Fh - 4e5), ElementType=AnchorLine, elementkwargs=ek(1))
    mb.addelement(m, mb.ElementConstraint, [n1, n3], λinod=1, λfield="λ", req=("Fh", "cr"), gap=lambda eleres, t: 1e-5 * eleres.Fh - 0.05 * eleres.cr + t,
                  mode=mb.equal, ElementType=AnchorLine, elementkwargs=ek(2))
    return m
