"""End-to-end on the GPU engine: the steel catenary riser of examples/DynamicBeamAnalysis.jl (BASELINE.json configs[1] at the reference's own
size, and the beam + SoilContact mesh of configs[4]): 100 EulerBeam3D in three segments with two cross-sections, 4 Hold, 2 DofConstraint with
time-dependent gaps, 101 DofLoad (weight ramp), 61 SoilContact — SweepX{0} through the reference's static loading sequence (time = −10:0.2:0),
then SweepX{2} Newmark steps under a forced top motion.  Compared with the same Newton loops driven by the oracle's CPU assembly.
The RIFLEX top-motion series (examples/SCR.csv) is replaced by a synthetic harmonic motion: the file does not travel to the GPU box."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import elements as OE
from oracle import pattern as OP

from muscade_b200.examples import (G_, RHO, xsection, X1, W1, X2, W2, NEL, SEGLEN, xmotion, zmotion, horiz_target, vert_target, ramp,   # noqa: F401
                                   scr_riser as build)


class OracleSCR:
    """assemble!{mission} of the whole model with the oracle's beams/soil and the boundary elements restated from src/BasicElements.jl"""

    def __init__(self, model, dis, weights):
        self.model, self.dis, self.weights = model, dis, weights
        self.ndof = model.getndof("X")
        odis = [dict(X=d.X, U=np.zeros((d.X.shape[0], 0), np.int64), A=np.zeros((d.X.shape[0], 0), np.int64)) for d in dis.dis]
        self.asm1, self.asm2, self.colptr, self.rowval = OP.prepare_sweepx(odis, self.ndof, 0, 0)
        self.kinds = [et.ElType.__name__ for et in model.ele]

    def assemble(self, OX, mission, X, nm, t):
        L = np.zeros(self.ndof); nz = np.zeros(len(self.rowval))
        iload = 0
        for k, (et, d) in enumerate(zip(self.model.ele, self.dis.dis)):
            a1, a2 = self.asm1[k], self.asm2[k]
            kind = self.kinds[k]
            if kind == "EulerBeam3D":
                OE.sweepx_assemble_beams(et.eleobj, d.X, a1.T, a2.T, OX, mission, X, np.ones(12), nm, L, nz)
            elif kind == "SoilContact":
                soil = et.eleobj
                OE.sweepx_addin_generic(lambda e, xv, sd: OE.soil_residual(soil[e], xv, sd)[:2], 3, d.X, a1.T, a2.T, OX, mission, X, np.ones(3), nm, L, nz)
            elif kind in ("Hold", "DofConstraint"):       # R = (−λ, −gap), gap = x − target(t): K = [[0,−1],[−1,0]]  (BasicElements.jl:425-438)
                target = 0. if kind == "Hold" else (horiz_target(t) if et.field[0] == "t1" else vert_target(t))
                for e in range(d.X.shape[0]):
                    ix = d.X[e] - 1; q = a2[:, e] - 1
                    L[ix[0]] += -X[0][ix[1]]; L[ix[1]] += -(X[0][ix[0]] - target)
                    nz[q[1]] += -1.; nz[q[2]] += -1.
            elif kind == "DofLoad":                        # R = −value(t)  (BasicElements.jl:275-284)
                F = self.weights[iload](t); iload += 1
                for e in range(d.X.shape[0]):
                    L[d.X[e, 0] - 1] += -F
            else:
                raise AssertionError(kind)
        return L, nz

    def solve(self, OX, X, times, t0, maxdx, maxiter=100):
        """the reference's SweepX loop (src/SweepX.jl:195-221) incl. Newmarkβdecrement! (:98-132)"""
        out = []; told = t0
        X = [x.copy() for x in X]
        while len(X) < OX + 1:
            X.append(np.zeros(self.ndof))
        for t in times:
            nm = OE.newmark_coefficients(OX, (t - told) if OX > 0 else 0.); told = t
            a1, a2, a3, b1, b2, b3 = nm[:6]
            for it in range(maxiter):
                mission = "step" if it == 0 else "iter"
                L, nz = self.assemble(OX, mission, X, nm, t)
                dx = spla.splu(sp.csc_matrix((nz, self.rowval - 1, self.colptr - 1), shape=(self.ndof, self.ndof))).solve(L)
                if OX == 2:
                    if it == 0:
                        dxp = a1 * dx + (a2 * X[1] + a3 * X[2]); dxpp = b1 * dx + (b2 * X[1] + b3 * X[2])
                    else:
                        dxp = a1 * dx; dxpp = b1 * dx
                    X[0] = X[0] - dx; X[1] = X[1] - dxp; X[2] = X[2] - dxpp
                else:
                    X[0] = X[0] - dx
                if dx @ dx <= maxdx ** 2:
                    break
            else:
                raise AssertionError("oracle loop: no convergence at t=%g" % t)
            out.append([x.copy() for x in X])
        return out


STATIC_TIMES = np.round(np.arange(-10., 0.0001, 0.2), 10)
DYN_TIMES = 0.1 + 0.3 * np.arange(4)


@pytest.mark.gpu
def test_scr_riser_static_then_dynamic(mb):
    model, node_lists, weights = build(mb)
    assert [e.ElType.__name__ for e in model.ele][:1] == ["EulerBeam3D"] and model.ele[0].nele == 100
    state0 = mb.initialize(model)
    assert model.getndof("X") == 6 * 101 + 4 + 2
    static = mb.sweepx.solve(0, state0, STATIC_TIMES, maxΔx=1e-6, maxiter=100)
    assert len(static) == len(STATIC_TIMES)
    top = node_lists[2][-1]
    # the constraints hold: the top end sits where the gap functions put it (DynamicBeamAnalysis.jl:113-116)
    assert abs(mb.getdof(static[-1], "t1", nodID=[top])[0] - horiz_target(0.)) < 1e-6
    assert abs(mb.getdof(static[-1], "t3", nodID=[top])[0] - vert_target(0.)) < 1e-6
    z = mb.getdof(static[-1], "t3", nodID=node_lists[0])
    assert (z < 0).any() and (z > 0).any()                      # part of the first segment rests on the soil springs, part has lifted off
    ora = OracleSCR(model, state0.dis, weights)
    ref = ora.solve(0, [np.zeros(ora.ndof)], STATIC_TIMES, -np.inf, 1e-6)
    for k in (0, 10, 25, 40, 50):
        r = ref[k][0]
        assert np.abs(static[k].X[0] - r).max() <= 1e-6 * max(1., np.abs(r).max()), k
    # dynamics from the static equilibrium, Newmark-β, forced top motion through the two DofConstraint gaps
    dyn = mb.sweepx.solve(2, static[-1], DYN_TIMES, maxΔx=1e-5)
    refd = ora.solve(2, [static[-1].X[0]], DYN_TIMES, 0.0, 1e-5)
    for k in range(len(DYN_TIMES)):
        for d in range(3):
            r = refd[k][d]
            assert np.abs(dyn[k].X[d] - r).max() <= 1e-6 * max(1., np.abs(r).max()), (k, d)
    assert np.abs(mb.getdof(dyn[-1], "t1", order=1)).max() > 1e-3      # it moves


@pytest.mark.gpu
@pytest.mark.parametrize("OX,mission", [(0, "iter"), (2, "step"), (2, "iter")])
def test_scr_riser_assembly_parity(mb, OX, mission):
    """one assemble! of the whole 7-type SCR model at a random state against the oracle: pattern bit-exact, values ≤ 1e-12"""
    model, node_lists, weights = build(mb)
    state = mb.initialize(model)
    dis = state.dis
    ndof = model.getndof("X")
    state = state.with_orders(1, OX + 1, 1)
    state.X[0] = mb.synthetic.uniform_pm1(3, ndof) * 0.2
    for d in range(1, OX + 1):
        state.X[d] = mb.synthetic.uniform_pm1(3 + d, ndof) * 0.3
    state.time = -6.1
    out, asm, gr = mb.sweepx.prepare(OX, model, dis)
    out.c = mb.synthetic.newmark_coefficients(OX, 0.3)
    mb.sweepx.assemble(mission, out, asm, dis, model, state, 0.3)
    ora = OracleSCR(model, dis, weights)
    assert np.array_equal(out.Lλx.indptr + 1, ora.colptr) and np.array_equal(out.Lλx.indices + 1, ora.rowval)
    L, nz = ora.assemble(OX, mission, state.X[: OX + 1], out.c, state.time)
    assert np.abs(out.Lλx.data - nz).max() <= 1e-12 * np.abs(nz).max()
    assert np.abs(out.Lλ - L).max() <= 1e-12 * max(np.abs(nz).max(), np.abs(L).max())
    out.engine.close()


def oracle_scr_direct_step(model, dis, weights, P, OX, OU, X, U, Lam, t, σu):
    """assemble!{:matrices}(out::AssemblyDirect{OX,OU,0},…) of the whole SCR model at one step: beams through the first-order path
    (DirectXUA.jl:85-120), SoilContact / Hold / DofConstraint / DofLoad through the second-order path (:152-171, L = Λ∘R), SingleDofCost on U
    (lagrangian = cost, BasicElements.jl:198-208).  Boundary elements restated by hand as in OracleSCR."""
    nX, nU = P["ndof"][0], P["ndof"][2]
    A = P["asm"]
    sL, sX, sU = dis.scaleΛ, dis.scaleX, dis.scaleU
    o = OE.direct_out_zeros(P, OX, OU)
    o["L1"][2] = np.zeros((OX + 1, nX)); o["L1"][3] = np.zeros((OU + 1, nU))
    huu = np.zeros(len(P["pat"][(3, 3)][3]))
    cpU, rvU = P["pat"][(3, 3)][2], P["pat"][(3, 3)][3]
    iload = 0
    for k, (et, d) in enumerate(zip(model.ele, dis.dis)):
        kind = et.ElType.__name__
        if kind == "EulerBeam3D":
            OE.direct_assemble_step_beams(et.eleobj, d.X, d.U, OX, OU, X[: OX + 1], [U], d.scaleX, d.scaleU, P, k, out=o)
        elif kind == "SoilContact":
            OE.direct_addin_soil_second_order(et.eleobj, d.X, OX, X[: OX + 1], Lam, d.scaleX, model.scaleΛ, P, k, o)
        elif kind in ("Hold", "DofConstraint", "DofLoad"):
            aL, aXv, aLX, aXL = A[OP.arrnum(1)][k], A[OP.arrnum(2)][k], A[OP.arrnum(1, 2)][k], A[OP.arrnum(2, 1)][k]
            if kind == "DofLoad":
                F = weights[iload](t); iload += 1
            for e in range(d.X.shape[0]):
                ix = d.X[e] - 1
                if kind == "DofLoad":                       # R = −value(t): only L1[Λ] (and zeros elsewhere)
                    o["L1"][1][aL[0, e] - 1] += -F * sL[ix[0]]
                    continue
                target = 0. if kind == "Hold" else (horiz_target(t) if et.field[0] == "t1" else vert_target(t))
                K = np.array([[0., -1.], [-1., 0.]])         # dofs (x, λc): R = (−λc, −(x − target))
                R = np.array([-X[0][ix[1]], -(X[0][ix[0]] - target)])
                for i in range(2):
                    o["L1"][1][aL[i, e] - 1] += R[i] * sL[ix[i]]
                    o["L1"][2][0, aXv[i, e] - 1] += (K[:, i] @ Lam[ix]) * sX[ix[i]]
                    for j in range(2):
                        o["L2"][(1, 2)][0, aLX[i + 2 * j, e] - 1] += K[i, j] * sL[ix[i]] * sX[ix[j]]
                        o["L2"][(2, 1)][0, aXL[j + 2 * i, e] - 1] += K[i, j] * sL[ix[i]] * sX[ix[j]]
        elif kind == "SingleDofCost":                        # ½(u/σu)² on a U dof
            for e in range(d.U.shape[0]):
                j = d.U[e, 0]
                o["L1"][3][0, j - 1] += U[j - 1] / σu ** 2 * sU[j - 1]
                q = cpU[j - 1] - 1 + np.nonzero(rvU[cpU[j - 1] - 1: cpU[j] - 1] == j)[0][0]
                huu[q] += sU[j - 1] ** 2 / σu ** 2
        else:
            raise AssertionError(kind)
    o["L2"][(3, 3)] = {(1, 1): huu}
    return o


@pytest.mark.gpu
def test_scr_directxua_assembly_parity(mb):
    """BASELINE.json configs[4] at test size: assemblebig!{:matrices} (DirectXUA.jl:316-356) of the SCR riser — 100 EulerBeam3D{Udof} in 5 types /
    2 cross-sections, 61 SoilContact, 4 Hold, 2 DofConstraint with moving gaps, 101 DofLoad, quadratic SingleDofCost on every U dof — over 7 steps
    at random (Λ, X, X′, X″, U): Lvv structure bit-exact, Lvv values and Lv ≤ 1e-12 against the oracle."""
    OX, OU, nstep, dt, t0, σu = 2, 0, 7, 0.3, -6.3, 50.
    model, node_lists, weights = build(mb, udof=True)
    unodes = np.concatenate([et.nodID[:, 2] for et in model.ele if et.ElType.__name__ == "EulerBeam3D"])
    for f in ("t1", "t2", "t3"):
        mb.addelement(model, mb.SingleDofCost, unodes[:, None], clas="U", field=f, cost=lambda u, t: 0.5 * (u / σu) ** 2)
    mb.setscale(model, scale=dict(X=dict(t1=2., t2=2., t3=2.), U=dict(t1=30., t2=30., t3=30.)), Λscale=1e3)
    st0 = mb.initialize(model); dis = st0.dis
    nX, nU, nA = model.getndof(("X", "U", "A"))
    assert nX == 612 and nU == 300
    time = t0 + dt * np.arange(nstep)                         # crosses t = −5: the weight ramp saturates (DynamicBeamAnalysis.jl:119-121)
    st = [([mb.synthetic.uniform_pm1(10 + 3 * s + d, nX) * (0.2 if d == 0 else 0.3) for d in range(3)], 20. * mb.synthetic.uniform_pm1(99 + s, nU)) for s in range(nstep)]
    Lam = [mb.synthetic.uniform_pm1(500 + s, nX) for s in range(nstep)]
    odis = [dict(X=d.X, U=d.U, A=d.A) for d in dis.dis]
    P = OP.prepare_direct(odis, nX, nU, nA, OX, OU, 0)
    big, bigasm, pgr, pgc = OP.preparebig(0, [nstep], P["nL2"], P["pat"])
    outs = [oracle_scr_direct_step(model, dis, weights, P, OX, OU, st[s][0], st[s][1], Lam[s], time[s], σu) for s in range(nstep)]
    nz, Lv = OP.assemblebig(0, nstep, dt, P, big, bigasm, pgr, outs)

    eng = mb.directxua.prepare(OX, OU, model, dis, nstep, dt, t0=t0)
    try:
        cp, rv = eng.big_pattern()
        assert np.array_equal(cp, big["colptr"]) and np.array_equal(rv, big["rowval"])
        for s, (X, U) in enumerate(st):
            eng.set_state(s, X, U)
            eng.set_lambda(s, Lam[s])
            eng.set_host_cost(s, *mb.directxua.host_costs(eng, s, X[0], U, time[s])[:4])
            mb.directxua.host_elements(eng, s, X, Lam[s], time[s], model.scaleΛ)
        Lvv = np.zeros(eng.nnzbig); Lvec = np.zeros(eng.ncol)
        eng.direct_assemble(Lvv=Lvv, Lv=Lvec)
        scale = np.abs(nz).max()
        assert np.abs(Lvv - nz).max() <= 1e-12 * scale
        assert np.abs(Lvec - Lv).max() <= 1e-12 * max(scale, np.abs(Lv).max())
        W = 2 * nX + nU
        assert np.abs(Lvec.reshape(nstep, W)[:, nX: 2 * nX]).max() > 0 and np.abs(Lvec.reshape(nstep, W)[:, 2 * nX:]).max() > 0
    finally:
        eng.close()


def test_scr_oracle_loop_converges_on_cpu():
    """CPU part: the model builds with the reference's dof numbering and the oracle-driven static sequence converges to a riser that hangs from
    the prescribed top position and rests on the sea bed (so the GPU comparison above compares meaningful states)"""
    import muscade_b200 as mb
    model, node_lists, weights = build(mb)
    state0 = mb.initialize(model)
    assert model.getndof("X") == 612 and len(model.ele) == 1 + 4 + 2 + 3 + 1
    ora = OracleSCR(model, state0.dis, weights)
    ref = ora.solve(0, [np.zeros(ora.ndof)], STATIC_TIMES[:6], -np.inf, 1e-6)
    top_t1 = model.nod_dof[model.getidoftyp("X", "t1") - 1][node_lists[2][-1]] - 1
    assert abs(ref[-1][0][top_t1] - horiz_target(STATIC_TIMES[5])) < 1e-6


@pytest.mark.gpu
def test_scr_general_form_assembly_parity(mb):
    """The same problem through the GENERAL DirectXUA form (mb_xua_*): every device type of the riser — EulerBeam3D{Udof} (first-order packets), SoilContact (second-order
    branch in closed form) — evaluated by mb_xua_eval_device, Hold / DofConstraint / DofLoad through their closed-form partials and the U-costs through D2 on the host:
    the structures are the oracle's bit for bit, Lvv values and Lv ≤ 1e-12."""
    from muscade_b200 import xua
    OX, OU, nstep, dt, t0, σu = 2, 0, 7, 0.3, -6.3, 50.
    model, node_lists, weights = build(mb, udof=True)
    unodes = np.concatenate([et.nodID[:, 2] for et in model.ele if et.ElType.__name__ == "EulerBeam3D"])
    for f in ("t1", "t2", "t3"):
        mb.addelement(model, mb.SingleDofCost, unodes[:, None], clas="U", field=f, cost=lambda u, t: 0.5 * (u / σu) ** 2)
    mb.setscale(model, scale=dict(X=dict(t1=2., t2=2., t3=2.), U=dict(t1=30., t2=30., t3=30.)), Λscale=1e3)
    st0 = mb.initialize(model); dis = st0.dis
    nX, nU, nA = model.getndof(("X", "U", "A"))
    time = t0 + dt * np.arange(nstep)
    st = [([mb.synthetic.uniform_pm1(10 + 3 * s + d, nX) * (0.2 if d == 0 else 0.3) for d in range(3)], 20. * mb.synthetic.uniform_pm1(99 + s, nU)) for s in range(nstep)]
    Lam = [mb.synthetic.uniform_pm1(500 + s, nX) for s in range(nstep)]
    odis = [dict(X=d.X, U=d.U, A=d.A) for d in dis.dis]
    P = OP.prepare_direct(odis, nX, nU, nA, OX, OU, 0)
    big, bigasm, pgr, pgc = OP.preparebig(0, [nstep], P["nL2"], P["pat"])
    outs = [oracle_scr_direct_step(model, dis, weights, P, OX, OU, st[s][0], st[s][1], Lam[s], time[s], σu) for s in range(nstep)]
    nz, Lv = OP.assemblebig(0, nstep, dt, P, big, bigasm, pgr, outs)
    eng = xua.XUAEngine(0)
    try:
        nbig, nnz = eng.prepare(model, dis, OX, OU, 0, [nstep], [dt])
        eng.set_time0(1, t0)
        cp, rv = eng.big_pattern()
        assert np.array_equal(cp, big["colptr"]) and np.array_equal(rv, big["rowval"])
        states = [[mb.State(float(time[s]), [Lam[s]], st[s][0], [st[s][1]], st0.A, None, model, dis) for s in range(nstep)]]
        eng.assemblebig(states)
        Lvv, Lvec = eng.big()
        scale = np.abs(nz).max()
        assert np.abs(Lvv - nz).max() <= 1e-12 * scale
        assert np.abs(Lvec - Lv).max() <= 1e-12 * max(scale, np.abs(Lv).max())
    finally:
        eng.close()


@pytest.mark.gpu
def test_gauged_scr_windowed_path_equals_general_form(mb):
    """The SCR riser with ElementCost{StrainGaugeOnEulerBeam3D} on every beam (load identification from strain gauges; bench.py --workload scr --gauged): the beam-specialised
    path with its step batching (mb_direct_set_gauge_cost; SoilContact, Hold / DofConstraint / DofLoad and the U-costs beside the costed beams) against the general form:
    identical structure, Lvv values and Lv ≤ 1e-12."""
    from muscade_b200 import xua
    OX, OU, nstep, dt, t0, σu = 2, 0, 7, 0.3, -6.3, 50.
    cost = mb.QuadraticGaugeCost(2e-5, lambda t: 1e-4 * np.cos(0.5 * t) * np.array([1., 0.5, -1., -0.5]))
    model, node_lists, weights = mb.examples.scr_riser(mb, udof=True, gauge_cost=cost)
    unodes = np.concatenate([et.nodID[:, 2] for et in model.ele if et.ElType.__name__ == "ElementCost"])
    for f in ("t1", "t2", "t3"):
        mb.addelement(model, mb.SingleDofCost, unodes[:, None], clas="U", field=f, cost=lambda u, t: 0.5 * (u / σu) ** 2)
    mb.setscale(model, scale=dict(X=dict(t1=2., t2=2., t3=2.), U=dict(t1=30., t2=30., t3=30.)), Λscale=1e3)
    st0 = mb.initialize(model); dis = st0.dis
    nX, nU = model.getndof("X"), model.getndof("U")
    time = t0 + dt * np.arange(nstep)
    st = [([mb.synthetic.uniform_pm1(10 + 3 * s + d, nX) * (0.2 if d == 0 else 0.3) for d in range(3)], 20. * mb.synthetic.uniform_pm1(99 + s, nU)) for s in range(nstep)]
    Lam = [mb.synthetic.uniform_pm1(500 + s, nX) for s in range(nstep)]
    spec = mb.directxua.prepare(OX, OU, model, dis, nstep, dt, t0=t0)
    gen = xua.XUAEngine(0)
    try:
        for s in range(nstep):
            spec.set_state(s, st[s][0], st[s][1]); spec.set_lambda(s, Lam[s])
            spec.set_host_cost(s, *mb.directxua.host_costs(spec, s, st[s][0][0], st[s][1], time[s])[:4])
            mb.directxua.host_elements(spec, s, st[s][0], Lam[s], time[s], model.scaleΛ)
        spec.set_gauge_times(time)
        Lvv = np.zeros(spec.nnzbig); Lv = np.zeros(spec.ncol)
        spec.direct_assemble(Lvv=Lvv, Lv=Lv)
        cp, rv = spec.big_pattern()
        nbig, nnz = gen.prepare(model, dis, OX, OU, 0, [nstep], [dt])
        gen.set_time0(1, t0)
        gcp, grv = gen.big_pattern()
        assert nbig == spec.ncol and np.array_equal(cp, gcp) and np.array_equal(rv, grv)
        states = [[mb.State(float(time[s]), [Lam[s]], st[s][0], [st[s][1]], st0.A, None, model, dis) for s in range(nstep)]]
        gen.assemblebig(states)
        gLvv, gLv = gen.big()
        scale = np.abs(gLvv).max()
        assert np.abs(gLvv - Lvv).max() <= 1e-12 * scale
        assert np.abs(gLv - Lv).max() <= 1e-12 * max(scale, np.abs(gLv).max())
    finally:
        spec.close(); gen.close()
