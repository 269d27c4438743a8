"""SweepX on the B200 engine: the reference's solver interface with `assemble!` replaced by the C-ABI call.

  DofGroup / allXdofs                 src/Assemble.jl:167-264
  prepare(AssemblySweepX{OX})         src/SweepX.jl:24-33      → device pattern/map build (mb_sweepx_prepare)
  assemble!{:step|:iter}              src/Assemble.jl:470-487  → mb_sweepx_assemble
  Newmarkβcoefficients / decrement!   src/SweepX.jl:3-15, 98-132
  solve(SweepX{OX})                   src/SweepX.jl:179-226    (driver; the LU is host SuperLU here — cuDSS is not in this image —
                                                                and is outside the graft target, as UMFPACK is outside the hot path)
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import toolbox
from ._lib import MuscadeB200Error
from .engine import Engine
from .model import State, muscadeerror
from .synthetic import newmark_coefficients


class DofGroup:
    """DofGroup(dis,iΛ,iX,iU,iA)  (Assemble.jl:195-204)"""

    def __init__(self, dis, iΛ=(), iX=(), iU=(), iA=()):
        self.nX, self.nU, self.nA = len(dis.scaleX), len(dis.scaleU), len(dis.scaleA)
        self.iΛ, self.iX, self.iU, self.iA = (np.asarray(list(v), np.int64) for v in (iΛ, iX, iU, iA))
        nλ, nx, nu, na = len(self.iΛ), len(self.iX), len(self.iU), len(self.iA)
        self.jΛ = np.arange(1, nλ + 1); self.jX = nλ + np.arange(1, nx + 1)
        self.jU = nλ + nx + np.arange(1, nu + 1); self.jA = nλ + nx + nu + np.arange(1, na + 1)
        self.scaleΛ = dis.scaleΛ[self.iΛ - 1]; self.scaleX = dis.scaleX[self.iX - 1]
        self.scaleU = dis.scaleU[self.iU - 1]; self.scaleA = dis.scaleA[self.iA - 1]

    def getndof(self):
        return len(self.iΛ) + len(self.iX) + len(self.iU) + len(self.iA)


def allXdofs(model, dis):
    return DofGroup(dis, iX=range(1, model.getndof("X") + 1))


def decrement(state, ider, y, gr):
    """decrement!(s,ider,y,gr)  (Assemble.jl:206-211); ider is 1-based like the reference"""
    if ider <= len(state.Λ) and len(gr.iΛ): state.Λ[ider - 1][gr.iΛ - 1] -= y[gr.jΛ - 1] * gr.scaleΛ
    if ider <= len(state.X) and len(gr.iX): state.X[ider - 1][gr.iX - 1] -= y[gr.jX - 1] * gr.scaleX
    if ider <= len(state.U) and len(gr.iU): state.U[ider - 1][gr.iU - 1] -= y[gr.jU - 1] * gr.scaleU
    if ider == 1 and len(gr.iA): state.A[gr.iA - 1] -= y[gr.jA - 1] * gr.scaleA


def getdof_group(state, ider, y, gr):
    """getdof!(s,ider,y,gr)  (Assemble.jl:229-233); ider 0-based derivative order as in the reference"""
    if len(gr.iX): y[gr.jX - 1] = state.X[ider][gr.iX - 1] / gr.scaleX


class AssemblySweepX:
    """mutable struct AssemblySweepX{OX}: Lλ, Lλx, c  (SweepX.jl:17-23) + the device engine behind it"""

    def __init__(self, OX, engine, Lλ, Lλx):
        self.OX, self.engine, self.Lλ, self.Lλx = OX, engine, Lλ, Lλx
        self.c = newmark_coefficients(OX, 0.) if OX == 0 else np.zeros(7)
        self.host_types = []      # (ieletyp, EleTyp, EletypDisassembler)


class Asm:
    """asm[iarray,ieletyp]: the reference's index maps, fetched from the device on demand (1-based, (ndof, nele) like Julia)"""

    def __init__(self, engine, neletyp):
        self.engine, self.neletyp, self._cache = engine, neletyp, {}

    def __getitem__(self, key):
        iarray, ieletyp = key
        if (ieletyp) not in self._cache:
            a1, a2 = self.engine.sweepx_asm(ieletyp)
            self._cache[ieletyp] = (np.ascontiguousarray(a1.T), np.ascontiguousarray(a2.T))
        return self._cache[ieletyp][iarray - 1]


def prepare(OX, model, dis, device=0):
    """out,asm,Xdofgr = prepare(AssemblySweepX{OX},model,dis)  (SweepX.jl:24-33)"""
    Xdofgr = allXdofs(model, dis)
    ndof = Xdofgr.getndof()
    eng = Engine(device)
    eng.set_ndofU(model.getndof("U"))
    host_types = []
    for k, (et, ed) in enumerate(zip(model.ele, dis.dis)):
        # ElementCost on a device beam: in an X-analysis R = ∂L/∂Λ of L = getlagrangian(target) + cost(eleres) is the target's own residual (src/BasicElements.jl:117-132,
        # src/Assemble.jl:623-680), so the wrapped beams run through the beam kernels unchanged
        target = et.extra.get("target") if (et.ElType.kind == "elementcost" and isinstance(et.extra, dict)) else None
        if et.ElType.kind == "eulerbeam3d" or (target is not None and target.kind == "eulerbeam3d"):
            udof = ed.U.shape[1] > 0
            eng.add_eulerbeam3d(et.eleobj, ed.X, ed.scaleX, udof=udof, idxU=ed.U if udof else None, scaleU=ed.scaleU if udof else None)
        elif et.ElType.kind == "bar3d":
            udof = ed.U.shape[1] > 0
            eng.add_bar3d(et.eleobj, ed.X, ed.scaleX, udof=udof, idxU=ed.U if udof else None, scaleU=ed.scaleU if udof else None)
        elif et.ElType.kind == "soilcontact":
            eng.add_soilcontact(et.eleobj, ed.X, ed.scaleX)
        elif et.ElType.kind == "hostcost":
            eng.add_host_elements(ed.X)      # costs have no Λ-dependence: zero residual in an X-analysis, but their dofs are in the pattern
        else:
            if (ed.U.shape[1] or ed.A.shape[1]) and not getattr(et.ElType, "takes_UA", False):
                muscadeerror("host-evaluated element types with U- or A-dofs must declare takes_UA (their residual then receives U and A as values): %s" % (et.key,))
            eng.add_host_elements(ed.X)
            host_types.append((k + 1, et, ed))
    nnz = eng.sweepx_prepare(ndof)
    colptr, rowval = eng.sweepx_pattern()
    Lλ = np.zeros(ndof)
    Lλx = sp.csc_matrix((np.ones(nnz), rowval - 1, colptr - 1), shape=(ndof, ndof))
    out = AssemblySweepX(OX, eng, Lλ, Lλx)
    out.host_types = host_types
    return out, Asm(eng, len(model.ele)), Xdofgr


def _host_elements(out, state, mission, t):
    """addin!{mission} for element types the host evaluates (SweepX.jl:45-96 applied to constant-tangent elements)."""
    OX = out.OX
    a1, a2, a3, b1, b2, b3 = out.c[:6]
    step = mission == "step" and OX > 0
    for ityp, et, ed in out.host_types:
        X = [state.X[d][ed.X - 1] for d in range(OX + 1)]
        Ue = state.U[0][ed.U - 1] if ed.U.shape[1] and len(state.U) else None
        Ae = state.A[ed.A - 1] if ed.A.shape[1] else None
        R, K0, K1, K2 = et.residual(X, t, U=Ue, A=Ae)
        s = ed.scaleX
        K = K0.copy()
        if K1 is not None and OX >= 1: K = K + a1 * K1
        if K2 is not None and OX >= 2: K = K + b1 * K2
        Ke = K * s[None, :, None] * s[None, None, :]                     # scale_i · ∂R_i/∂X_j · scale_j
        Re = R * s[None, :]
        Rp = None
        if step:
            Rp = np.zeros_like(Re)
            if K1 is not None: Rp += np.einsum("eij,ej->ei", K1, a2 * X[1] + (a3 * X[2] if OX >= 2 else 0.)) * s[None, :]
            if K2 is not None and OX >= 2: Rp += np.einsum("eij,ej->ei", K2, b2 * X[1] + b3 * X[2]) * s[None, :]
        nx = Re.shape[1]
        out.engine.set_host_elements(ityp, Re, Rp, np.ascontiguousarray(Ke.transpose(0, 2, 1).reshape(-1, nx * nx)))   # entry i+nx·j


def assemble(mission, out, asm, dis, model, state, Δt, dbg=None):
    """assemble!{mission}(out,asm,dis,model,state,Δt,dbg)  (Assemble.jl:470-476) — one C-ABI call for all device element types."""
    _host_elements(out, state, mission, state.time)
    OX = out.OX
    X = [state.X[d] for d in range(OX + 1)]
    U0 = state.U[0] if len(state.U) and state.U[0].size else None
    out.engine.sweepx_assemble(OX, mission, X, out.c, U0=U0, t=state.time, Llambda=out.Lλ, nzval=out.Lλx.data, dbg=dbg)


def newmark_decrement(OX, state, Δx, Xdofgr, c, firstiter, buf):
    """Newmarkβdecrement!{OX}  (SweepX.jl:98-132)"""
    a1, a2, a3, b1, b2, b3 = c[:6]
    if OX == 0:
        decrement(state, 1, Δx, Xdofgr); return
    xp, xpp = buf
    if OX == 2:
        if firstiter:
            getdof_group(state, 1, xp, Xdofgr); getdof_group(state, 2, xpp, Xdofgr)
            a = a2 * xp + a3 * xpp; b = b2 * xp + b3 * xpp
            Δxp = a1 * Δx + a; Δxpp = b1 * Δx + b
        else:
            Δxp = a1 * Δx; Δxpp = b1 * Δx
        decrement(state, 1, Δx, Xdofgr); decrement(state, 2, Δxp, Xdofgr); decrement(state, 3, Δxpp, Xdofgr)
    else:
        if firstiter:
            getdof_group(state, 1, xp, Xdofgr)
            Δxp = a1 * Δx + a2 * xp
        else:
            Δxp = a1 * Δx
        decrement(state, 1, Δx, Xdofgr); decrement(state, 2, Δxp, Xdofgr)


def solve(OX, initialstate, time, β=0.25, γ=0.5, maxiter=50, maxΔx=1e-5, maxLλ=np.inf, verbose=False, device=0, dbg=None, device_state=False):
    """solve(SweepX{OX};initialstate,time,β,γ,maxiter,maxΔx,maxLλ)  (SweepX.jl:179-226) → list of states, one per time step.
    device_state=True keeps state.X on the GPU between iterations: Newmarkβdecrement! runs there (mb_sweepx_newmark_decrement) and X is
    downloaded only for saved states (and for host-evaluated element types)."""
    model, dis = initialstate.model, initialstate.dis
    out, asm, Xdofgr = prepare(OX, model, dis, device)
    n = Xdofgr.getndof()
    buf = (np.zeros(n), np.zeros(n))
    cΔx2, cLλ2 = maxΔx ** 2, maxLλ ** 2
    state = initialstate.copy().with_orders(1, OX + 1, 1)
    states = []
    citer = 0
    eng = out.engine
    if device_state:
        eng.set_dof_scale(Xdofgr.scaleX)
        eng.set_state(state.X[: OX + 1])
    try:
        for step, t in enumerate(time, 1):
            oldt = state.time
            state.time = float(t)
            Δt = t - oldt
            if Δt <= 0 and OX > 0:
                muscadeerror("Time step length not strictly positive at step=%3d" % step)
            out.c = newmark_coefficients(OX, Δt if OX > 0 else 0., β, γ)
            for iiter in range(1, maxiter + 1):
                citer += 1
                firstiter = iiter == 1
                mission, dbgi = "step" if firstiter else "iter", dict(dbg or {}, solver="SweepX", step=step, iiter=iiter)
                if device_state:
                    if out.host_types:
                        state.X[: OX + 1] = eng.get_state(OX)
                        _host_elements(out, state, mission, state.time)
                    U0 = state.U[0] if len(state.U) and state.U[0].size else None
                    eng.sweepx_assemble_resident(OX, mission, out.c, U0=U0, t=state.time, Llambda=out.Lλ, nzval=out.Lλx.data, dbg=dbgi)
                else:
                    assemble(mission, out, asm, dis, model, state, Δt, dbgi)
                try:
                    lu = spla.splu(out.Lλx.tocsc())
                except RuntimeError:
                    muscadeerror("matrix factorization failed at step=%i, iiter=%i" % (step, iiter))
                Δx = lu.solve(out.Lλ)
                if device_state:
                    Δx2, Lλ2 = eng.newmark_decrement(OX, firstiter, Δx, out.c)
                else:
                    Δx2, Lλ2 = float(Δx @ Δx), float(out.Lλ @ out.Lλ)
                    newmark_decrement(OX, state, Δx, Xdofgr, out.c, firstiter, buf)
                if Δx2 <= cΔx2 and Lλ2 <= cLλ2:
                    if device_state:
                        state.X[: OX + 1] = eng.get_state(OX)
                    if verbose:
                        print("    step %3d converged in %3d iterations. |Δx|=%7.1e |Lλ|=%7.1e" % (step, iiter, Δx2 ** .5, Lλ2 ** .5))
                    states.append(State(state.time, state.Λ, [x.copy() for x in state.X], state.U, state.A, state.SP, model, dis))
                    break
                if iiter == maxiter:
                    muscadeerror("no convergence of step %3d after %3d iterations |Δx|=%g / %g, |Lλ|=%g / %g" % (step, iiter, Δx2 ** .5, maxΔx, Lλ2 ** .5, maxLλ))
    finally:
        out.engine.close()
    return states
