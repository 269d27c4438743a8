"""DirectXUA on the B200 engine, beam-specialised path (IA = 0; xua.py is the general form with IA = 1, A-dofs, several experiments): host-side mirror of

  prepare(AssemblyDirect{OX,OU,IA})     src/DirectXUA.jl:22-56    → class-pair patterns / maps built on the device
  makepattern / preparebig              src/DirectXUA.jl:245-315  → block pattern here (a few entries per step), Lvv structure on the device
  assemblebig!{:matrices}               src/DirectXUA.jl:316-356  → mb_direct_assemble
  finitediff                            src/FiniteDifferences.jl:2-31

Time steps are 0-based in this module's arguments (`lo`, `hi`, `step`); block numbers follow the reference: block of (step s, class α)
is 3·s + α with α = 0 Λ, 1 X, 2 U.  A handle owns the block columns of steps [lo,hi) — the unit of sharding over GPUs.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ErrInfo, check, ptr, MuscadeB200Error
from .engine import Engine, _f64
from .model import muscadeerror

_FD = [[[(0, 1.)]],
       [[(0, -1.), (1, 1.)], [(-1, -1.), (0, 1.)], [(-1, -.5), (1, .5)]],
       [[(0, 1.), (1, -2.), (2, 1.)], [(-2, 1.), (-1, -2.), (0, 1.)], [(-1, 1.), (0, -2.), (1, 1.)]]]


def finitediff(order, n, s):
    """finitediff(order,n,s): list of (Δs,w); s is the 1-based step as in the reference (src/FiniteDifferences.jl:8-31)"""
    if order > 0 and n < 6:
        muscadeerror("Number of steps must be ≥6")
    if order == 0:
        return _FD[0][0]
    return _FD[order][0 if s == 1 else (1 if s == n else 2)]


def block_pattern(OX, OU, nstep, lo, hi):
    """makepattern(IA=0,[nstep],out) restricted to the block columns of steps [lo,hi): CSC over blocks (bcolptr, browval), 0-based block
    numbers 3·step+class.  Which blocks exist depends only on (OX,OU) and the finite-difference stencils (src/DirectXUA.jl:253-275)."""
    nder = (1, OX + 1, OU + 1)
    cols = {}
    for istep in range(max(1, lo + 1 - 2), min(nstep, hi + 2) + 1):            # evaluation steps that can reach the owned columns
        for a in range(3):
            for b in range(3):
                if a == 0 and b == 0:
                    continue                                                  # Lλλ is always zero (DirectXUA.jl:41)
                for ad in range(1, nder[a] + 1):
                    for bd in range(1, nder[b] + 1):
                        for (das, _) in finitediff(ad - 1, nstep, istep):
                            for (dbs, _) in finitediff(bd - 1, nstep, istep):
                                sc = istep + dbs - 1
                                if lo <= sc < hi:
                                    cols.setdefault(3 * sc + b, set()).add(3 * (istep + das - 1) + a)
    bcolptr = [0]; browval = []
    for bc in range(3 * lo, 3 * hi):
        rows = sorted(cols.get(bc, ()))
        browval.extend(rows); bcolptr.append(len(browval))
    return np.asarray(bcolptr, np.int32), np.asarray(browval, np.int32)


class DirectEngine(Engine):
    """Engine + the DirectXUA entry points of the C ABI"""

    def direct_prepare(self, OX, OU, ndofX, ndofU, nstep, lo, hi, dt):
        bcolptr, browval = block_pattern(OX, OU, nstep, lo, hi)
        ncol, nnz = C.c_int64(), C.c_int64()
        check(self.h, self.L.mb_direct_prepare(self.h, OX, OU, int(ndofX), int(ndofU), int(nstep), int(lo), int(hi), float(dt), ptr(bcolptr), ptr(browval),
                                               C.byref(ncol), C.byref(nnz)))
        self.OX, self.OU, self.ndofX, self.ndofU, self.nstep, self.lo, self.hi = OX, OU, int(ndofX), int(ndofU), nstep, lo, hi
        self.ncol, self.nnzbig = ncol.value, nnz.value
        return self.ncol, self.nnzbig

    def class_pattern(self, which):
        """which: 0 X×X (=Λ×X, X×Λ), 1 X×U, 2 U×X, 3 U×U → (colptr, rowval) 1-based"""
        n = C.c_int64()
        check(self.h, self.L.mb_direct_class_pattern(self.h, which, C.byref(n), None, None))
        ncols = self.ndofU if which in (1, 3) else self.ndofX
        colptr = np.zeros(ncols + 1, np.int64); rowval = np.zeros(n.value, np.int64)
        check(self.h, self.L.mb_direct_class_pattern(self.h, which, C.byref(n), ptr(colptr), ptr(rowval)))
        return colptr, rowval

    def direct_asm(self, ityp, which):
        nele = self.groups[ityp - 1][1]
        nx, nu = self.groups[ityp - 1][2], 3
        ni = nu if which in (2, 3) else nx; nj = nu if which in (1, 3) else nx
        out = np.zeros((nele, ni * nj), np.int64)
        check(self.h, self.L.mb_direct_get_asm(self.h, ityp, which, ptr(out)))
        return out

    def set_time0(self, t0):
        """state[step].time = t0 + step·dt"""
        check(self.h, self.L.mb_direct_set_time0(self.h, float(t0)))

    def set_lambda_scale(self, Λscale):
        check(self.h, self.L.mb_direct_set_lambda_scale(self.h, float(Λscale)))

    def set_state(self, step, X, U0=None):
        X = [_f64(x) for x in X]
        check(self.h, self.L.mb_direct_set_state(self.h, int(step), ptr(X[0]), ptr(X[1]) if len(X) > 1 else None, ptr(X[2]) if len(X) > 2 else None, ptr(_f64(U0))))

    # ---- Newton update on the device
    def set_lambda(self, step, Lam):
        check(self.h, self.L.mb_direct_set_lambda(self.h, int(step), ptr(_f64(Lam))))

    def get_state(self, step):
        """→ (X[0..OX], U0, Λ) of a stored step"""
        X = [np.empty(self.ndofX) for _ in range(self.OX + 1)]
        U = np.empty(self.ndofU); Lam = np.empty(self.ndofX)
        check(self.h, self.L.mb_direct_get_state(self.h, int(step), ptr(X[0]), ptr(X[1]) if self.OX >= 1 else None, ptr(X[2]) if self.OX >= 2 else None,
                                                 ptr(U) if self.ndofU else None, ptr(Lam)))
        return X, U, Lam

    def set_dof_scale(self, scaleL=None, scaleX=None, scaleU=None):
        check(self.h, self.L.mb_direct_set_dof_scale(self.h, ptr(_f64(scaleL)), ptr(_f64(scaleX)), ptr(_f64(scaleU))))

    def decrement(self, dv, s0=0, s1=None):
        """decrementbig! (DirectXUA.jl:357-383) on the stored steps; dv = Δv rows of steps [s0,s1) → Δ² (Λ,X,U), max over the owned steps"""
        s1 = self.nstep if s1 is None else s1
        dv = _f64(dv)
        assert dv.shape == ((s1 - s0) * (2 * self.ndofX + self.ndofU),)
        d2 = np.zeros(3)
        check(self.h, self.L.mb_direct_decrement(self.h, int(s0), int(s1), ptr(dv), ptr(d2)))
        return d2

    def rebase(self, new_lo):
        """slide an interior window (3 ≤ lo, hi ≤ nstep−3) to [new_lo, new_lo+hi−lo): same Lvv structure, rows shifted → row shift relative to prepare"""
        sh = C.c_int64()
        check(self.h, self.L.mb_direct_rebase(self.h, int(new_lo), C.byref(sh)))
        d = int(new_lo) - self.lo
        self.lo += d; self.hi += d
        return sh.value

    def stored_range(self):
        return max(0, self.lo - 2), min(self.nstep, self.hi + 2)

    def set_state_dev(self, step, X_ptrs, U_ptr=None):
        """set_state from device memory (raw device addresses, e.g. torch .data_ptr())"""
        p = [C.c_void_p(int(a)) for a in X_ptrs] + [None] * (3 - len(X_ptrs))
        check(self.h, self.L.mb_direct_set_state(self.h, int(step), p[0], p[1], p[2], C.c_void_p(int(U_ptr)) if U_ptr else None))

    def direct_set_host_elements(self, step, ityp, R, dR, GX):
        check(self.h, self.L.mb_direct_set_host_elements(self.h, int(step), int(ityp), ptr(_f64(R)), ptr(_f64(dR)), ptr(_f64(GX))))

    def set_host_xx(self, step, i, j, v):
        """L2[X,X][1,1] entries (1-based X dofs, one per (i,j), already scaled) of host-evaluated types that are not linear in X; empty lists clear"""
        i = np.ascontiguousarray(i, np.int64); j = np.ascontiguousarray(j, np.int64); v = _f64(np.asarray(v, float))
        check(self.h, self.L.mb_direct_set_host_xx(self.h, int(step), len(i), ptr(i), ptr(j), ptr(v)))

    def set_gauge_measurements(self, step, ityp, epsm):
        """measured strains of one stored step for a costed beam type: (Ngauge,) for all elements or (nele,Ngauge)"""
        em = np.ascontiguousarray(epsm, np.float64)
        check(self.h, self.L.mb_direct_set_gauge_measurements(self.h, int(step), int(ityp), ptr(em), 1 if em.ndim == 2 else 0))

    def set_gauge_times(self, time):
        """measurements of every stored step from the costs' own measured(t) (time[k] = time of step k)"""
        elo, ehi = self.stored_range()
        for ityp, cost in self.gauge_costs:
            for k in range(elo, ehi):
                self.set_gauge_measurements(k, ityp, cost.measured(float(time[k])))

    def set_host_cost(self, step, gX=None, hX=None, gU=None, hU=None):
        check(self.h, self.L.mb_direct_set_host_cost(self.h, int(step), ptr(_f64(gX)), ptr(_f64(hX)), ptr(_f64(gU)), ptr(_f64(hU))))

    def sparser(self, rtol=1e-9):
        """sparser!(cLvv,Lvv,rtol) (SparseTools.jl:196-199) on the device → number of entries kept"""
        n = C.c_int64()
        check(self.h, self.L.mb_direct_sparser(self.h, float(rtol), C.byref(n)))
        self.cnnz = n.value
        return self.cnnz

    def sparse(self):
        colptr = np.zeros(self.ncol + 1, np.int64); rowval = np.zeros(self.cnnz, np.int64); nzval = np.zeros(self.cnnz)
        check(self.h, self.L.mb_direct_get_sparse(self.h, ptr(colptr), ptr(rowval), ptr(nzval)))
        return colptr, rowval, nzval

    def direct_assemble(self, eval_range=None, build_big=True, Lvv=None, Lv=None):
        where = ErrInfo()
        lo, hi = (-1, -1) if eval_range is None else eval_range
        rc = self.L.mb_direct_assemble(self.h, lo, hi, int(build_big), ptr(Lvv), ptr(Lv), C.byref(where))
        if rc == _lib.MB_ERR_NAN:
            raise MuscadeB200Error("residual(...) returned NaN in R, FB or derivatives", dict(step=where.step, ieletyp=where.ieletyp, iele=where.iele))
        check(self.h, rc)

    def big_pattern(self):
        colptr = np.zeros(self.ncol + 1, np.int64); rowval = np.zeros(self.nnzbig, np.int64)
        check(self.h, self.L.mb_direct_big_pattern(self.h, ptr(colptr), ptr(rowval)))
        return colptr, rowval

    def step_block(self, step, which, der=0):
        n = {0: self.ndofX}.get(which)
        if n is None:
            nn = C.c_int64()
            check(self.h, self.L.mb_direct_class_pattern(self.h, {1: 0, 2: 0, 3: 1, 4: 2}[which], C.byref(nn), None, None))
            n = nn.value
        out = np.zeros(n)
        check(self.h, self.L.mb_direct_get_step_block(self.h, int(step), which, der, ptr(out)))
        return out

    def step_ptrs(self, step):
        p = [C.c_void_p() for _ in range(3)]; n = [C.c_int64() for _ in range(3)]
        check(self.h, self.L.mb_direct_step_ptrs(self.h, int(step), C.byref(p[0]), C.byref(n[0]), C.byref(p[1]), C.byref(n[1]), C.byref(p[2]), C.byref(n[2])))
        return [(p[i].value, n[i].value) for i in range(3)]

    def halo_exchange(self):
        """time shards: the per-step blocks of the two steps at each end of the owned range go to the neighbours, the halo steps' blocks come in
        (NCCL inside the shim, asynchronous on the handle's stream)"""
        check(self.h, self.L.mb_direct_halo_exchange(self.h))

    def direct_time(self, reps=3):
        ms = np.zeros(3, np.float32)
        check(self.h, self.L.mb_direct_time_dev(self.h, reps, ms))
        self.last_element_ms = float(ms[2])          # element kernels of the owned steps alone
        return float(ms[0]), float(ms[1])


def prepare(OX, OU, model, dis, nstep, dt, lo=0, hi=None, device=0, t0=0.):
    """prepare(AssemblyDirect{OX,OU,0},model,dis) + preparebig(0,[nstep],out) for the owned steps [lo,hi)"""
    hi = nstep if hi is None else hi
    eng = DirectEngine(device)
    eng.host_costs = []
    eng.host_types = []
    eng.gauge_costs = []                             # (ityp, QuadraticGaugeCost): ElementCost on strain-gauged beams, evaluated by the device
    for et, ed in zip(model.ele, dis.dis):
        udof = ed.U.shape[1] > 0
        target = et.extra.get("target") if (et.ElType.kind == "elementcost" and isinstance(et.extra, dict)) else None
        if et.ElType.kind == "eulerbeam3d":
            eng.add_eulerbeam3d(et.eleobj, ed.X, ed.scaleX, udof=udof, idxU=ed.U if udof else None, scaleU=ed.scaleU if udof else None)
        elif target is not None and target.kind == "eulerbeam3d":
            cost = et.extra["cost"]
            if not (hasattr(cost, "sigma") and hasattr(cost, "measured") and "G" in et.extra and et.extra["req"] == ("ε",)):
                muscadeerror("ElementCost on the device: ElementType = StrainGaugeOnEulerBeam3D, req = ('ε',), cost = QuadraticGaugeCost")
            ityp = eng.add_eulerbeam3d(et.eleobj, ed.X, ed.scaleX, udof=udof, idxU=ed.U if udof else None, scaleU=ed.scaleU if udof else None)
            G = np.ascontiguousarray(et.extra["G"], np.float64)
            check(eng.h, eng.L.mb_direct_set_gauge_cost(eng.h, ityp, G.shape[0], ptr(G), float(cost.sigma)))
            eng.gauge_costs.append((ityp, cost))
        elif et.ElType.kind == "bar3d":
            eng.add_bar3d(et.eleobj, ed.X, ed.scaleX, udof=udof, idxU=ed.U if udof else None, scaleU=ed.scaleU if udof else None)
        elif et.ElType.kind == "soilcontact":
            eng.add_soilcontact(et.eleobj, ed.X, ed.scaleX)
        elif et.ElType.kind == "hostcost":
            eng.host_costs.append((et, ed))          # evaluated by the host, merged by the device (set_host_cost)
        elif et.ElType.kind == "host" and ed.U.shape[1] == 0 and ed.A.shape[1] == 0:
            ityp = eng.add_host_elements(ed.X)       # Hold, DofLoad, …: second-order branch, evaluated by the host (set_host_elements)
            eng.host_types.append((ityp, et, ed))
        else:
            muscadeerror("DirectXUA on the device supports EulerBeam3D, Bar3D, SoilContact and SingleDofCost element types in this version: %s" % (et.key,))
    eng.direct_prepare(OX, OU, model.getndof("X"), model.getndof("U"), nstep, lo, hi, dt)
    eng.set_time0(t0)
    eng.set_lambda_scale(model.scaleΛ)
    return eng


def host_costs(eng, step, X0, U0, t):
    """lagrangian of the host-evaluated SingleDofCost types at one step → (gX,hX,gU,hU), scaled (∂/∂(dof/scale)), and Σ cost"""
    nX, nU = eng.ndofX, eng.ndofU
    g = dict(X=np.zeros(nX), U=np.zeros(nU)); h = dict(X=np.zeros(nX), U=np.zeros(nU))
    total = 0.
    for et, ed in eng.host_costs:
        clas = et.extra["clas"]
        idx = (ed.X if clas == "X" else ed.U)[:, 0] - 1
        sc = (ed.scaleX if clas == "X" else ed.scaleU)[0]
        c, c1, c2 = et.cost_derivs((X0 if clas == "X" else U0)[idx], t)
        np.add.at(g[clas], idx, c1 * sc); np.add.at(h[clas], idx, c2 * sc * sc)
        total += float(c.sum())
    return g["X"], h["X"], g["U"], h["U"], total


def host_elements(eng, step, X, Lam, t, Λscale):
    """second-order branch (DirectXUA.jl:152-171) of the host-evaluated X-class types (Hold, DofLoad, DofConstraint{:X}; L2[X,X] through `hessian` where R is not linear in X):
    L = Λ∘R ⇒ L1[Λ] = R·sΛ, L1[X][der] = (∂R/∂X_der)ᵀΛ·sX, L2[Λ,X][1,der] = ∂R/∂X_der·sΛ·sX (and its transpose in L2[X,Λ])"""
    nd = eng.OX + 1
    xx = {}                                     # (i,j) → Σ over elements, in element order (types whose residual is not linear in X)
    for ityp, et, ed in eng.host_types:
        Xe = [X[d][ed.X - 1] for d in range(nd)]
        R, K0, K1, K2 = et.residual(Xe, t)
        H = et.hessian(Xe, Lam[ed.X - 1], t)
        if H is not None:                       # L2[X,X][1,1] += Λ·∂²R/∂X²·sX·sX  (DirectXUA.jl:121-150)
            H = H * ed.scaleX[None, :, None] * ed.scaleX[None, None, :]
            for e in range(H.shape[0]):
                for a in range(H.shape[1]):
                    for b in range(H.shape[2]):
                        key = (int(ed.X[e, a]), int(ed.X[e, b]))
                        xx[key] = xx.get(key, 0.) + H[e, a, b]
        nele, nx = R.shape
        sX = ed.scaleX; sL = ed.scaleX * Λscale
        lam = Lam[ed.X - 1]
        dR = np.zeros((nele, nd * nx, nx)); GX = np.zeros((nele, nx, nd))
        for der, K in enumerate([K0, K1, K2][:nd]):
            if K is None:
                continue
            dR[:, nx * der: nx * (der + 1), :] = (K * sL[None, :, None] * sX[None, None, :]).transpose(0, 2, 1)      # [e][(nx·der+j)][i]
            GX[:, :, der] = np.einsum("ek,eki->ei", lam, K) * sX[None, :]
        eng.direct_set_host_elements(step, ityp, R * sL[None, :], dR, GX)
    if xx or getattr(eng, "_had_xx", False):
        keys = list(xx.keys())
        eng.set_host_xx(step, [k[0] for k in keys], [k[1] for k in keys], [xx[k] for k in keys])
        eng._had_xx = True


def solve(OX, OU, initialstate, time, maxiter=50, maxΔλ=1e-5, maxΔx=1e-5, maxΔu=1e-5, verbose=False, device=0, sparser_rtol=1e-20):
    """solve(DirectXUA{OX,OU,0};initialstate=[s],time=[t0:Δt:t1],…) (src/DirectXUA.jl:440-508), one experiment: Newton iterations on the
    all-steps KKT system.  Element evaluation, Lvv/Lv assembly, sparser! and decrementbig! run on the device; the factorisation (UMFPACK in
    the reference) is SuperLU on the host.  Returns the list of converged states, one per time step."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from .model import State
    model, dis = initialstate.model, initialstate.dis
    time = np.asarray(time, float)
    nstep = len(time)
    dt = float(time[1] - time[0])
    if OU != 0:
        muscadeerror("device decrementbig! stores ∂0(U) only: OU must be 0")
    eng = prepare(OX, OU, model, dis, nstep, dt, t0=float(time[0]), device=device)
    try:
        s0 = initialstate.with_orders(1, OX + 1, OU + 1)
        for k in range(nstep):
            eng.set_state(k, s0.X, s0.U[0])
            eng.set_lambda(k, s0.Λ[0])
        eng.set_dof_scale(dis.scaleΛ, dis.scaleX, dis.scaleU)
        maxΔ2 = np.array([maxΔλ, maxΔx, maxΔu]) ** 2
        Lv = np.zeros(eng.ncol)
        eng.set_gauge_times(time)
        for it in range(1, maxiter + 1):
            if eng.host_costs or eng.host_types:
                for k in range(nstep):
                    X, U, Lam = eng.get_state(k)
                    if eng.host_costs:
                        eng.set_host_cost(k, *host_costs(eng, k, X[0], U, time[k])[:4])
                    if eng.host_types:
                        host_elements(eng, k, X, Lam, time[k], model.scaleΛ)
            eng.direct_assemble(Lv=Lv)
            eng.sparser(sparser_rtol)
            colptr, rowval, nzval = eng.sparse()
            try:
                LU = spla.splu(sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(eng.ncol, eng.ncol)))
            except RuntimeError:
                muscadeerror("Lvv matrix factorization failed")
            Δ2 = eng.decrement(LU.solve(Lv))
            if verbose:
                print("    iteration %3d  maxₜ|ΔΛ|=%7.1e |ΔX|=%7.1e |ΔU|=%7.1e" % ((it,) + tuple(np.sqrt(Δ2))))
            if (Δ2 <= maxΔ2).all():
                break
            if it == maxiter:
                muscadeerror("no convergence after %3d iterations." % it)
        out = []
        for k in range(nstep):
            X, U, Lam = eng.get_state(k)
            out.append(State(float(time[k]), [Lam], X, [U], s0.A, dict(γ=0., iter=it), model, dis))
        return out
    finally:
        eng.close()
