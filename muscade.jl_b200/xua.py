"""DirectXUA{OX,OU,IA} in its GENERAL form on the B200 engine (csrc/mb_xua.cu): A-dofs, U-costs, user Lagrangians, several experiments.

  prepare(AssemblyDirect{OX,OU,IA})          src/DirectXUA.jl:22-56     → XUAEngine.prepare            (mb_xua_prepare: maps, patterns, Lvv on the device)
  addin!{:matrices}(out::AssemblyDirect,…)   src/DirectXUA.jl:70-171    → packets()                    (host: the element's closure, differentiated with D2)
  assembleA! / assemblebig!{:matrices}       src/DirectXUA.jl:316-356   → XUAEngine.assemblebig        (device: mb_xua_add_A / mb_xua_add_step)
  sparser! / decrementbig!                   src/DirectXUA.jl:357-383   → XUAEngine.sparser / decrement (device)
  solve(DirectXUA{OX,OU,IA};…)               src/DirectXUA.jl:438-513   → solve()

Element types on this path are Python classes with the reference's element API, written against adiff2.D2 (the host's dual numbers):

    class El(ElementType):
        kind = "lagrangian"            # evaluated by the host, assembled by the device
        no_second_order = False        # Muscade.no_second_order(::Type{El}) = Val(false|true)   (src/ElementAPI.jl)
        acost = False                  # True: an Acost — skipped by assemble!, taken by assembleA!  (src/Assemble.jl:477,515)
        residual(eleobj, extra, X, U, A, t, SP)      → [R₁…R_nx]      X[der][i], U[der][i], A[i]: D2 (or floats)
        lagrangian(eleobj, extra, Λ, X, U, A, t, SP) → L              (either one; `lagrangian` wins, as getlagrangian does, src/Assemble.jl:714-726)

Classes are numbered 1 Λ, 2 X, 3 U, 4 A (`ind`, src/DirectXUA.jl:14); experiments and steps are 1-based in the engine calls, as in the reference."""
import ctypes as C

import numpy as np

from ._lib import check, ptr, ErrInfo, MB_ERR_NAN, MuscadeB200Error
from .adiff2 import D2
from .engine import Engine, _f64, _i64
from .model import State, muscadeerror


def packets(et, ed, OX, OU, IA, Λ, X, U, A, t, SP=None, assembleA=False):
    """∇L (nele,Np), ∇²L (nele,Np,Np) of every element of type `et` at one state, in the order of partials Λ, X₀…, U₀…, A (src/DirectXUA.jl:127-135), scaled by
    revariate's scales — filled as the reference's addin! method for the type fills out (Acost :70-84, no_second_order :85-120, second order :152-171).
    Λ, X[der], U[der], A: model vectors of the state.
    assembleA: the packet of assembleA! (Acost types only; Np = na, unscaled).  NB: as the reference is written, Acost types ALSO go through assemble! at every step —
    `assemble_!(…,eleobj::Acost,…) = nothing` (src/Assemble.jl:477) never matches the Vector{<:Acost} it is called with, so they take the second-order branch like any
    element; test/TestDirectXUA.jl:107 pins that (out.L2[A,A] = 2e-14·I after the last step).  This path reproduces it."""
    T = et.ElType
    nele = et.nele
    nx, nu, na = ed.X.shape[1], ed.U.shape[1], ed.A.shape[1]
    parts_g, parts_H = [], []
    for a, b, ex in et._runs():
        eo = et.eleobj[a:b] if et.eleobj is not None else None
        n = b - a
        Ae = A[ed.A[a:b] - 1] if na else np.zeros((n, 0))
        if assembleA:                                                      # addin!(…,eleobj::Acost,A) :70-84 — ∂A = revariate{2}(A), no scale
            v = D2.variables(Ae)
            L = D2.lift(T.lagrangian(eo, ex, None, None, None, v, t, SP))
            parts_g.append(L.grad(n, na)); parts_H.append(L.hess(n, na))
            continue
        Xe = [X[d][ed.X[a:b] - 1] if nx else np.zeros((n, 0)) for d in range(OX + 1)]
        Ue = [U[d][ed.U[a:b] - 1] if nu else np.zeros((n, 0)) for d in range(OU + 1)]
        Le = Λ[ed.X[a:b] - 1] if nx else np.zeros((n, 0))
        Np = nx + nx * (OX + 1) + nu * (OU + 1) + na * IA
        if getattr(T, "kind", None) == "host":
            # X-class toolbox types with closed-form partials (Hold, DofLoad, DofConstraint: `residual(extra,X,t)` → R, ∂R/∂X₀, ∂R/∂X′, ∂R/∂X″ and `hessian` → Σ Λₖ∂²Rₖ/∂X²):
            # the second-order branch L = Λ∘R written out (src/DirectXUA.jl:152-171) — ∇L[Λ] = R·sΛ, ∇L[X_d] = (∂R/∂X_d)ᵀΛ·sX, ∇²L[Λ,X_d] = ∂R/∂X_d·sΛ·sX, ∇²L[X₀,X₀] = hessian·sX·sX
            Xs = [x for x in Xe]
            R, K0, K1, K2 = T.residual(ex, Xs, t)
            Hx = T.hessian(ex, Xs, Le, t) if hasattr(T, "hessian") else None
            g = np.zeros((n, Np)); H = np.zeros((n, Np, Np))
            sL, sX = ed.scaleΛ, ed.scaleX
            g[:, :nx] = R * sL[None, :]
            for d, K in enumerate([K0, K1, K2][:OX + 1]):
                if K is None:
                    continue
                c0 = nx + nx * d
                g[:, c0:c0 + nx] = np.einsum("ek,eki->ei", Le, K) * sX[None, :]
                blk = K * sL[None, :, None] * sX[None, None, :]
                H[:, :nx, c0:c0 + nx] = blk; H[:, c0:c0 + nx, :nx] = blk.transpose(0, 2, 1)
            if Hx is not None:
                H[:, nx:2 * nx, nx:2 * nx] = Hx * sX[None, :, None] * sX[None, None, :]
            parts_g.append(g); parts_H.append(H)
            continue
        if getattr(T, "no_second_order", False) and not hasattr(T, "lagrangian"):
            # :85-120 — revariate{1}((;X,U,A),scale); Lλ += R, Lλβ += ∂R/∂β, Lβλ += its transpose; nothing else
            vals = np.concatenate(Xe + Ue + ([Ae] if IA else []), axis=1)
            sc = np.concatenate([ed.scaleX] * (OX + 1) + [ed.scaleU] * (OU + 1) + ([ed.scaleA] if IA else []))
            v = D2.variables(vals, sc)
            p = 0
            Xv = []
            for d in range(OX + 1): Xv.append(v[p:p + nx]); p += nx
            Uv = []
            for d in range(OU + 1): Uv.append(v[p:p + nu]); p += nu
            Av = v[p:p + na] if IA else [Ae[:, i] for i in range(na)]
            R = T.residual(eo, ex, Xv, Uv, Av, t, SP)
            g = np.zeros((n, Np)); H = np.zeros((n, Np, Np))
            for i in range(nx):
                Ri = D2.lift(R[i])
                g[:, i] = np.broadcast_to(Ri.v, (n,))
                dR = Ri.grad(n, Np - nx)
                H[:, i, nx:] = dR; H[:, nx:, i] = dR
            parts_g.append(g); parts_H.append(H)
            continue
        # :152-171 — revariate{2}((;Λ,X,U,A),scale); L = getlagrangian(…)
        vals = np.concatenate([Le] + Xe + Ue + ([Ae] if IA else []), axis=1)
        sc = np.concatenate([ed.scaleΛ] + [ed.scaleX] * (OX + 1) + [ed.scaleU] * (OU + 1) + ([ed.scaleA] if IA else []))
        v = D2.variables(vals, sc)
        p = nx
        Lv_ = v[0:nx]
        Xv = []
        for d in range(OX + 1): Xv.append(v[p:p + nx]); p += nx
        Uv = []
        for d in range(OU + 1): Uv.append(v[p:p + nu]); p += nu
        Av = v[p:p + na] if IA else [Ae[:, i] for i in range(na)]
        if hasattr(T, "lagrangian"):
            L = D2.lift(T.lagrangian(eo, ex, Lv_, Xv, Uv, Av, t, SP))
        else:                                                               # L = Λ ∘₁ R  (src/Assemble.jl:721-726)
            R = T.residual(eo, ex, Xv, Uv, Av, t, SP)
            L = D2(np.zeros(n))
            for i in range(nx):
                L = L + Lv_[i] * R[i]
        parts_g.append(L.grad(n, Np)); parts_H.append(L.hess(n, Np))
    if not parts_g:
        Np = na if assembleA else nx + nx * (OX + 1) + nu * (OU + 1) + na * IA
        return np.zeros((0, Np)), np.zeros((0, Np, Np))
    g, H = np.concatenate(parts_g), np.concatenate(parts_H)
    if not (np.isfinite(g).all() and np.isfinite(H).all()):
        bad = int(np.flatnonzero(~(np.isfinite(g).all(axis=1) & np.isfinite(H).all(axis=(1, 2))))[0]) + 1
        muscadeerror("lagrangian(...) returned NaN in L, FB or derivatives", dict(iele=bad, eletyp=T.__name__))
    assert g.shape[0] == nele
    return g, H


def shard_steps(nstep, rank, world):
    """time shards of the general form: the (iexp, istep) pairs (0-based) rank `rank` of `world` evaluates — the steps of all experiments, flattened, dealt round-robin
    (every rank gets steps of every part of the time axis: the end-of-series stencils cost the same as interior ones, so no balancing is needed beyond the count)"""
    out, k = [], 0
    for iexp, n in enumerate(nstep):
        for istep in range(n):
            if k % world == rank:
                out.append((iexp, istep))
            k += 1
    return out


class XUAEngine(Engine):
    """Engine + the mb_xua_* entry points"""

    def prepare(self, model, dis, OX, OU, IA, nstep, dt, Xwhite=False, XUindep=False, UAindep=False, XAindep=False):
        self.model, self.dis, self.OX, self.OU, self.IA = model, dis, OX, OU, IA
        self.nstep = [int(n) for n in nstep]; self.dt = [float(d) for d in dt]; self.nexp = len(self.nstep)
        self.nX, self.nU, self.nA = model.getndof(("X", "U", "A"))
        it = C.c_int32()
        self.devtypes = {}                       # ityp (1-based) → QuadraticGaugeCost or None: element types evaluated by the device kernels
        for ityp, (et, ed) in enumerate(zip(model.ele, dis.dis), 1):
            T = et.ElType
            target = et.extra.get("target") if (T.kind == "elementcost" and isinstance(et.extra, dict)) else None
            if T.kind in ("bar3d", "soilcontact"):
                udof = ed.U.shape[1] > 0
                idev = self.add_soilcontact(et.eleobj, ed.X, ed.scaleX) if T.kind == "soilcontact" else \
                    self.add_bar3d(et.eleobj, ed.X, ed.scaleX, udof=udof, idxU=ed.U if udof else None, scaleU=ed.scaleU if udof else None)
                check(self.h, self.L.mb_xua_add_device_eletyp(self.h, idev, C.byref(it)))
                self.devtypes[ityp] = None
                continue
            if T.kind == "eulerbeam3d" or (target is not None and target.kind == "eulerbeam3d"):
                udof = ed.U.shape[1] > 0
                idev = self.add_eulerbeam3d(et.eleobj, ed.X, ed.scaleX, udof=udof, idxU=ed.U if udof else None, scaleU=ed.scaleU if udof else None)
                check(self.h, self.L.mb_xua_add_device_eletyp(self.h, idev, C.byref(it)))
                cost = None
                if target is not None:
                    cost = et.extra["cost"]
                    if not (hasattr(cost, "sigma") and hasattr(cost, "measured") and "G" in et.extra and et.extra["req"] == ("ε",)):
                        muscadeerror("ElementCost on the device: ElementType = StrainGaugeOnEulerBeam3D, req = ('ε',), cost = QuadraticGaugeCost")
                    G = _f64(et.extra["G"])
                    check(self.h, self.L.mb_xua_set_gauge_cost(self.h, it.value, G.shape[0], ptr(G), cost.sigma, float(model.scaleΛ)))
                self.devtypes[ityp] = cost
                continue
            if "lagrangian" not in (getattr(T, "kind", None), getattr(T, "kind_general", None)) and T.kind != "host":
                muscadeerror("the general DirectXUA path takes EulerBeam3D (device), ElementCost on strain-gauged beams (device) and host-evaluated element types "
                             "written against adiff2.D2 (kind = 'lagrangian'): %s" % (et.key,))
            iX, iU, iA = _i64(ed.X), _i64(ed.U), _i64(ed.A)
            check(self.h, self.L.mb_xua_add_eletyp(self.h, et.nele, iX.shape[1], iU.shape[1], iA.shape[1], ptr(iX), ptr(iU), ptr(iA),
                                                   1 if getattr(T, "acost", False) else 0, C.byref(it)))
        flags = (1 if Xwhite else 0) | (2 if XUindep else 0) | (4 if UAindep else 0) | (8 if XAindep else 0)
        nbig, nnz = C.c_int64(), C.c_int64()
        ns = np.asarray(self.nstep, np.int64); dts = np.asarray(self.dt, float)
        check(self.h, self.L.mb_xua_prepare(self.h, OX, OU, IA, self.nX, self.nU, self.nA, self.nexp, ptr(ns), ptr(dts), flags, C.byref(nbig), C.byref(nnz)))
        self.nbig, self.nnzbig = nbig.value, nnz.value
        self.ndofc = {1: self.nX, 2: self.nX, 3: self.nU, 4: self.nA}
        check(self.h, self.L.mb_xua_set_dof_scale(self.h, ptr(_f64(dis.scaleΛ)), ptr(_f64(dis.scaleX)), ptr(_f64(dis.scaleU)), ptr(_f64(dis.scaleA))))
        check(self.h, self.L.mb_xua_set_lambda_scale(self.h, float(model.scaleΛ)))
        return self.nbig, self.nnzbig

    def set_time0(self, iexp, t0):
        """state[iexp][istep].time = t0 + (istep−1)·Δt[iexp] for the device types that read the time (Bar3D's weight ramp)"""
        check(self.h, self.L.mb_xua_set_time0(self.h, int(iexp), float(t0)))

    # ---- structures (checked against the reference's goldens)
    def class_pattern(self, α, β):
        n = C.c_int64()
        check(self.h, self.L.mb_xua_class_pattern(self.h, α, β, C.byref(n), None, None))
        colptr = np.zeros(self.ndofc[β] + 1, np.int64); rowval = np.zeros(n.value, np.int64)
        check(self.h, self.L.mb_xua_class_pattern(self.h, α, β, C.byref(n), ptr(colptr), ptr(rowval)))
        return colptr, rowval

    def asm(self, ityp, α, β=0):
        """asm[arrnum(α),ityp] (β = 0) or asm[arrnum(α,β),ityp] as the reference stores it: (n, nele), 1-based"""
        et, ed = self.model.ele[ityp - 1], self.dis.dis[ityp - 1]
        n = {1: ed.X.shape[1], 2: ed.X.shape[1], 3: ed.U.shape[1], 4: ed.A.shape[1]}
        rows = n[α] * (n[β] if β else 1)
        out = np.zeros((et.nele, rows), np.int64)
        check(self.h, self.L.mb_xua_get_asm(self.h, ityp, α, β, ptr(out)))
        return np.ascontiguousarray(out.T)

    def big_pattern(self):
        colptr = np.zeros(self.nbig + 1, np.int64); rowval = np.zeros(self.nnzbig, np.int64)
        check(self.h, self.L.mb_xua_big_pattern(self.h, ptr(colptr), ptr(rowval)))
        return colptr, rowval

    def big_asm(self):
        """Lvvasm as dict(colptr, rowval, nzval = list of index vectors) and Lvasm = pgr"""
        nb, nblock = C.c_int64(), C.c_int64()
        check(self.h, self.L.mb_xua_big_asm(self.h, C.byref(nb), C.byref(nblock), None, None, None, None, None))
        bcolptr = np.zeros(nb.value + 1, np.int64); browval = np.zeros(nblock.value, np.int64); boff = np.zeros(nblock.value + 1, np.int64)
        basm = np.zeros(self.nnzbig, np.int64); pgr = np.zeros(nb.value + 1, np.int64)
        check(self.h, self.L.mb_xua_big_asm(self.h, C.byref(nb), C.byref(nblock), ptr(bcolptr), ptr(browval), ptr(boff), ptr(basm), ptr(pgr)))
        return dict(colptr=bcolptr, rowval=browval, nzval=[basm[boff[b]:boff[b + 1]] for b in range(nblock.value)]), pgr

    def out_shape(self, α, β):
        na, nb = C.c_int32(), C.c_int32()
        check(self.h, self.L.mb_xua_out_shape(self.h, α, β, C.byref(na), C.byref(nb)))
        return na.value, nb.value

    def get_out(self, α, β=0, αder=1, βder=1):
        """out.L1[α][αder] (β = 0) or out.L2[α,β][αder,βder].nzval of the last step added"""
        if β == 0:
            out = np.zeros(self.ndofc[α])
        else:
            n = C.c_int64()
            check(self.h, self.L.mb_xua_class_pattern(self.h, α, β, C.byref(n), None, None))
            out = np.zeros(n.value)
        check(self.h, self.L.mb_xua_get_out(self.h, α, β, αder, βder, ptr(out)))
        return out

    # ---- assembly
    def zero(self):
        check(self.h, self.L.mb_xua_zero(self.h))

    def set_packet(self, ityp, g, H):
        check(self.h, self.L.mb_xua_set_packet(self.h, ityp, ptr(_f64(g)), ptr(_f64(H))))

    def get_packet(self, ityp):
        et, ed = self.model.ele[ityp - 1], self.dis.dis[ityp - 1]
        Np = ed.X.shape[1] * (self.OX + 2) + ed.U.shape[1] * (self.OU + 1) + ed.A.shape[1] * self.IA
        g = np.zeros((et.nele, Np)); H = np.zeros((et.nele, Np, Np))
        check(self.h, self.L.mb_xua_get_packet(self.h, ityp, ptr(g), ptr(H)))
        return g, H

    def get_gauge(self, ityp):
        """(εₐₓ,κ) (nele,4), ∂(εₐₓ,κ)/∂X₀ (nele,4,12), cost (nele,) of a costed device type at the last evaluated step"""
        n = self.model.ele[ityp - 1].nele
        e4 = np.zeros((n, 4)); J = np.zeros((n, 4, 12)); c = np.zeros(n)
        check(self.h, self.L.mb_xua_get_gauge(self.h, ityp, ptr(e4), ptr(J), ptr(c)))
        return e4, J, c

    def add_A(self):
        check(self.h, self.L.mb_xua_add_A(self.h))

    def add_step(self, iexp, istep):
        check(self.h, self.L.mb_xua_add_step(self.h, iexp, istep))

    def assembleA(self, state, SP=None):
        """assembleA!{:matrices}(out,asm,dis,model,state,dbg) and its addition into the A block (src/DirectXUA.jl:320-326)"""
        for ityp, (et, ed) in enumerate(zip(self.model.ele, self.dis.dis), 1):
            if getattr(et.ElType, "acost", False):
                g, H = packets(et, ed, self.OX, self.OU, self.IA, None, None, None, state.A, state.time, SP, assembleA=True)
                self.set_packet(ityp, g, H)
        self.add_A()

    def assemble_step(self, iexp, istep, state, SP=None):
        """assemble!{:matrices}(out,…,state[iexp][istep],…) and its weighted addition into Lvv / Lv (src/DirectXUA.jl:328-353)"""
        if self.devtypes:                        # device element types read the device-resident state
            self.put_state(iexp, istep, state)
            for ityp, cost in self.devtypes.items():
                if cost is not None:
                    em = _f64(cost.measured(state.time))
                    check(self.h, self.L.mb_xua_set_gauge_measurements(self.h, ityp, ptr(em), 1 if em.ndim == 2 else 0))
            where = ErrInfo()
            rc = self.L.mb_xua_eval_device(self.h, iexp, istep, C.byref(where))
            if rc == MB_ERR_NAN:
                raise MuscadeB200Error("residual(...) returned NaN in R, FB or derivatives", dict(ieletyp=where.ieletyp, iele=where.iele, step=istep))
            check(self.h, rc)
        for ityp, (et, ed) in enumerate(zip(self.model.ele, self.dis.dis), 1):           # Acost types included: see packets()
            if ityp in self.devtypes:
                continue
            g, H = packets(et, ed, self.OX, self.OU, self.IA, state.Λ[0], state.X, state.U, state.A, state.time, SP)
            self.set_packet(ityp, g, H)
        self.add_step(iexp, istep)

    def assemblebig(self, states, SP=None):
        """assemblebig!{:matrices}(Lvv,Lv,…,state,nstep,Δt,SP,dbg): states[iexp][istep] (0-based lists)"""
        self.zero()
        if self.IA == 1:
            self.assembleA(states[0][0], SP)
        for iexp in range(self.nexp):
            for istep in range(self.nstep[iexp]):
                self.assemble_step(iexp + 1, istep + 1, states[iexp][istep], SP)

    def assemblebig_sharded(self, states, rank, world, SP=None):
        """time shards: rank r of `world` evaluates and adds the steps (flattened over the experiments) s with s mod world == r — and the A step on rank 0 — then the
        copies of Lvv / Lv are summed over the ranks (mb_xua_allreduce_big; comm_init first).  Every rank ends with the whole system."""
        self.zero()
        if self.IA == 1 and rank == 0:
            self.assembleA(states[0][0], SP)
        for iexp, istep in shard_steps(self.nstep, rank, world):
            self.assemble_step(iexp + 1, istep + 1, states[iexp][istep], SP)
        check(self.h, self.L.mb_xua_allreduce_big(self.h))

    def big(self):
        nz = np.zeros(self.nnzbig); Lv = np.zeros(self.nbig)
        check(self.h, self.L.mb_xua_get_big(self.h, ptr(nz), ptr(Lv)))
        return nz, Lv

    def sparser(self, rtol=1e-20):
        n = C.c_int64()
        check(self.h, self.L.mb_xua_sparser(self.h, float(rtol), C.byref(n)))
        colptr = np.zeros(self.nbig + 1, np.int64); rowval = np.zeros(n.value, np.int64); nzval = np.zeros(n.value)
        check(self.h, self.L.mb_xua_get_sparse(self.h, ptr(colptr), ptr(rowval), ptr(nzval)))
        return colptr, rowval, nzval

    # ---- states / decrementbig!
    def put_state(self, iexp, istep, s):
        X = np.concatenate([_f64(x) for x in s.X[:self.OX + 1]]) if self.nX else None
        U = np.concatenate([_f64(u) for u in s.U[:self.OU + 1]]) if self.nU else None
        check(self.h, self.L.mb_xua_set_state(self.h, iexp, istep, ptr(_f64(s.Λ[0])) if self.nX else None, ptr(X), ptr(U), ptr(_f64(s.A)) if self.nA else None))

    def fetch_state(self, iexp, istep, s):
        """device state[iexp][istep] → the State object (in place; A into the shared vector)"""
        Lam = np.zeros(self.nX); X = np.zeros((self.OX + 1, self.nX)); U = np.zeros((self.OU + 1, self.nU)); A = np.zeros(self.nA)
        check(self.h, self.L.mb_xua_get_state(self.h, iexp, istep, ptr(Lam) if self.nX else None, ptr(X) if self.nX else None, ptr(U) if self.nU else None,
                                              ptr(A) if self.nA else None))
        s.Λ[0][:] = Lam
        for d in range(self.OX + 1): s.X[d][:] = X[d]
        for d in range(self.OU + 1): s.U[d][:] = U[d]
        s.A[:] = A

    def decrement(self, dv):
        d2 = np.zeros(4)
        check(self.h, self.L.mb_xua_decrement(self.h, ptr(_f64(dv)), ptr(d2)))
        return d2


def solve(OX, OU, IA, initialstate, time, maxiter=50, maxΔλ=1e-5, maxΔx=1e-5, maxΔu=1e-5, maxΔa=1e-5, verbose=False, device=0, sparser_rtol=1e-20, **flags):
    """solve(DirectXUA{OX,OU,IA};initialstate=[s₁,…],time=[t₁,…],…) (src/DirectXUA.jl:438-513): Newton iterations on the all-steps, all-experiments KKT system.
    `time[iexp]`: equally spaced times of experiment iexp.  Assembly, sparser! and decrementbig! on the device; the element closures and the factorisation
    (UMFPACK in the reference, SuperLU here) on the host.  Returns state[iexp][istep]."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    if len(initialstate) != len(time):
        muscadeerror("initialstate and time must be of the same length")
    time = [np.asarray(t, float) for t in time]
    nstep = [len(t) for t in time]; dt = [float(t[1] - t[0]) for t in time]
    model, dis = initialstate[0].model, initialstate[0].dis
    A = initialstate[0].A.copy()                                  # all state[iexp][istep].A are === (src/DirectXUA.jl:452)
    states = []
    for s0, t in zip(initialstate, time):
        s0 = s0.with_orders(1, OX + 1, OU + 1)
        states.append([State(float(ti), [v.copy() for v in s0.Λ], [v.copy() for v in s0.X], [v.copy() for v in s0.U], A, dict(γ=0., iter=1), model, dis) for ti in t])
    eng = XUAEngine(device)
    try:
        eng.prepare(model, dis, OX, OU, IA, nstep, dt, **flags)
        for ie in range(len(time)):
            for k in range(nstep[ie]):
                eng.put_state(ie + 1, k + 1, states[ie][k])
        maxΔ2 = np.array([maxΔλ, maxΔx, maxΔu, maxΔa][:4 if IA else 3]) ** 2
        for it in range(1, maxiter + 1):
            SP = dict(γ=0., iter=it)
            eng.assemblebig(states, SP)
            _, Lv = eng.big()
            colptr, rowval, nzval = eng.sparser(sparser_rtol)
            try:
                LU = spla.splu(sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(eng.nbig, eng.nbig)))
            except RuntimeError:
                muscadeerror("Lvv matrix factorization failed")
            Δ2 = eng.decrement(LU.solve(Lv))[:len(maxΔ2)]
            for ie in range(len(time)):
                for k in range(nstep[ie]):
                    eng.fetch_state(ie + 1, k + 1, states[ie][k])
                    states[ie][k].SP = SP
            if verbose:
                print("    iteration %3d  " % it + "  ".join("|Δ%s|=%7.1e" % (c, np.sqrt(d)) for c, d in zip("ΛXUA", Δ2)))
            if (Δ2 <= maxΔ2).all():
                return states
        muscadeerror("no convergence after %3d iterations. \n" % maxiter)
    finally:
        eng.close()
