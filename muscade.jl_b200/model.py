"""Host-side mirror of the reference's model description and disassembler (setup time; integer structures reproduced exactly).

  Model / addnode! / addelement! / setscale! / initialize!      src/ModelDescription.jl:15-324
  Disassembler / EletypDisassembler / State                     src/Assemble.jl:3-159

Julia's `!` names become `addnode`, `addelement`, `setscale`, `initialize`.  All dof / node / element numbers are 1-based
as in the reference.  `addelement` is vectorised over the rows of `nodID` (a 10M-element mesh is one call), but assigns dof
numbers in exactly the reference's order of first appearance (element-major, then element-dof order, per class;
ModelDescription.jl:226-249).  Per-node / per-dof element lists (Node.eleID, Dof.eleID), only used by `describe`, are not kept.
"""
import numpy as np

from ._lib import MuscadeB200Error

CLASSES = ("X", "U", "A")


def muscadeerror(msg, dbg=None):
    raise MuscadeB200Error(msg, dbg)


class DofTyp:
    def __init__(self, clas, field):
        self.clas, self.field, self.scale = clas, field, 1.0
        self.dofID = []          # list of arrays of idof (appended per addelement call)


class EleTyp:
    """One entry of model.ele / model.eleobj: all elements of one concrete element type."""

    def __init__(self, ElType, key, doflist):
        self.ElType, self.key = ElType, key
        self.inod, self.clas, self.field = doflist
        self.nodID = np.zeros((0, max(self.inod) if self.inod else 0), np.int64)
        self.dofID = np.zeros((0, len(self.inod)), np.int64)      # idof within its class, per element dof
        self.eleobj = None                                        # ElType-specific batch (e.g. (nele,69) array)
        self.extra = None

    @property
    def nele(self):
        return self.dofID.shape[0]

    def idof(self, clas):
        return [j for j, c in enumerate(self.clas) if c == clas]

    # ---- host-evaluated element types (user closures) -----------------------------------------------------------------------------------------
    # The reference stores args / costargs / gargs in every element object (src/BasicElements.jl:198-208,275-284,406-438), so two addelement! calls
    # with the same closure and different arguments land in ONE element type whose elements differ.  `runs` keeps the keyword data of each call with
    # the element range it covers; the evaluators below call the element type once per run and stitch the results in element order.
    def _runs(self):
        runs = getattr(self, "runs", None)
        return runs if runs else [(0, self.nele, self.extra if self.extra is not None else self.eleobj)]

    def residual(self, X, t, U=None, A=None):
        """ElType.residual over all elements: X = [X₀,X′,…] of shape (nele,nx) each → (R, K0, K1, K2), K1/K2 None when no run has them.
        Element types that read U- or A-dofs (as plain values: an X-analysis does not differentiate with respect to them, src/SweepX.jl:45-96)
        declare `takes_UA = True` and receive U (nele,nu), A (nele,na) as keyword arguments."""
        k = getattr(self.ElType, "kind", None)
        if k == "lagrangian" or (k in ("elementcost", "elementconstraint") and getattr(self.ElType, "kind_general", None) == "lagrangian"):
            return self._residual_d2(X, t, U, A)      # host-evaluated wrappers: R = ∂L/∂Λ is the target's residual, costs and multipliers do not enter an X-analysis
        if getattr(self.ElType, "takes_UA", False):
            parts = [self.ElType.residual(ex, [x[a:b] for x in X], t, U=None if U is None else U[a:b], A=None if A is None else A[a:b]) for a, b, ex in self._runs()]
        else:
            parts = [self.ElType.residual(ex, [x[a:b] for x in X], t) for a, b, ex in self._runs()]
        if len(parts) == 1:
            return parts[0]
        out = []
        for k in range(4):
            cols = [p[k] for p in parts]
            if all(c is None for c in cols):
                out.append(None); continue
            ref = next(c for c in cols if c is not None)
            cols = [c if c is not None else np.zeros((b - a,) + ref.shape[1:]) for c, (a, b, _) in zip(cols, self._runs())]
            out.append(np.concatenate(cols))
        return tuple(out)

    def _residual_d2(self, X, t, U, A):
        """An X-analysis (SweepX) of an element type written against adiff2.D2 (toolbox.LagrangianElement): R and ∂R/∂X₀, ∂R/∂X′, ∂R/∂X″ by the host's dual numbers, U and A
        as plain values.  A type that only has `lagrangian` gives R = ∂L/∂Λ, as getresidual derives it (src/Assemble.jl:660-680)."""
        from .adiff2 import D2
        T = self.ElType
        nd, nx = len(X), X[0].shape[1]
        outs = []
        for a, b, ex in self._runs():
            n = b - a
            eo = self.eleobj[a:b] if self.eleobj is not None else None
            Uv = [[U[a:b][:, i] for i in range(U.shape[1])]] if U is not None else [[]]
            Av = [A[a:b][:, i] for i in range(A.shape[1])] if A is not None else []
            Xval = np.concatenate([x[a:b] for x in X], axis=1)
            R = np.zeros((n, nx)); K = [np.zeros((n, nx, nx)) for _ in range(nd)]
            if hasattr(T, "lagrangian"):
                v = D2.variables(np.concatenate([np.zeros((n, nx)), Xval], axis=1))
                Xv = [v[nx + d * nx: nx + (d + 1) * nx] for d in range(nd)]
                L = D2.lift(T.lagrangian(eo, ex, v[:nx], Xv, Uv, Av, t, None))
                g, H = L.grad(n, nx * (nd + 1)), L.hess(n, nx * (nd + 1))
                R = g[:, :nx]
                for d in range(nd):
                    K[d] = H[:, :nx, nx + d * nx: nx + (d + 1) * nx]
            else:
                v = D2.variables(Xval)
                Xv = [v[d * nx: (d + 1) * nx] for d in range(nd)]
                Rv = T.residual(eo, ex, Xv, Uv, Av, t, None)
                for i in range(nx):
                    Ri = D2.lift(Rv[i])
                    R[:, i] = np.broadcast_to(Ri.v, (n,))
                    gi = Ri.grad(n, nx * nd)
                    for d in range(nd):
                        K[d][:, i, :] = gi[:, d * nx: (d + 1) * nx]
            outs.append((R, K))
        R = np.concatenate([o[0] for o in outs])
        Ks = [np.concatenate([o[1][d] for o in outs]) for d in range(nd)] + [None] * (3 - nd)
        return R, Ks[0], Ks[1], Ks[2]

    def hessian(self, X, lam_e, t):
        fn = getattr(self.ElType, "hessian", None)
        if fn is None:
            return None
        parts = [fn(ex, [x[a:b] for x in X], lam_e[a:b], t) for a, b, ex in self._runs()]
        if all(p is None for p in parts):
            return None
        ref = next(p for p in parts if p is not None)
        return np.concatenate([p if p is not None else np.zeros((b - a,) + ref.shape[1:]) for p, (a, b, _) in zip(parts, self._runs())])

    def cost_derivs(self, x, t):
        parts = [self.ElType.cost_derivs(ex, x[a:b], t) for a, b, ex in self._runs()]
        return tuple(np.concatenate([p[k] for p in parts]) for k in range(3))


class Model:
    def __init__(self, ID="muscade_model"):
        self.ID = ID
        self.coord = []                      # model.nod[inod].coord
        self.coord_chunks = []               # (first inod, 2-D array) for nodes added as a matrix: vectorised lookup
        self.ele = []                        # list of EleTyp  (model.ele / model.eleobj)
        self.ndof = dict(X=0, U=0, A=0)      # length(model.dof[class])
        self.dof_nod = dict(X=[], U=[], A=[])   # model.dof[class][idof].nodID   (list of arrays)
        self.dof_typ = dict(X=[], U=[], A=[])   # model.dof[class][idof].idoftyp
        self.doftyp = []
        self.nod_dof = []                    # per idoftyp: array over nodes → idof (0 = node has no such dof)
        self.scaleΛ = 1.0
        self.locked = False

    # -- accessors (ModelDescription.jl:84-133)
    def getndof(self, clas=None):
        if clas is None:
            return sum(self.ndof.values())
        if isinstance(clas, (tuple, list)):
            return tuple(self.ndof[c] for c in clas)
        return self.ndof[clas]

    def getnele(self, ieletyp=None):
        return sum(e.nele for e in self.ele) if ieletyp is None else self.ele[ieletyp - 1].nele

    def getneletyp(self):
        return len(self.ele)

    def getidoftyp(self, clas, field):
        for i, d in enumerate(self.doftyp):
            if d.clas == clas and d.field == field:
                return i + 1
        return 0

    def getdoftyp(self, clas, field):
        i = self.getidoftyp(clas, field)
        if i == 0:
            muscadeerror("The model has no dof of class %s and field %s." % (clas, field))
        return self.doftyp[i - 1]

    def nnod(self):
        return len(self.coord)

    def assert_unlocked(self):
        if self.locked:
            muscadeerror("model %s is initialized and can no longer be edited" % self.ID)


def addnode(model, coord):
    """nodid = addnode!(model,coord)  (ModelDescription.jl:166-173): a vector adds one node, a matrix one node per row."""
    model.assert_unlocked()
    coord = np.asarray(coord, float)
    if coord.ndim == 1:
        model.coord.append(coord.copy())
        return len(model.coord)
    first = len(model.coord) + 1
    model.coord.extend(list(coord))
    model.coord_chunks.append((first, coord.copy()))
    return np.arange(first, first + coord.shape[0], dtype=np.int64)


def _node_coords(model, ids):
    """coordinates of many nodes at once: (n, dim) array; ragged dims (e.g. coordinate-less U/A nodes) give (n,0)"""
    if len(ids) == 0:
        return np.zeros((0, 0))
    for first, arr in model.coord_chunks:
        if ids.min() >= first and ids.max() < first + arr.shape[0]:
            return arr[ids - first]
    dims = {len(model.coord[n - 1]) for n in ids}
    if len(dims) != 1:
        return np.zeros((len(ids), 0))
    return np.asarray([model.coord[n - 1] for n in ids], float).reshape(len(ids), dims.pop())


def addelement(model, ElType, nodID, **kwargs):
    """eleid = addelement!(model,ElType,nodid;kwargs...)  (ModelDescription.jl:195-263).

    nodID: vector → one element, returns (ieletyp,iele); matrix (nele × nnod) → many, returns (ieletyp, array of iele)."""
    model.assert_unlocked()
    nodID = np.asarray(nodID, np.int64)
    single = nodID.ndim == 1
    if single:
        nodID = nodID[None, :]
    nele_new, nnod = nodID.shape
    key = ElType.typekey(**kwargs)
    doflist = ElType.doflist(**kwargs)
    inod, clas, field = doflist
    if nnod != (max(inod) if inod else 0):
        muscadeerror("Connecting element of type %s: Second dimension of inod (%d) must be equal to element's nnod (%d)" %
                     (ElType.__name__, nnod, max(inod) if inod else 0))
    # element objects first (the reference constructs ele1 before touching the model, so constructor errors leave it intact)
    coords = [_node_coords(model, nodID[:, k]) for k in range(nnod)]      # coords[k][iele] = coord(nod)[k] of element iele
    built = ElType.construct(coords, **kwargs)
    eleobj, extra = built if isinstance(built, tuple) else (built, None)
    # element type (getieletyp, ModelDescription.jl:205-213)
    ieletyp = 0
    for i, e in enumerate(model.ele):
        if e.key == key:
            ieletyp = i + 1
    if ieletyp == 0:
        model.ele.append(EleTyp(ElType, key, doflist))
        ieletyp = len(model.ele)
    et = model.ele[ieletyp - 1]
    iele_sofar = et.nele
    neledof = len(inod)
    # dof types, in doflist order (created while visiting the first element)
    idoftyp = np.zeros(neledof, np.int64)
    for j in range(neledof):
        t = model.getidoftyp(clas[j], field[j])
        if t == 0:
            model.doftyp.append(DofTyp(clas[j], field[j]))
            model.nod_dof.append(np.zeros(0, np.int64))
            t = len(model.doftyp)
        idoftyp[j] = t
    nn = model.nnod()
    for t in set(idoftyp.tolist()):
        tab = model.nod_dof[t - 1]
        if tab.size < nn + 1:
            model.nod_dof[t - 1] = np.concatenate([tab, np.zeros(nn + 1 - tab.size, np.int64)])
    # dofs: existing ones are looked up, new ones numbered per class in order of first appearance (element-major, eledof-minor)
    node = nodID[:, np.asarray(inod, np.int64) - 1] if neledof else np.zeros((nele_new, 0), np.int64)      # (nele, neledof)
    typ = np.broadcast_to(idoftyp[None, :], node.shape)
    existing = np.zeros(node.shape, np.int64)
    for j in range(neledof):
        existing[:, j] = model.nod_dof[idoftyp[j] - 1][node[:, j]]
    dof = existing.copy()
    newmask = existing == 0
    if newmask.any():
        code = (typ.astype(np.int64) * (nn + 1) + node).reshape(-1)
        flat_new = np.flatnonzero(newmask.reshape(-1))
        uniq, first = np.unique(code[flat_new], return_index=True)
        order = np.argsort(first, kind="stable")
        uniq = uniq[order]                                 # new (doftyp,node) keys in order of first appearance
        utyp = uniq // (nn + 1); unod = uniq % (nn + 1)
        uclass = np.array([model.doftyp[t - 1].clas for t in utyp])
        newid = np.zeros(uniq.size, np.int64)
        for c in CLASSES:
            sel = np.flatnonzero(uclass == c)
            newid[sel] = model.ndof[c] + 1 + np.arange(sel.size)
            if sel.size:
                model.ndof[c] += sel.size
                model.dof_nod[c].append(unod[sel]); model.dof_typ[c].append(utyp[sel])
        for t in np.unique(utyp):
            sel = utyp == t
            model.nod_dof[t - 1][unod[sel]] = newid[sel]
            model.doftyp[t - 1].dofID.append(newid[sel])
        for j in range(neledof):
            dof[:, j] = model.nod_dof[idoftyp[j] - 1][node[:, j]]
    et.nodID = np.concatenate([et.nodID, nodID]) if et.nodID.shape[1] == nnod else nodID
    et.dofID = np.concatenate([et.dofID, dof])
    if et.eleobj is None:
        et.eleobj, et.extra = eleobj, extra
    else:
        et.eleobj = np.concatenate([et.eleobj, eleobj])
    if extra is not None:                                   # per-call keyword data (args, costargs, gargs …) with the elements it belongs to
        if not hasattr(et, "runs"):
            et.runs = []
        et.runs.append((iele_sofar, iele_sofar + nele_new, extra))
    iele = iele_sofar + np.arange(1, nele_new + 1)
    return (ieletyp, int(iele[0])) if single else (ieletyp, iele)


def setscale(model, scale=None, Λscale=None):
    """setscale!(model;scale,Λscale)  (ModelDescription.jl:287-299); scale = dict(X=dict(tx=10,rx=1), A=dict(drag=3.))"""
    model.assert_unlocked()
    if scale is not None:
        for d in model.doftyp:
            if d.clas in scale and d.field in scale[d.clas]:
                d.scale = float(scale[d.clas][d.field])
    if Λscale is not None:
        model.scaleΛ = float(Λscale)


class EletypDisassembler:
    """dis.dis[ieletyp]: index[iele].X|U|A (here (nele,n) int64 arrays) and scale.Λ|X|U|A  (Assemble.jl:14-17)."""

    def __init__(self, X, U, A, sΛ, sX, sU, sA):
        self.X, self.U, self.A = X, U, A
        self.scaleΛ, self.scaleX, self.scaleU, self.scaleA = sΛ, sX, sU, sA


class Disassembler:
    """Disassembler(model)  (Assemble.jl:32-99)"""

    def __init__(self, model):
        NX, NU, NA = model.getndof(("X", "U", "A"))
        self.scaleΛ = np.zeros(NX); self.scaleX = np.zeros(NX); self.scaleU = np.zeros(NU); self.scaleA = np.zeros(NA)
        self.fieldX = [None] * NX; self.fieldU = [None] * NU; self.fieldA = [None] * NA
        self.dis = []
        for et in model.ele:
            jX, jU, jA = et.idof("X"), et.idof("U"), et.idof("A")
            sc = np.array([model.getdoftyp(c, f).scale for c, f in zip(et.clas, et.field)])
            sX = sc[jX]; sΛ = sX * model.scaleΛ; sU = sc[jU]; sA = sc[jA]
            iX = np.ascontiguousarray(et.dofID[:, jX]); iU = np.ascontiguousarray(et.dofID[:, jU]); iA = np.ascontiguousarray(et.dofID[:, jA])
            if et.nele:
                self.scaleΛ[iX - 1] = sΛ[None, :]; self.scaleX[iX - 1] = sX[None, :]
                self.scaleU[iU - 1] = sU[None, :]; self.scaleA[iA - 1] = sA[None, :]
                for cls, jj, idx, fl in (("X", jX, iX, self.fieldX), ("U", jU, iU, self.fieldU), ("A", jA, iA, self.fieldA)):
                    for k, j in enumerate(jj):
                        for d in np.unique(idx[:, k]):
                            fl[d - 1] = et.field[j]
            self.dis.append(EletypDisassembler(iX, iU, iA, sΛ, sX, sU, sA))


class State:
    """State{nΛder,nXder,nUder,TSP}  (Assemble.jl:106-159): time, Λ, X, U (tuples of vectors), A, SP, model, dis."""

    def __init__(self, time, Λ, X, U, A, SP, model, dis):
        self.time, self.Λ, self.X, self.U, self.A, self.SP, self.model, self.dis = time, list(Λ), list(X), list(U), A, SP, model, dis

    @classmethod
    def zeros(cls, model, dis, nΛder=1, nXder=1, nUder=1, time=-np.inf):
        nX, nU, nA = model.getndof(("X", "U", "A"))
        return cls(time, [np.zeros(nX) for _ in range(nΛder)], [np.zeros(nX) for _ in range(nXder)], [np.zeros(nU) for _ in range(nUder)],
                   np.zeros(nA), None, model, dis)

    def with_orders(self, nΛder, nXder, nUder, SP=None):
        """State{nΛder,nXder,nUder}(s): shallow copy that drops or zero-pads derivatives (Assemble.jl:131-144, ∂n → zeros)."""
        def fit(v, n):
            return [v[i] if i < len(v) else np.zeros_like(v[0]) for i in range(n)]
        return State(self.time, fit(self.Λ, nΛder), fit(self.X, nXder), fit(self.U, nUder), self.A, self.SP if SP is None else SP, self.model, self.dis)

    def copy(self, time=None, SP=None):
        """Base.copy(s::State): deep copy except SP, model, dis (Assemble.jl:159)"""
        return State(self.time if time is None else time, [v.copy() for v in self.Λ], [v.copy() for v in self.X], [v.copy() for v in self.U],
                     self.A.copy(), self.SP if SP is None else SP, self.model, self.dis)


def initialize(model, nΛder=1, nXder=1, nUder=1, time=-np.inf):
    """initialstate = initialize!(model)  (ModelDescription.jl:319-324)"""
    model.assert_unlocked()
    model.locked = True
    dis = Disassembler(model)
    return State.zeros(model, dis, nΛder, nXder, nUder, time)


def getdof(state, field, nodID=None, clas="X", order=0):
    """getdof(state;class,field,nodID,order)  (src/Output.jl:21-54), minimal: values of the dofs of one field at the given nodes."""
    model = state.model
    t = model.getidoftyp(clas, field)
    if t == 0:
        muscadeerror("The model has no dof of class %s and field %s." % (clas, field))
    tab = model.nod_dof[t - 1]
    nodes = np.arange(1, model.nnod() + 1) if nodID is None else np.atleast_1d(np.asarray(nodID, np.int64))
    idof = tab[nodes]
    if (idof == 0).any():
        muscadeerror("some nodes have no dof of field %s" % field)
    vec = {"X": state.X, "U": state.U}[clas][order] if clas in ("X", "U") else state.A
    return vec[idof - 1]
