"""EigX over the device assembly (src/EigX.jl:28-48,101-128): the caller on the other side of the hot path that reads its per-step blocks directly.

`solve(EigX{ℝ|ℂ};state,nmod)` makes ONE `assemble!{:matrices}(out::AssemblyDirect{2,0,0},…,state₀,…)` and takes K = out.L2[Λ,X][1,1], C = …[1,2], M = …[1,3] —
the blocks this path's element kernels and reductions produce for every DirectXUA time step.  `assemble_matrices` returns them as scipy CSC matrices (structure = the
reference's `asmmat!` pattern of the Λ-X class pair), from the beam-specialised path (`mb_direct_*`) when every element type of the model is one it evaluates, else from the
general form (`mb_xua_*`: user elements written against adiff2.D2, device beams beside them).  `sparser!(·,droptol)` (src/SparseTools.jl:172-199) is applied as the reference
does.  The generalised eigenproblem itself (`geneig`, src/Eigenmodes.jl:40-75: shift-invert at 0 with KrylovKit, vectors normalised to largest entry 1) is host work outside the
path; `solve` does it with scipy's ARPACK in the same shift-invert form so that the assembly can be checked against the reference's own goldens (test/TestEigX.jl:34-39, 48-53).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import directxua as _dx
from . import xua as _xua
from .model import State, muscadeerror

_NSTEP = 6          # finitediff needs ≥ 6 steps for time derivatives (src/FiniteDifferences.jl:8-31); only step 0 is evaluated


def _specialised(model, dis):
    for et, ed in zip(model.ele, dis.dis):
        k = et.ElType.kind
        target = et.extra.get("target") if (k == "elementcost" and isinstance(et.extra, dict)) else None
        if k in ("eulerbeam3d", "bar3d", "soilcontact", "hostcost"):
            continue
        if target is not None and target.kind == "eulerbeam3d":
            continue
        if k == "host" and ed.U.shape[1] == 0 and ed.A.shape[1] == 0:
            continue
        return False
    return True


def sparser(M, droptol):
    """sparser!(M,droptol): drop the entries below droptol·max|M| (src/SparseTools.jl:172-199)"""
    M = M.tocsc(copy=True)
    if M.nnz:
        M.data[np.abs(M.data) < droptol * np.abs(M.data).max()] = 0.
        M.eliminate_zeros()
    return M


def assemble_matrices(state, device=0):
    """→ K, C, M = out.L2[Λ,X][1,1], [1,2], [1,3] and L1[Λ] of assemble!{:matrices}(AssemblyDirect{2,0,0}) at `state` (src/EigX.jl:33-40): scipy CSC, scaled as the reference's"""
    model, dis = state.model, state.dis
    OX, OU = 2, 0
    s0 = state.with_orders(1, OX + 1, OU + 1)
    nX = model.getndof("X")
    if _specialised(model, dis):
        eng = _dx.prepare(OX, OU, model, dis, _NSTEP, 1., 0, 1, device=device, t0=float(s0.time) if np.isfinite(s0.time) else 0.)
        try:
            eng.set_state(0, s0.X, s0.U[0]); eng.set_lambda(0, s0.Λ[0])
            for ityp, cost in eng.gauge_costs:           # a strain cost does not enter L2[Λ,X]; its kernels still want measurements
                eng.set_gauge_measurements(0, ityp, cost.measured(float(s0.time) if np.isfinite(s0.time) else 0.))
            if eng.host_types:
                _dx.host_elements(eng, 0, s0.X, s0.Λ[0], s0.time, model.scaleΛ)
            eng.direct_assemble(eval_range=(0, 1), build_big=False)
            colptr, rowval = eng.class_pattern(0)
            blocks = [eng.step_block(0, 1, der) for der in range(OX + 1)]
            L1 = eng.step_block(0, 0)
        finally:
            eng.close()
    else:
        eng = _xua.XUAEngine(device)
        try:
            eng.prepare(model, dis, OX, OU, 0, [_NSTEP], [1.])
            eng.zero()
            eng.assemble_step(1, 1, s0)
            colptr, rowval = eng.class_pattern(1, 2)
            blocks = [eng.get_out(1, 2, 1, der + 1) for der in range(OX + 1)]
            L1 = eng.get_out(1)
        finally:
            eng.close()
    K, C, M = (sp.csc_matrix((b, rowval - 1, colptr - 1), shape=(nX, nX)) for b in blocks)
    return K, C, M, L1


def _normalize_inf(v):
    """normalize∞!: the largest term becomes 1 (src/Eigenmodes.jl:18-21)"""
    return v / v[np.argmax(np.abs(v) ** 2)]


class EigXRincrement:
    def __init__(self, dis, ω, Δx):
        self.dis, self.ω, self.Δx = dis, ω, Δx


class EigXCincrement:
    def __init__(self, dis, p, Δx):
        self.dis, self.p, self.Δx = dis, p, Δx


def solve(state, nmod=5, droptol=1e-9, complex_modes=False, device=0, **kw):
    """solve(EigX{ℝ};state,nmod,droptol) / solve(EigX{ℂ};…) (src/EigX.jl:28-48, 101-128) → EigXRincrement(ω, Δx) / EigXCincrement(p, Δx)"""
    K, C, M, _ = assemble_matrices(state, device)
    nX = K.shape[0]
    K, M = sparser(K, droptol), sparser(M, droptol)
    if nmod >= nX - 1:
        muscadeerror("nmod must be smaller than the number of X-dofs − 1 (Krylov solver)")
    if not complex_modes:
        # geneig{:symmetric}(K,M,nmod): the nmod eigenvalues of (K − λM)v = 0 closest to 0, shift-invert (src/Eigenmodes.jl:55-64)
        lu = spla.splu(K)
        op = spla.LinearOperator((nX, nX), matvec=lambda x: M @ lu.solve(x), dtype=float)
        val, vec = spla.eigs(op, k=nmod, which="LR", **kw)
        order = np.argsort(-val.real)
        val, vec = val[order], vec[:, order]
        lam = (1. / val).real
        Δx = [_normalize_inf(lu.solve(vec[:, i].real)) for i in range(nmod)]
        return EigXRincrement(state.dis, np.sqrt(lam), Δx)
    C = sparser(C, droptol)
    I = sp.identity(nX, format="csc")
    A = sp.bmat([[I, None], [None, K]], format="csc")               # blkasm([1,2],[1,2],[I,K])      (src/EigX.jl:115-116)
    B = sp.bmat([[None, M], [-I, C]], format="csc")                 # blkasm([2,1,2],[1,2,2],[-I,M,C])
    lu = spla.splu(A)
    op = spla.LinearOperator((2 * nX, 2 * nX), matvec=lambda x: B @ lu.solve(x), dtype=float)      # A, B are real: ARPACK's real non-symmetric driver returns the complex pairs
    val, vec = spla.eigs(op, k=nmod, which="LR", **kw)
    order = np.argsort(-val.real)
    val, vec = val[order], vec[:, order]
    p = 1. / val
    full = [_normalize_inf(lu.solve(vec[:, i].real) + 1j * lu.solve(vec[:, i].imag)) for i in range(nmod)]
    return EigXCincrement(state.dis, p, [f[nX:2 * nX] for f in full])


def increment(initialstate, eiginc, imod, A, OX=2):
    """increment{OX}(initialstate,eiginc,imod,A) (src/EigX.jl:70-84, 130-143): snapshot of the vibrating structure; imod 1-based"""
    if len(imod) != len(A):
        muscadeerror("imod and A must be of same length.")
    s = initialstate.copy().with_orders(1, OX + 1, 1)
    s.X = [x.copy() for x in s.X]
    scale = np.asarray(initialstate.dis.scaleX, float)
    real = isinstance(eiginc, EigXRincrement)
    n = len(eiginc.ω if real else eiginc.p)
    if max(imod) > n:
        muscadeerror("eiginc only has %d modes." % n)
    for i, a in zip(imod, A):
        Δx = eiginc.Δx[i - 1]
        rate = 1j * eiginc.ω[i - 1] if real else eiginc.p[i - 1]
        for d in range(OX + 1):
            s.X[d] = s.X[d] + (rate ** d * a * Δx).real * scale      # increment!(state,ider,Δ,dofgr) descales (src/Assemble.jl:206-211)
    return s
