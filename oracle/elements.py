"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes front-end of oracle/liboracle.so (built by oracle/Makefile).

Each function cites the reference code it restates (paths relative to the reference root).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the C++ restatement (g++ only; no reference sources are compiled or copied)."""
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = C.CDLL(so)
        L.orc_beam_ctor.argtypes = [f64p, f64p, f64p, f64p, f64p]
        L.orc_beam_ctor.restype = C.c_int
        L.orc_beam_residual.argtypes = [f64p, C.c_int, C.c_int, f64p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, f64p, f64p]
        L.orc_beam_residual.restype = C.c_int
        L.orc_sinc1k.argtypes = [C.c_int, C.c_double]
        L.orc_sinc1k.restype = C.c_double
        L.orc_scac_d.argtypes = [C.c_int, C.c_double]
        L.orc_scac_d.restype = C.c_double
        L.orc_rodrigues_roundtrip.argtypes = [f64p, f64p, f64p, f64p]
        L.orc_sweepx_assemble_beams.argtypes = [C.c_int64, f64p, i64p, i64p, i64p, C.c_int, C.c_int, f64p, C.c_void_p, C.c_void_p,
                                                f64p, f64p, f64p, f64p]
        L.orc_sweepx_assemble_beams.restype = C.c_int
        L.orc_beams_iter_elementwise.argtypes = [C.c_int64, f64p, i64p, C.c_int, f64p, C.c_void_p, C.c_void_p, f64p, f64p, f64p, f64p, C.c_int]
        L.orc_beams_iter_elementwise.restype = C.c_int
        L.orc_max_threads.restype = C.c_int
        L.orc_sweepx_assemble_beams_mt.argtypes = [C.c_int64, f64p, i64p, i64p, i64p, C.c_int, f64p, C.c_void_p, C.c_void_p, f64p, f64p, f64p, f64p, C.c_int]
        L.orc_sweepx_assemble_beams_mt.restype = C.c_int
        L.orc_bar_ctor.argtypes = [f64p, f64p, f64p, C.c_double, f64p]
        L.orc_bar_residual.argtypes = [f64p, C.c_int, C.c_int, f64p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, f64p, f64p]
        L.orc_bar_residual.restype = C.c_int
        L.orc_soil_residual.argtypes = [f64p, C.c_int, C.c_int, f64p, C.c_void_p, f64p, f64p]
        L.orc_soil_residual.restype = C.c_int
        L.orc_direct_addin_beams.argtypes = [C.c_int64, f64p, i64p, C.c_void_p, C.c_int, C.c_int, C.c_int, f64p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, f64p, C.c_void_p, i64p, i64p, i64p, C.c_void_p, C.c_void_p,
                                             f64p, f64p, C.c_int64, f64p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        L.orc_direct_addin_beams.restype = C.c_int
        _LIB = L
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# BeamCrossSection field order, toolbox/BeamElement.jl:6-23
MAT_FIELDS = ("EA", "EI2", "EI3", "GJ", "mu", "iota1", "w", "Ca1", "Cl1", "Cq1", "Ca2", "Cl2", "Cq2", "Ca3", "Cl3", "Cq3")


def beam_cross_section(**kw):
    """BeamCrossSection(;EA,EI₂,EI₃,GJ,μ,ι₁,w=0,...)  toolbox/BeamElement.jl:25 → 16 doubles."""
    m = np.zeros(16)
    for k, v in kw.items():
        m[MAT_FIELDS.index(k)] = v
    return m


# layout of the 69-double EulerBeam3D struct, toolbox/BeamElement.jl:87-103
BEAM_LAYOUT = dict(cm=(0, 3), rm=(3, 12), zgp=(12, 16), znod=(16, 18), tgm=(18, 21), tge=(21, 24), ya=(24, 28), yu=(28, 32),
                   yv=(32, 36), ka=(36, 40), ku=(40, 44), kv=(44, 48), L=(48, 49), dL=(49, 53), mat=(53, 69))


def beam_ctor(c1, c2, mat, orient2=(0., 1., 0.)):
    """EulerBeam3D(nod; mat, orient2)  toolbox/BeamElement.jl:121-148 → 69 doubles (rm column-major)."""
    out = np.zeros(69)
    rc = lib().orc_beam_ctor(np.ascontiguousarray(c1, float), np.ascontiguousarray(c2, float), np.ascontiguousarray(orient2, float),
                             np.ascontiguousarray(mat, float), out)
    if rc:
        raise ValueError("Provide a 'orient' input that is not nearly parallel to the element")
    return out


def beam_field(e, name):
    a, b = BEAM_LAYOUT[name]
    v = e[..., a:b]
    if name == "rm":
        return v.reshape(v.shape[:-1] + (3, 3)).swapaxes(-1, -2)  # column-major → [i,j]
    return v


def beam_residual(elem, X, Xseed=None, U=None, Useed=None):
    """Muscade.residual(o::EulerBeam3D,X,U,A,t,SP,dbg)  toolbox/BeamElement.jl:151-174 with first-order seeds.

    X: (nd,12) values; Xseed: (nd,12,np) partials; U: (3,) or None; Useed (3,np).  Returns R (12,), dR (12,np)."""
    X = np.ascontiguousarray(np.atleast_2d(X), float)
    nd = X.shape[0]
    if Xseed is None:
        Xseed = np.zeros((nd, 12, 0))
    Xseed = np.ascontiguousarray(Xseed, float)
    np_ = Xseed.shape[2]
    udof = U is not None
    if udof:
        U = np.ascontiguousarray(U, float)
        Useed = np.zeros((3, np_)) if Useed is None else np.ascontiguousarray(Useed, float)
    R = np.zeros(12)
    dR = np.zeros((12, max(np_, 1)))
    rc = lib().orc_beam_residual(np.ascontiguousarray(elem, float), nd, np_, X, _ptr(Xseed) if np_ else None, int(udof),
                                 _ptr(U), _ptr(Useed), R, dR)
    if rc < 0:
        raise ValueError("bad arguments")
    return R, dR[:, :np_], rc


def diffed_residual(elem, X, U=None):
    """Muscade.diffed_residual(ele;X,U,A)  src/Diagnostic.jl:929-961: revariate{1}((;X,U,A)) unit seeds, flat ordering
    X₀ X₁ X₂ U₀ ; returns R and the list ∇R[X][ider] (12×12 each) (+ ∇R[U][0] 12×3)."""
    X = np.atleast_2d(np.asarray(X, float))
    nd = X.shape[0]
    nu = 0 if U is None else 3
    np_ = 12 * nd + nu
    Xseed = np.zeros((nd, 12, np_))
    for d in range(nd):
        for i in range(12):
            Xseed[d, i, 12 * d + i] = 1.
    Useed = None
    if U is not None:
        Useed = np.zeros((3, np_))
        for i in range(3):
            Useed[i, 12 * nd + i] = 1.
    R, dR, rc = beam_residual(elem, X, Xseed, U, Useed)
    gX = [dR[:, 12 * d:12 * d + 12] for d in range(nd)]
    gU = dR[:, 12 * nd:] if U is not None else None
    return R, gX, gU


def sinc1k(k, x):
    """sinc1, sinc1′, … sinc1⁗  toolbox/Rotations.jl:13-44"""
    return lib().orc_sinc1k(k, float(x))


def scac_d(n, x):
    """n-th derivative of scac by nested duals  toolbox/Rotations.jl:61-68, test/TestRotations.jl:35-60"""
    return lib().orc_scac_d(n, float(x))


def rodrigues_roundtrip(v):
    """Rodrigues, Rodrigues⁻¹ and ∂w/∂v  toolbox/Rotations.jl:105,131-135"""
    M = np.zeros(9); w = np.zeros(3); dw = np.zeros(9)
    lib().orc_rodrigues_roundtrip(np.ascontiguousarray(v, float), M, w, dw)
    return M.reshape(3, 3).T, w, dw.reshape(3, 3)


def sweepx_assemble_beams(elems, idx, asm1, asm2, OX, mission, X, scaleX, newmark, Llambda, nzval):
    """assemble_!{mission} + addin!{mission}(::AssemblySweepX{OX}) over one EulerBeam3D element type
    (src/Assemble.jl:478-487, src/SweepX.jl:45-96). idx/asm1/asm2 are (nele,12|12|144) int64, 1-based; accumulates."""
    nele = elems.shape[0]
    X = [np.ascontiguousarray(x, float) for x in X]
    rc = lib().orc_sweepx_assemble_beams(nele, np.ascontiguousarray(elems), np.ascontiguousarray(idx, np.int64),
                                         np.ascontiguousarray(asm1, np.int64), np.ascontiguousarray(asm2, np.int64), OX,
                                         {"step": 0, "iter": 1}[mission], X[0], _ptr(X[1]) if OX >= 1 else None,
                                         _ptr(X[2]) if OX >= 2 else None, np.ascontiguousarray(scaleX, float),
                                         np.ascontiguousarray(newmark, float), Llambda, nzval)
    if rc:
        raise FloatingPointError("residual(EulerBeam3D,...) returned NaN in R, FB or derivatives (iele=%d)" % rc)


_LIB12 = None


def lib_np12():
    """liboracle_np12.so: the same sources built with static-size duals (12 partials, the SweepX :iter seeding) and -O3 — the CPU baseline build."""
    global _LIB12
    if _LIB12 is None:
        so = os.path.join(_HERE, "liboracle_np12.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle_np12.so"])
        L = C.CDLL(so)
        L.orc_sweepx_assemble_beams.argtypes = lib().orc_sweepx_assemble_beams.argtypes
        L.orc_sweepx_assemble_beams.restype = C.c_int
        L.orc_sweepx_assemble_beams_mt.argtypes = lib().orc_sweepx_assemble_beams_mt.argtypes
        L.orc_sweepx_assemble_beams_mt.restype = C.c_int
        L.orc_max_threads.restype = C.c_int
        _LIB12 = L
    return _LIB12


def sweepx_assemble_beams_mt(elems, idx, asm1, asm2, OX, X, scaleX, newmark, Llambda, nzval, nthreads, static_duals=False):
    """:iter assembly with the element loop spread over host threads (bench.py --impl reference); static_duals: the -O3 build with 12 compile-time partials"""
    X = [np.ascontiguousarray(x, float) for x in X]
    rc = (lib_np12() if static_duals else lib()).orc_sweepx_assemble_beams_mt(elems.shape[0], np.ascontiguousarray(elems), np.ascontiguousarray(idx, np.int64),
                                            np.ascontiguousarray(asm1, np.int64), np.ascontiguousarray(asm2, np.int64), OX, X[0],
                                            _ptr(X[1]) if OX >= 1 else None, _ptr(X[2]) if OX >= 2 else None,
                                            np.ascontiguousarray(scaleX, float), np.ascontiguousarray(newmark, float), Llambda, nzval, nthreads)
    if rc:
        raise FloatingPointError("NaN in %d elements" % rc)


def max_threads():
    return lib().orc_max_threads()


def beams_iter_elementwise(elems, idx, OX, X, scaleX, newmark, nthreads=1):
    nele = elems.shape[0]
    Re = np.zeros((nele, 12)); Ke = np.zeros((nele, 144))
    X = [np.ascontiguousarray(x, float) for x in X]
    bad = lib().orc_beams_iter_elementwise(nele, np.ascontiguousarray(elems), np.ascontiguousarray(idx, np.int64), OX, X[0],
                                           _ptr(X[1]) if OX >= 1 else None, _ptr(X[2]) if OX >= 2 else None,
                                           np.ascontiguousarray(scaleX, float), np.ascontiguousarray(newmark, float), Re, Ke, nthreads)
    return Re, Ke, bad


def newmark_coefficients(OX, dt, beta=0.25, gamma=0.5):
    """Newmarkβcoefficients{OX}(Δt,β,γ)  src/SweepX.jl:12-15 → (a1,a2,a3,b1,b2,b3,Δt)"""
    if OX == 0:
        return np.array([0., 0., 0., 0., 0., 0., dt])
    if OX == 1:
        return np.array([1 / (gamma * dt), 1 / gamma, 0., 0., 0., 0., dt])
    return np.array([gamma / (beta * dt), gamma / beta, (gamma / (2 * beta) - 1) * dt, 1 / (beta * dt ** 2), 1 / (beta * dt), 1 / (2 * beta), dt])


# ------------------------------------------------------------------------------------------------ Bar3D / SoilContact
BAR_MAT_FIELDS = ("EA", "mu", "w", "Cat", "Clt", "Cqt", "Can", "Cln", "Cqn")


def bar_cross_section(**kw):
    """AxisymmetricBarCrossSection  toolbox/BarElement.jl:36-48"""
    m = np.zeros(9)
    for k, v in kw.items():
        m[BAR_MAT_FIELDS.index(k)] = v
    return m


def bar_ctor(c1, c2, mat, eps_s=np.finfo(float).eps):
    """Bar3D(nod;mat,ϵₛ)  toolbox/BarElement.jl:117-133 → 38 doubles"""
    out = np.zeros(38)
    lib().orc_bar_ctor(np.ascontiguousarray(c1, float), np.ascontiguousarray(c2, float), np.ascontiguousarray(mat, float), float(eps_s), out)
    return out


def bar_residual(elem, X, Xseed=None, U=None, Useed=None, t=0.):
    """residual(o::Bar3D,…)  toolbox/BarElement.jl:167-202 ; X (nd,6), Xseed (nd,6,np)"""
    X = np.ascontiguousarray(np.atleast_2d(X), float); nd = X.shape[0]
    Xseed = np.zeros((nd, 6, 0)) if Xseed is None else np.ascontiguousarray(Xseed, float)
    np_ = Xseed.shape[2]
    udof = U is not None
    if udof:
        U = np.ascontiguousarray(U, float); Useed = np.zeros((3, np_)) if Useed is None else np.ascontiguousarray(Useed, float)
    R = np.zeros(6); dR = np.zeros((6, max(np_, 1)))
    rc = lib().orc_bar_residual(np.ascontiguousarray(elem, float), nd, np_, X, _ptr(Xseed) if np_ else None, int(udof), _ptr(U), _ptr(Useed), float(t), R, dR)
    return R, dR[:, :np_], rc


def soil_residual(elem5, X, Xseed=None):
    """residual(o::SoilContact,…)  toolbox/SoilContact.jl:10-20 ; X (nd,3)"""
    X = np.ascontiguousarray(np.atleast_2d(X), float); nd = X.shape[0]
    Xseed = np.zeros((nd, 3, 0)) if Xseed is None else np.ascontiguousarray(Xseed, float)
    np_ = Xseed.shape[2]
    R = np.zeros(3); dR = np.zeros((3, max(np_, 1)))
    c = lib().orc_soil_residual(np.ascontiguousarray(elem5, float), nd, np_, X, _ptr(Xseed) if np_ else None, R, dR)
    return R, dR[:, :np_], c


def sweepx_addin_generic(residual_fn, nx, idx, asm1, asm2, OX, mission, X, scaleX, newmark, Llambda, nzval):
    """addin!{mission}(out::AssemblySweepX{OX},…) (src/SweepX.jl:45-96) for any element given `residual_fn(iele, Xval (nd,nx), Xseed (nd,nx,np))`
    → (R, dR); python loop over elements — small cases only.  idx/asm1 (nele,nx), asm2 (nele,nx²), 1-based."""
    a1, a2, a3, b1, b2, b3 = newmark[:6]
    step = mission == "step" and OX > 0
    np_ = nx + (1 if step else 0)
    nd = OX + 1
    for e in range(idx.shape[0]):
        xv = np.stack([X[d][idx[e] - 1] for d in range(nd)])
        seed = np.zeros((nd, nx, np_))
        for i in range(nx):
            seed[0, i, i] = scaleX[i]
            if OX >= 1: seed[1, i, i] = a1 * scaleX[i]
            if OX >= 2: seed[2, i, i] = b1 * scaleX[i]
        if step:
            seed[1, :, nx] = a2 * xv[1] + (a3 * xv[2] if OX >= 2 else 0.)
            if OX >= 2: seed[2, :, nx] = b2 * xv[1] + b3 * xv[2]
        R, dR = residual_fn(e, xv, seed)
        R = R * scaleX; dR = dR * scaleX[:, None]
        for i in range(nx):
            if asm1[e, i]:
                Llambda[asm1[e, i] - 1] += R[i]
        if step:
            for i in range(nx):
                if asm1[e, i]:
                    Llambda[asm1[e, i] - 1] -= dR[i, nx]
        for i in range(nx):
            for j in range(nx):
                k = asm2[e, i + nx * j]
                if k:
                    nzval[k - 1] += dR[i, j]


# ------------------------------------------------------------------------------------------------ DirectXUA
def direct_assemble_step_beams(elems, idxX, idxU, OX, OU, X, U, scaleX, scaleU, P, ityp, out=None):
    """assemble!{:matrices}(out::AssemblyDirect{OX,OU,0},…) for one state (one time step), EulerBeam3D type `ityp` (0-based) of the
    model whose prepare_direct() result is P (src/DirectXUA.jl:85-120, src/Assemble.jl:470-487).
    Returns dict with L1[1] (Λ) and L2[(1,2)],L2[(2,1)] (lists over X derivative), L2[(1,3)],L2[(3,1)] (lists over U derivative)."""
    from .pattern import arrnum
    nd, ndu = OX + 1, OU + 1
    udof = idxU is not None and idxU.shape[1] == 3
    asm = P["asm"]
    T = lambda a: np.ascontiguousarray(a.T, dtype=np.int64)
    nnz = {ab: len(P["pat"][ab][3]) for ab in P["pat"]}
    if out is None:
        out = dict(L1={1: np.zeros(P["ndof"][0])}, L2={(1, 2): np.zeros((nd, nnz[(1, 2)])), (2, 1): np.zeros((nd, nnz[(2, 1)])),
                                                      (1, 3): np.zeros((ndu, nnz[(1, 3)])), (3, 1): np.zeros((ndu, nnz[(3, 1)]))})
    X = [np.ascontiguousarray(x, float) for x in X]; U = [np.ascontiguousarray(u, float) for u in U]
    aLU = T(asm[arrnum(1, 3)][ityp]) if udof else None; aUL = T(asm[arrnum(3, 1)][ityp]) if udof else None
    rc = lib().orc_direct_addin_beams(
        elems.shape[0], np.ascontiguousarray(elems), np.ascontiguousarray(idxX, np.int64), _ptr(np.ascontiguousarray(idxU, np.int64)) if udof else None,
        int(udof), OX, OU, X[0], _ptr(X[1]) if OX >= 1 else None, _ptr(X[2]) if OX >= 2 else None,
        _ptr(U[0]) if udof else None, _ptr(U[1]) if (udof and OU >= 1) else None, _ptr(U[2]) if (udof and OU >= 2) else None,
        np.ascontiguousarray(scaleX, float), _ptr(np.ascontiguousarray(scaleU, float)) if udof else None,
        T(asm[arrnum(1)][ityp]), T(asm[arrnum(1, 2)][ityp]), T(asm[arrnum(2, 1)][ityp]), _ptr(aLU), _ptr(aUL),
        out["L1"][1], out["L2"][(1, 2)], nnz[(1, 2)], out["L2"][(2, 1)], nnz[(2, 1)],
        _ptr(out["L2"][(1, 3)]), nnz[(1, 3)], _ptr(out["L2"][(3, 1)]), nnz[(3, 1)])
    if rc:
        raise FloatingPointError("NaN or bad arguments (%d)" % rc)
    return out


def direct_out_zeros(P, OX, OU):
    """empty out.L1 / out.L2 of prepare(AssemblyDirect{OX,OU,0}) for the blocks the first-order path fills (DirectXUA.jl:85-120)"""
    nd, ndu = OX + 1, OU + 1
    nnz = {ab: len(P["pat"][ab][3]) for ab in P["pat"]}
    return dict(L1={1: np.zeros(P["ndof"][0])}, L2={(1, 2): np.zeros((nd, nnz[(1, 2)])), (2, 1): np.zeros((nd, nnz[(2, 1)])),
                                                   (1, 3): np.zeros((ndu, nnz[(1, 3)])), (3, 1): np.zeros((ndu, nnz[(3, 1)]))})


def direct_addin_generic(residual_fn, nx, nu, idxX, idxU, OX, OU, X, U, scaleX, scaleU, P, ityp, out):
    """addin!{:matrices}(out::AssemblyDirect{OX,OU,0},…) first-order path (src/DirectXUA.jl:85-120) for any element type given
    `residual_fn(iele, Xval (nd,nx), Xseed (nd,nx,np), Uval (nu,), Useed (nu,np))` → (R, dR (nx,np)); seeds = revariate{1}((;X,U),scale)
    (Taylor.jl:158-166).  Adds into `out` (direct_out_zeros); python loop over elements — small cases only."""
    from .pattern import arrnum
    nd = OX + 1
    np_ = nx * nd + nu
    asm = P["asm"]
    aL = asm[arrnum(1)][ityp]; aLX = asm[arrnum(1, 2)][ityp]; aXL = asm[arrnum(2, 1)][ityp]          # (n, nele), 1-based, 0 = absent
    aLU = asm[arrnum(1, 3)][ityp] if nu else None; aUL = asm[arrnum(3, 1)][ityp] if nu else None
    for e in range(idxX.shape[0]):
        xv = np.stack([X[d][idxX[e] - 1] for d in range(nd)])
        seed = np.zeros((nd, nx, np_))
        for d in range(nd):
            for i in range(nx):
                seed[d, i, nx * d + i] = scaleX[i]
        uv = U[0][idxU[e] - 1] if nu else np.zeros(0)
        useed = np.zeros((nu, np_))
        for i in range(nu):
            useed[i, nx * nd + i] = scaleU[i]
        R, dR = residual_fn(e, xv, seed, uv, useed)
        for i in range(nx):
            if aL[i, e]: out["L1"][1][aL[i, e] - 1] += R[i]
        for d in range(nd):
            for i in range(nx):
                for j in range(nx):
                    v = dR[i, nx * d + j]
                    k = aLX[i + nx * j, e]
                    if k: out["L2"][(1, 2)][d, k - 1] += v
                    k = aXL[j + nx * i, e]
                    if k: out["L2"][(2, 1)][d, k - 1] += v
        for i in range(nx):
            for j in range(nu):
                v = dR[i, nx * nd + j]
                k = aLU[i + nx * j, e]
                if k: out["L2"][(1, 3)][0, k - 1] += v
                k = aUL[j + nu * i, e]
                if k: out["L2"][(3, 1)][0, k - 1] += v
    return out


def beam_results(elem, X, U=None):
    """getresult for one EulerBeam3D (src/Output.jl:131-181): dict of the @espy requestables (BeamElement.jl:151-174, :28-64), values only.
    X (nd,12)."""
    X = np.ascontiguousarray(np.atleast_2d(X), float)
    out = np.zeros(77)
    f = lib().orc_beam_results
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    U = None if U is None else np.ascontiguousarray(U, float)
    rc = f(_ptr(np.ascontiguousarray(elem, float)), X.shape[0], _ptr(X), int(U is not None), _ptr(U), _ptr(out))
    if rc:
        raise ValueError("orc_beam_results: %d" % rc)
    return out


def direct_addin_soil_second_order(soil, idxX, OX, X, Lam, scaleX, lam_scale, P, ityp, out):
    """addin!{:matrices}(out::AssemblyDirect{OX,OU,0},…,no_second_order::Val{false},…) (src/DirectXUA.jl:152-171 → :121-150) for SoilContact, which does
    not declare `no_second_order`: L = Λ∘R(X) (Assemble.jl:721-726) differentiated to second order over (Λ, X₀, X′, X″) seeded by
    revariate{2} with scales (Λ=scale.Λ, X=scale.X), scale.Λ = scale.X·Λscale (Assemble.jl:55).  R (SoilContact.jl:10-20) is linear in (x,x′) inside
    the contact branch and identically zero outside, so the partials are written out instead of carried by nested duals:
        ∂L/∂Λᵢ = Rᵢ·sΛᵢ            ∂L/∂X₀ᵢ = Λᵢ·Kᵢ·sXᵢ       ∂L/∂X′ᵢ = Λᵢ·Cᵢ·sXᵢ      ∂L/∂X″ᵢ = 0
        ∂²L/∂Λᵢ∂X₀ᵢ = Kᵢ·sΛᵢ·sXᵢ     ∂²L/∂Λᵢ∂X′ᵢ = Cᵢ·sΛᵢ·sXᵢ   all other second partials 0
    scattered by DirectXUA_lagrangian_addition! into L1[Λ][1], L1[X][der], L2[Λ,X][1,der], L2[X,Λ][der,1] (the zero blocks add nothing).
    `out` as direct_out_zeros, plus out["L1"][2] of shape (OX+1, nX) created on demand."""
    from .pattern import arrnum
    nd = OX + 1
    nX = P["ndof"][0]
    if 2 not in out["L1"]:
        out["L1"][2] = np.zeros((nd, nX))
    asm = P["asm"]
    aL = asm[arrnum(1)][ityp]; aX = asm[arrnum(2)][ityp]; aLX = asm[arrnum(1, 2)][ityp]; aXL = asm[arrnum(2, 1)][ityp]
    for e in range(idxX.shape[0]):
        z0, Kh, Kv, Ch, Cv = soil[e]
        ix = idxX[e] - 1
        x = X[0][ix]; xp = X[1][ix] if OX >= 1 else np.zeros(3)
        lam = Lam[ix]
        contact = x[2] < z0
        K = np.array([Kh, Kh, Kv]); C = np.array([Ch, Ch, Cv])
        R = (K * (x - np.array([0., 0., z0])) + C * xp) if contact else np.zeros(3)
        sX = scaleX; sL = scaleX * lam_scale
        coef = [K if contact else np.zeros(3), C if contact else np.zeros(3), np.zeros(3)]
        for i in range(3):
            if aL[i, e]: out["L1"][1][aL[i, e] - 1] += R[i] * sL[i]
            for d in range(nd):
                if aX[i, e]: out["L1"][2][d, aX[i, e] - 1] += lam[i] * coef[d][i] * sX[i]
                v = coef[d][i] * sL[i] * sX[i]
                k = aLX[i + 3 * i, e]
                if k: out["L2"][(1, 2)][d, k - 1] += v
                k = aXL[i + 3 * i, e]
                if k: out["L2"][(2, 1)][d, k - 1] += v
    return out
