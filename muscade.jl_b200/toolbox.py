"""Host-side mirror of the reference toolbox element *constructors* (setup time, numpy-vectorised over elements).

The element physics (`residual`) lives in the CUDA kernels (csrc/beam_math.cuh …); here are only the as-meshed data the
kernels consume, laid out exactly like the reference's isbits structs so the same arrays could come from Julia.
"""
import numpy as np

# ------------------------------------------------------------------------------------------------ EulerBeam3D
MAT_FIELDS = ("EA", "EI2", "EI3", "GJ", "mu", "iota1", "w", "Ca1", "Cl1", "Cq1", "Ca2", "Cl2", "Cq2", "Ca3", "Cl3", "Cq3")
_MAT_ALIASES = {"EI₂": "EI2", "EI₃": "EI3", "μ": "mu", "ι₁": "iota1", "Ca₁": "Ca1", "Cl₁": "Cl1", "Cq₁": "Cq1", "Ca₂": "Ca2",
                "Cl₂": "Cl2", "Cq₂": "Cq2", "Ca₃": "Ca3", "Cl₃": "Cl3", "Cq₃": "Cq3"}


def BeamCrossSection(**kw):
    """BeamCrossSection(;EA,EI₂,EI₃,GJ,μ,ι₁,w=0.,Ca₁=0.,…)  toolbox/BeamElement.jl:6-25 → 16 Float64 in field order."""
    m = np.zeros(16)
    for k, v in kw.items():
        m[MAT_FIELDS.index(_MAT_ALIASES.get(k, k))] = v
    for req in ("EA", "EI2", "EI3", "GJ", "mu", "iota1"):
        if _MAT_ALIASES.get(req, req) not in [_MAT_ALIASES.get(k, k) for k in kw]:
            raise TypeError("BeamCrossSection: keyword argument %s not assigned" % req)
    return m


BEAM_STRUCT_LEN = 69  # cₘ3 rₘ9 ζgp4 ζnod2 tgₘ3 tgₑ3 yₐ4 yᵤ4 yᵥ4 κₐ4 κᵤ4 κᵥ4 L dL4 mat16   (toolbox/BeamElement.jl:87-103)


def eulerbeam3d_structs(c1, c2, mat, orient2=(0., 1., 0.)):
    """EulerBeam3D{Udof}(nod;mat,orient2) for many elements at once  (toolbox/BeamElement.jl:121-148).

    c1, c2: (nele,3) node coordinates.  Returns (nele,69) Float64, the memory image of Vector{EulerBeam3D{BeamCrossSection,·}}."""
    c1 = np.atleast_2d(np.asarray(c1, float)); c2 = np.atleast_2d(np.asarray(c2, float))
    n = c1.shape[0]
    out = np.zeros((n, BEAM_STRUCT_LEN))
    cm = (c1 + c2) / 2
    tgm = c2 - c1
    L = np.sqrt((tgm[:, 0] * tgm[:, 0] + tgm[:, 1] * tgm[:, 1]) + tgm[:, 2] * tgm[:, 2])
    t = tgm / L[:, None]
    o2 = np.asarray(orient2, float)
    o2 = o2 / np.sqrt((o2[0] * o2[0] + o2[1] * o2[1]) + o2[2] * o2[2])
    d = (o2[0] * t[:, 0] + o2[1] * t[:, 1]) + o2[2] * t[:, 2]
    nn_ = o2[None, :] - t * d[:, None]
    nn = np.sqrt((nn_[:, 0] * nn_[:, 0] + nn_[:, 1] * nn_[:, 1]) + nn_[:, 2] * nn_[:, 2])
    if not np.all(nn > 1e-3):
        raise ValueError("Provide a 'orient' input that is not nearly parallel to the element")
    nv = nn_ / nn[:, None]
    b = np.stack([t[:, 1] * nv[:, 2] - t[:, 2] * nv[:, 1], t[:, 2] * nv[:, 0] - t[:, 0] * nv[:, 2], t[:, 0] * nv[:, 1] - t[:, 1] * nv[:, 0]], axis=1)
    out[:, 0:3] = cm
    out[:, 3:6] = t; out[:, 6:9] = nv; out[:, 9:12] = b          # SMatrix(t...,n...,b...): column-major
    s65 = np.sqrt(6. / 5); s30 = np.sqrt(30.)
    zgp = np.array([-1. / 2 * np.sqrt(3. / 7 + 2. / 7 * s65), -1. / 2 * np.sqrt(3. / 7 - 2. / 7 * s65),
                    +1. / 2 * np.sqrt(3. / 7 - 2. / 7 * s65), +1. / 2 * np.sqrt(3. / 7 + 2. / 7 * s65)])
    out[:, 12:16] = zgp
    out[:, 16:18] = (-0.5, 0.5)
    out[:, 18:21] = tgm
    out[:, 21] = L
    out[:, 24:28] = 2 * zgp
    out[:, 28:32] = -4 * (zgp * zgp * zgp) + 3 * zgp
    out[:, 32:36] = ((zgp * zgp) - 1. / 4)[None, :] * L[:, None]
    out[:, 36:40] = (2. / L)[:, None]
    out[:, 40:44] = (-24 * zgp)[None, :] / (L * L)[:, None]
    out[:, 44:48] = (2. / L)[:, None]
    out[:, 48] = L
    out[:, 49] = L / 2 * (18 - s30) / 36; out[:, 50] = L / 2 * (18 + s30) / 36
    out[:, 51] = L / 2 * (18 + s30) / 36; out[:, 52] = L / 2 * (18 - s30) / 36
    out[:, 53:69] = np.asarray(mat, float)
    return out


class ElementType:
    """Base of the Python stand-ins for `E<:AbstractElement` (src/ModelDescription.jl:14).  Subclasses give `doflist()`
    (src/ElementAPI.jl:99), a vectorised constructor `construct(coords, **kw)` and a `typekey` that plays the role of
    the concrete Julia type (one element *type* per distinct key, src/ModelDescription.jl:205-213)."""
    kind = "host"

    @classmethod
    def doflist(cls, **kw):
        raise NotImplementedError("method 'doflist' must be provided for elements of type %s" % cls.__name__)


class EulerBeam3D(ElementType):
    kind = "eulerbeam3d"

    @classmethod
    def doflist(cls, Udof=False, **kw):
        inod = (1,) * 6 + (2,) * 6
        clas = ("X",) * 12
        field = ("t1", "t2", "t3", "r1", "r2", "r3") * 2
        if Udof:
            inod += (3, 3, 3); clas += ("U",) * 3; field += ("t1", "t2", "t3")
        return inod, clas, field

    @classmethod
    def typekey(cls, Udof=False, **kw):
        return ("EulerBeam3D", "BeamCrossSection", bool(Udof))

    @classmethod
    def construct(cls, coords, mat, orient2=(0., 1., 0.), Udof=False):
        return eulerbeam3d_structs(coords[0], coords[1], mat, orient2)


# ------------------------------------------------------------------------------------------------ Bar3D, SoilContact
BAR_MAT_FIELDS = ("EA", "mu", "w", "Cat", "Clt", "Cqt", "Can", "Cln", "Cqn")
_BAR_ALIASES = {"μ": "mu", "Caₜ": "Cat", "Clₜ": "Clt", "Cqₜ": "Cqt", "Caₙ": "Can", "Clₙ": "Cln", "Cqₙ": "Cqn"}
BAR_STRUCT_LEN = 38  # cₘ3 tgₘ3 tgₑ3 L₀ Lₛ mat9 wgp4 ζgp4 ζnod2 ψ₁4 ψ₂4   (toolbox/BarElement.jl:89-101)


def AxisymmetricBarCrossSection(**kw):
    """AxisymmetricBarCrossSection(;EA,μ,w=0.,Caₜ=0.,…)  toolbox/BarElement.jl:36-48 → 9 Float64"""
    m = np.zeros(9)
    for k, v in kw.items():
        m[BAR_MAT_FIELDS.index(_BAR_ALIASES.get(k, k))] = v
    return m


def bar3d_structs(c1, c2, mat, eps_s=np.finfo(float).eps):
    """Bar3D{Udof}(nod;mat,ϵₛ=eps())  toolbox/BarElement.jl:117-133 → (nele,38)"""
    c1 = np.atleast_2d(np.asarray(c1, float)); c2 = np.atleast_2d(np.asarray(c2, float))
    n = c1.shape[0]
    out = np.zeros((n, BAR_STRUCT_LEN))
    tgm = c2 - c1
    L0 = np.sqrt((tgm[:, 0] * tgm[:, 0] + tgm[:, 1] * tgm[:, 1]) + tgm[:, 2] * tgm[:, 2])
    out[:, 0:3] = (c1 + c2) / 2; out[:, 3:6] = tgm; out[:, 6] = L0; out[:, 9] = L0; out[:, 10] = (1 - eps_s) * L0
    out[:, 11:20] = np.asarray(mat, float)
    s65 = np.sqrt(6. / 5); s30 = np.sqrt(30.)
    zgp = np.array([-1. / 2 * np.sqrt(3. / 7 + 2. / 7 * s65), -1. / 2 * np.sqrt(3. / 7 - 2. / 7 * s65),
                    +1. / 2 * np.sqrt(3. / 7 - 2. / 7 * s65), +1. / 2 * np.sqrt(3. / 7 + 2. / 7 * s65)])
    out[:, 20] = L0 / 2 * (18 - s30) / 36; out[:, 21] = L0 / 2 * (18 + s30) / 36; out[:, 22] = L0 / 2 * (18 + s30) / 36; out[:, 23] = L0 / 2 * (18 - s30) / 36
    out[:, 24:28] = zgp; out[:, 28:30] = (-0.5, 0.5); out[:, 30:34] = -zgp + 1. / 2; out[:, 34:38] = zgp + 1. / 2
    return out


class Bar3D(ElementType):
    kind = "bar3d"

    @classmethod
    def doflist(cls, Udof=False, **kw):
        inod = (1, 1, 1, 2, 2, 2); clas = ("X",) * 6; field = ("t1", "t2", "t3") * 2
        if Udof:
            inod += (3, 3, 3); clas += ("U",) * 3; field += ("t1", "t2", "t3")
        return inod, clas, field

    @classmethod
    def typekey(cls, Udof=False, **kw):
        return ("Bar3D", "AxisymmetricBarCrossSection", bool(Udof))

    @classmethod
    def construct(cls, coords, mat, ϵₛ=np.finfo(float).eps, Udof=False):
        return bar3d_structs(coords[0], coords[1], mat, ϵₛ)


class SoilContact(ElementType):
    """SoilContact(nod;z₀=0.,Kh=0.,Kv=0.,Ch=0.,Cv=0.)  toolbox/SoilContact.jl:2-9 → (nele,5)"""
    kind = "soilcontact"

    @classmethod
    def doflist(cls, **kw):
        return (1, 1, 1), ("X", "X", "X"), ("t1", "t2", "t3")

    @classmethod
    def typekey(cls, **kw):
        return ("SoilContact",)

    @classmethod
    def construct(cls, coords, z0=0., Kh=0., Kv=0., Ch=0., Cv=0.):
        return np.tile(np.array([z0, Kh, Kv, Ch, Cv], float), (coords[0].shape[0], 1))


# ------------------------------------------------------------------------------------------------ boundary elements (host evaluated)
class Hold(ElementType):
    """Hold(nod;field,λfield=Symbol(:λ,field)) = DofConstraint{:X,1,…} with gap(v,t)=v[1], mode=equal (src/BasicElements.jl:484-488).
    residual (:425-438): R = (−λ, −x), ∂R/∂X = [[0,−1],[−1,0]]."""

    @classmethod
    def doflist(cls, field, λfield=None, **kw):
        return (1, 1), ("X", "X"), (field, λfield or "λ" + field)

    @classmethod
    def typekey(cls, field, λfield=None, **kw):
        return ("DofConstraint", "X", 1, 0, 0, (1,), (field,), 1, λfield or "λ" + field, "Hold.gap", "equal")

    @classmethod
    def construct(cls, coords, **kw):
        return np.zeros((coords[0].shape[0], 0))

    @staticmethod
    def residual(eleobj, X, t):
        """X: (nder, nele, 2) element dofs. Returns R (nele,2), dR/dX0 (nele,2,2), dR/dX1, dR/dX2 (None = zero)."""
        x = X[0]
        R = np.stack([-x[:, 1], -x[:, 0]], axis=1)
        K = np.zeros((x.shape[0], 2, 2)); K[:, 0, 1] = -1.; K[:, 1, 0] = -1.
        return R, K, None, None


class DofConstraint(ElementType):
    """DofConstraint(nod;xinod,xfield,λinod,λclass=:X,λfield,gap,gargs,mode)  (src/BasicElements.jl:406-438), λclass = :X only.
    residual (:425-438):  R = (−∂g/∂x·λ, −g) [equal],  (−∂g/∂x·λ, −(g·λ−γ)) [positive, γ = 0],  (0, −λ) [off].
    The reference differentiates `gap` with its own dual numbers; here `gap(x,t,*gargs)` returns (g, ∂g/∂x) or (g, ∂g/∂x, ∂²g/∂x²) for
    x of shape (nele,Nx) — arrays (nele,), (nele,Nx), (nele,Nx,Nx); evaluated on the host like Hold and DofLoad.  mode: "equal" | "positive" |
    "off" or a callable t → one of them."""

    @classmethod
    def doflist(cls, xinod=(), xfield=(), λinod=1, λclass="X", λfield=None, **kw):
        if λclass != "X":
            raise ValueError("DofConstraint: only λclass = :X is on this path")
        return tuple(xinod) + (λinod,), ("X",) * (len(xinod) + 1), tuple(xfield) + (λfield,)

    @classmethod
    def typekey(cls, xinod=(), xfield=(), λinod=1, λclass="X", λfield=None, gap=None, gargs=(), mode="equal", **kw):
        return ("DofConstraint", λclass, len(xinod), tuple(xinod), tuple(xfield), λinod, λfield, id(gap), id(mode) if callable(mode) else mode)

    @classmethod
    def construct(cls, coords, xinod=(), xfield=(), λinod=1, λclass="X", λfield=None, gap=None, gargs=(), mode="equal"):
        return np.zeros((coords[0].shape[0], 0)), dict(gap=gap, gargs=gargs, mode=mode, Nx=len(xinod))

    @staticmethod
    def residual(extra, X, t):
        Nx = extra["Nx"]
        x, lam = X[0][:, :Nx], X[0][:, Nx]
        n = x.shape[0]
        m = extra["mode"](t) if callable(extra["mode"]) else extra["mode"]
        res = extra["gap"](x, t, *extra["gargs"])
        g, dg = np.broadcast_to(np.asarray(res[0], float), (n,)), np.broadcast_to(np.asarray(res[1], float), (n, Nx))
        d2g = np.broadcast_to(np.asarray(res[2], float), (n, Nx, Nx)) if len(res) > 2 else np.zeros((n, Nx, Nx))
        R = np.zeros((n, Nx + 1)); K = np.zeros((n, Nx + 1, Nx + 1))
        if m == "equal":
            R[:, :Nx] = -dg * lam[:, None]; R[:, Nx] = -g
            K[:, :Nx, :Nx] = -d2g * lam[:, None, None]; K[:, :Nx, Nx] = -dg; K[:, Nx, :Nx] = -dg
        elif m == "positive":                                  # S(λ,g,γ) = g·λ − γ, γ = 0 (SP.γ is a solver parameter of other solvers)
            R[:, :Nx] = -dg * lam[:, None]; R[:, Nx] = -g * lam
            K[:, :Nx, :Nx] = -d2g * lam[:, None, None]; K[:, :Nx, Nx] = -dg; K[:, Nx, :Nx] = -dg * lam[:, None]; K[:, Nx, Nx] = -g
        elif m == "off":
            R[:, Nx] = -lam; K[:, Nx, Nx] = -1.
        else:
            raise ValueError("DofConstraint: mode must be equal, positive or off")
        return R, K, None, None

    @staticmethod
    def hessian(extra, X, lam_e, t):
        """H[e,i,j] = Σₖ Λₖ·∂²Rₖ/∂Xᵢ∂Xⱼ for the second-order branch of DirectXUA (L = Λ∘R, src/DirectXUA.jl:152-171; src/Assemble.jl:721-726), element dofs
        (x₁…x_Nx, λ), lam_e (nele,Nx+1) = Λ at the element dofs.  Exact for gaps up to second degree (∂³g/∂x³ is not asked from the closure); None when R is
        linear in (x,λ)."""
        Nx = extra["Nx"]
        x, lam = X[0][:, :Nx], X[0][:, Nx]
        n = x.shape[0]
        m = extra["mode"](t) if callable(extra["mode"]) else extra["mode"]
        res = extra["gap"](x, t, *extra["gargs"])
        curved = len(res) > 2 and np.any(np.asarray(res[2]) != 0)
        if m == "off" or (m == "equal" and not curved):
            return None
        dg = np.broadcast_to(np.asarray(res[1], float), (n, Nx))
        d2g = np.broadcast_to(np.asarray(res[2], float), (n, Nx, Nx)) if len(res) > 2 else np.zeros((n, Nx, Nx))
        H = np.zeros((n, Nx + 1, Nx + 1))
        Lx, Ll = lam_e[:, :Nx], lam_e[:, Nx]
        cross = -np.einsum("ei,eij->ej", Lx, d2g)                        # Σᵢ Λᵢ·∂²(−∂g/∂xᵢ·λ)/∂xⱼ∂λ
        if m == "equal":                                                   # R = (−∂g/∂x·λ, −g)
            H[:, :Nx, :Nx] = -d2g * Ll[:, None, None]
        else:                                                              # positive: R = (−∂g/∂x·λ, −g·λ)
            H[:, :Nx, :Nx] = -d2g * (Ll * lam)[:, None, None]
            cross = cross - dg * Ll[:, None]
        H[:, :Nx, Nx] = cross; H[:, Nx, :Nx] = cross
        return H


class DofLoad(ElementType):
    """DofLoad(nod;field,value)  (src/BasicElements.jl:275-284): R = (−value(t)); plain Float64 residual ⇒ no tangent."""

    @classmethod
    def doflist(cls, field, **kw):
        return (1,), ("X",), (field,)

    @classmethod
    def typekey(cls, field, value, **kw):
        return ("DofLoad", field, id(value))

    @classmethod
    def construct(cls, coords, field, value, args=()):
        return np.zeros((coords[0].shape[0], 0)), dict(value=value, args=args)

    @staticmethod
    def residual(extra, X, t):
        n = X[0].shape[0]
        F = float(extra["value"](t, *extra["args"]))
        return np.full((n, 1), -F), np.zeros((n, 1, 1)), None, None


class Taylor2:
    """Second-order Taylor scalar (value, d/dx, d²/dx²), vectorised over elements: the part of ∂ℝ{2,1,∂ℝ{1,1}} (src/Adiff.jl) a single-dof
    cost function needs.  Supports + − * / ** (real exponent) and unary minus; `Taylor2.apply(f, f′, f″)` lifts any other function."""
    __array_priority__ = 100

    def __init__(self, v, d1=0., d2=0.):
        self.v, self.d1, self.d2 = np.asarray(v, float), d1, d2

    @staticmethod
    def lift(x):
        return x if isinstance(x, Taylor2) else Taylor2(x)

    def apply(self, f, f1, f2):
        a, b, c = f(self.v), f1(self.v), f2(self.v)
        return Taylor2(a, b * self.d1, c * self.d1 * self.d1 + b * self.d2)

    def __add__(self, o): o = Taylor2.lift(o); return Taylor2(self.v + o.v, self.d1 + o.d1, self.d2 + o.d2)
    __radd__ = __add__
    def __neg__(self): return Taylor2(-self.v, -self.d1, -self.d2)
    def __sub__(self, o): return self + (-Taylor2.lift(o))
    def __rsub__(self, o): return Taylor2.lift(o) + (-self)
    def __mul__(self, o):
        o = Taylor2.lift(o)
        return Taylor2(self.v * o.v, self.d1 * o.v + self.v * o.d1, self.d2 * o.v + 2. * self.d1 * o.d1 + self.v * o.d2)
    __rmul__ = __mul__
    def __truediv__(self, o):
        o = Taylor2.lift(o)
        return self * o.apply(lambda x: 1. / x, lambda x: -1. / (x * x), lambda x: 2. / (x * x * x))
    def __rtruediv__(self, o): return Taylor2.lift(o) / self
    def __pow__(self, p):
        p = float(p)
        return self.apply(lambda x: x ** p, lambda x: p * x ** (p - 1.), lambda x: p * (p - 1.) * x ** (p - 2.))


class SingleDofCost(ElementType):
    """SingleDofCost(nod;class,field,cost,costargs) (src/BasicElements.jl:198-208), derivative=0: lagrangian L = cost(dof,t,costargs...).
    `clas` replaces the reserved word `class`.  The cost is a Python closure, so the host evaluates it (value, gradient and second
    derivative through Taylor2) and the device merges the result (mb_direct_set_host_cost)."""
    kind = "hostcost"

    @classmethod
    def doflist(cls, clas, field, **kw):
        if clas not in ("X", "U"):
            raise ValueError("'class' must be :X or :U")
        return (1,), (clas,), (field,)

    @classmethod
    def typekey(cls, clas, field, cost, **kw):
        return ("SingleDofCost", clas, field, id(cost))

    @classmethod
    def construct(cls, coords, clas, field, cost, costargs=()):
        return np.zeros((coords[0].shape[0], 0)), dict(clas=clas, cost=cost, args=tuple(costargs))

    @staticmethod
    def cost_derivs(extra, x, t):
        """x (nele,) → cost, ∂cost/∂x, ∂²cost/∂x² per element"""
        c = Taylor2.lift(extra["cost"](Taylor2(x, 1., 0.), t, *extra["args"]))
        bc = lambda a: np.broadcast_to(np.asarray(a, float), np.shape(x)).copy()
        return bc(c.v), bc(c.d1), bc(c.d2)

    @staticmethod
    def lagrangian(eleobj, extra, Λ, X, U, A, t, SP):
        """lagrangian(o::SingleDofCost,Λ,X,U,A,t,SP,dbg) (src/BasicElements.jl:203-208), derivative = 0 — the general DirectXUA path (xua.py)"""
        dof = X[0][0] if extra["clas"] == "X" else U[0][0]
        return extra["cost"](dof, t, *extra["args"])


# ------------------------------------------------------------------------------------------------ element types of the general DirectXUA path (xua.py)
class LagrangianElement(ElementType):
    """Base of host-evaluated element types written against adiff2.D2: `residual(eleobj, extra, X, U, A, t, SP)` → [R₁…] or
    `lagrangian(eleobj, extra, Λ, X, U, A, t, SP)` → L, with X[der][i], U[der][i], A[i], Λ[i] per-element arrays or D2 (xua.packets)."""
    kind = "lagrangian"
    no_second_order = False
    takes_UA = True            # in an X-analysis (SweepX) U and A reach the element as plain values (model.EleTyp._residual_d2)

    @classmethod
    def typekey(cls, **kw):
        return (cls.__name__,) + tuple(sorted((k, id(v) if callable(v) else (tuple(v) if isinstance(v, (list, tuple)) else v)) for k, v in kw.items()
                                              if k in getattr(cls, "type_parameters", ())))

    @classmethod
    def construct(cls, coords, **kw):
        return np.zeros((coords[0].shape[0] if coords else 1, 0)), dict(kw)


SingleDofCost.kind_general = "lagrangian"


class SingleUdof(LagrangianElement):
    """SingleUdof(nod;Xfield,Ufield,cost,costargs) (src/BasicElements.jl:239-248): L = cost(u,t,costargs...) − λ·u"""
    type_parameters = ("Xfield", "Ufield", "cost")

    @classmethod
    def doflist(cls, Xfield, Ufield, **kw):
        return (1, 1), ("X", "U"), (Xfield, Ufield)

    @classmethod
    def construct(cls, coords, Xfield, Ufield, cost, costargs=()):
        return np.zeros((coords[0].shape[0], 0)), dict(cost=cost, args=tuple(costargs))

    @staticmethod
    def lagrangian(eleobj, extra, Λ, X, U, A, t, SP):
        u = U[0][0]
        return extra["cost"](u, t, *extra["args"]) - Λ[0] * u


class Acost(LagrangianElement):
    """Acost(nod;inod,field,cost,costargs) (src/BasicElements.jl:77-86): L = cost(A,costargs...), added once (assembleA!, src/Assemble.jl:507-520)"""
    acost = True
    type_parameters = ("inod", "field", "cost")

    @classmethod
    def doflist(cls, inod=(), field=(), **kw):
        return tuple(inod), ("A",) * len(inod), tuple(field)

    @classmethod
    def construct(cls, coords, inod=(), field=(), cost=None, costargs=()):
        return np.zeros((coords[0].shape[0] if coords else 1, 0)), dict(cost=cost, args=tuple(costargs))

    @staticmethod
    def lagrangian(eleobj, extra, Λ, X, U, A, t, SP):
        return extra["cost"](A, *extra["args"])


class SingleAcost(Acost):
    """SingleAcost(nod;field,cost,costargs) (src/BasicElements.jl:163-167): an Acost on one A-dof, cost(a,costargs...)"""
    type_parameters = ("field", "cost")

    @classmethod
    def doflist(cls, field, **kw):
        return (1,), ("A",), (field,)

    @classmethod
    def construct(cls, coords, field, cost, costargs=()):
        return np.zeros((coords[0].shape[0], 0)), dict(cost=cost, args=tuple(costargs))

    @staticmethod
    def lagrangian(eleobj, extra, Λ, X, U, A, t, SP):
        return extra["cost"](A[0], *extra["args"])


# ------------------------------------------------------------------------------------------------ strain gauges on beams, ElementCost (device accelerator, xua.py)
class StrainGaugeOnEulerBeam3D(EulerBeam3D):
    """StrainGaugeOnEulerBeam3D(nod;P,D,ElementType=EulerBeam3D,elementkwargs) (toolbox/StrainGaugeOnBeamElement.jl:47-67): wraps an EulerBeam3D, adds the requestables
    εₐₓ, κ and ε = E·εₐₓ + K1·κ₁ + K2·κ₂ + K3·κ₃ (:70-76).  P, D: (3, Ngauge) position offsets (P[0,:] = 0) and directions of the gauges."""

    @classmethod
    def doflist(cls, P=None, D=None, elementkwargs=None, **kw):
        return EulerBeam3D.doflist(**(elementkwargs or {}))

    @classmethod
    def typekey(cls, P=None, D=None, elementkwargs=None, **kw):
        # P and D are type data (the gauge matrix G lives in the type's `extra`): two calls with different gauges must not merge into one type
        return ("StrainGaugeOnEulerBeam3D", np.asarray(P, float).shape[1], np.asarray(P, float).tobytes(), np.asarray(D, float).tobytes()) + EulerBeam3D.typekey(**(elementkwargs or {}))

    @staticmethod
    def gauge_matrix(P, D):
        """(Ngauge,4): E, K1, K2, K3 per gauge (:62-65)"""
        P = np.asarray(P, float); D = np.asarray(D, float)
        if not (P[0, :] == 0.).all():
            raise ValueError("In arguments of StrainGaugeOnEulerBeam3D, P[1,:] must all be zero")
        E = D[0, :] ** 2
        K1 = D[0, :] * (D[2, :] * P[1, :] - D[1, :] * P[2, :])
        K2 = -D[0, :] ** 2 * P[1, :]
        K3 = -D[0, :] ** 2 * P[2, :]
        return np.ascontiguousarray(np.stack([E, K1, K2, K3], axis=1))

    @classmethod
    def construct(cls, coords, P, D, elementkwargs):
        return EulerBeam3D.construct(coords, **elementkwargs), dict(G=cls.gauge_matrix(P, D), P=np.asarray(P, float), D=np.asarray(D, float))


class QuadraticGaugeCost:
    """One of the cost functors the device evaluates for ElementCost: cost(eleres,t) = Δε·Δε/(2σ²), Δε = eleres.ε − εm(t) (the `straincost` of
    test/TestBeamElementStrainGauge.jl:90-97).  measured(t) → (Ngauge,) for all elements, or (nele,Ngauge)."""

    def __init__(self, sigma, measured):
        self.sigma, self.measured = float(sigma), measured


class _Eleres:
    """the requested element results handed to a cost / gap functor: attributes by name (`eleres.Fh`), only what `req` names"""

    def __init__(self, d, req):
        missing = [k for k in req if k not in d]
        if missing:
            from .model import muscadeerror
            muscadeerror("the target element does not provide the requested result(s) %s" % (missing,))
        for k in req:
            setattr(self, k, d[k])


def _target_lagrangian(T, o, extra, Λ, X, U, A, t, SP):
    """getlagrangian(target,…,req) (src/Assemble.jl:690-740) for a host-evaluated target written against adiff2.D2: L and the dict of its element results.
    The target provides `lagrangian_req` / `residual_req` (→ (L | [R…], dict)); plain `lagrangian` / `residual` give no results."""
    if hasattr(T, "lagrangian_req"):
        return T.lagrangian_req(o, extra, Λ, X, U, A, t, SP)
    if hasattr(T, "residual_req"):
        R, res = T.residual_req(o, extra, X, U, A, t, SP)
    elif hasattr(T, "lagrangian"):
        return T.lagrangian(o, extra, Λ, X, U, A, t, SP), {}
    else:
        R, res = T.residual(o, extra, X, U, A, t, SP), {}
    L = 0.
    for i, Ri in enumerate(R):                     # L = Λ ∘₁ R (src/Assemble.jl:721-726)
        L = L + Λ[i] * Ri
    return L, res


class ElementCost(ElementType):
    """ElementCost(nod;req,cost,costargs,ElementType,elementkwargs) (src/BasicElements.jl:117-132): L = getlagrangian(target) + cost(eleres,t,costargs...).
    On the device (csrc/mb_xua.cu, csrc/mb_direct.cu: the accelerator of src/DirectXUA.jl:172-198): ElementType = StrainGaugeOnEulerBeam3D, req = ("ε",), cost = QuadraticGaugeCost.
    Around a host-evaluated target (a LagrangianElement with `lagrangian_req` / `residual_req`): the generic second-order path, differentiated by adiff2.D2 (xua.packets);
    pinned to test/TestElementCost.jl:29-33."""
    kind = "elementcost"
    kind_general = "lagrangian"
    takes_UA = True

    @classmethod
    def doflist(cls, ElementType=None, elementkwargs=None, **kw):
        return ElementType.doflist(**elementkwargs)

    @classmethod
    def typekey(cls, req=None, cost=None, ElementType=None, elementkwargs=None, **kw):
        return ("ElementCost", tuple(req or ()), id(cost)) + ElementType.typekey(**elementkwargs)

    @classmethod
    def construct(cls, coords, req, cost, ElementType, elementkwargs, costargs=()):
        built = ElementType.construct(coords, **elementkwargs)
        eleobj, extra = built if isinstance(built, tuple) else (built, {})
        return eleobj, dict(extra, req=tuple(req), cost=cost, costargs=tuple(costargs), target=ElementType)

    @staticmethod
    def lagrangian(o, extra, Λ, X, U, A, t, SP):
        L, res = _target_lagrangian(extra["target"], o, extra, Λ, X, U, A, t, SP)
        return L + extra["cost"](_Eleres(res, extra["req"]), t, *extra["costargs"])


def equal(t): return "equal"          # src/BasicElements.jl:310-345: the three constant mode functors
def positive(t): return "positive"
def off(t): return "off"


class ElementConstraint(ElementType):
    """ElementConstraint(nod;λinod,λfield,req,gap,gargs,mode,ElementType,elementkwargs) (src/BasicElements.jl:493-575): the target's dofs plus one U-dof λ on node λinod;
    L = getlagrangian(target) − gap(eleres,t,gargs...)·λ (`equal`), − KKT(λ,gap,γ) (`positive`), − λ²/2 (`off`).  Host-evaluated through adiff2.D2 (the generic
    second-order path; the reference's accelerator of src/DirectXUA.jl:199-241 carries its own TODO and is not restated); pinned to test/TestElementCost.jl:63-85."""
    kind = "elementconstraint"
    kind_general = "lagrangian"
    takes_UA = True

    @classmethod
    def doflist(cls, λinod=None, λfield=None, ElementType=None, elementkwargs=None, **kw):
        inod, clas, field = ElementType.doflist(**elementkwargs)
        return tuple(inod) + (λinod,), tuple(clas) + ("U",), tuple(field) + (λfield,)

    @classmethod
    def typekey(cls, λinod=None, λfield=None, req=None, gap=None, mode=None, ElementType=None, elementkwargs=None, **kw):
        return ("ElementConstraint", λinod, λfield, tuple(req or ()), id(gap), id(mode)) + ElementType.typekey(**elementkwargs)

    @classmethod
    def construct(cls, coords, λinod, λfield, req, gap, mode, ElementType, elementkwargs, gargs=()):
        built = ElementType.construct(coords, **elementkwargs)
        eleobj, extra = built if isinstance(built, tuple) else (built, {})
        nu = sum(1 for c in ElementType.doflist(**elementkwargs)[1] if c == "U")
        return eleobj, dict(extra, req=tuple(req), gap=gap, gargs=tuple(gargs), mode=mode, target=ElementType, Nu=nu)

    @staticmethod
    def lagrangian(o, extra, Λ, X, U, A, t, SP):
        from .adiff2 import KKT
        Nu = extra["Nu"]
        u = [list(ud[:Nu]) for ud in U]
        λ = U[0][Nu]
        L, res = _target_lagrangian(extra["target"], o, extra, Λ, X, u, A, t, SP)
        gap = extra["gap"](_Eleres(res, extra["req"]), t, *extra["gargs"])
        m = extra["mode"](t)
        γ = float((SP or {}).get("γ", 0.)) if isinstance(SP, dict) or SP is None else float(getattr(SP, "γ", 0.))
        if m == "equal":
            return L - gap * λ
        if m == "positive":
            return L - KKT(λ, gap, γ)
        if m == "off":
            return L - 0.5 * (λ * λ)
        from .model import muscadeerror
        muscadeerror("mode(t) must return 'equal', 'positive' or 'off'")
