"""Second-order forward differentiation for host-evaluated element types, vectorised over the elements of a type.

`D2` carries (value, gradient, Hessian) with respect to Np seeded variables — what ∂ℝ{2,Np,∂ℝ{1,Np}} carries in the reference (src/Adiff.jl) — for
arrays of elements: v (n,), g (n,Np), H (n,Np,Np).  User elements written against it (lagrangian / residual as plain arithmetic) are differentiated
by the host as the reference differentiates its own with dual numbers; the device assembles the result (xua.py → mb_xua_set_packet).
Supports + − * / ** (real exponent), unary minus, and the functions below; `apply(f, f′, f″)` lifts any other scalar function."""
import numpy as np


class D2:
    __array_priority__ = 100
    __slots__ = ("v", "g", "H")

    def __init__(self, v, g=None, H=None):
        self.v = np.asarray(v, float)
        self.g, self.H = g, H                     # None = identically zero

    # ---- construction
    @staticmethod
    def variables(x, scale=None):
        """x (n,Np) values → list of Np D2, variable k seeded with gradient scale[k]·e_k (revariate{2} with scales, src/Taylor.jl:158-166)"""
        x = np.asarray(x, float)
        n, Np = x.shape
        sc = np.ones(Np) if scale is None else np.asarray(scale, float)
        out = []
        for k in range(Np):
            g = np.zeros((n, Np)); g[:, k] = sc[k]
            out.append(D2(x[:, k].copy(), g, None))
        return out

    @staticmethod
    def lift(x):
        return x if isinstance(x, D2) else D2(x)

    def grad(self, n, Np):
        return np.zeros((n, Np)) if self.g is None else np.broadcast_to(self.g, (n, Np)).copy()

    def hess(self, n, Np):
        return np.zeros((n, Np, Np)) if self.H is None else np.broadcast_to(self.H, (n, Np, Np)).copy()

    # ---- chain rule for a scalar function
    def apply(self, f, f1, f2):
        a, b = f(self.v), f1(self.v)
        if self.g is None:
            return D2(a)
        c = f2(self.v)
        H = c[..., None, None] * (self.g[..., :, None] * self.g[..., None, :])
        if self.H is not None:
            H = H + b[..., None, None] * self.H
        return D2(a, b[..., None] * self.g, H)

    # ---- arithmetic
    def __add__(self, o):
        o = D2.lift(o)
        g = o.g if self.g is None else (self.g if o.g is None else self.g + o.g)
        H = o.H if self.H is None else (self.H if o.H is None else self.H + o.H)
        return D2(self.v + o.v, g, H)
    __radd__ = __add__

    def __neg__(self):
        return D2(-self.v, None if self.g is None else -self.g, None if self.H is None else -self.H)

    def __sub__(self, o): return self + (-D2.lift(o))
    def __rsub__(self, o): return D2.lift(o) + (-self)

    def __mul__(self, o):
        o = D2.lift(o)
        v = self.v * o.v
        g = None
        if self.g is not None: g = self.g * o.v[..., None]
        if o.g is not None: g = (o.g * self.v[..., None]) if g is None else g + o.g * self.v[..., None]
        H = None
        def acc(H, t): return t if H is None else H + t
        if self.H is not None: H = acc(H, self.H * o.v[..., None, None])
        if o.H is not None: H = acc(H, o.H * self.v[..., None, None])
        if self.g is not None and o.g is not None:
            c = self.g[..., :, None] * o.g[..., None, :]
            H = acc(H, c + np.swapaxes(c, -1, -2))
        return D2(v, g, H)
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = D2.lift(o)
        return self * o.apply(lambda x: 1. / x, lambda x: -1. / (x * x), lambda x: 2. / (x * x * x))

    def __rtruediv__(self, o): return D2.lift(o) / self

    def __pow__(self, p):
        p = float(p)
        if p == 2.:
            return self * self
        return self.apply(lambda x: x ** p, lambda x: p * x ** (p - 1.), lambda x: p * (p - 1.) * x ** (p - 2.))


LN10 = np.log(10.)
def exp(x): x = D2.lift(x); return x.apply(np.exp, np.exp, np.exp)
def exp10(x): x = D2.lift(x); return x.apply(lambda v: 10. ** v, lambda v: LN10 * 10. ** v, lambda v: LN10 * LN10 * 10. ** v)
def log(x): x = D2.lift(x); return x.apply(np.log, lambda v: 1. / v, lambda v: -1. / (v * v))
def sqrt(x): x = D2.lift(x); return x.apply(np.sqrt, lambda v: .5 / np.sqrt(v), lambda v: -.25 / (v * np.sqrt(v)))
def sin(x): x = D2.lift(x); return x.apply(np.sin, np.cos, lambda v: -np.sin(v))
def cos(x): x = D2.lift(x); return x.apply(np.cos, lambda v: -np.sin(v), lambda v: -np.cos(v))
def absolute(x): x = D2.lift(x); return x.apply(np.abs, np.sign, lambda v: np.zeros_like(v))


def KKT(lam, g, gamma):
    """KKT(λ,g,γ) (src/BasicElements.jl:289-307): "a pseudo-potential with strange derivatives" — value 0, gradient λ·∇g + S·∇λ with the complementary slackness
    S(λ,g,γ) = g·λ − γ, and as second derivative the derivative of THAT expression, as the reference's nested duals form it (entry [i][j] = ∂ⱼ of gradient entry i):
    λⱼ·gᵢ + λ·gᵢⱼ + (gⱼ·λ + g·λⱼ)·λᵢ + S·λᵢⱼ — not symmetric."""
    lam, g = D2.lift(lam), D2.lift(g)
    v = np.zeros(np.broadcast(lam.v, g.v).shape)
    if lam.g is None and g.g is None:
        return D2(v)
    S = g.v * lam.v - gamma
    col = lambda a: np.asarray(a, float)[..., None]
    Np = (lam.g if lam.g is not None else g.g).shape[-1]
    zg = np.zeros(v.shape + (Np,))
    lg = lam.g if lam.g is not None else zg
    gg = g.g if g.g is not None else zg
    grad = col(lam.v) * gg + col(S) * lg
    H = gg[..., :, None] * lg[..., None, :] + (col(lam.v) * gg + col(g.v) * lg)[..., None, :] * lg[..., :, None]
    if g.H is not None:
        H = H + col(lam.v)[..., None] * g.H
    if lam.H is not None:
        H = H + col(S)[..., None] * lam.H
    return D2(v, grad, H)
