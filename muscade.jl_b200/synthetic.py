"""Deterministic synthetic models and states: generated identically for the oracle and the GPU path."""
import numpy as np

from . import toolbox

MASK = (1 << 64) - 1


def splitmix64(seed, counter):
    """splitmix64 of (seed + counter·golden) → uint64 array; counter is an integer array."""
    z = (np.uint64(seed) + np.asarray(counter, np.uint64) * np.uint64(0x9E3779B97F4A7C15)) & np.uint64(MASK)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform_pm1(seed, n, offset=0):
    """U(−1,1) from splitmix64(seed, dof index); `offset` = first (0-based) dof index, so a shard can generate its own slice"""
    with np.errstate(over="ignore"):
        u = splitmix64(seed, np.arange(offset, offset + n, dtype=np.uint64))
    return (u >> np.uint64(11)).astype(np.float64) * (2.0 / (1 << 53)) - 1.0


def chain(N, h=1.0, mat=None, dynamic=False):
    """`chain(N,h)`: N EulerBeam3D elements along (0.8,0.6,0), nodes p_k = k·h·(0.8,0.6,0), orient2=(0,1,0), added in one
    addelement! call ⇒ X-dof numbers 6k+1…6k+6 per node (t1,t2,t3,r1,r2,r3).  Returns (eleobj (N,69), idxX (N,12) 1-based, ndofX)."""
    if mat is None:
        kw = dict(EA=10., EI2=3., EI3=3., GJ=4., mu=1., iota1=1.)          # inspect/PerformanceEulerBeam3D.jl:17
        if dynamic:
            kw.update(Ca2=169.6, Ca3=169.6, Cq2=235.2, Cq3=235.2)            # ≈ examples/DynamicBeamAnalysis.jl:34-37
        mat = toolbox.BeamCrossSection(**kw)
    k = np.arange(N + 1, dtype=np.float64)[:, None]
    p = k * h * np.array([0.8, 0.6, 0.0])[None, :]
    eleobj = toolbox.eulerbeam3d_structs(p[:-1], p[1:], mat)
    e = np.arange(N, dtype=np.int64)[:, None]
    idx = 6 * e + np.arange(1, 13, dtype=np.int64)[None, :]
    return eleobj, idx, 6 * (N + 1)


def state(ndofX, h=1.0, seed=0x5EED, nder=1, zero=False, offset=0):
    """State of SURVEY.md §8d: translations 0.05·h·u, rotations 0.1·u; x′ = 0.1·u′, x″ = 0.1·u″ (seeds +1, +2)."""
    X = []
    for d in range(nder):
        if zero:
            X.append(np.zeros(ndofX)); continue
        u = uniform_pm1(seed + d, ndofX, offset)
        if d == 0:
            amp = np.where(((np.arange(ndofX) + offset) % 6) < 3, 0.05 * h, 0.1)
            X.append(u * amp)
        else:
            X.append(0.1 * u)
    return X


def newmark_coefficients(OX, dt, beta=0.25, gamma=0.5):
    """Newmarkβcoefficients{OX}(Δt,β,γ)  src/SweepX.jl:12-15 → (a1,a2,a3,b1,b2,b3,Δt)"""
    if OX == 0:
        return np.array([0., 0., 0., 0., 0., 0., dt])
    if OX == 1:
        return np.array([1 / (gamma * dt), 1 / gamma, 0., 0., 0., 0., dt])
    return np.array([gamma / (beta * dt), gamma / beta, (gamma / (2 * beta) - 1) * dt, 1 / (beta * dt ** 2), 1 / (beta * dt), 1 / (2 * beta), dt])
