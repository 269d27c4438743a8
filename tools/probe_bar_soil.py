"""K4/K5 against the HBM roofline (Bar3D and SoilContact are HBM-bound): a chain of N Bar3D elements (3 translation dofs per node) and one
SoilContact per node, SweepX{0} and SweepX{2} `:iter`, element kernels and reduction timed with CUDA events on the engine's stream.
Algorithmic bytes per element — Bar3D: geometry 64 + dof indices 24 + state 48·(OX+1) + Ke 288 + Re 48; SoilContact: parameters 40 + indices 12 + state 24·(OX+1)
+ Ke 72 + Re 24.   usage: python tools/probe_bar_soil.py [N]"""
import json, os, sys
import numpy as np
sys.path.insert(0, ".")
import muscade_b200 as mb
from muscade_b200 import toolbox

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
try:
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    peak = 6650.0
p = np.arange(N + 1, dtype=np.float64)[:, None] * np.array([0.8, 0.6, 0.0])[None, :]
mat = toolbox.AxisymmetricBarCrossSection(EA=500., mu=1.5, w=3., Cat=.2, Clt=.3, Cqt=.4, Can=2., Cln=.6, Cqn=1.2)
bars = toolbox.bar3d_structs(p[:-1], p[1:], mat)
e = np.arange(N, dtype=np.int64)[:, None]
idxb = 3 * e + np.arange(1, 7, dtype=np.int64)[None, :]
ndof = 3 * (N + 1)
for what in ("bar3d", "soilcontact"):
    eng = mb.Engine(0)
    if what == "bar3d":
        eng.add_bar3d(bars, idxb, np.ones(6)); nele = N
        alg = lambda OX: 64 + 24 + 48 * (OX + 1) + 288 + 48
    else:
        n = N + 1
        soil = np.tile(np.array([0., 30., 200., 3., 7.]), (n, 1))
        eng.add_soilcontact(soil, 3 * np.arange(n, dtype=np.int64)[:, None] + np.arange(1, 4, dtype=np.int64)[None, :], np.ones(3)); nele = n
        alg = lambda OX: 40 + 12 + 24 * (OX + 1) + 72 + 24
    nnz = eng.sweepx_prepare(ndof)
    X = [mb.synthetic.uniform_pm1(7 + d, ndof) * (0.05 if d == 0 else 0.1) for d in range(3)]
    for OX in (0, 2):
        nm = mb.synthetic.newmark_coefficients(OX, 0.3)
        eng.sweepx_assemble(OX, "iter", X[: OX + 1], nm)
        el, ga = eng.time_dev(OX, "iter", nm, reps=5)
        gbs = nele * alg(OX) / (el * 1e-3) / 1e9
        print(json.dumps({"kernel": what, "elements": nele, "OX": OX, "element_ms": el, "reduction_ms": ga, "alg_bytes_per_element": alg(OX),
                          "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak, "elements_per_s": nele / (el * 1e-3)}), flush=True)
    eng.close()
