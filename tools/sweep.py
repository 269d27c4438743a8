"""Size sweep of SURVEY.md §8d on one GPU: SweepX{0} :iter, SweepX{2} :iter/:step for N ∈ {8,1e3,1e5,1e6,1e7} (element kernels + segmented
reduction, CUDA events, state resident in HBM).  Prints a markdown table (profiles/README.md)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import muscade_b200 as mb

print("| N | OX | mission | element kernels (ms) | reduction (ms) | element-assemblies/s |")
print("|---|---|---|---|---|---|")
for N in (8, 1000, 100000, 1000000, 10000000):
    eng = mb.Engine(0)
    eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=True)
    eng.add_eulerbeam3d(eleobj, idx, np.ones(12)); eng.sweepx_prepare(ndof)
    X = mb.synthetic.state(ndof, nder=3)
    for OX, mission in ((0, "iter"), (2, "iter"), (2, "step")):
        nm = mb.synthetic.newmark_coefficients(OX, 0.3)
        eng.sweepx_assemble(OX, mission, X, nm)
        eng.time_dev(OX, mission, nm, reps=1)           # warm-up: the whole-model cotangent workspace is allocated at the first device-path launch
        el, ga = eng.time_dev(OX, mission, nm, reps=5 if N >= 1000000 else 20)
        print("| %d | %d | %s | %.3f | %.3f | %.3e |" % (N, OX, mission, el, ga, N / (el + ga) * 1e3), flush=True)
    eng.close()
