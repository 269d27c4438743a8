"""Quick device probe: kernel timings of the SweepX assembly at a given size (not the bench)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import muscade_b200 as mb

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
eng = mb.Engine(0)
print("fp64 TFLOP/s", eng.fp64_tflops(), "copy GB/s", eng.copy_gbs(), flush=True)
t = time.time(); eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=True); print("gen", time.time() - t, flush=True)
t = time.time(); eng.add_eulerbeam3d(eleobj, idx, np.ones(12)); print("add", time.time() - t, flush=True)
t = time.time(); nnz = eng.sweepx_prepare(ndof); print("prepare", time.time() - t, "nnz", nnz, flush=True)
X = mb.synthetic.state(ndof, nder=3)
for OX, mission in [(0, "iter"), (2, "iter"), (2, "step")]:
    nm = mb.synthetic.newmark_coefficients(OX, 0.3)
    L, nz = eng.sweepx_assemble(OX, mission, X, nm)
    eng.time_dev(OX, mission, nm, reps=1)
    el, ga = eng.time_dev(OX, mission, nm, reps=3)
    print(f"OX={OX} {mission}: element {el:.3f} ms, gather {ga:.3f} ms -> {N/(el+ga)*1e3:.3e} el/s (kernels only), |L|={np.abs(L).max():.3e}", flush=True)
t = time.time(); L, nz = eng.sweepx_assemble(0, "iter", X, mb.synthetic.newmark_coefficients(0, 0.)); print("e2e host call (pageable)", time.time() - t)
