"""Static element kernel against the shared-memory carve-out asked for it (MB_AP_CARVE, %; the kernel uses no shared memory: what is not carved out is L1 for its spill lines)."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import muscade_b200 as mb
N = int(float(sys.argv[1])); os.environ["MB_E2E_CHUNKS"] = "1"
eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=False)
X = mb.synthetic.state(ndof, nder=1); nm = mb.synthetic.newmark_coefficients(0, 0.)
eng = mb.Engine(0); eng.add_eulerbeam3d(eleobj, idx, np.ones(12)); eng.sweepx_prepare(ndof)
eng.set_state(X)
for _ in range(4): eng.sweepx_assemble_dev(0, "iter", nm)
eng.sync(); print("carve", os.environ.get("MB_AP_CARVE"), eng.time_dev(0, "iter", nm, reps=5), flush=True)
