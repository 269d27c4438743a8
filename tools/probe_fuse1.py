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
eng.sync(); print(eng.time_dev(0, "iter", nm, reps=3))
