import sys, os, time
import numpy as np
sys.path.insert(0, ".")
import muscade_b200 as mb
N = int(float(sys.argv[1])); oxs = [int(x) for x in sys.argv[2].split(",")]
eng = mb.Engine(0)
eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=True)
eng.add_eulerbeam3d(eleobj, idx, np.ones(12)); eng.sweepx_prepare(ndof)
X = mb.synthetic.state(ndof, nder=3)
for OX in oxs:
    for mission in (["iter"] if OX == 0 else ["iter", "step"]):
        nm = mb.synthetic.newmark_coefficients(OX, 0.3)
        eng.sweepx_assemble(OX, mission, X, nm)
        eng.time_dev(OX, mission, nm, reps=1)
        el, ga = eng.time_dev(OX, mission, nm, reps=5)
        print(f"{os.environ.get('MB_LIB','default'):40s} OX={OX} {mission}: element {el:.3f} ms gather {ga:.3f} ms -> {N/el*1e3:.3e} el/s", flush=True)
