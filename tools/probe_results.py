"""one batched getresult of N EulerBeam3D elements at the device-resident state (run under `ncu --metrics gpu__time_duration.sum -k regex:beam_results`)"""
import sys
import numpy as np
sys.path.insert(0, ".")
import muscade_b200 as mb
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
eng = mb.Engine(0)
eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=True)
ityp = eng.add_eulerbeam3d(eleobj, idx, np.ones(12)); eng.sweepx_prepare(ndof)
X = mb.synthetic.state(ndof, nder=3)
eng.set_state(X)
for OX in (0, 2):
    r = eng.beam_results(ityp, OX)
    print(OX, r["raw"].shape, float(np.abs(r["raw"]).max()))
