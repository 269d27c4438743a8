"""Fused epilogue of the static kernel (MB_FUSE=1, default) against the unfused path (MB_FUSE=0): timing at N elements and BIT-equality of Lλ / nzval."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import muscade_b200 as mb

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000
eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=False)
X = mb.synthetic.state(ndof, nder=1)
nm = mb.synthetic.newmark_coefficients(0, 0.)
ref = None
for fuse in (0, 1):
    os.environ["MB_FUSE"] = str(fuse)
    eng = mb.Engine(0)
    eng.add_eulerbeam3d(eleobj, idx, np.ones(12)); eng.sweepx_prepare(ndof)
    L, nz = eng.sweepx_assemble(0, "iter", X, nm)
    eng.time_dev(0, "iter", nm, reps=2)
    el, ga = eng.time_dev(0, "iter", nm, reps=5)
    st = eng.time_step_dev(0, "iter", nm, reps=5)
    msg = ""
    if ref is None:
        ref = (L.copy(), nz.copy())
    else:
        msg = " bit-equal L %s nzval %s  max|dnz| %.3e" % (np.array_equal(L, ref[0]), np.array_equal(nz, ref[1]), np.abs(nz - ref[1]).max())
    print("fuse %d: element %.3f ms  gather %.3f ms  step %.3f ms -> %.3e el/s%s" % (fuse, el, ga, st, N / st * 1e3, msg), flush=True)
    eng.close()
