"""Static element kernel variants (MB_STATIC_SYM = 1 symmetric, 2 active/passive, 3 active/passive + shared packs): timing at N elements and
agreement of Lλ / nzval with variant 1 (not the bench)."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import muscade_b200 as mb

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 3]
eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=False)
X = mb.synthetic.state(ndof, nder=1)
nm = mb.synthetic.newmark_coefficients(0, 0.)
ref = None
for v in variants:
    os.environ["MB_STATIC_SYM"] = str(v)
    eng = mb.Engine(0)
    eng.add_eulerbeam3d(eleobj, idx, np.ones(12)); eng.sweepx_prepare(ndof)
    L, nz = eng.sweepx_assemble(0, "iter", X, nm)
    eng.time_dev(0, "iter", nm, reps=2)
    el, ga = eng.time_dev(0, "iter", nm, reps=5)
    msg = ""
    if ref is None:
        ref = (L.copy(), nz.copy())
    else:
        s = max(np.abs(ref[1]).max(), np.abs(ref[0]).max())
        msg = " |dL|/s=%.2e |dnz|/s=%.2e" % (np.abs(L - ref[0]).max() / s, np.abs(nz - ref[1]).max() / s)
    print("variant %d: element %.3f ms  gather %.3f ms  -> %.3e el/s%s" % (v, el, ga, N / el * 1e3, msg), flush=True)
    eng.close()
