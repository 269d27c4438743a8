"""e2e of mb_sweepx_assemble (pinned host state in, host Lλ out, nzval left in HBM) against the number of pipeline chunks (MB_E2E_CHUNKS)."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import muscade_b200 as mb

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000
eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=False)
X = mb.synthetic.state(ndof, nder=1)
nm = mb.synthetic.newmark_coefficients(0, 0.)
for chunks in (1, 4, 8, 16, 32, 64):
    os.environ["MB_E2E_CHUNKS"] = str(chunks)
    eng = mb.Engine(0)
    eng.add_eulerbeam3d(eleobj, idx, np.ones(12)); eng.sweepx_prepare(ndof)
    Lh = np.empty(ndof)
    for a in X + [Lh]: eng.pin(a)
    for _ in range(3): eng.sweepx_assemble(0, "iter", X, nm, Llambda=Lh, nzval_on_device=True)
    ts = []
    for _ in range(5):
        t1 = time.perf_counter(); eng.sweepx_assemble(0, "iter", X, nm, Llambda=Lh, nzval_on_device=True); ts.append(time.perf_counter() - t1)
    print("chunks %2d: e2e %.2f ms (min %.2f) -> %.3e el/s" % (chunks, 1e3 * np.mean(ts), 1e3 * min(ts), N / np.mean(ts)), flush=True)
    for a in X + [Lh]: eng.unpin(a)
    eng.close()
