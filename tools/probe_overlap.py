"""Segmented reduction of element chunk j on a second stream while the element kernels of chunk j+1 run (MB_DEV_OVERLAP = 1: above, 2: below the engine's stream priority;
MB_E2E_CHUNKS chunks; MB_GATHER_BLOCK threads per CTA of the reduction).  Prints whole-step times.  The runs recorded in profiles/r2_probe_overlap*.txt also had builds of
the static kernel held to 240 / 224 / 208 registers (`__maxnreg__`, first column) so that the reduction's CTAs fit beside its two CTAs per SM; those variants lost
(18.5 / 19.1 / 19.6 ms against 18.2) and are no longer in the tree — the first column is kept for reading the old logs."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import muscade_b200 as mb

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000000
eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=False)
X = mb.synthetic.state(ndof, nder=1)
nm = mb.synthetic.newmark_coefficients(0, 0.)
ref = None
cfgs = [(255, 0, 8, 256), (255, 1, 8, 256), (255, 2, 8, 256), (255, 2, 16, 256), (255, 2, 32, 128)]
if len(sys.argv) > 2: cfgs = [tuple(int(x) for x in c.split(",")) for c in sys.argv[2:]]
for maxt, ov, ch, gb in cfgs:
    os.environ.update(MB_FUSE="0", MB_DEV_OVERLAP=str(ov), MB_E2E_CHUNKS=str(ch), MB_GATHER_BLOCK=str(gb))
    eng = mb.Engine(0)
    eng.add_eulerbeam3d(eleobj, idx, np.ones(12)); eng.sweepx_prepare(ndof)
    eng.set_state(X)
    for _ in range(3): eng.sweepx_assemble_dev(0, "iter", nm)
    eng.sync()
    st = eng.time_step_dev(0, "iter", nm, reps=5)
    st = min(st, eng.time_step_dev(0, "iter", nm, reps=5))
    L, nz = eng.get_results() if hasattr(eng, "get_results") else (None, None)
    msg = ""
    if L is not None:
        if ref is None: ref = (L.copy(), nz.copy())
        else: msg = " bit-equal L %s nzval %s" % (np.array_equal(L, ref[0]), np.array_equal(nz, ref[1]))
    print("regs %d overlap %d chunks %d gather-block %d: step %.3f ms -> %.3e el/s%s" % (maxt, ov, ch, gb, st, N / st * 1e3, msg), flush=True)
    eng.close()
