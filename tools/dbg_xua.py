import sys; sys.path.insert(0,"tests"); sys.path.insert(0,".")
import numpy as np, muscade_b200 as mb, scipy.sparse as sp, scipy.sparse.linalg as spla
import xua_models as XM
from muscade_b200 import xua
m = XM.model_testdirectxua(); s0 = mb.initialize(m); dis = s0.dis
OX, OU, IA = 2, 0, 1
times = [np.arange(0., 6.), np.arange(6., 13.)]
st = s0.with_orders(1, 3, 1); A = st.A.copy()
states = [[mb.State(float(t), [v.copy() for v in st.Λ], [v.copy() for v in st.X], [v.copy() for v in st.U], A, None, m, dis) for t in tt] for tt in times]
eng = xua.XUAEngine(0)
eng.prepare(m, dis, OX, OU, IA, [6, 7], [1., 1.])
for ie in range(2):
    for k in range(len(times[ie])): eng.put_state(ie + 1, k + 1, states[ie][k])
for it in range(3):
    eng.assemblebig(states)
    nz, Lv = eng.big()
    print("Lv[A]", Lv[-6:])
    cp, rv, v = eng.sparser(1e-20)
    dv = spla.splu(sp.csc_matrix((v, rv - 1, cp - 1), shape=(eng.nbig, eng.nbig))).solve(Lv)
    print("dv[A]", dv[-6:])
    print("d2", eng.decrement(dv))
    for ie in range(2):
        for k in range(len(times[ie])): eng.fetch_state(ie + 1, k + 1, states[ie][k])
    print("A", A, states[0][0].A is A)
