"""General DirectXUA path with DEVICE element types: N strain-gauged EulerBeam3D{Udof} (ElementCost accelerator) x nstep steps, DirectXUA{OX,0,0}.
Times one assemblebig! (eval_device + add_step per step) with CUDA-synchronised wall clock; prints element-step assemblies/s and the share of the phases."""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import muscade_b200 as mb
from muscade_b200 import xua
import ctypes as C
from muscade_b200._lib import check, ErrInfo

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20000
OX = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nstep = int(sys.argv[3]) if len(sys.argv) > 3 else 6
P5 = np.array([[0., .5, 0.], [0., 0, .5], [0., -.5, 0.], [0., 0, -.5], [0., .5, 0.]]).T
D5 = np.array([[1., 0., 0.], [1., 0., 0.], [1., 0., 0.], [1., 0., 0.], [1 / np.sqrt(2), 0, 1 / np.sqrt(2)]]).T
m = mb.Model("gauged")
nod = mb.addnode(m, np.stack([np.arange(N + 1, dtype=float), np.zeros(N + 1), np.zeros(N + 1)], axis=1))
un = mb.addnode(m, np.zeros((N, 3)))
nodes = np.stack([nod[:-1], nod[1:], un], axis=1)
cost = mb.QuadraticGaugeCost(1e-5, lambda t: np.zeros(5))
mb.addelement(m, mb.ElementCost, nodes, req=("ε",), cost=cost, ElementType=mb.StrainGaugeOnEulerBeam3D,
              elementkwargs=dict(P=P5, D=D5, elementkwargs=dict(mat=mb.BeamCrossSection(EA=1e4, EI2=300., EI3=300., GJ=400., mu=1., iota1=1.), orient2=(0., 1., 0.), Udof=True)))
s0 = mb.initialize(m); dis = s0.dis
nX, nU, nA = m.getndof(("X", "U", "A"))
eng = xua.XUAEngine(0)
t0 = time.perf_counter()
nbig, nnz = eng.prepare(m, dis, OX, 0, 0, [nstep], [0.1])
print("prepare %.2f s: Lv %d, Lvv nnz %d (%.1f GB of values + maps)" % (time.perf_counter() - t0, nbig, nnz, nnz * 24 / 1e9), flush=True)
rng = np.random.default_rng(0)
st = s0.with_orders(1, OX + 1, 1)
for k in range(nstep):
    s = mb.State(0.1 * k, [rng.normal(0, 1., nX)], [rng.normal(0, 0.01, nX) for _ in range(OX + 1)], [rng.normal(0, 0.1, nU)], st.A, None, m, dis)
    eng.put_state(1, k + 1, s)
em = np.zeros(5)
def one_pass():
    eng.zero()
    t_ev = t_add = 0.
    for k in range(nstep):
        check(eng.h, eng.L.mb_xua_set_gauge_measurements(eng.h, 1, em.ctypes.data_as(C.c_void_p), 0))
        t1 = time.perf_counter()
        check(eng.h, eng.L.mb_xua_eval_device(eng.h, 1, k + 1, C.byref(ErrInfo())))
        t2 = time.perf_counter()
        eng.add_step(1, k + 1); eng.sync()
        t3 = time.perf_counter()
        t_ev += t2 - t1; t_add += t3 - t2
    return t_ev, t_add
one_pass()
ev, ad = one_pass()
print("N %d OX %d nstep %d: eval_device %.2f ms/step, add_step %.2f ms/step -> %.3e element-step assemblies/s, launches %d" %
      (N, OX, nstep, 1e3 * ev / nstep, 1e3 * ad / nstep, N * nstep / (ev + ad), eng.launch_count()), flush=True)
eng.close()
