"""2-GPU check of the time-sharded general DirectXUA form (torchrun --nproc-per-node 2): each rank adds its share of the steps, mb_xua_allreduce_big sums Lvv / Lv;
compared with the single-rank assemblebig! of the same problem (IA = 1, two experiments)."""
import os, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, torch.distributed as dist
import muscade_b200 as mb
from muscade_b200 import xua
import xua_models as XM

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
rng = np.random.default_rng(5)
m = XM.model_chain(23, rng, ox=2); s0 = mb.initialize(m); dis = s0.dis
OX, OU, IA, nsteps, dts = 2, 0, 1, [6, 7], [1., 0.5]
st = s0.with_orders(1, OX + 1, OU + 1)
A = rng.normal(0, 0.1, st.A.shape)
states = [[mb.State(3. + dt * k, [rng.normal(0, .3, v.shape) for v in st.Λ], [rng.normal(0, .3, v.shape) for v in st.X], [rng.normal(0, .3, v.shape) for v in st.U], A, None, m, dis)
           for k in range(n)] for n, dt in zip(nsteps, dts)]
eng = xua.XUAEngine(local); eng.prepare(m, dis, OX, OU, IA, nsteps, dts)
uid = torch.from_numpy(mb.Engine.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)); dist.broadcast(uid, src=0)
eng.comm_init(uid.numpy(), rank, world)
eng.assemblebig_sharded(states, rank, world)
nz, Lv = eng.big()
ref = xua.XUAEngine(local); ref.prepare(m, dis, OX, OU, IA, nsteps, dts); ref.assemblebig(states)
nz0, Lv0 = ref.big()
e1, e2 = np.abs(nz - nz0).max() / np.abs(nz0).max(), np.abs(Lv - Lv0).max() / np.abs(nz0).max()
print("rank %d: sharded vs single-rank assemblebig!: rel err Lvv %.2e Lv %.2e" % (rank, e1, e2), flush=True)
assert e1 < 1e-14 and e2 < 1e-14
eng.close(); ref.close()
