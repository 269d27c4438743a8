"""Regenerates tests/golden/*.npz.  Run from the repository root:  python tests/golden/make_fixtures.py

The reference is a Julia package and Julia is not in this image, so these vectors are NOT outputs of the reference: they are outputs of the
oracle (oracle/, the literal C++ restatement of the reference algorithm), frozen here after the oracle was pinned against the reference's
own golden values (tests/test_oracle_goldens.py, transcribed from test/Test*.jl of the reference).  They serve two purposes:
  * `-m "not gpu"`: the oracle must keep reproducing them bit-for-bit on this toolchain (regression of the checker itself);
  * `-m gpu`: the CUDA path is compared with them on the GPU box, where /root/reference does not exist.
Inputs are the deterministic synthetic ones of SURVEY.md §8d (splitmix64 states, chain mesh), so only outputs are stored."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import muscade_b200 as mb                 # host-side model building only (no CUDA needed)
from oracle import elements as OE
from oracle import pattern as OP

HERE = os.path.dirname(os.path.abspath(__file__))


def sweepx_chain(N=8):
    """BASELINE.json configs[0]-sized chain (8 EulerBeam3D): Lλ and nzval for every (OX, mission), random and zero state."""
    out = {}
    for dynamic in (False, True):
        eleobj, idx, ndof = mb.synthetic.chain(N, dynamic=dynamic)
        dis = [dict(X=idx, U=np.zeros((N, 0), np.int64), A=np.zeros((N, 0), np.int64))]
        asm1, asm2, colptr, rowval = OP.prepare_sweepx(dis, ndof, 0, 0)
        out["colptr"], out["rowval"] = colptr, rowval
        for OX, mission in [(0, "iter")] if not dynamic else [(1, "iter"), (1, "step"), (2, "iter"), (2, "step")]:
            for zero in (False, True):
                X = mb.synthetic.state(ndof, nder=OX + 1, zero=zero)
                nm = mb.synthetic.newmark_coefficients(OX, 0.3)
                L = np.zeros(ndof); nz = np.zeros(len(rowval))
                OE.sweepx_assemble_beams(eleobj, idx, asm1[0].T, asm2[0].T, OX, mission, X, np.ones(12), nm, L, nz)
                key = "OX%d_%s_%s" % (OX, mission, "zero" if zero else "rand")
                out[key + "_L"], out[key + "_nz"] = L, nz
    return out


def direct_chain(N=4, nstep=6, dt=0.1):
    """DirectXUA{2,0,0} on a 4-element Udof chain, 6 steps (the minimum finitediff accepts): Lvv structure + values and Lv."""
    model = mb.Model()
    coord = np.arange(N + 1)[:, None] * np.array([.8, .6, 0.])[None, :]
    nod = mb.addnode(model, coord)
    unod = mb.addnode(model, np.zeros((N, 0)))
    mat = mb.BeamCrossSection(EA=10., EI2=3., EI3=2.5, GJ=4., mu=1., iota1=1.2, w=.3, Ca2=2., Ca3=1.5, Cq2=1., Cq3=.7, Cl1=.2)
    mb.addelement(model, mb.EulerBeam3D, np.stack([nod[:-1], nod[1:], unod], axis=1), mat=mat, Udof=True)
    st0 = mb.initialize(model); dis = st0.dis
    nX, nU, nA = model.getndof(("X", "U", "A"))
    st = [([mb.synthetic.uniform_pm1(10 + 3 * s + d, nX) * (0.1 if d == 0 else 0.3) for d in range(3)], mb.synthetic.uniform_pm1(99 + s, nU)) for s in range(nstep)]
    odis = [dict(X=d.X, U=d.U, A=d.A) for d in dis.dis]
    P = OP.prepare_direct(odis, nX, nU, nA, 2, 0, 0)
    big, bigasm, pgr, pgc = OP.preparebig(0, [nstep], P["nL2"], P["pat"])
    outs = [OE.direct_assemble_step_beams(model.ele[0].eleobj, dis.dis[0].X, dis.dis[0].U, 2, 0, X, [U], dis.dis[0].scaleX, dis.dis[0].scaleU, P, 0) for (X, U) in st]
    nz, Lv = OP.assemblebig(0, nstep, dt, P, big, bigasm, pgr, outs)
    return dict(colptr=big["colptr"], rowval=big["rowval"], nzval=nz, Lv=Lv)


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "sweepx_chain8.npz"), **sweepx_chain())
    np.savez_compressed(os.path.join(HERE, "directxua_chain4x6.npz"), **direct_chain())
    print("written", os.listdir(HERE))
