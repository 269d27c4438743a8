"""Reference example models rebuilt on the host mirror (used by tests/ and by `bench.py --workload scr`).

`scr_riser`: the steel catenary riser of examples/DynamicBeamAnalysis.jl:62-135 (BASELINE.json configs[1] at the reference's own size and the beam +
SoilContact mesh of configs[4]): 100 EulerBeam3D in three segments with two cross-sections, 4 Hold, 2 DofConstraint with time-dependent gaps, 101 DofLoad
(weight ramp), 61 SoilContact.  The RIFLEX top-motion series (examples/SCR.csv) is replaced by a synthetic harmonic motion: the file does not travel."""
import numpy as np

G_, RHO = 9.81, 1025.


def xsection(D, t, EA, EI, GJ, rg, Dh):
    steel = (D - t) * np.pi * t; inner = (D - 2 * t) ** 2 * np.pi / 4
    mu = steel * 7850. + inner * 200.
    return dict(EA=EA, EI2=EI, EI3=EI, GJ=GJ, mu=mu, iota1=rg ** 2 * steel * 7850., Ca2=RHO * np.pi * Dh ** 2 / 4, Ca3=RHO * np.pi * Dh ** 2 / 4,
                Cq2=0.5 * RHO * Dh, Cq3=0.5 * RHO * Dh), mu * G_ - np.pi * D ** 2 / 4 * RHO * G_


X1, W1 = xsection(0.429, 0.022, 5.823e9, 1.209e8, 9.347e7, 0.2053, 0.459)
X2, W2 = xsection(0.441, 0.028, 7.520e9, 1.611e8, 1.245e8, 0.2084, 0.471)
NEL, SEGLEN = [60, 30, 10], [300., 300., 80.]


def xmotion(t):            # stands for the interpolated RIFLEX series (DynamicBeamAnalysis.jl:9-13): zero up to t = 0, then harmonic
    return -1.5 * np.sin(2 * np.pi * max(t, 0.) / 9.0)


def zmotion(t):
    return 0.8 * np.sin(2 * np.pi * max(t, 0.) / 11.0)


def horiz_target(t): return 1.0 - np.exp(min(t, 0.)) * 181.0 + xmotion(t)       # DynamicBeamAnalysis.jl:113
def vert_target(t): return np.exp(min(t, 0.)) * 303.1 + zmotion(t)               # :114
def ramp(t): return (min(t, -5.) + 10.) / 5.                                     # :119-121


def scr_riser(mb, udof=False, gauge_cost=None):
    """udof: EulerBeam3D{Udof=true} with one U-node per element (unknown distributed loads, the XUA set-up of configs[4]).
    gauge_cost: a QuadraticGaugeCost — every beam is wrapped in ElementCost{StrainGaugeOnEulerBeam3D} with four gauges round the pipe wall (load identification from strains)"""
    acc = np.concatenate([[0.], np.cumsum(SEGLEN)])

    def addbeams(nodes, mat):
        if gauge_cost is None:
            mb.addelement(model, mb.EulerBeam3D, nodes, mat=mat, orient2=(0., 1., 0.), Udof=udof)
        else:
            r = 0.15
            P = np.array([[0., r, 0.], [0., 0., r], [0., -r, 0.], [0., 0., -r]]).T
            D = np.array([[1., 0., 0.]] * 4).T
            mb.addelement(model, mb.ElementCost, nodes, req=("ε",), cost=gauge_cost, ElementType=mb.StrainGaugeOnEulerBeam3D,
                          elementkwargs=dict(P=P, D=D, elementkwargs=dict(mat=mat, orient2=(0., 1., 0.), Udof=udof)))

    def mesh(n1, n2):
        n1, n2 = np.atleast_1d(n1), np.atleast_1d(n2)
        cols = [n1, n2] + ([mb.addnode(model, np.zeros((len(n1), 0)))] if udof else [])
        return np.stack(cols, axis=1)
    model = mb.Model("CatenaryRiser")
    mats = [mb.BeamCrossSection(**X1), mb.BeamCrossSection(**X2), mb.BeamCrossSection(**X1)]
    node_lists = []
    for seg in range(3):
        nn = NEL[seg] + 1
        c = np.stack([acc[seg] + np.arange(nn) / (nn - 1) * SEGLEN[seg], np.zeros(nn), -300. + np.zeros(nn)], axis=1)
        if seg == 0:
            nod = mb.addnode(model, c)
            addbeams(mesh(nod[:-1], nod[1:]), mats[0])
            last = nod[-1]
        else:
            nod = mb.addnode(model, c[1:])
            addbeams(mesh(last, nod[0]), mats[seg])
            addbeams(mesh(nod[:-1], nod[1:]), mats[seg])
            last = nod[-1]
        node_lists.append(nod)
    first = node_lists[0][0]
    for f in ["t1", "t2", "t3", "r1"]:
        mb.addelement(model, mb.Hold, [first], field=f)
    gap_h = lambda x, t: (x[:, 0] - horiz_target(t), np.ones_like(x))
    gap_v = lambda x, t: (x[:, 0] - vert_target(t), np.ones_like(x))
    mb.addelement(model, mb.DofConstraint, [last], xinod=(1,), xfield=("t1",), λinod=1, λclass="X", λfield="λt1", gap=gap_h, mode="equal")
    mb.addelement(model, mb.DofConstraint, [last], xinod=(1,), xfield=("t3",), λinod=1, λclass="X", λfield="λt3", gap=gap_v, mode="equal")
    weights = []
    for seg, w in enumerate([W1, W2, W1]):
        f = (lambda ww, s: (lambda t: -ramp(t) * ww * SEGLEN[s] / NEL[s]))(w, seg)
        weights.append(f)
        for n in node_lists[seg]:
            mb.addelement(model, mb.DofLoad, [n], field="t3", value=f)
    mb.addelement(model, mb.SoilContact, node_lists[0][:, None], z0=0., Kh=1.0e3, Kv=1.0e4, Ch=0., Cv=0.)
    return model, node_lists, weights
