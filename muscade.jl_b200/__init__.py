"""muscade_b200 — B200-native element-evaluation-and-assembly path of SINTEF/Muscade.jl.

Host-side mirror of the reference interface for this path (Model / addnode! / addelement! / initialize! / prepare /
assemble! / solve(SweepX)), over the C ABI in include/muscade_b200.h.  The directory name contains a dot, so the package
is imported through the loader `muscade_b200.py` at the repository root (`import muscade_b200`).
"""
from . import _lib, toolbox, synthetic, model, sweepx, directxua, sharding, examples, adiff2, xua, eigx  # noqa: F401
from ._lib import MuscadeB200Error, build  # noqa: F401
from .engine import Engine  # noqa: F401
from .model import Model, addnode, addelement, setscale, initialize, Disassembler, State, getdof  # noqa: F401
from .toolbox import (BeamCrossSection, EulerBeam3D, AxisymmetricBarCrossSection, Bar3D, SoilContact, Hold, DofLoad, DofConstraint,
                      ElementType, SingleDofCost, Taylor2, LagrangianElement, SingleUdof, Acost, SingleAcost, StrainGaugeOnEulerBeam3D, QuadraticGaugeCost, ElementCost, ElementConstraint, equal, positive, off)  # noqa: F401
from .adiff2 import D2  # noqa: F401
