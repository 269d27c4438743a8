"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's (SINTEF/Muscade.jl v0.7.0) algorithm for the element-evaluation-and-assembly
hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; nothing under ``muscade.jl_b200/`` does.

Parity pinning: the oracle is checked against every golden vector the reference's own tests hold for this path
(``tests/test_oracle_*.py``: TestAdiff, TestRotations, TestBeamElement, TestBarElement, TestAssemble, TestSparseTools,
TestDirectXUA, TestFiniteDifferences, TestModelDescription integer/FP pins, and the Longva/Crisfield tip positions of
examples/StaticBeamAnalysis.jl).  The reference itself (Julia) cannot be run in this image; floating point parity
at 1e-12 is therefore oracle↔GPU, with the oracle pinned to the reference at the ≈1e-8 level of its goldens.

 * ``oracle.elements``  – ctypes front-end of the C++ literal restatement (adiff.hpp, rotations.hpp, beam.hpp, bar.hpp)
 * ``oracle.pattern``   – numpy restatement of Disassembler / asmvec! / asmmat! / SparseTools.prepare / finitediff
"""
from . import elements, pattern  # noqa: F401
