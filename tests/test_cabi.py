"""The C-ABI shared library loads without a GPU and exports every symbol include/muscade_b200.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "muscade_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mb_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(mb):
    so = mb._lib.SO_PATH
    assert os.path.exists(so), "run __graft_entry__.build()"
    L = C.CDLL(so)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    # the Python binding covers the same set
    assert set(names) == set(mb._lib.SYMBOLS)


def test_no_cpu_fallback(mb):
    """without a CUDA device the engine refuses to exist; there is no CPU path in the product"""
    import subprocess, sys
    code = "import os; os.environ['CUDA_VISIBLE_DEVICES']=''; import sys; sys.path.insert(0, %r); import muscade_b200 as mb\n" \
           "try:\n    mb.Engine(0); print('CREATED')\nexcept mb.MuscadeB200Error as e:\n    print('REFUSED', e)" % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout
    assert "REFUSED" in out and "CREATED" not in out


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "muscade.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".jl")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, os.path.join(dirpath, f)
