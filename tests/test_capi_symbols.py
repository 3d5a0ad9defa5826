"""The C-ABI library loads and exports every symbol include/cudacollision.h declares (not gpu, no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cudacollision.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cc_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_header():
    from openrave_b200 import checker
    if not os.path.isfile(checker.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(checker.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
    from openrave_b200 import planning
    assert sorted(checker.EXPORTS + planning.PLANNER_EXPORTS) == syms


def test_no_device_fails_loudly():
    """Without a CUDA device the product must refuse to work (no CPU path)."""
    import subprocess
    import sys
    code = ("import os,sys; os.environ['CUDA_VISIBLE_DEVICES']=''; sys.path.insert(0,%r);"
            "from openrave_b200.checker import CudaCollisionChecker, CudaCollisionError\n"
            "try:\n    CudaCollisionChecker(None,0); print('CREATED')\n"
            "except CudaCollisionError as e:\n    print('REFUSED', e)\n") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120).stdout
    assert "REFUSED" in out and "no CUDA device" in out


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "openrave_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f
