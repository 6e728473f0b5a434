"""CPU-only: libgcp.so builds, loads, and exports every symbol include/gcp.h declares;
the engine refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gcp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gcp_[a-z_0-9]+)\s*\(", src)))


def test_build_and_symbols():
    from genetic_coverage_planning_b200 import build, _lib
    build.build()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "libgcp.so does not export %s" % n
        assert n in _lib.SIGNATURES, "ctypes binding lacks %s" % n
    assert lib.gcp_abi_version() == _lib.GCP_ABI_VERSION


def test_struct_layouts_match_header():
    from genetic_coverage_planning_b200 import _lib
    assert C.sizeof(_lib.GcpParams) == 48
    assert C.sizeof(_lib.GcpMapConsts) == 48
    assert C.sizeof(_lib.GcpEvalOut) == 7 * C.sizeof(C.c_void_p)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from genetic_coverage_planning_b200 import _lib
    lib = _lib.load()
    ctx = C.c_void_p()
    assert lib.gcp_create(C.byref(ctx), 0) != 0
    assert b"no CUDA device" in lib.gcp_last_error(None)
    from genetic_coverage_planning_b200.engine import FitnessEngine, GcpError
    with pytest.raises(GcpError):
        FitnessEngine(None, {}, 1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "genetic_coverage_planning_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "fitness_oracle" not in txt and "ref_shim" not in txt, fn


def test_option_names_documented_and_implemented():
    """Every knob gcp_set_option accepts is described in include/gcp.h, and every knob the header
    describes is accepted (the header is the only documentation a C caller has)."""
    api = open(os.path.join(ROOT, "genetic_coverage_planning_b200", "csrc", "gcp_api.cu")).read()
    implemented = set(re.findall(r'strcmp\(name, "([a-z_0-9]+)"\)', api))
    hdr = open(os.path.join(ROOT, "include", "gcp.h")).read()
    i = hdr.index("tuning / testing knobs")
    documented = set(re.findall(r'"([a-z_0-9]+)"', hdr[i:hdr.index("int gcp_set_option", i)]))
    assert implemented and implemented == documented, (sorted(implemented - documented), sorted(documented - implemented))
