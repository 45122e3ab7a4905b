"""CPU tests: the C-ABI library loads, exports every symbol glba.h declares, struct layouts agree."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import gl_slam_b200 as g
from gl_slam_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "glba.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(glba_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = g.lib()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"libglba.so does not export {n}"
    assert sorted(g.ABI_SYMBOLS) == names
    assert L.glba_version() == 100


def test_struct_layouts_match_header(tmp_path):
    prog = tmp_path / "layout.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "glba.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                    "sizeof(glba_device_cfg),sizeof(glba_problem),sizeof(glba_options),sizeof(glba_summary),sizeof(glba_linearization),"
                    "offsetof(glba_summary,cost),offsetof(glba_summary,accepted),offsetof(glba_options,cg_rel_tol));return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = list(map(int, subprocess.check_output([str(exe)]).split()))
    want = [C.sizeof(_abi.DeviceCfg), C.sizeof(_abi.Problem), C.sizeof(_abi.Options), C.sizeof(_abi.Summary),
            C.sizeof(_abi.Linearization), _abi.Summary.cost.offset, _abi.Summary.accepted.offset, _abi.Options.cg_rel_tol.offset]
    assert got == want


def test_defaults_are_the_reference_constants():
    o = g.options()       # slam_core.cpp:814 CauchyLoss(1.0), :846 max_num_iterations = 30, Ceres defaults otherwise
    assert (o.loss, o.loss_scale, o.max_iters) == (_abi.LOSS_CAUCHY, 1.0, 30)
    assert (o.function_tol, o.gradient_tol, o.parameter_tol) == (1e-6, 1e-10, 1e-8)
    assert (o.initial_radius, o.max_radius, o.min_radius, o.min_relative_decrease) == (1e4, 1e16, 1e-32, 1e-3)
    assert (o.min_lm_diagonal, o.max_lm_diagonal, o.jacobi_scaling) == (1e-6, 1e32, 1)
    from oracle import oracle
    oo = oracle.options()
    for f, _ in _abi.Options._fields_:
        assert getattr(o, f) == getattr(oo, f), f


def test_error_strings_and_null_handling():
    L = g.lib()
    assert b"no CPU fallback" in L.glba_strerror(_abi.E_NO_DEVICE)
    assert L.glba_create(None, None) == _abi.E_INVALID_ARG
    L.glba_destroy(None)       # must not crash
    assert L.glba_kernel_launch_count() >= 0


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(g.GlbaError) as e:
        g.Context(device=0)
    assert e.value.status == _abi.E_NO_DEVICE


def test_product_never_touches_the_oracle():
    """The product path must not import, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "gl_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "glbao_" not in text and "libglba_oracle" not in text, os.path.join(dirpath, f)
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), os.path.join(dirpath, f)
    out = subprocess.check_output(["ldd", g.LIB_PATH]).decode()
    assert "oracle" not in out


def test_docs_name_only_exported_symbols():
    """Every glba_* function INTEGRATION.md / README.md / DESIGN.md mention exists in include/glba.h (docs do not drift)."""
    import re
    header = open(os.path.join(ROOT, "include", "glba.h")).read()
    declared = set(re.findall(r"\b(glba_[a-z0-9_]+)\s*\(", header))
    types = set(re.findall(r"\b(glba_[a-z0-9_]+)\b", header)) - declared
    for doc in ("INTEGRATION.md", "README.md", "DESIGN.md"):
        text = open(os.path.join(ROOT, doc)).read()
        for name in set(re.findall(r"\b(glba_[a-z0-9_]+)\s*\(", text)):
            assert name in declared or name in types, f"{doc} mentions {name}(), which include/glba.h does not declare"
