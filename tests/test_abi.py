"""The drop-in boundary: libmgcfd_b200.so loads, exports every symbol include/*.h declares, and -- with no GPU --
refuses to compute instead of falling back to the CPU."""
import ctypes as C
import os
import re

import pytest

import mgcfd_b200 as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in ("mgcfd_b200.h", "mgcfd_mesh.h", "mgcfd_dist.h"):
        p = os.path.join(ROOT, "include", h)
        if not os.path.exists(p):
            continue
        src = re.sub(r"/\*.*?\*/", "", open(p).read(), flags=re.S)
        names += re.findall(r"\b(mgcfd_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_every_declared_symbol_is_exported():
    L = M.lib()
    syms = declared_symbols()
    assert len(syms) >= 45
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_header_compiles_as_plain_c(tmp_path):
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "mgcfd_b200.h"\n#include "mgcfd_mesh.h"\nint main(void){mgcfd_options o; mgcfd_default_options(&o); return o.device;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_version_and_defaults():
    L = M.lib()
    assert b"sm_100a" in L.mgcfd_version()
    o = M.Options()
    L.mgcfd_default_options(C.byref(o))
    assert (o.flux_mode, o.ordering, o.tile_nodes, o.use_graph) == (M.FLUX_SORTED_SEGMENT, M.ORDER_PARTITION_RCM, 0, 1)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(M.MgcfdError) as e:
        M.Solver(1, M.MESH_M6_WING)
    assert e.value.code == 4 and "no CPU fallback" in str(e.value)


def test_argument_errors():
    L = M.lib()
    h = C.c_void_p()
    assert L.mgcfd_create(0, 2, None, C.byref(h)) == 2
    assert L.mgcfd_create(1, 2, None, None) == 2
    assert L.mgcfd_run_cycles(None, 1, None, None) == 2
    assert L.mgcfd_destroy(None) == 0


def test_product_does_not_touch_the_oracle():
    """only tests/, smoke() and bench.py's CPU legs may use oracle/ (the product must not)."""
    pkg = os.path.join(ROOT, "mg-cfd-app-plain_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("oracle's", "").replace("the oracle", ""), os.path.join(dirpath, f)
