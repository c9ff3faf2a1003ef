"""The C-ABI library loads on a CPU-only machine and exports every symbol include/prs_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "prs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(prs_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def library():
    from pyresample_b200 import _cabi
    if _cabi.needs_rebuild():
        _cabi.build()
    return _cabi.load_library()


def test_header_symbols_are_exported(library):
    from pyresample_b200 import _cabi
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(library, name), name
    assert set(declared) == set(_cabi.EXPORTED_SYMBOLS)
    assert library.prs_abi_version() == 1


def test_struct_layouts_match_header():
    """ctypes mirrors vs the C structs: sizes computed by compiling a probe with gcc."""
    import subprocess
    import tempfile
    from pyresample_b200 import _cabi
    from pyresample_b200.proj import prs_proj_params
    src = r'''
#include <stdio.h>
#include "prs_b200.h"
int main(void) { printf("%zu %zu %zu %zu\n", sizeof(prs_proj_params), sizeof(prs_area_grid), sizeof(prs_reduce_box), sizeof(prs_target)); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "probe.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "probe")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(prs_proj_params), ctypes.sizeof(_cabi.prs_area_grid),
                     ctypes.sizeof(_cabi.prs_reduce_box), ctypes.sizeof(_cabi.prs_target)]


def test_no_cpu_fallback_without_gpu(library):
    """Compute entry points refuse to run without a CUDA device instead of falling back."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pyresample_b200 import _cabi, geometry, kd_tree
    swath = geometry.SwathDefinition(np.arange(10.), np.arange(10.))
    with pytest.raises(_cabi.PrsError):
        kd_tree.resample_nearest(swath, np.arange(10.), swath, 1000)
    with pytest.raises(_cabi.PrsError):
        kd_tree.get_neighbour_info(swath, swath, 1000, neighbours=1)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under pyresample_b200/ may reference it."""
    pkg = os.path.join(ROOT, "pyresample_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "scipy" not in text, f
