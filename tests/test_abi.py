"""The C-ABI library loads and exports every symbol include/tfhe_b200.h
declares (no compute calls: there is no GPU in the build container)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "tfhe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b([a-z_0-9]+)\s*\(", text)
    skip = {"defined", "sizeof"}
    return sorted({n for n in names if n not in skip and (n.startswith(("cuda_", "scratch_", "cleanup_", "has_", "b200_")))})


def test_library_exports_every_declared_symbol():
    import tfhe_rs_b200
    from tfhe_rs_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        tfhe_rs_b200.build()
    L = tfhe_rs_b200.lib()
    declared = _header_functions()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/tfhe_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(declared)
    assert L.b200_pbs_uses_fast_path(918, 1, 2048, 1) == 1
    assert L.b200_pbs_uses_fast_path(918, 1, 2048, 2) == 0
    assert L.b200_pbs_uses_fast_path(805, 3, 512, 2) == 0


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under tfhe-rs_b200/ may
    reference it."""
    pkg = os.path.join(ROOT, "tfhe-rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pbs_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_missing_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tfhe_rs_b200 import gpu

    with pytest.raises(RuntimeError):
        gpu.CudaStreams.new_single_gpu(0)
