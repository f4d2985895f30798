"""The C-ABI library builds, loads without a GPU and exports every symbol include/uvlt.h declares."""
import ctypes
import os
import re

from util import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "uvlt.h")).read()
    return sorted(set(re.findall(r"UVLT_API\s+[\w\s\*]+?\b(uvlt_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from uvltrack_b200 import _cabi

    assert os.path.exists(_cabi.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/uvlt.h but not exported"


def test_ctypes_table_matches_header():
    from uvltrack_b200 import _cabi

    assert sorted(_cabi.SIGNATURES) == _declared()
    lib = _cabi.load()
    assert lib.uvlt_abi_version() == 1


def test_structs_match_header_layout():
    from uvltrack_b200 import _cabi

    # uvlt_config: 16 int32 + int32[32]; uvlt_outputs: 8 pointers + 6 int32
    assert ctypes.sizeof(_cabi.UvltConfig) == 4 * (16 + 32)
    assert ctypes.sizeof(_cabi.UvltOutputs) == 8 * 8 + 4 * 6


def test_no_gpu_fails_loudly():
    """Without a CUDA device the product must raise, not fall back."""
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import uvltrack_b200 as u

    with pytest.raises(RuntimeError):
        u.registry.MODELS["uvltrack"](u.config.baseline_cfg())
    from uvltrack_b200 import _cabi

    lib = _cabi.load()
    h = ctypes.c_void_p()
    cfg = _cabi.UvltConfig()
    assert lib.uvlt_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"CUDA" in lib.uvlt_last_error() or b"device" in lib.uvlt_last_error()


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under uvltrack_b200/ may import it."""
    pkg = os.path.join(ROOT, "uvltrack_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
