"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/lc_gpu.h
declares, and refuses to work without a GPU (no CPU fallback). No compute calls here."""
import ctypes

import pytest


def test_library_exports_every_declared_symbol():
    from liquid_cache_b200 import _native

    lib = _native.lib()
    declared = _native.exported_symbols_declared_in_header()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"liblc_gpu.so does not export {missing}"


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; the no-device refusal cannot be exercised")
    from liquid_cache_b200 import LiquidCacheBuilder
    from liquid_cache_b200 import _native as N

    with pytest.raises(N.NativeError) as ei:
        LiquidCacheBuilder.new().build()
    assert ei.value.code == N.LC_ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_product_package_never_imports_the_oracle():
    import os
    import re

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "liquid_cache_b200")
    for dirpath, _d, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cc", ".cu", ".h", ".cuh", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f"{f} imports the oracle"
                assert "oracle/" not in text or f.endswith((".md",)), f"{f} references oracle/"


def test_struct_layouts_match_header():
    from liquid_cache_b200 import _native as N

    assert ctypes.sizeof(N.Predicate) == 40
    assert ctypes.sizeof(N.Stats) == 56
