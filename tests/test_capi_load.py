"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/vlite_b200.h declares, and fails LOUDLY (no CPU fallback) when asked to compute
without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from vlite_b200 import _capi


def test_library_exports_every_declared_symbol():
    lib = _capi.load()
    declared = _capi.declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert set(lib._signatures) == set(declared), "ctypes signatures out of sync with the header"
    assert b"sm_100a" in lib.vl_version()


def test_header_cites_the_reference_interfaces_it_replaces():
    text = open(_capi.HEADER_PATH).read()
    for cite in ("vlite/model.py:76-90", "vlite/model.py:82", "vlite/main.py:143-146", "vlite/ctx.py:116-119",
                 "vlite/main.py:198", "vlite/main.py:159-165"):
        assert cite in text, cite
    assert "torch" not in re.sub(r"/\*.*?\*/", "", text, flags=re.S).lower()   # no torch types in signatures


def test_library_is_built_for_sm_100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_a_device():
    if _capi.device_count() > 0:
        pytest.skip("a B200 is visible; the loud-failure path is for CPU boxes")
    with pytest.raises(_capi.VliteB200Error) as ei:
        _capi.Corpus(128)
    assert ei.value.code == _capi.ENODEV
    assert "no CPU fallback" in ei.value.msg
    with pytest.raises(_capi.VliteB200Error):
        _capi.microbench(0, 16)
    with pytest.raises(_capi.VliteB200Error):
        _capi.quantize_binary(np.zeros((1, 64), np.float32))


def test_argument_validation_does_not_need_a_device():
    lib = _capi.load()
    h = ctypes.c_void_p()
    assert lib.vl_corpus_create(0, 100, 0, ctypes.byref(h)) == _capi.EINVAL      # bad width (32/64/128 only)
    assert b"width_bytes" in lib.vl_last_error()
    assert lib.vl_search(None, None, 1, 1, 0, None, None, None) == _capi.EINVAL
    assert lib.vl_corpus_destroy(None) == 0
    assert lib.vl_launch_count() >= 0


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "vlite_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle_scan" not in src, f
