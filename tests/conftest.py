import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        from vlite_b200 import _capi
        return _capi.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """GPU tests must FAIL (not skip) when the native library or the device is missing:
    there is no CPU fallback to hide behind."""
    from vlite_b200 import _capi
    _capi.load()
    assert _capi.device_count() > 0, "no sm_100 device visible to libvlite_b200.so"
    return _capi


@pytest.fixture(autouse=True)
def _device_assertions(request):
    """With VLITE_B200_LIB pointing at the -DVL_DEBUG_BOUNDS build, every GPU test also asserts that no
    device-side bounds / protocol check fired while it ran (the stand-in for compute-sanitizer)."""
    yield
    if "gpu" in request.keywords and os.environ.get("VLITE_B200_LIB"):
        from vlite_b200 import _capi
        for d in range(_capi.device_count()):
            got = _capi.debug_checks(d)
            assert got is not None, "VLITE_B200_LIB is not a debug build"
            assert got[0] == 0, f"{got[0]} device-side check(s) failed on GPU {d}, first id {got[1]}"
