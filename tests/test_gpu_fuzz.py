"""-m gpu: short runs of the two randomized differential drivers under scripts/ (their long runs
are how the round's code was soaked; these keep them from rotting)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("script,args,needle", [
    ("fuzz_search.py", ["120", "21"], "120 cases, 0 mismatches"),
    ("fuzz_collection.py", ["120", "22"], "fuzz_collection: 120 steps"),
    ("fuzz_collection.py", ["100", "23", "0,0"], "fuzz_collection: 100 steps"),
])
def test_fuzz_driver(gpu, script, args, needle):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script)] + args, capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert needle in res.stdout
