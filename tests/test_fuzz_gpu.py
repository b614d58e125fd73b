"""Randomised differential test of the CUDA path against the C oracle (tools/fuzz_parity.py): random geometries, scales, quirks,
tap formats, filter kernels, gray and BGRA batches.  Two fixed seeds; any unexcused difference fails."""
import json
import os
import subprocess
import sys

import pytest

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [7, 8])
def test_random_cases_against_oracle(seed):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "--cases", "30", "--seed", str(seed)],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    tot = json.loads(p.stdout.strip().splitlines()[-1])
    print("fuzz seed %d: %s" % (seed, tot))
    assert tot["cases"] == 30 and tot["excused"] < 1e-5 * tot["pixels"]
