"""Random shapes / metrics / data scales: the tensor-core path against the on-device fp64 scan.  -m gpu."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [1, 2])
def test_random_shapes_equal_the_fp64_scan(cuda_device, seed):
    """scripts/fuzz_shapes.py: 120 random (nq, nb, d, k, metric, data kind) combinations per seed -- every filter
    variant, key form and plan the host picks -- must reproduce the fp64 scan bit for bit."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fuzz_shapes.py"), "120", str(seed)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_random_running_searches_equal_the_one_shot_result(cuda_device):
    """scripts/fuzz_stream.py: random shapes, slab splits and arrival orders through nwk_topk_state_search."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fuzz_stream.py"), "100", "4"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_random_maxsim_shapes_equal_the_oracle(cuda_device):
    """scripts/fuzz_maxsim.py: random token counts (1..128 per query, 1..256 per document), dimensions, ragged
    documents and query scales against oracle/maxsim_oracle.py."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fuzz_maxsim.py"), "30", "2"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
