"""A short run of dev/fuzz_gpu.py: random shapes / label kinds / R lists / remove_first through the top-R paths
(tensor-core select + candidate ranking, sampled and exact, two-level sample) against the CPU oracle."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_random_top_r_cases_match_oracle():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "dev", "fuzz_gpu.py"), "16", "7"], cwd=ROOT,
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "bad 0" in res.stdout
