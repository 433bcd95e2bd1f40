"""The two training-loss helpers the reference imports from the same missing ``utils.hashing`` module
(models/loss/dpsh.py:4, models/loss/hashnet.py:5, models/loss/adsh.py:5): a drop-in module must export them.
CPU tests; their semantics are pinned by the call sites and the formula kept beside them as a comment."""
import os
import re

import pytest
import torch

from concepthash_b200.hashing import get_sim, log_trick

REF = "/root/reference"


def test_every_name_the_reference_imports_from_utils_hashing_is_exported():
    """import lines collected from the reference tree when it is present (this container), else the committed list"""
    names = {"calculate_mAP", "calculate_pr_curve", "get_hamm_dist", "get_sim", "log_trick"}
    if os.path.isdir(REF):
        found = set()
        for root, _, files in os.walk(REF):
            for f in files:
                if f.endswith(".py"):
                    for line in open(os.path.join(root, f), errors="ignore"):
                        m = re.match(r"\s*from utils\.hashing import (.+)", line)
                        if m:
                            found |= {n.strip() for n in m.group(1).split("#")[0].replace("(", "").replace(")", "").split(",")
                                      if n.strip()}
        assert found == names, found ^ names
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_dropin_utils_hashing", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "utils", "hashing.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for n in names:
        assert callable(getattr(mod, n)), n


def test_get_sim_onehot_multihot_and_ids():
    a = torch.tensor([[1., 0, 0], [0, 1, 1], [0, 0, 0]])
    b = torch.tensor([[1., 1, 0], [0, 0, 1]])
    s = get_sim(a, b)
    assert s.dtype == torch.bool and s.tolist() == [[True, False], [True, True], [False, False]]
    assert get_sim(a, b).float().sum().item() == 3.0                       # the callers' `.float()`
    ia, ib = torch.tensor([3, 1, 2]), torch.tensor([1, 3])
    assert get_sim(ia, ib, False).tolist() == [[False, True], [True, False], [False, False]]
    assert get_sim(ia.view(-1, 1), ib.view(-1, 1), onehot=False).shape == (3, 2)
    # the relevance notion of calculate_mAP (oracle.relevance_matrix) is the same statement
    from oracle import map_oracle as mo
    g = torch.Generator().manual_seed(0)
    y1 = (torch.rand(17, 9, generator=g) < 0.2).float()
    y2 = (torch.rand(23, 9, generator=g) < 0.2).float()
    assert torch.equal(get_sim(y1, y2), mo.relevance_matrix(y1, y2).bool())


def test_log_trick_is_a_stable_differentiable_softplus():
    x = torch.linspace(-30, 30, 240, dtype=torch.float64, requires_grad=True)   # (no exact 0: |x| has no gradient there)
    y = log_trick(x)
    ref = torch.log1p(torch.exp(x.detach()))
    assert torch.allclose(y.detach(), ref, rtol=1e-12, atol=1e-12)
    # the commented expression at models/loss/hashnet.py:79 / dpsh.py:64, verbatim
    d = x.detach()
    assert torch.equal(y.detach(), (1 + (-d.abs()).exp()).log() + d.clamp(min=0))
    y.sum().backward()
    assert torch.allclose(x.grad, torch.sigmoid(x.detach()), atol=1e-12)
    big = torch.tensor([-1e4, 1e4, 0.0])
    out = log_trick(big)
    assert torch.isfinite(out).all() and out[1].item() == pytest.approx(1e4) and out[0].item() == pytest.approx(0.0, abs=1e-30)
    assert out[2].item() == pytest.approx(0.6931471805599453, rel=1e-6)


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree (this container only)")
def test_reference_losses_run_with_the_drop_in_module():
    """the reference's own pairwise losses, imported UNMODIFIED with this repo's ``utils/hashing.py`` standing in for
    the module they import their helpers from, run forward and backward (a subprocess: its own ``sys.path``)"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = """
import torch, utils.hashing, utils.metrics
assert utils.hashing.__file__.startswith(%r) and utils.metrics.__file__.startswith(%r)
from models.loss.dpsh import DPSHLoss
from models.loss.hashnet import HashNetLoss
torch.manual_seed(0)
u = torch.randn(12, 16, requires_grad=True)
y = torch.nn.functional.one_hot(torch.randint(4, (12,)), 4)
for loss in (DPSHLoss(), HashNetLoss()):
    out = loss(u, y)
    out = out[0] if isinstance(out, (tuple, list)) else out
    assert torch.isfinite(out).all()
    out.sum().backward()
assert torch.isfinite(u.grad).all() and u.grad.abs().sum() > 0
print("ok")
""" % (root, REF)
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + REF)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp", timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout + res.stderr
