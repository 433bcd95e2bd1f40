"""per-step wall time of the public call with pinned host tensors: python dev/e2e_steps.py <workload> [steps]"""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from concepthash_b200 import hashing  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5s"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
w, d, dl, q, ql = bench.make_workload(name, "cuda")
import os
PIN = os.environ.get("PIN", "1") == "1"
hd, hdl, hq, hql = ((t.cpu().pin_memory() if PIN else t.cpu()) for t in (d, dl, q, ql))
del d, dl, q, ql
ev = hashing.get_evaluator()
for i in range(steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m = hashing.calculate_mAP(hd, hdl, hq, hql, w["R"])
    torch.cuda.synchronize()
    print(i, "ms %.2f" % ((time.perf_counter() - t0) * 1e3), ev.stats["mode"], ev.stats.get("sample", {}).get("fallback"),
          "alloc GB %.1f reserved %.1f" % (torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9))
