"""cProfile of the public call with PAGEABLE host tensors: python dev/e2e_profile.py <workload> [stream_chunks]"""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from concepthash_b200 import hashing  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
w, d, dl, q, ql = bench.make_workload(name, "cuda")
hd, hdl, hq, hql = (t.cpu() for t in (d, dl, q, ql))
del d, dl, q, ql
ev = hashing.get_evaluator()
if len(sys.argv) > 2:
    ev.stream_chunks = int(sys.argv[2])
f = lambda: hashing.calculate_mAP(hd, hdl, hq, hql, w["R"])
for _ in range(4):
    f()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    f()
torch.cuda.synchronize()
print("ms per step", (time.perf_counter() - t0) * 100, ev.stats["mode"], ev.stats["geometry"])
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    f()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
