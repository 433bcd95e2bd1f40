"""Host-side cost of one evaluation on a small shard (the per-rank share of cfg4 at 8 GPUs):
python dev/host_profile.py  -> wall time per step and the cProfile top list"""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
from concepthash_b200 import hashing, synth  # noqa: E402

import os
NDB = int(os.environ.get("NDB", "125000"))
d, dl, q, ql, ncls = synth.make_random_case(25000, NDB, 128, 101, p=0.30, seed=0, device="cuda")
ev = hashing.get_evaluator()
f = lambda: ev.evaluate(d, dl, q, ql, [1000], 0.0, [], False)
for _ in range(3):
    f()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    f()
torch.cuda.synchronize()
print("ms per step", (time.perf_counter() - t0) * 100, ev.stats["mode"], "launches", ev.b.launch_count())
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    f()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
