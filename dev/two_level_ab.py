"""A/B of the two-level sample on single-GPU shards: python dev/two_level_ab.py"""
import sys
import time

import torch

sys.path.insert(0, ".")
from concepthash_b200 import hashing, synth  # noqa: E402

ev = hashing.get_evaluator()
for ndb in (250_000, 500_000, 1_000_000):
    d, dl, q, ql, ncls = synth.make_random_case(25000, ndb, 128, 101, p=0.30, seed=0, device="cuda")
    for two in (False, True):
        ev.sample_two_level = two
        ev.sample2_min_work = 0.0 if two else 1e30
        f = lambda: ev.evaluate(d, dl, q, ql, [1000], 0.0, [], False)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            f()
        torch.cuda.synchronize()
        print(ndb, "two-level" if two else "single   ", "ms %.3f" % ((time.perf_counter() - t0) * 100), ev.stats["mode"],
              "sample2" in ev.stats)
