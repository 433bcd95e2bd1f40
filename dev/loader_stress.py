"""Stress of the host-loader path: many evaluations of random shapes from HOST tensors (pageable / pinned / column
slices, a ring smaller than the gallery on some of them), each compared with the device-resident evaluation of the
same data.  python dev/loader_stress.py [iterations] [seed]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from concepthash_b200 import hashing, synth  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
ev = hashing.get_evaluator()
bad = 0
t0 = time.time()
for it in range(iters):
    nbit = int(rng.choice([32, 64, 128, 256]))
    ndb = int(rng.integers(200_000, 700_000))
    nq = int(rng.integers(100, 3000))
    ncls = int(rng.integers(5, 300))
    R = [int(rng.choice([50, 100, 1000]))]
    d, dl, q, ql, _ = synth.make_random_case(nq, ndb, nbit, ncls, p=0.30, seed=1000 + it, device="cuda")
    ref = ev.evaluate(d, dl, q, ql, R, 0.0, [1, 10], False, return_ap=True)
    kind = int(rng.integers(0, 4))
    hd, hq = d.cpu(), q.cpu()
    if kind == 1:
        hd = hd.pin_memory()
    elif kind == 2:
        wide = torch.empty(ndb, nbit + 32)
        wide[:, 16:16 + nbit] = hd
        hd = wide[:, 16:16 + nbit]
    elif kind == 3:
        hq = hq.pin_memory()
    if rng.random() < 0.4:
        os.environ["CH_LOADER_RING_BYTES"] = str(int(rng.choice([1 << 20, 4 << 20])))
    try:
        for rep in range(3):                                     # no hint, hint, hint again
            out = ev.evaluate(hd, dl.cpu(), hq, ql.cpu(), R, 0.0, [1, 10], False, return_ap=True)
            same = (out[0] == ref[0] and out[1] == ref[1] and out[2] == ref[2] and torch.equal(out[3], ref[3]))
            if not same or not ev.stats["mode"].endswith("streamed"):
                bad += 1
                print("MISMATCH" if not same else "NOT STREAMED", it, rep, nbit, ndb, nq, ncls, R, kind, ev.stats["mode"],
                      out[0], ref[0])
    finally:
        os.environ.pop("CH_LOADER_RING_BYTES", None)
    del d, dl, q, ql, hd, hq
print("iterations", iters, "bad", bad, "seconds %.1f" % (time.time() - t0))
sys.exit(1 if bad else 0)
