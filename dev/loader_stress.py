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

import faulthandler

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
ev = hashing.get_evaluator()
bad = 0
t0 = time.time()
# ---- the loader alone with a ring of 4 slots, many times (slots refilled all the time; a thread waiting for a slot
# depends on thread 0 sending chunks to the very end)
b = ev.b._b
side = torch.cuda.Stream(priority=-1)
x = torch.randn(600_000, 64)
x[x == 0] = 1.0
small = x[:3000]
bits = torch.empty((b.padded_rows(x.shape[0]), 2), dtype=torch.int32, device="cuda")
bits0 = torch.empty((b.padded_rows(3000), 2), dtype=torch.int32, device="cuda")
flags = torch.zeros(1, dtype=torch.int32, device="cuda")
ref_bits, _ = b.pack_sign(x.cuda(), 0.0, flags.clone(), False)
os.environ["CH_LOADER_RING_BYTES"] = str(1 << 20)
faulthandler.dump_traceback_later(120, exit=True)
for rep in range(400):
    ld = b.host_loader_start([(small, bits0, flags), (x, bits, flags)], side)
    ld.wait(1, x.shape[0], torch.cuda.current_stream())
    if rep % 50 == 0 and not torch.equal(bits, ref_bits):
        bad += 1
        print("LOADER MISMATCH", rep)
    ld.join()
faulthandler.cancel_dump_traceback_later()
os.environ.pop("CH_LOADER_RING_BYTES", None)
print("loader alone, ring of 4 slots: 400 runs ok, %.1f s" % (time.time() - t0), flush=True)
del x, small, bits, bits0, ref_bits
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
    print("case", it, nbit, ndb, nq, ncls, R, kind, os.environ.get("CH_LOADER_RING_BYTES"), flush=True)
    faulthandler.dump_traceback_later(40, exit=True)             # a hang: where is this thread?
    try:
        for rep in range(3):                                     # no hint, hint, hint again
            out = ev.evaluate(hd, dl.cpu(), hq, ql.cpu(), R, 0.0, [1, 10], False, return_ap=True)
            # (1e-12: a query repaired by the exact path in one run and not in the other sums its AP in another order)
            same = (np.allclose(out[0], ref[0], rtol=0, atol=1e-12) and np.allclose(out[1], ref[1], rtol=0, atol=1e-12)
                    and np.allclose(out[2], ref[2], rtol=0, atol=1e-12)
                    and torch.allclose(out[3], ref[3], rtol=0, atol=1e-12))
            if not same or not ev.stats["mode"].endswith("streamed"):
                bad += 1
                print("MISMATCH" if not same else "NOT STREAMED", it, rep, nbit, ndb, nq, ncls, R, kind, ev.stats["mode"],
                      out[0], ref[0], float((out[3] - ref[3]).abs().max()))
    finally:
        faulthandler.cancel_dump_traceback_later()
        os.environ.pop("CH_LOADER_RING_BYTES", None)
    del d, dl, q, ql, hd, hq
print("iterations", iters, "bad", bad, "seconds %.1f" % (time.time() - t0))
sys.exit(1 if bad else 0)
