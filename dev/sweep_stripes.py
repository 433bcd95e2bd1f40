"""Dev: time of the full-ranking POPC pass (hist_count_rec) and of the whole step against the stripe length."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from concepthash_b200 import hashing, synth

ev = hashing.get_evaluator()
for name, nbit in (("cub200", 64), ("cars196", 64), ("cars196", 16), ("nabirds", 64)):
    d, dl, q, ql, ncls = synth.make_dataset_case(name, nbit=nbit, p=0.15, seed=0, device="cuda")
    ndb = d.shape[0]
    for rows in (None, 256, 512, 768, 1024, 1536, 2048, 3072, 4096, 6144, 8192, 12288):
        if rows is not None and rows > ndb * 1.3:
            continue
        ev.stripe_rows_override = rows
        ev._hints.clear()
        for _ in range(4):
            ev.evaluate(d, dl, q, ql, [-1], 0.0, [], False)
        torch.cuda.synchronize()
        ev.events, ev.profile = [], True
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ev.evaluate(d, dl, q, ql, [-1], 0.0, [], False)
        e1.record()
        torch.cuda.synchronize()
        ev.profile = False
        hist = sum(a.elapsed_time(b) for k, u, a, b in ev.events if k == "hist_count_rec") / 10
        print(f"{name}-{nbit} rows/stripe={rows} geo={ev.stats['geometry']} hist={hist*1e3:.1f} us step={e0.elapsed_time(e1)/10*1e3:.1f} us", flush=True)
    ev.stripe_rows_override = None
