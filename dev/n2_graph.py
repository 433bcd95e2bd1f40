"""Debug helper (dev only): graph replay in group mode under torchrun: python -m torch.distributed.run ... dev/n2_graph.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from concepthash_b200 import hashing, synth  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ev = hashing.get_evaluator(dev, dist.group.WORLD)
for nq, ndb, nbit, ncls, R in [(5794, 3000, 64, 200, -1)]:
    for it in range(6):
        d, dl, q, ql, _ = synth.make_random_case(nq, ndb, nbit, ncls, p=0.30, seed=it, device=dev, shard=rank)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        m = ev.evaluate(d, dl, q, ql, [R], 0.0, [], False)
        b.record()
        torch.cuda.synchronize()
        ev.use_graphs = False
        m2 = ev.evaluate(d, dl, q, ql, [R], 0.0, [], False)
        ev.use_graphs = True
        print(f"[rank {rank}] {nq}x{ndb} it={it} mAP={m[0][0]:.9f} same={m == m2} spec={ev.stats.get('speculation')} "
              f"ms={a.elapsed_time(b):.3f} err={getattr(ev, 'stats_graph_error', None)}", flush=True)
if os.environ.get("CLEAN", "1") == "1":
    ev._graphs.clear()
    import gc
    gc.collect()
    torch.cuda.synchronize()
print(f"[rank {rank}] before destroy", flush=True)
dist.destroy_process_group()
print(f"[rank {rank}] after destroy", flush=True)
