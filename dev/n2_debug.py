"""Debug helper (dev only): cfg5-like weak-scaled run under torchrun, printing the status words of every attempt."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from concepthash_b200 import hashing, synth  # noqa: E402
from concepthash_b200 import evaluator as E  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
nq, ndb = int(sys.argv[1]), int(sys.argv[2])
d, dl, q, ql, _ = synth.make_random_case(nq, ndb, 64, 1000, p=0.30, seed=0, device=dev, shard=rank)
ev = hashing.get_evaluator(dev, dist.group.WORLD)
ev.speculate = len(sys.argv) < 4 or sys.argv[3] != "nospec"
orig = ev._finish


def finish(c, st):
    res = orig(c, st)
    if "cand" in st:
        cand = st["cand"]
        over = (cand["cnt"][:, :nq].to(torch.int64) >= cand["cap"][:, :nq].to(torch.int64)).sum().item()
        print(f"[rank {rank}] attempt mode={ev.stats.get('mode')} flags={res[4][:8]} slices at cap={over} "
              f"cnt max={int(cand['cnt'][:, :nq].max())} cap min={int(cand['cap'][:, :nq].min())} "
              f"sample={ev.stats.get('sample')} s2={ev.stats.get('sample2')}", flush=True)
    else:
        print(f"[rank {rank}] attempt mode={ev.stats.get('mode')} flags={res[4][:8]} (records)", flush=True)
    return res


ev._finish = finish
for it in range(3):
    m = ev.evaluate(d, dl, q, ql, [1000], 0.0, [], False)
    print(f"[rank {rank}] it={it} mAP={m[0][0]:.6f} mode={ev.stats['mode']} spec={ev.stats.get('speculation')} "
          f"syncs={ev.stats.get('host_syncs')}", flush=True)
dist.destroy_process_group()
