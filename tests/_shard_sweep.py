"""Random row-sharded evaluations over N gloo ranks on CPU (the real Evaluator + DistComm over the numpy backend
emulation) against the oracle over the whole gallery: random cuts (empty shards included), code widths, R / R lists,
PRs, remove_first, ternary thresholds, exact / sampled, records / candidate lists.

    python tests/_shard_sweep.py WORLD SEED_LO SEED_HI      -> "world W bad 0 of K"
"""
import os, sys, socket, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist, torch.multiprocessing as mp
from concepthash_b200 import synth
from concepthash_b200.evaluator import DistComm, Evaluator
from oracle import map_oracle as mo
from tests._emu_backend import EmuBackend

def case(seed, world):
    rng = np.random.default_rng(seed)
    nq, ndb = int(rng.integers(1, 30)), int(rng.integers(30, 500))
    nbit = int(rng.choice([16, 32, 64, 128])); ncls = int(rng.integers(2, 10))
    d, dl, q, ql, _ = synth.make_random_case(nq, ndb, nbit, ncls, p=float(rng.choice([0.0, 0.3])), seed=seed)
    rf = bool(rng.integers(0, 2)) and nq <= ndb
    if rf: q, ql = d[:nq].clone(), dl[:nq].clone()
    thr = float(rng.choice([0.0, 0.0, 0.25]))
    top = max(1, ndb - int(rf))
    small = lambda: int(rng.integers(1, max(2, top // 6)))
    kind = int(rng.integers(0, 4))
    R = [-1, small(), small(), [small(), small()]][kind]
    PRs = sorted({small() for _ in range(int(rng.integers(0, 3)))})
    cuts = np.sort(rng.integers(0, ndb + 1, world - 1)).tolist()
    if rng.integers(0, 4) == 0: cuts[0] = 0                     # an empty first shard
    cut = [0] + cuts + [ndb]
    knobs = dict(tc=bool(rng.integers(0, 2)), sampled=bool(rng.integers(0, 2)), rps=int(rng.choice([32, 64])))
    return d, dl, q, ql, R, PRs, rf, thr, cut, knobs

def worker(rank, world, port, lo, hi):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bad = 0
    for seed in range(lo, hi):
        d, dl, q, ql, R, PRs, rf, thr, cut, kn = case(seed, world)
        ds, dls = d[cut[rank]:cut[rank + 1]], dl[cut[rank]:cut[rank + 1]]
        be = EmuBackend(rows_per_stripe=kn["rps"], threads=128, tensor_cores=True) if kn["tc"] else EmuBackend(rows_per_stripe=kn["rps"])
        ev = Evaluator(be, DistComm())
        if kn["sampled"]: ev.sample_stride, ev.sample_min_rows, ev.sample_min_ratio = 2, 0, 4
        else: ev.sample_stride = 0
        r_list = R if isinstance(R, list) else [R]
        try:
            maps, rec, prec = ev.evaluate(ds, dls, q, ql, r_list, thr, PRs, rf)
            ids, keys, tern = ev.retrieve(ds, q, min(20, max(1, d.shape[0] - int(rf))), thr, rf)
        except Exception as e:
            print("SEED", seed, "rank", rank, "EXC", type(e).__name__, str(e)[:200], "cut", cut, flush=True); traceback.print_exc(limit=4); bad += 1
            raise
        if rank == 0:
            om, orec, oprec = mo.calculate_mAP(d, dl, q, ql, R, threshold=thr, PRs=PRs, remove_first_retrieved=rf)
            om = om if isinstance(om, list) else [om]
            oids, odist = mo.topk_ids(q, d, min(20, max(1, d.shape[0] - int(rf))), threshold=thr, remove_first_retrieved=rf)
            ok = (np.allclose(maps, om, atol=1e-12) and np.allclose(rec, orec, atol=1e-12) and np.allclose(prec, oprec, atol=1e-12)
                  and torch.equal(ids, oids) and torch.equal(keys.float() * (0.5 if tern else 1.0), odist))
            if not ok:
                bad += 1; print("SEED", seed, "MISMATCH", "cut", cut, kn, "R", R, "mode", ev.stats.get("mode"), flush=True)
    if rank == 0: print("world", world, "bad", bad, "of", hi - lo, flush=True)
    dist.destroy_process_group()

if __name__ == "__main__":
    world, lo, hi = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(worker, args=(world, port, lo, hi), nprocs=world, join=True)
