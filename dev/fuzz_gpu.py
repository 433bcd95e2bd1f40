"""Randomised GPU-vs-oracle check of the top-R paths (tensor-core select + candidate ranking, sampled and exact,
two-level sample on/off, ids / one-hot / multi-hot labels, remove_first, R lists, PRs).
usage (on a B200): python dev/fuzz_gpu.py [cases] [seed]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from concepthash_b200 import hashing, synth  # noqa: E402
from oracle import map_oracle as mo  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
ev = hashing.get_evaluator()
saved = (ev.sample_min_rows, ev.sample_min_ratio, ev.sample2_min_rows, ev.sample2_min_work, ev.sample_stride,
         ev.sample_two_level)
bad = 0
modes = {}
for it in range(cases):
    nq = int(rng.integers(1, 600))
    ndb = int(rng.integers(130, 30000))
    nbit = int(rng.choice([8, 16, 24, 31, 32, 33, 48, 64, 96, 100, 127, 128]))
    ncls = int(rng.integers(2, 12))
    d, dl, q, ql, _ = synth.make_random_case(nq, ndb, nbit, ncls, p=float(rng.choice([0.15, 0.3, 0.45])), seed=1000 + it)
    kind = int(rng.integers(0, 3))
    if kind == 1:
        dl, ql = synth.one_hot(dl, ncls), synth.one_hot(ql, ncls)
    elif kind == 2:
        g = torch.Generator().manual_seed(it)
        dl = (torch.rand(ndb, 40, generator=g) < 0.05).float()
        ql = (torch.rand(nq, 40, generator=g) < 0.07).float()
    rf = bool(rng.integers(0, 4) == 0)
    if rf:
        q, ql = d[:nq].clone(), dl[:nq].clone()
        nq = q.shape[0]
    R = [int(rng.integers(1, max(2, ndb // 5)))]
    if rng.integers(0, 3) == 0:
        R.append(int(rng.integers(1, max(2, ndb // 5))))
    PRs = [1, 5, 10] if rng.integers(0, 2) else []
    # force the sampled path (and its two-level form) on small galleries half of the time
    ev.sample_min_rows, ev.sample_min_ratio = (0, 4) if rng.integers(0, 2) else saved[:2]
    ev.sample2_min_rows, ev.sample2_min_work = (0, 0) if rng.integers(0, 2) else saved[2:4]
    ev.sample_stride = int(rng.choice([4, 8, 32]))
    ev.sample_two_level = bool(rng.integers(0, 4) != 0)
    try:
        m, rec, prec = hashing.calculate_mAP(d.cuda(), dl.cuda(), q.cuda(), ql.cuda(), R if len(R) > 1 else R[0],
                                             PRs=PRs, remove_first_retrieved=rf)
        om, orec, oprec = mo.calculate_mAP(d, dl, q, ql, R if len(R) > 1 else R[0], PRs=PRs, remove_first_retrieved=rf)
        ok = np.allclose(m, om, atol=1e-9) and np.allclose(rec, orec, atol=1e-9) and np.allclose(prec, oprec, atol=1e-9)
        k = min(R[0], 50)
        ids, dist = hashing.retrieve_topk(q.cuda(), d.cuda(), k, remove_first_retrieved=rf)
        oids, odist = mo.topk_ids(q, d, k, remove_first_retrieved=rf)
        ok = ok and torch.equal(ids.cpu(), oids) and torch.equal(dist.cpu(), odist)
    except Exception as e:  # noqa: BLE001
        ok = False
        print("EXC", repr(e)[:300])
    key = (ev.stats.get("mode"), ev.stats.get("select_kernel"), "sample2" in ev.stats)
    modes[key] = modes.get(key, 0) + 1
    if not ok:
        bad += 1
        print("MISMATCH case", it, dict(nq=nq, ndb=ndb, nbit=nbit, kind=kind, rf=rf, R=R, PRs=PRs), ev.stats)
(ev.sample_min_rows, ev.sample_min_ratio, ev.sample2_min_rows, ev.sample2_min_work, ev.sample_stride,
 ev.sample_two_level) = saved
print("cases", cases, "bad", bad, "modes", modes)
sys.exit(1 if bad else 0)
