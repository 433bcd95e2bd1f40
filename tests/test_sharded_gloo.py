"""world_size-2 and -3 (gloo, CPU) runs of the REAL multi-rank orchestration (evaluator + DistComm) over the numpy
backend emulation: a row-sharded gallery must give the 1-rank / oracle answers exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from concepthash_b200 import synth
from concepthash_b200.evaluator import DistComm, Evaluator
from oracle import map_oracle as mo
from tests._emu_backend import EmuBackend


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out, extras=True):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        results = {}
        for case, (R, PRs, rf, thr, zero) in enumerate([(-1, [1, 5, 10], False, 0.0, False),
                                                         (15, [1, 5], False, 0.0, False),
                                                         ([4, 30, -1], [3], False, 0.0, True),
                                                         (9, [2], True, 0.0, False),
                                                         (-1, [2], True, 0.3, False)]):
            d, dl, q, ql, _ = synth.make_random_case(12, 230, 16, 4, p=0.3, seed=20 + case)
            if rf:
                q, ql = d[:12].clone(), dl[:12].clone()
            if zero:
                d[::7, 3] = 0
            # uneven contiguous split in rank order
            cut = [0, 97, 230] if world == 2 else np.linspace(0, 230, world + 1).astype(int).tolist()
            ds, dls = d[cut[rank]:cut[rank + 1]], dl[cut[rank]:cut[rank + 1]]
            # odd cases run the candidate-list path (tensor-core select pass) where it applies
            be = EmuBackend(rows_per_stripe=32, threads=128, tensor_cores=True) if case % 2 else \
                EmuBackend(rows_per_stripe=32)
            ev = Evaluator(be, DistComm())
            r_list = R if isinstance(R, list) else [R]
            maps, rec, prec = ev.evaluate(ds, dls, q, ql, r_list, thr, PRs, rf)
            ids, keys, tern = ev.retrieve(ds, q, 20, thr, rf)
            results[case] = (maps, rec, prec, ids.clone(), keys.clone(), tern)
            if case == 1 and extras:
                # sampled top-R with two-level thresholds across ranks
                ev2 = Evaluator(EmuBackend(rows_per_stripe=32, threads=128, tensor_cores=True), DistComm())
                ev2.sample_stride, ev2.sample_min_rows, ev2.sample_min_ratio = 2, 0, 4
                ev2.sample2_min_rows, ev2.sample2_min_work, ev2.sample2_sub = 0, 0, 2
                results["s2"] = ev2.evaluate(ds, dls, q, ql, [15], thr, PRs, rf) + (ev2.stats["mode"],
                                                                                     "sample2" in ev2.stats)
                # the same call again: both ranks hold a hint -> speculative run (one agreement round trip at the
                # idle start + the final read), identical numbers; then a smaller slot limit on ONE rank's shard
                # only must still make BOTH ranks chunk the query set together
                again = ev2.evaluate(ds, dls, q, ql, [15], thr, PRs, rf)
                results["spec"] = (again, ev2.stats["speculation"], ev2.stats["host_syncs"])
                ev3 = Evaluator(EmuBackend(rows_per_stripe=32, threads=128, tensor_cores=True), DistComm())
                ev3.sample_stride, ev3.sample_min_rows, ev3.sample_min_ratio = 2, 0, 4
                ev3.speculate = False
                ev3.evaluate(ds, dls, q, ql, [15], thr, PRs, rf)
                mine = torch.tensor([ev3.stats["record_slots"]], dtype=torch.int64)
                both = [torch.zeros_like(mine) for _ in range(world)]
                dist.all_gather(both, mine)
                lo, hi = min(int(t) for t in both), max(int(t) for t in both)
                assert lo < hi                       # uneven shards: only ONE rank's slices exceed the limit below
                ev3.max_slots = hi                   # (the same constant on every rank, as in the product)
                results["chunk"] = (ev3.evaluate(ds, dls, q, ql, [15], thr, PRs, rf), ev3.stats.get("query_chunks", 0))
                # a list slice overflows on ONE rank only (rank 1, two queries): both ranks must agree on the marked
                # queries and re-rank exactly those together
                ev4 = Evaluator(EmuBackend(rows_per_stripe=32, threads=128, tensor_cores=True), DistComm())
                ev4.sample_stride, ev4.sample_min_rows, ev4.sample_min_ratio, ev4.sample_two_level = 2, 0, 4, False
                if rank == 1:
                    orig = ev4.b._b.record_caps

                    def starved(source, a0, a1, nstripes, nb, nq, nq_pad, mwp, cap, sample_stride=0, replicate=False):
                        orig(source, a0, a1, nstripes, nb, nq, nq_pad, mwp, cap, sample_stride=sample_stride,
                             replicate=replicate)
                        if sample_stride > 1:
                            cap[:, [2, 9]] = 0
                    ev4.b._b.record_caps = starved
                results["repair"] = (ev4.evaluate(ds, dls, q, ql, [15], thr, PRs, rf), ev4.stats["mode"],
                                     ev4.stats["sample"].get("repaired_queries"))
                # zero_mean_eval: the column mean is that of the WHOLE gallery (sums all-reduced over the ranks)
                results["zm"] = ev.evaluate(ds + 0.3, dls, q + 0.3, ql, [15], 0.0, [1, 5], False, zero_mean=True)
        if rank == 0:
            torch.save(results, out)
    finally:
        dist.destroy_process_group()


def _check_base_cases(results, extras):
    for case, (R, PRs, rf, thr, zero) in enumerate([(-1, [1, 5, 10], False, 0.0, False),
                                                     (15, [1, 5], False, 0.0, False),
                                                     ([4, 30, -1], [3], False, 0.0, True),
                                                     (9, [2], True, 0.0, False),
                                                     (-1, [2], True, 0.3, False)]):
        d, dl, q, ql, _ = synth.make_random_case(12, 230, 16, 4, p=0.3, seed=20 + case)
        if rf:
            q, ql = d[:12].clone(), dl[:12].clone()
        if zero:
            d[::7, 3] = 0
        om, orec, oprec = mo.calculate_mAP(d, dl, q, ql, R, threshold=thr, PRs=PRs, remove_first_retrieved=rf)
        om = om if isinstance(om, list) else [om]
        maps, rec, prec, ids, keys, tern = results[case]
        assert np.allclose(maps, om, atol=1e-12), case
        assert np.allclose(rec, orec, atol=1e-12), case
        assert np.allclose(prec, oprec, atol=1e-12), case
        oids, odist = mo.topk_ids(q, d, 20, threshold=thr, remove_first_retrieved=rf)
        assert torch.equal(ids, oids), case
        assert torch.equal(keys.float() * (0.5 if tern else 1.0), odist), case
        if case == 1 and extras:
            s2 = results["s2"]
            assert s2[4] and s2[3] in ("topR-sampled", "topR"), s2
            assert np.allclose(s2[0], om, atol=1e-12) and np.allclose(s2[1], orec, atol=1e-12)
            again, spec, syncs = results["spec"]
            assert spec == "hit" and syncs == 2, (spec, syncs)
            assert np.allclose(again[0], om, atol=1e-12) and np.allclose(again[2], oprec, atol=1e-12)
            chunked, nchunks = results["chunk"]
            assert nchunks >= 2, nchunks
            assert np.allclose(chunked[0], om, atol=1e-12) and np.allclose(chunked[1], orec, atol=1e-12)
            rep, mode, nrep = results["repair"]
            assert mode == "topR-sampled" and nrep == 2, (mode, nrep)
            assert np.allclose(rep[0], om, atol=1e-12) and np.allclose(rep[1], orec, atol=1e-12)
            dz, qz = mo.zero_mean(d + 0.3, q + 0.3)
            om, orec, oprec = mo.calculate_mAP(dz, dl, qz, ql, 15, PRs=[1, 5])
            zm = results["zm"]
            assert np.allclose(zm[0], [om], atol=1e-12) and np.allclose(zm[1], orec, atol=1e-12)
            assert np.allclose(zm[2], oprec, atol=1e-12)


@pytest.mark.timeout(300)
def test_two_ranks_match_oracle(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    _check_base_cases(torch.load(out), True)


@pytest.mark.timeout(300)
def test_three_ranks_match_oracle(tmp_path):
    """Three shards: ranks >= 2 add the totals of SEVERAL lower ranks to their stable-tie prefixes (the two-rank run
    only ever adds one), and the candidate totals are gathered from more than one peer."""
    out = str(tmp_path / "res3.pt")
    mp.spawn(_worker, args=(3, _free_port(), out, False), nprocs=3, join=True)
    _check_base_cases(torch.load(out), False)


@pytest.mark.timeout(300)
def test_random_shardings_match_oracle():
    """A short run of tests/_shard_sweep.py: random cuts over three ranks (empty shards included), random R / PRs /
    remove_first / thresholds / path knobs, each checked against the oracle over the whole gallery."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tests", "_shard_sweep.py"), "3", "0", "24"], cwd=root,
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "bad 0 of 24" in res.stdout, res.stdout[-2000:]
