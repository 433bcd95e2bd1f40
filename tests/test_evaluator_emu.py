"""The real orchestration (concepthash_b200.evaluator) over the numpy emulation of the backend must
reproduce the oracle: validates the sort-free rank algebra (stripes, bases, thresholds, records,
remove_first, R lists, P@k / R@k, multi-hot, ternary) on CPU."""
import os

import numpy as np
import pytest
import torch

from concepthash_b200 import synth
from concepthash_b200.evaluator import Evaluator
from oracle import map_oracle as mo
from tests._emu_backend import EmuBackend


def run_case(d, dl, q, ql, R, PRs=(), rf=False, thr=0.0, rps=64, sampled=False, tc=False, two_level=False):
    # tc: the candidate-list path of the tensor-core select pass (whole 128-query tiles)
    ev = Evaluator(EmuBackend(rows_per_stripe=rps, threads=128, tensor_cores=True) if tc
                   else EmuBackend(rows_per_stripe=rps))
    ev.sample_two_level = two_level
    if two_level:
        ev.sample2_min_rows, ev.sample2_min_work, ev.sample2_sub = 0, 0, 4
    if sampled:
        ev.sample_stride, ev.sample_min_rows, ev.sample_min_ratio = 4, 0, 4
    else:
        ev.sample_stride = 0
    r_list = R if isinstance(R, list) else [R]
    maps, rec, prec, ap = ev.evaluate(d, dl, q, ql, r_list, thr, list(PRs), rf, return_ap=True)
    om, orec, oprec, oaps = mo.calculate_mAP(d, dl, q, ql, R, threshold=thr, PRs=list(PRs),
                                             remove_first_retrieved=rf, return_per_query=True)
    om = om if isinstance(om, list) else [om]
    assert np.allclose(ap.numpy(), oaps, atol=1e-12), ev.stats
    assert np.allclose(maps, om, atol=1e-12)
    assert np.allclose(rec, orec, atol=1e-12)
    assert np.allclose(prec, oprec, atol=1e-12)
    return ev


def test_golden_cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "oracle_cases.npz"))
    modes = set()
    for name in z["names"]:
        R = z[f"{name}_R"].tolist()
        R = R if bool(z[f"{name}_R_is_list"]) else R[0]
        ev = run_case(torch.from_numpy(z[f"{name}_db_codes"]), torch.from_numpy(z[f"{name}_db_labels"]),
                      torch.from_numpy(z[f"{name}_q_codes"]), torch.from_numpy(z[f"{name}_q_labels"]),
                      R, z[f"{name}_PRs"].tolist(), bool(z[f"{name}_rf"]), float(z[f"{name}_thr"]))
        modes.add((ev.stats["mode"], ev.stats["ternary"]))
    assert ("all", False) in modes and ("topR", False) in modes and ("topR", True) in modes


@pytest.mark.parametrize("seed", range(6))
def test_random_heavy_ties(seed):
    g = torch.Generator().manual_seed(seed)
    nbit = [8, 16, 33, 64][seed % 4]
    nq, ndb, ncls = 17, 150 + 37 * seed, 4
    d, dl, q, ql, _ = synth.make_random_case(nq, ndb, nbit, ncls, p=0.3, seed=seed)
    onehot = seed % 2 == 0
    if onehot:
        dl, ql = synth.one_hot(dl, ncls), synth.one_hot(ql, ncls)
    R = [-1, 5, 20, ndb + 10][seed % 4]
    run_case(d, dl, q, ql, R, PRs=[1, 5, 10], rf=False, rps=32)
    run_case(d, dl, q, ql, [3, 11, -1], PRs=[2], rf=False, rps=48)


def test_remove_first_topr_and_all():
    d, dl, _, _, _ = synth.make_random_case(5, 120, 16, 3, p=0.3, seed=11)
    run_case(d, dl, d.clone(), dl.clone(), -1, PRs=[1, 5], rf=True, rps=32)
    run_case(d, dl, d.clone(), dl.clone(), 7, PRs=[1, 5], rf=True, rps=32)
    run_case(d, dl, d.clone(), dl.clone(), [2, 9], rf=True, rps=32)


def test_multi_hot_labels():
    g = torch.Generator().manual_seed(3)
    d, _, q, _, _ = synth.make_random_case(13, 140, 32, 5, p=0.3, seed=5)
    dl = (torch.rand(140, 40, generator=g) < 0.06).float()      # 40 classes -> 2 mask words
    ql = (torch.rand(13, 40, generator=g) < 0.08).float()
    ev = run_case(d, dl, q, ql, -1, PRs=[1, 10], rps=32)
    assert ev.stats["label_mode"] == 2
    run_case(d, dl, q, ql, 9, PRs=[1, 10], rps=32)


def test_ternary_zero_codes_and_threshold():
    g = torch.Generator().manual_seed(4)
    d, dl, q, ql, _ = synth.make_random_case(11, 130, 24, 4, p=0.3, seed=6)
    d[torch.rand(d.shape, generator=g) < 0.1] = 0
    q[torch.rand(q.shape, generator=g) < 0.1] = 0
    ev = run_case(d, dl, q, ql, -1, PRs=[3], rps=32)
    assert ev.stats["ternary"]
    run_case(d, dl, q, ql, 10, PRs=[3], thr=0.4, rps=32)


def test_all_identical_codes_single_bucket():
    d = torch.ones(90, 8)
    q = torch.ones(6, 8)
    dl = torch.arange(90) % 3
    ql = torch.arange(6) % 3
    run_case(d, dl, q, ql, -1, PRs=[1, 4], rps=32)
    run_case(d, dl, q, ql, 10, PRs=[1, 4], rps=32)


def test_queries_without_labels_and_unseen_classes():
    d, dl, q, ql, _ = synth.make_random_case(9, 100, 16, 4, p=0.3, seed=8)
    dl2, ql2 = synth.one_hot(dl, 6), synth.one_hot(ql, 6)
    ql2[0] = 0            # no positive class at all
    ql2[1] = 0
    ql2[1, 5] = 1         # class never seen in the gallery
    run_case(d, dl2, q, ql2, -1, PRs=[1, 5], rps=32)
    run_case(d, dl2, q, ql2, 12, PRs=[1, 5], rps=32)


def test_retrieve_matches_stable_sort():
    d, _, q, _, _ = synth.make_random_case(14, 170, 16, 4, p=0.3, seed=9)
    for R, rf in [(25, False), (-1, False), (400, False), (10, True)]:
        ev = Evaluator(EmuBackend(rows_per_stripe=32))
        ids, keys, tern = ev.retrieve(d, q, R, 0.0, rf)
        oids, odist = mo.topk_ids(q, d, R, remove_first_retrieved=rf)
        assert torch.equal(ids, oids)
        assert torch.equal(keys.float(), odist)


def test_errors():
    ev = Evaluator(EmuBackend())
    d, dl, q, ql, _ = synth.make_random_case(4, 40, 16, 3, seed=1)
    with pytest.raises(ValueError):
        ev.evaluate(d, dl, q[:, :8], ql, [-1])
    with pytest.raises(ValueError):
        ev.evaluate(d, dl[:10], q, ql, [-1])
    qn = q.clone()
    qn[0, 0] = float("nan")
    with pytest.raises(ValueError):
        ev.evaluate(d, dl, qn, ql, [-1])
    with pytest.raises(ValueError):
        ev.evaluate(d, dl, q, ql, [0])


@pytest.mark.parametrize("seed", range(4))
def test_sampled_single_pass_topr(seed):
    """Threshold from a 1-in-4 row sample + one full pass: exact whenever it verifies, else falls back."""
    nbit = [16, 32, 64, 24][seed]
    d, dl, q, ql, _ = synth.make_random_case(9, 900 + 50 * seed, nbit, 5, p=0.3, seed=40 + seed)
    if seed == 3:
        d[::9, 1] = 0                                     # ternary keys
    ev = run_case(d, dl, q, ql, 20, PRs=[1, 5, 10], rps=64, sampled=True)
    assert ev.stats["mode"] == "topR-sampled", ev.stats
    ev = run_case(d, dl, q, ql, [5, 40], PRs=[], rf=(seed == 1), rps=128, sampled=True)
    assert ev.stats["mode"] in ("topR-sampled", "topR")


def test_sampled_falls_back_on_adversarial_order():
    """35 near rows sit exactly on the sampled positions (every 4th row), everything else is far: the sample
    over-represents them 4x, the sample threshold is too low, and the verification must catch it and fall
    back to the exact two-pass path."""
    nbit = 16
    q = torch.ones(3, nbit)
    d = -torch.ones(800, nbit)                            # distance 16
    d[0:140:4] = 1.0
    d[0:140:4, 0] = -1.0                                  # 35 rows at distance 1, all on sampled positions
    dl = torch.arange(800) % 2
    ql = torch.tensor([0, 1, 0])
    ev = run_case(d, dl, q, ql, 40, PRs=[1, 5], rps=64, sampled=True)
    assert ev.stats["mode"] == "topR" and ev.stats["sample"]["fallback"]
    # benign order: verifies and stays in the one-pass mode
    ev = run_case(d, dl, q, ql, 30, PRs=[1, 5], rps=64, sampled=True)
    assert ev.stats["mode"] == "topR-sampled"


# ---------------------------------------------------------------- candidate-list path (tensor-core select pass)
@pytest.mark.parametrize("seed", range(4))
def test_candidate_path_exact_two_pass(seed):
    nbit = [16, 32, 64, 128][seed]
    d, dl, q, ql, _ = synth.make_random_case(11, 300 + 31 * seed, nbit, 4, p=0.3, seed=60 + seed)
    ev = run_case(d, dl, q, ql, 20, PRs=[1, 5, 10], rps=64, tc=True)
    assert ev.stats["mode"] == "topR" and ev.stats["select_kernel"] == "tcgen05"
    run_case(d, dl, q, ql, [3, 11, 40], PRs=[2], rps=48, tc=True)
    ev = run_case(d, dl, d[:9].clone(), dl[:9].clone(), 7, PRs=[1, 5], rf=True, rps=32, tc=True)
    assert ev.stats["select_kernel"] == "tcgen05"


def test_candidate_path_multi_hot():
    g = torch.Generator().manual_seed(3)
    d, _, q, _, _ = synth.make_random_case(13, 240, 32, 5, p=0.3, seed=5)
    dl = (torch.rand(240, 40, generator=g) < 0.06).float()
    ql = (torch.rand(13, 40, generator=g) < 0.08).float()
    ev = run_case(d, dl, q, ql, 9, PRs=[1, 10], rps=32, tc=True)
    assert ev.stats["label_mode"] == 2 and ev.stats["select_kernel"] == "tcgen05"


@pytest.mark.parametrize("seed", range(3))
def test_candidate_path_sampled(seed):
    nbit = [16, 64, 128][seed]
    d, dl, q, ql, _ = synth.make_random_case(9, 900 + 50 * seed, nbit, 5, p=0.3, seed=40 + seed)
    ev = run_case(d, dl, q, ql, 20, PRs=[1, 5, 10], rps=64, sampled=True, tc=True)
    assert ev.stats["mode"] == "topR-sampled" and ev.stats["select_kernel"] == "tcgen05", ev.stats
    ev = run_case(d, dl, q, ql, [5, 40], PRs=[], rf=(seed == 1), rps=128, sampled=True, tc=True)
    assert ev.stats["mode"] in ("topR-sampled", "topR")


def test_candidate_path_sampled_falls_back():
    nbit = 16
    q = torch.ones(3, nbit)
    d = -torch.ones(800, nbit)
    d[0:140:4] = 1.0
    d[0:140:4, 0] = -1.0
    dl = torch.arange(800) % 2
    ql = torch.tensor([0, 1, 0])
    ev = run_case(d, dl, q, ql, 40, PRs=[1, 5], rps=64, sampled=True, tc=True)
    # the sample is defeated for every query: they are re-ranked by the exact path (few queries: per-query repair)
    assert ev.stats["sample"].get("repaired_queries") == 3 or ev.stats["sample"].get("fallback"), ev.stats


def test_candidate_path_retrieve():
    d, _, q, _, _ = synth.make_random_case(14, 170, 16, 4, p=0.3, seed=9)
    for R, rf in [(25, False), (-1, False), (400, False), (10, True)]:
        ev = Evaluator(EmuBackend(rows_per_stripe=32, threads=128, tensor_cores=True))
        ids, keys, tern = ev.retrieve(d, q, R, 0.0, rf)
        assert ev.stats["select_kernel"] == "tcgen05"
        oids, odist = mo.topk_ids(q, d, R, remove_first_retrieved=rf)
        assert torch.equal(ids, oids)
        assert torch.equal(keys.float(), odist)


# ---------------------------------------------------------------- zero_mean_eval fused into the pack step
@pytest.mark.parametrize("tc", [False, True])
def test_zero_mean_eval_matches_caller_side_subtraction(tc):
    d, dl, q, ql, _ = synth.make_random_case(11, 400, 32, 4, p=0.3, seed=71)
    d, q = d + 0.35, q + 0.35                                   # a real offset: the mean matters
    be = EmuBackend(rows_per_stripe=64, threads=128, tensor_cores=True) if tc else EmuBackend(rows_per_stripe=64)
    ev = Evaluator(be)
    ev.sample_stride = 0
    dz, qz = mo.zero_mean(d, q)
    for R, thr in [(-1, 0.0), (25, 0.0), (25, 0.2)]:
        maps, rec, prec = ev.evaluate(d, dl, q, ql, [R], thr, [1, 5], False, zero_mean=True)
        om, orec, oprec = mo.calculate_mAP(dz, dl, qz, ql, R, threshold=thr, PRs=[1, 5])
        assert np.allclose(maps, [om], atol=1e-12) and np.allclose(rec, orec, atol=1e-12)
        assert np.allclose(prec, oprec, atol=1e-12)
        # and it is not a no-op on this data
        m0, _, _ = ev.evaluate(d, dl, q, ql, [R], thr, [], False)
        assert abs(m0[0] - maps[0]) > 1e-6
    ids, keys, _ = ev.retrieve(d, q, 30, 0.0, False, zero_mean=True)
    oids, odist = mo.topk_ids(qz, dz, 30)
    assert torch.equal(ids, oids) and torch.equal(keys.float(), odist)


@pytest.mark.parametrize("seed", range(3))
def test_candidate_path_two_level_sample(seed):
    """thresholds from the sample by a select pass over the sample itself (level 0: every 4th sample row)"""
    nbit = [16, 64, 128][seed]
    d, dl, q, ql, _ = synth.make_random_case(9, 1800 + 50 * seed, nbit, 5, p=0.3, seed=80 + seed)
    ev = run_case(d, dl, q, ql, 20, PRs=[1, 5, 10], rps=64, sampled=True, tc=True, two_level=True)
    assert ev.stats["mode"] == "topR-sampled" and "sample2" in ev.stats, ev.stats
    ev = run_case(d, dl, q, ql, [5, 40], PRs=[], rf=(seed == 1), rps=128, sampled=True, tc=True, two_level=True)
    assert ev.stats["mode"] in ("topR-sampled", "topR")
    # adversarial order: falls back to the exact path, still the oracle's answer
    qq = torch.ones(3, 16)
    dd = -torch.ones(1600, 16)
    dd[0:280:4] = 1.0
    dd[0:280:4, 0] = -1.0
    ev = run_case(dd, torch.arange(1600) % 2, qq, torch.tensor([0, 1, 0]), 80, PRs=[1, 5], rps=64, sampled=True,
                  tc=True, two_level=True)
    assert ev.stats["sample"].get("repaired_queries") == 3 or ev.stats["sample"].get("fallback"), ev.stats


# ------------------------------------------------------------------ speculative re-evaluation + query chunking
def _spec_ev(tc=True):
    ev = Evaluator(EmuBackend(rows_per_stripe=64, threads=128, tensor_cores=True) if tc else EmuBackend(rows_per_stripe=64))
    ev.sample_stride, ev.sample_min_rows, ev.sample_min_ratio = 4, 0, 4
    ev.sample2_min_rows, ev.sample2_min_work, ev.sample2_sub = 0, 0, 4
    return ev


def _oracle(d, dl, q, ql, R, PRs):
    om, orec, oprec, oaps = mo.calculate_mAP(d, dl, q, ql, R, PRs=PRs, return_per_query=True)
    return (om if isinstance(om, list) else [om]), orec, oprec, oaps


def _check(out, ora):
    assert np.allclose(out[0], ora[0], atol=1e-12) and np.allclose(out[1], ora[1], atol=1e-12)
    assert np.allclose(out[2], ora[2], atol=1e-12) and np.allclose(out[3].numpy(), ora[3], atol=1e-12)


@pytest.mark.parametrize("R", [20, -1])
@pytest.mark.parametrize("tc", [True, False])
def test_second_evaluation_of_a_shape_is_speculative_and_sync_free(R, tc):
    """The first evaluation of a shape asks the device (label form, list sizes, largest threshold); the second one
    assumes them, verifies on the device and needs ONE host round trip -- same numbers."""
    ev = _spec_ev(tc)
    d, dl, q, ql, _ = synth.make_random_case(23, 700, 32, 5, p=0.3, seed=1)
    a = ev.evaluate(d, dl, q, ql, [R], 0.0, [1, 5], False, return_ap=True)
    first = dict(ev.stats)
    b = ev.evaluate(d, dl, q, ql, [R], 0.0, [1, 5], False, return_ap=True)
    assert first["speculation"] == "none" and first["host_syncs"] >= 2
    assert ev.stats["speculation"] == "hit" and ev.stats["host_syncs"] == 1, ev.stats
    assert ev.stats["mode"] == first["mode"]
    ora = _oracle(d, dl, q, ql, R, [1, 5])
    _check(a, ora)
    _check(b, ora)
    # new data of the same shape (another seed): still exact, whether the hint held or the run was repeated
    d2, dl2, q2, ql2, _ = synth.make_random_case(23, 700, 32, 5, p=0.3, seed=2)
    c = ev.evaluate(d2, dl2, q2, ql2, [R], 0.0, [1, 5], False, return_ap=True)
    assert ev.stats["speculation"] in ("hit", "retried")
    _check(c, _oracle(d2, dl2, q2, ql2, R, [1, 5]))


def test_stale_hints_are_detected_and_the_run_repeated():
    ev = _spec_ev()
    d, dl, q, ql, _ = synth.make_random_case(23, 700, 32, 5, p=0.05, seed=3)      # tight clusters: short lists
    ev.evaluate(d, dl, q, ql, [20], 0.0, [1, 5], False)
    # (a) far more candidates under the same shape: p = 0.5 -> distances concentrate, lists grow, thresholds move
    d2, dl2, q2, ql2, _ = synth.make_random_case(23, 700, 32, 5, p=0.5, seed=4)
    out = ev.evaluate(d2, dl2, q2, ql2, [20], 0.0, [1, 5], False, return_ap=True)
    assert ev.stats["speculation"] == "retried", ev.stats
    _check(out, _oracle(d2, dl2, q2, ql2, 20, [1, 5]))
    # (b) exact zeros appear (ternary keys)
    ev.evaluate(d, dl, q, ql, [20], 0.0, [1, 5], False)
    dz = d.clone()
    dz[::9, 2] = 0.0
    out = ev.evaluate(dz, dl, q, ql, [20], 0.0, [1, 5], False, return_ap=True)
    assert ev.stats["speculation"] == "retried" and ev.stats["ternary"], ev.stats
    _check(out, _oracle(dz, dl, q, ql, 20, [1, 5]))
    # (c) the label form changes under the same shapes: one-hot -> multi-hot
    oh_d, oh_q = synth.one_hot(dl, 5), synth.one_hot(ql, 5)
    ev.evaluate(d, oh_d, q, oh_q, [20], 0.0, [1, 5], False)
    mh_d = oh_d.clone()
    mh_d[::4, 0] = 1
    out = ev.evaluate(d, mh_d, q, oh_q, [20], 0.0, [1, 5], False, return_ap=True)
    assert ev.stats["speculation"] == "retried" and ev.stats["label_mode"] == 2, ev.stats
    _check(out, _oracle(d, mh_d, q, oh_q, 20, [1, 5]))
    # (d) NaN under a hint still raises
    ev.evaluate(d, dl, q, ql, [20], 0.0, [1, 5], False)
    dn = d.clone()
    dn[5, 5] = float("nan")
    with pytest.raises(ValueError, match="NaN"):
        ev.evaluate(dn, dl, q, ql, [20], 0.0, [1, 5], False)


@pytest.mark.parametrize("R", [20, -1])
def test_query_chunking_when_slots_exceed_the_offset_range(R):
    """A query set whose slices would not fit the 32-bit slot index is evaluated in query chunks, combined exactly
    (forced here by a tiny slot limit)."""
    ev = _spec_ev()
    d, dl, q, ql, _ = synth.make_random_case(37, 700, 32, 5, p=0.3, seed=6)
    ref = ev.evaluate(d, dl, q, ql, [R], 0.0, [1, 5], False, return_ap=True)
    slots = ev.stats["record_slots"]
    ev2 = _spec_ev()
    ev2.max_slots = max(64, slots // 3)
    out = ev2.evaluate(d, dl, q, ql, [R], 0.0, [1, 5], False, return_ap=True)
    assert ev2.stats.get("query_chunks", 0) >= 2, ev2.stats
    assert out[3].shape == ref[3].shape
    _check(out, _oracle(d, dl, q, ql, R, [1, 5]))
    # and again (the chunks now run speculatively)
    out = ev2.evaluate(d, dl, q, ql, [R], 0.0, [1, 5], False, return_ap=True)
    _check(out, _oracle(d, dl, q, ql, R, [1, 5]))


@pytest.mark.parametrize("mode", ["exact", "sampled", "two_level"])
def test_candidate_path_ternary_codes(mode):
    """ternary codes (threshold / exact zeros) through the candidate-list path: keys on the doubled scale"""
    d, dl, q, ql, _ = synth.make_random_case(21, 500, 24, 4, p=0.3, seed=31)
    d[::5, 3] = 0.0
    for thr in (0.0, 0.4):
        ev = run_case(d, dl, q, ql, [7, 60], PRs=[1, 5], thr=thr, tc=True, sampled=mode != "exact",
                      two_level=mode == "two_level")
        assert ev.stats["ternary"] and ev.stats.get("select_kernel") == "tcgen05", ev.stats
    ev = Evaluator(EmuBackend(rows_per_stripe=64, threads=128, tensor_cores=True))
    ids, keys, tern = ev.retrieve(d, q, 40, 0.4)
    oids, odist = mo.topk_ids(q, d, 40, threshold=0.4)
    assert tern and torch.equal(ids, oids) and torch.equal(keys.float() * 0.5, odist)


@pytest.mark.parametrize("R,rf", [(20, False), (35, True)])
def test_failed_lists_are_repaired_per_query(R, rf):
    """A list slice that overflows its capacity (forced here for 3 of 150 queries) or a list shorter than R marks the
    QUERY; only the marked queries are re-ranked by the exact path and their per-query sums replaced -- the other
    147 keep the result of the one-pass path.  Same numbers as the oracle, AP and P@k / R@k alike."""
    d, dl, q, ql, _ = synth.make_random_case(150, 900, 32, 5, p=0.3, seed=41)
    if rf:
        q, ql = d[:150].clone(), dl[:150].clone()
    ev = Evaluator(EmuBackend(rows_per_stripe=64, threads=128, tensor_cores=True))
    ev.sample_stride, ev.sample_min_rows, ev.sample_min_ratio, ev.sample_two_level = 4, 0, 4, False
    victims = [7, 64, 149]
    orig = ev.b._b.record_caps

    def starved(source, a0, a1, nstripes, nb, nq, nq_pad, min_with_prev, cap, sample_stride=0, replicate=False):
        orig(source, a0, a1, nstripes, nb, nq, nq_pad, min_with_prev, cap, sample_stride=sample_stride, replicate=replicate)
        if sample_stride > 1:
            cap[:, victims] = 1          # far too small: these queries' slices overflow in the full pass
    ev.b._b.record_caps = starved
    out = ev.evaluate(d, dl, q, ql, [R], 0.0, [1, 5], rf, return_ap=True)
    assert ev.stats["mode"] == "topR-sampled" and ev.stats["sample"].get("repaired_queries") == 3, ev.stats
    om, orec, oprec, oaps = mo.calculate_mAP(d, dl, q, ql, R, PRs=[1, 5], remove_first_retrieved=rf, return_per_query=True)
    assert np.allclose(out[3].numpy(), oaps, atol=1e-12) and np.allclose(out[0], [om], atol=1e-12)
    assert np.allclose(out[1], orec, atol=1e-12) and np.allclose(out[2], oprec, atol=1e-12)
    # too many failures -> the whole evaluation is redone exactly
    victims[:] = list(range(0, 150, 2))
    out = ev.evaluate(d, dl, q, ql, [R], 0.0, [1, 5], rf, return_ap=True)
    assert ev.stats["mode"] == "topR" and ev.stats["sample"].get("fallback"), ev.stats
    assert np.allclose(out[3].numpy(), oaps, atol=1e-12)


@pytest.mark.parametrize("seed", range(24))
def test_random_sweep_of_paths_and_shapes(seed):
    """CPU analogue of dev/fuzz_gpu.py: random shapes, code widths, label kinds, R (single / list / all), PRs,
    remove_first, ternary thresholds, stripe sizes and path knobs (exact / sampled / two-level, POPC records /
    candidate lists) -- every combination must give the oracle's per-query AP, means and curves."""
    rng = np.random.default_rng(1000 + seed)
    nq, ndb = int(rng.integers(1, 40)), int(rng.integers(20, 700))
    nbit = int(rng.choice([8, 16, 32, 48, 64, 96, 128]))
    ncls = int(rng.integers(2, 12))
    d, dl, q, ql, _ = synth.make_random_case(nq, ndb, nbit, ncls, p=float(rng.choice([0.0, 0.3, 0.6])),
                                             seed=2000 + seed)
    rf = bool(rng.integers(0, 2)) and nq <= ndb
    if rf:
        q, ql = d[:nq].clone(), dl[:nq].clone()
    if rng.integers(0, 3) == 0:                                   # multi-hot rows
        dl, ql = synth.one_hot(dl, ncls), synth.one_hot(ql, ncls)
        dl[rng.integers(0, ndb, ndb // 5), rng.integers(0, ncls)] = 1
    thr = float(rng.choice([0.0, 0.0, 0.25]))
    top = max(1, ndb - int(rf))
    small = lambda: int(rng.integers(1, max(2, top // 6)))        # short lists: the top-R paths proper
    kind = int(rng.integers(0, 6))
    R = [-1, small(), small(), int(rng.integers(1, top + 1)), [small(), small()],
         [small(), -1, int(rng.integers(1, top + 1))]][kind]
    # P@k / R@k cut-offs lengthen the list a query needs: short ones beside a short R, any beside a long one
    PRs = sorted({small() if kind in (1, 2, 4) else int(rng.integers(1, top + 1))
                  for _ in range(int(rng.integers(0, 4)))})
    sampled = bool(rng.integers(0, 2))
    run_case(d, dl, q, ql, R, PRs, rf, thr, rps=int(rng.choice([32, 64, 256])), sampled=sampled,
             tc=bool(rng.integers(0, 2)), two_level=sampled and bool(rng.integers(0, 2)))


@pytest.mark.parametrize("seed", [1, 3, 19, 30, 174, 183, 5, 8])
def test_repeated_evaluations_with_changing_data(seed):
    """Four evaluations of one shape on ONE evaluator while the data change under it (tie density, zeros appear ->
    ternary keys, labels become multi-hot, one huge tie bucket, everything relevant): buffer sizes and key ranges
    assumed from the previous evaluation must be verified and the run repeated when they fail -- never an
    out-of-range key (the hinted key range of the POPC select pass once let a larger threshold write past its
    histogram: seeds 1, 3, 19, 30, 174, 183 of the sweep this comes from) and never a different number."""
    rng = np.random.default_rng(seed)
    nq, ndb = int(rng.integers(1, 40)), int(rng.integers(40, 700))
    nbit = int(rng.choice([16, 32, 64, 128]))
    ncls = int(rng.integers(2, 12))
    tc = bool(rng.integers(0, 2))
    rps = int(rng.choice([32, 64, 256]))
    ev = Evaluator(EmuBackend(rows_per_stripe=rps, threads=128, tensor_cores=True) if tc
                   else EmuBackend(rows_per_stripe=rps))
    if bool(rng.integers(0, 2)):
        ev.sample_stride, ev.sample_min_rows, ev.sample_min_ratio = 4, 0, 4
    else:
        ev.sample_stride = 0
    small = lambda: int(rng.integers(1, max(2, ndb // 6)))
    R = [-1, small(), small(), [small(), small()]][int(rng.integers(0, 4))]
    PRs = sorted({small() for _ in range(int(rng.integers(0, 3)))})
    for rep in range(4):
        p = float(rng.choice([0.0, 0.3, 0.8]))
        d, dl, q, ql, _ = synth.make_random_case(nq, ndb, nbit, ncls, p=p, seed=seed * 10 + rep)
        kind = int(rng.integers(0, 5))
        if kind == 0:
            d[:: int(rng.integers(2, 9)), int(rng.integers(0, nbit))] = 0
        if kind == 1:
            dl, ql = synth.one_hot(dl, ncls), synth.one_hot(ql, ncls)
            dl[::3, 0] = 1
        if kind == 2:
            d[:] = d[0]
        if kind == 3:
            dl[:] = dl[0]
            ql[:] = dl[0]
        maps, rec, prec, ap = ev.evaluate(d, dl, q, ql, R if isinstance(R, list) else [R], 0.0, PRs, False,
                                          return_ap=True)
        om, orec, oprec, oaps = mo.calculate_mAP(d, dl, q, ql, R, PRs=PRs, return_per_query=True)
        om = om if isinstance(om, list) else [om]
        assert np.allclose(ap.numpy(), oaps, atol=1e-12), (rep, kind, ev.stats)
        assert np.allclose(maps, om, atol=1e-12) and np.allclose(rec, orec, atol=1e-12)
        assert np.allclose(prec, oprec, atol=1e-12)


def test_hinted_key_range_outgrown_by_the_next_evaluation():
    """Tight clusters first (thresholds of 0-2 keys), near-random codes of the same shape next (thresholds ~20): the
    POPC select pass of the second evaluation runs with slabs as narrow as the FIRST one's largest threshold + 3.  The
    kernel clamps, the device check flags it, the evaluation is repeated -- and gives the oracle's numbers."""
    ev = Evaluator(EmuBackend(rows_per_stripe=256))
    ev.sample_stride, ev.sample_min_rows, ev.sample_min_ratio = 4, 0, 4
    seen = []
    for p, seed in ((0.02, 1), (0.02, 1), (0.45, 2)):
        d, dl, q, ql, _ = synth.make_random_case(24, 1500, 64, 5, p=p, seed=seed)
        maps, rec, prec, ap = ev.evaluate(d, dl, q, ql, [10], 0.0, [1, 5], False, return_ap=True)
        om, orec, oprec, oaps = mo.calculate_mAP(d, dl, q, ql, 10, PRs=[1, 5], return_per_query=True)
        assert np.allclose(ap.numpy(), oaps, atol=1e-12) and abs(maps[0] - om) < 1e-12
        assert np.allclose(rec, orec, atol=1e-12) and np.allclose(prec, oprec, atol=1e-12)
        seen.append((ev.stats["mode"], ev.stats["select_kernel"], ev.stats["speculation"],
                     ev.stats["sample"]["key_limit"]))
    assert [s[:3] for s in seen] == [("topR-sampled", "popc", "none"), ("topR-sampled", "popc", "hit"),
                                     ("topR-sampled", "popc", "retried")], seen
    assert seen[2][3] > seen[1][3] + 3, seen          # the key range really was outgrown
