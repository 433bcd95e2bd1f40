"""The reference's two call sites, issued verbatim against the drop-in module path ``utils.hashing``:

  call A  experiments/train_helper.py:228-234   (periodic evaluation inside training)
  call B  experiments/test_hashing.py:106-119   (stand-alone validation; also calculate_pr_curve, :152-162)

with the inputs their producers really hand over: CPU tensors concatenated from ``.cpu()`` batches
(trainers/base.py:291-304), numpy label arrays (``np.concatenate`` branch, base.py:303), ``F.one_hot`` int64 labels
(test_hashing.py:83-85), an OmegaConf-style list for ``R`` / ``PRs`` (test_hashing.py:124-131), a tensor ``db_id``
(train_helper.py:232-233), ``multiclass=`` and ``landmark_gt=None``.  Results are checked against the CPU oracle.
"""
import collections.abc
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from concepthash_b200 import synth
from oracle import map_oracle as mo

pytestmark = pytest.mark.gpu
TOL = 1e-9


class ListConfig(collections.abc.Sequence):
    """stand-in for omegaconf.ListConfig (not installed here): a Sequence that is neither list nor tuple"""

    def __init__(self, items):
        self._items = list(items)

    def __getitem__(self, i):
        return self._items[i]

    def __len__(self):
        return len(self._items)


@pytest.fixture(scope="module")
def calc():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from utils.hashing import calculate_mAP, calculate_pr_curve      # the import of test_hashing.py:15
    return calculate_mAP, calculate_pr_curve


def _outputs(nq=400, ndb=3000, nbit=64, ncls=20, seed=0):
    d, dl, q, ql, _ = synth.make_random_case(nq, ndb, nbit, ncls, p=0.3, seed=seed)
    # inference_one_epoch: per-batch .cpu() then torch.cat -> pageable CPU tensors; labels one-hot float (OneHot)
    db_out = {"codes": torch.cat(list(d.split(64))), "labels": synth.one_hot(dl, ncls)}
    test_out = {"codes": torch.cat(list(q.split(64))), "labels": synth.one_hot(ql, ncls)}
    return db_out, test_out, (d, dl, q, ql, ncls)


def test_call_a_train_helper_verbatim(calc):
    calculate_mAP, _ = calc
    db_out, test_out, (d, dl, q, ql, ncls) = _outputs()
    config = types.SimpleNamespace(dataset=types.SimpleNamespace(R=-1, multiclass=False), dist_metric="hamming")
    landmark_gt = None
    codes_name = "codes"
    for R in (-1, 1000):                         # cub200.yaml:4 / inat_birds.yaml:4
        config.dataset.R = R
        mAP, recalls, precisions = calculate_mAP(db_out[codes_name], db_out['labels'],
                                                 test_out[codes_name], test_out['labels'],
                                                 config.dataset.R, dist_metric=config.dist_metric,
                                                 PRs=[1, 5, 10], landmark_gt=landmark_gt,
                                                 db_id=db_out.get('id'),
                                                 test_id=test_out.get('id'),
                                                 multiclass=config.dataset.multiclass)
        om, orec, oprec = mo.calculate_mAP(d, dl, q, ql, R, PRs=[1, 5, 10])
        assert isinstance(mAP, float) and f"{mAP:.6f}" == f"{om:.6f}" and abs(mAP - om) < TOL
        assert np.allclose(recalls, orec, atol=TOL) and np.allclose(precisions, oprec, atol=TOL)
        assert isinstance(recalls[-1], float) and isinstance(precisions[-1], float)    # train_helper.py:240-241
    # 'id' outputs present (a tensor / numpy array per sample) while landmark_gt is None: accepted and unused
    db_out["id"] = torch.arange(d.shape[0])
    test_out["id"] = np.arange(q.shape[0])
    m2, _, _ = calculate_mAP(db_out[codes_name], db_out['labels'], test_out[codes_name], test_out['labels'],
                             config.dataset.R, dist_metric=config.dist_metric, PRs=[1, 5, 10], landmark_gt=None,
                             db_id=db_out.get('id'), test_id=test_out.get('id'), multiclass=True)
    assert m2 == mAP
    # numpy label arrays (np.concatenate branch of inference_one_epoch) and CUDA codes (trainers/shallow.py:86-88)
    m3, _, _ = calculate_mAP(db_out[codes_name].cuda(), db_out['labels'].numpy(), test_out[codes_name].cuda(),
                             test_out['labels'].numpy(), config.dataset.R, dist_metric="hamming", PRs=[1, 5, 10])
    assert m3 == mAP
    with pytest.raises(NotImplementedError):
        calculate_mAP(db_out[codes_name], db_out['labels'], test_out[codes_name], test_out['labels'], -1,
                      landmark_gt=object(), db_id=db_out.get('id'), test_id=test_out.get('id'))


def test_call_b_test_hashing_verbatim(calc):
    calculate_mAP, calculate_pr_curve = calc
    db_out, test_out, (d, dl, q, ql, ncls) = _outputs(seed=1)
    # this trainer returned 1-D labels -> the caller one-hots them (int64), after cloning (test_hashing.py:80-85)
    db_out["labels"], test_out["labels"] = dl.clone(), ql.clone()
    config = types.SimpleNamespace(R=ListConfig([100, 1000, -1]), ternary_threshold=0.0, dist_metric="hamming",
                                   PRs=ListConfig([1, 5, 10]), dataset=types.SimpleNamespace(nclass=ncls))
    codes_name = "codes"
    db_labels = db_out['labels'].clone()
    test_labels = test_out['labels'].clone()
    if len(db_labels.size()) == 1:
        db_labels = F.one_hot(db_labels, config.dataset.nclass)
        test_labels = F.one_hot(test_labels, config.dataset.nclass)
    snap = test_labels.clone()
    for thr in (0.0, 0.3):                       # configs/val.yaml:12 ternary_threshold
        config.ternary_threshold = thr
        mAPs, recalls, precisions = calculate_mAP(db_out[codes_name], db_labels,
                                                  test_out[codes_name], test_labels,
                                                  config.R,
                                                  threshold=config.ternary_threshold,
                                                  dist_metric=config.dist_metric,
                                                  PRs=config.PRs)
        assert isinstance(mAPs, list) and len(mAPs) == 3                  # test_hashing.py:124-128
        om, orec, oprec = mo.calculate_mAP(d, dl, q, ql, [100, 1000, -1], threshold=thr, PRs=[1, 5, 10])
        assert np.allclose(mAPs, om, atol=TOL)
        assert np.allclose(recalls, orec, atol=TOL) and np.allclose(precisions, oprec, atol=TOL)
        for R, recall, precision in zip(config.PRs, recalls, precisions):   # test_hashing.py:130-131
            assert isinstance(recall, float) and isinstance(precision, float)
    assert torch.equal(test_labels, snap)        # never mutated (upstream mutates; the clone at :80-81 guards it)
    # test_as_database (test_hashing.py:105-112): the query set is its own gallery, rank-0 item dropped
    config.R = 50
    mAPs, recalls, precisions = calculate_mAP(test_out[codes_name], test_labels,
                                              test_out[codes_name], test_labels,
                                              config.R,
                                              threshold=config.ternary_threshold,
                                              dist_metric=config.dist_metric,
                                              PRs=config.PRs,
                                              remove_first_retrieved=True)
    om, orec, oprec = mo.calculate_mAP(q, ql, q, ql, 50, threshold=0.3, PRs=[1, 5, 10], remove_first_retrieved=True)
    assert not isinstance(mAPs, list) and abs(mAPs - om) < TOL
    assert np.allclose(recalls, orec, atol=TOL) and np.allclose(precisions, oprec, atol=TOL)
    # compute_mAP=False branch (test_hashing.py:152-162)
    recalls, precisions, Rs = calculate_pr_curve(db_out[codes_name], db_labels,
                                                 test_out[codes_name], test_labels,
                                                 threshold=config.ternary_threshold,
                                                 dist_metric=config.dist_metric)
    orec, oprec, ors = mo.calculate_pr_curve(d, dl, q, ql, threshold=0.3)
    assert list(Rs) == ors and np.allclose(recalls, orec, atol=TOL) and np.allclose(precisions, oprec, atol=TOL)


def test_sub_code_and_zero_mean_as_the_callers_do_it(calc):
    """test_hashing.py:87-103: column slice / random bit subset (fancy indexing -> a new tensor) and the callers'
    own zero-mean subtraction, handed to the plain call."""
    calculate_mAP, _ = calc
    db_out, test_out, (d, dl, q, ql, ncls) = _outputs(seed=2)
    g = torch.Generator().manual_seed(0)
    bit_idxs = torch.randperm(64, generator=g)[:24]
    for sel in (slice(8, 40), bit_idxs):
        dbc, tc = db_out["codes"][:, sel], test_out["codes"][:, sel]
        db_mean = dbc.mean(dim=0, keepdim=True)
        dbc, tc = dbc - db_mean, tc - db_mean
        m, rec, prec = calculate_mAP(dbc, db_out["labels"], tc, test_out["labels"], -1, threshold=0.0,
                                     dist_metric="hamming", PRs=[1, 5, 10])
        om, orec, oprec = mo.calculate_mAP(dbc, dl, tc, ql, -1, PRs=[1, 5, 10])
        assert abs(m - om) < TOL and np.allclose(rec, orec, atol=TOL) and np.allclose(prec, oprec, atol=TOL)
