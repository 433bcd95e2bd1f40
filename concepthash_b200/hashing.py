"""Public Python surface of the B200-native retrieval evaluation -- the drop-in for the reference's
missing ``utils.hashing`` module (imported at ``experiments/test_hashing.py:15`` and
``experiments/train_helper.py:18``; called at ``test_hashing.py:106-119,153-162`` and
``train_helper.py:228-234``).  Same names, argument meaning, return types and error behaviour.

Positional order is the call sites': GALLERY first, QUERY second.
"""
from __future__ import annotations

import torch

from . import _lib
from .evaluator import DistComm, Evaluator, LocalComm

_EVALUATORS = {}


def get_evaluator(device=None, group=None):
    """Process-wide evaluator for ``device`` (one ``ch_ws`` workspace per GPU), optionally bound to a
    ``torch.distributed`` process group for the row-sharded gallery mode."""
    from .backend_cuda import CudaBackend   # raises loudly if CUDA or the native library is missing
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    key = (str(device), id(group) if group is not None else None)
    ev = _EVALUATORS.get(key)
    if ev is None:
        comm = DistComm(group) if group is not None else LocalComm()
        ev = Evaluator(CudaBackend(device), comm)
        _EVALUATORS[key] = ev
    return ev


def _device_of(*tensors):
    for t in tensors:
        if getattr(t, "is_cuda", False):
            return t.device
    return None


def _check_common(dist_metric, landmark_gt, db_id, test_id):
    if dist_metric != "hamming":
        # only "hamming" is ever configured (configs/train.yaml:21, configs/val.yaml:11)
        raise NotImplementedError(f"dist_metric={dist_metric!r}: only 'hamming' is implemented")
    if landmark_gt is not None:
        raise NotImplementedError("GLDv2 landmark ground truth (landmark_gt/db_id/test_id) is not implemented")


def _as_tensor(x):
    from .codes_io import PackedCodes
    if isinstance(x, PackedCodes):
        return x
    return x.detach() if isinstance(x, torch.Tensor) else torch.as_tensor(x)


def calculate_mAP(db_codes, db_labels, test_codes, test_labels, R, threshold=0., dist_metric="hamming",
                  PRs=None, multiclass=False, landmark_gt=None, db_id=None, test_id=None,
                  remove_first_retrieved=False, group=None, zero_mean_eval=False, empty_queries="zero", **_ignored):
    """mAP@R (+ R@k, P@k for ``PRs``) of a Hamming ranking on sign-binarised codes.

    ``empty_queries`` (not in the reference signature) names the convention for a query whose list holds NO relevant
    item: ``"zero"`` (default; the OrthoHash-lineage ``APx.append(0)``) scores it AP = 0 and keeps it in the mean,
    ``"skip"`` (DeepHash-lineage ``if rel != 0: APx.append(...)``) leaves it out of the mean (mAP = 0 if every query is
    empty).  The upstream function is not in the reference tree, so the choice is the caller's.

    ``zero_mean_eval=True`` (not in the reference signature) fuses the callers' preprocessing
    ``db_mean = db.mean(0); db -= db_mean; test -= db_mean`` (experiments/train_helper.py:223-226,
    test_hashing.py:100-103) into the sign/bit-pack kernel: no (N, nbit) temporaries, no extra passes.  Bit slices
    (``sub_code_eval``, test_hashing.py:87-98) need nothing: strided column views are packed in place.

    ``R``: int, ``-1`` = whole gallery (configs/val.yaml:7), or a list (test_hashing.py:124-128).
    Returns ``(mAP | [mAP, ...], recalls, precisions)`` as Python floats / lists of floats.
    Ties are broken by ascending gallery row index.  Inputs are neither retained nor mutated.
    With ``group`` (a torch.distributed process group) ``db_*`` is this rank's contiguous gallery block.
    """
    _check_common(dist_metric, landmark_gt, db_id, test_id)
    if empty_queries not in ("zero", "skip"):
        raise ValueError(f"empty_queries={empty_queries!r}: 'zero' or 'skip'")
    skip = empty_queries == "skip"
    db_codes, db_labels, test_codes, test_labels = map(_as_tensor, (db_codes, db_labels, test_codes, test_labels))
    r_is_list = isinstance(R, (list, tuple)) or (hasattr(R, "__iter__") and not isinstance(R, (str, bytes)))
    r_list = [int(r) for r in R] if r_is_list else [int(R)]
    pr_list = [] if PRs is None else [int(k) for k in PRs]
    ev = get_evaluator(_device_of(db_codes, test_codes, db_labels, test_labels), group)
    maps, recalls, precisions = [], [], []
    # one pass serves up to CH_MAX_R values of R and CH_MAX_PR cut-offs; longer lists take several passes
    nr, npr = max(1, -(-len(r_list) // _lib.CH_MAX_R)), max(1, -(-len(pr_list) // _lib.CH_MAX_PR))
    for i in range(max(nr, npr)):
        rs = r_list[i * _lib.CH_MAX_R:(i + 1) * _lib.CH_MAX_R]
        ks = pr_list[i * _lib.CH_MAX_PR:(i + 1) * _lib.CH_MAX_PR]
        if not rs and not ks:
            continue
        if skip and rs:
            # per-query AP (device, fp64) -> mean over the queries whose list holds a relevant item (AP > 0 <=> it does)
            m, r, p, ap = ev.evaluate(db_codes, db_labels, test_codes, test_labels, rs, threshold, ks,
                                      bool(remove_first_retrieved), return_ap=True, zero_mean=bool(zero_mean_eval))
            hit = (ap > 0).sum(dim=1)
            m = (ap.sum(dim=1) / hit.clamp(min=1)).tolist()
        else:
            m, r, p = ev.evaluate(db_codes, db_labels, test_codes, test_labels, rs, threshold, ks,
                                  bool(remove_first_retrieved), zero_mean=bool(zero_mean_eval))
        maps += m
        recalls += r
        precisions += p
    return (maps if r_is_list else maps[0]), recalls, precisions


def map_at_r(*, query_codes, db_codes, query_labels, db_labels, R, **kw):
    """Keyword-only convenience in BASELINE.json's wording (query first)."""
    return calculate_mAP(db_codes, db_labels, query_codes, query_labels, R, **kw)[0]


def default_pr_cutoffs(n):
    out, k = [], 1
    while k < n:
        out.append(k)
        k *= 2
    if n > 0:
        out.append(int(n))
    return out


def calculate_pr_curve(db_codes, db_labels, test_codes, test_labels, threshold=0., dist_metric="hamming",
                       remove_first_retrieved=False, Rs=None, group=None, zero_mean_eval=False, **_ignored):
    """Precision / recall at a set of cut-offs (second symbol imported at test_hashing.py:15; used at
    :152-168).  Returns ``(recalls, precisions, Rs)``; default ``Rs`` = powers of two up to the list
    length, plus the length itself.  Hit counting as in ``calculate_mAP``."""
    _check_common(dist_metric, None, None, None)
    db_codes = _as_tensor(db_codes)
    n = int(db_codes.shape[0])
    if group is not None:
        import torch.distributed as dist
        t = torch.tensor([n], dtype=torch.int64, device=_device_of(db_codes) or "cuda")
        dist.all_reduce(t, group=group)
        n = int(t.item())
    n -= 1 if remove_first_retrieved else 0
    rs = default_pr_cutoffs(n) if Rs is None else [int(r) for r in Rs]
    _, recalls, precisions = calculate_mAP(db_codes, db_labels, test_codes, test_labels, [], threshold=threshold,
                                           PRs=rs, remove_first_retrieved=remove_first_retrieved, group=group,
                                           zero_mean_eval=zero_mean_eval)       # (chunks of CH_MAX_PR cut-offs)
    return recalls, precisions, rs


def retrieve_topk(query_codes, db_codes, R, threshold=0., remove_first_retrieved=False, group=None,
                  zero_mean_eval=False):
    """Exact ranked retrieval: ``(ids int64 (nq, L), dist float32 (nq, L))``, ascending (distance, gallery
    row) -- bit-identical to a stable sort of the dense distance matrix."""
    query_codes, db_codes = _as_tensor(query_codes), _as_tensor(db_codes)
    ev = get_evaluator(_device_of(db_codes, query_codes), group)
    ids, keys, ternary = ev.retrieve(db_codes, query_codes, int(R), threshold, bool(remove_first_retrieved),
                                     zero_mean=bool(zero_mean_eval))
    dist = keys.to(torch.float32) * (0.5 if ternary else 1.0)
    return ids, dist


def pack_codes(codes):
    """Real-valued codes -> ``PackedCodes`` through the sign/bit-pack KERNEL (K1): what ``codes_io.save_packed``
    writes and what every entry point accepts in place of real-valued codes.  Raises if some sign is 0."""
    from .codes_io import PackedCodes
    codes = _as_tensor(codes)
    ev = get_evaluator(_device_of(codes))
    if hasattr(ev.b, "begin"):
        ev.b.begin()
    flags = ev.b.zeros((1,), torch.int32)
    bits, _ = ev.b.pack_sign(codes, 0.0, flags, False)
    fl = int(flags.cpu()[0])
    if fl & 2:
        raise ValueError("codes contain NaN")
    if fl & 1:
        raise ValueError("codes with exact zeros cannot be stored as packed bits")
    return PackedCodes(bits[:codes.shape[0]].contiguous(), int(codes.shape[1]))


def get_hamm_dist(codes, centroids, margin=0., normalize=False):
    """Code -> codebook Hamming distance matrix (callers: trainers/orthohash.py:362,397,430,465,
    trainers/dpn.py:30,62; identity: trainers/orthohash.py:263-264).  float32, returned on the device of ``codes``
    (the trainers hand the matrix straight to ``calculate_accuracy_hamm_dist(hamm_dist, labels)``,
    utils/metrics.py:18-29, with labels on that device); the arithmetic runs on the GPU either way.

    ``margin`` zeroes ``|codes| < margin`` before the sign (ternary keys, half-integer distances) and ONE matrix is
    returned; no in-tree caller passes it (every call is ``get_hamm_dist(codes, codebook, normalize=True)``), and
    the upstream helper's behaviour for ``margin != 0`` is unpinned -- see INTEGRATION.md."""
    codes, centroids = _as_tensor(codes), _as_tensor(centroids)
    if codes.dim() != 2 or centroids.dim() != 2 or codes.shape[1] != centroids.shape[1]:
        raise ValueError("nbit mismatch")
    ev = get_evaluator(_device_of(codes, centroids))
    keys, ternary = ev.hamming_matrix(codes, centroids, margin)
    d = keys.to(torch.float32) * (0.5 if ternary else 1.0)
    d = d / codes.shape[1] if normalize else d
    return d.to(codes.device) if isinstance(codes, torch.Tensor) and d.device != codes.device else d


# ---- training-loss helpers of the same missing module -------------------------------------------------------------
# Not part of the retrieval hot path: two small autograd-visible torch helpers that the reference's pairwise losses
# import from ``utils.hashing`` (models/loss/dpsh.py:4, models/loss/hashnet.py:5, models/loss/adsh.py:5).  A drop-in
# ``utils/hashing.py`` has to export them or those modules stop importing.  Their behaviour is pinned by the call sites.
def get_sim(label_a, label_b, onehot=True):
    """Boolean pairwise similarity ``(N, M)``: ``s_ij = 1`` iff samples i and j share a class.

    ``onehot=True``: one-/multi-hot ``(N, C)`` / ``(M, C)`` label matrices -> ``label_a @ label_b.T >= 1`` (the
    "share >= 1 positive class" relevance of ``calculate_mAP``); ``onehot=False``: class ids -> equality.  Callers:
    ``get_sim(y1, y2).float()`` (models/loss/dpsh.py:58, models/loss/hashnet.py:73),
    ``get_sim(Y.cpu(), Y_train.cpu(), onehot).float() * 2. - 1.`` (models/loss/adsh.py:26,41,65)."""
    label_a, label_b = torch.as_tensor(label_a), torch.as_tensor(label_b)
    if onehot:
        return torch.matmul(label_a.float(), label_b.float().t()) >= 1
    return label_a.reshape(-1, 1) == label_b.reshape(1, -1)


def log_trick(dot_product):
    """``log(1 + exp(x))`` without overflow: ``log(1 + exp(-|x|)) + max(x, 0)`` -- the expression the callers keep
    beside the call as a comment (models/loss/dpsh.py:64, models/loss/hashnet.py:79).  Differentiable."""
    return torch.log(1 + torch.exp(-torch.abs(dot_product))) + dot_product.clamp(min=0)
