"""CPU oracle for the retrieval-evaluation hot path (codes -> Hamming ranking -> mAP@R).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``concepthash_b200`` / ``utils.hashing``) never does.

PARITY UNPINNED.  The function this restates, ``utils.hashing.calculate_mAP``, is NOT in
the reference tree: ``experiments/test_hashing.py:15`` and ``experiments/train_helper.py:18``
import it from ``utils.hashing``, but ``/root/reference/utils/`` ships only ``metrics.py``
and ``README.md:11`` points at another, unpinned repository (kamwoh/sdc, no commit hash).
The reference has no tests and no golden vectors for this path.  The
oracle therefore follows (a) the two call sites (``test_hashing.py:106-119``,
``train_helper.py:228-234``), (b) the in-tree statements of each step, cited per function
below, and (c) the normative definition in SURVEY.md §8(c).  What IS pinned against the
reference are the sub-steps that exist in tree (sign, the Hamming identity, the
nearest-codeword accuracy helper, the 32-row chunk iterator) -- see ``oracle/make_golden.py``.

Tie rule (normative for this repo): ascending distance, then ascending
gallery row index (= a stable sort).  ``tie="torch_topk"`` reproduces upstream's
unspecified ``torch.topk`` order purely to report the delta.
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = [
    "sign_codes",
    "hamming_distance_matrix",
    "ranked_lists",
    "relevance_matrix",
    "calculate_mAP",
    "calculate_mAP_upstream_style",
    "calculate_pr_curve",
    "get_hamm_dist",
    "topk_ids",
]


def _as_2d_labels(labels, nclass=None):
    """1-D integer class ids -> one-hot (N, C); 2-D passes through.

    Mirrors ``experiments/test_hashing.py:83-85`` (``F.one_hot`` for 1-D labels)."""
    labels = torch.as_tensor(labels)
    if labels.dim() == 1:
        ids = labels.to(torch.int64)
        c = int(nclass) if nclass is not None else (int(ids.max()) + 1 if ids.numel() else 1)
        out = torch.zeros(ids.numel(), c, dtype=torch.float32)
        ok = (ids >= 0) & (ids < c)
        out[torch.nonzero(ok).squeeze(1), ids[ok]] = 1.0
        return out
    return labels


def _pair_labels(q_labels, d_labels):
    """Both label sets as 2-D float arrays with a common class count."""
    ql = _as_2d_labels(q_labels).to(torch.float32)
    dl = _as_2d_labels(d_labels).to(torch.float32)
    c = max(ql.shape[1], dl.shape[1])
    if ql.shape[1] < c:
        ql = torch.nn.functional.pad(ql, (0, c - ql.shape[1]))
    if dl.shape[1] < c:
        dl = torch.nn.functional.pad(dl, (0, c - dl.shape[1]))
    return ql, dl


def sign_codes(codes, threshold=0.0):
    """Step 1: optional ternary zeroing, then ``torch.sign``.

    ``x[|x| < threshold] = 0`` iff ``threshold != 0`` (``test_hashing.py:109``,
    ``configs/val.yaml:12``); sign as in ``models/layers/signhash.py:11`` /
    ``trainers/orthohash.py:78``.  ``sign(0) == 0`` is kept (half-integer distances)."""
    x = torch.as_tensor(codes).detach().to("cpu").clone()
    if not x.dtype.is_floating_point:
        x = x.to(torch.float32)
    if threshold != 0:
        x[x.abs() < threshold] = 0          # in the dtype of the codes, as the reference's in-place statement does
    return torch.sign(x.to(torch.float32))


def hamming_distance_matrix(q_sign, d_sign):
    """Step 2: ``dist = 0.5 * (nbit - q @ d.T)`` in fp32 (exact for |values| <= 2**24).

    The identity is stated in tree at ``trainers/orthohash.py:263-264`` (``get_hd``) and in
    inequality-count form at ``trainers/orthohash.py:49``."""
    nbit = q_sign.shape[1]
    return 0.5 * (nbit - q_sign.to(torch.float32) @ d_sign.to(torch.float32).t())


def relevance_matrix(q_labels, d_labels):
    """Step 5: ``rel[i, j] = any_c (q_lab[i, c] > 0) & (d_lab[j, c] > 0)``.

    Same notion as ``get_sim(y1, y2)`` used at ``models/loss/dpsh.py:58`` and the id equality
    at ``trainers/adsh.py:144``; covers one-hot and multi-hot alike."""
    ql, dl = _pair_labels(q_labels, d_labels)
    return ((ql > 0).to(torch.float32) @ (dl > 0).to(torch.float32).t()) > 0


def _stable_order(dist):
    """Step 4 (canonical): ascending distance, ties by ascending gallery row index."""
    return torch.sort(dist, dim=1, stable=True)[1]


def ranked_lists(q_codes, d_codes, threshold=0.0, tie="stable", topk=None):
    """Full (or top-``topk``) ranking of the gallery for every query.

    Returns ``(ids int64 (nq, L), dist fp32 (nq, L))`` in ranked order."""
    qs = sign_codes(q_codes, threshold)
    ds = sign_codes(d_codes, threshold)
    if qs.shape[1] != ds.shape[1]:
        raise ValueError("nbit mismatch between query and gallery codes")
    dist = hamming_distance_matrix(qs, ds)
    ndb = ds.shape[0]
    L = ndb if topk is None else min(int(topk), ndb)
    if tie == "stable":
        ids = _stable_order(dist)[:, :L]
    elif tie == "torch_topk":
        # upstream: torch.topk(dist, R, dim=1, largest=False)[1] -- tie order unspecified
        ids = torch.topk(dist, L, dim=1, largest=False)[1]
    else:
        raise ValueError(f"unknown tie rule {tie!r}")
    return ids, torch.gather(dist, 1, ids)


def topk_ids(q_codes, d_codes, R, threshold=0.0, remove_first_retrieved=False):
    """Ranked id / distance lists exactly as the GPU selector must emit them."""
    ndb = torch.as_tensor(d_codes).shape[0]
    extra = 1 if remove_first_retrieved else 0
    L = ndb if R == -1 else min(int(R) + extra, ndb)
    ids, dist = ranked_lists(q_codes, d_codes, threshold, "stable", L)
    if remove_first_retrieved:
        ids, dist = ids[:, 1:], dist[:, 1:]
    return ids, dist


def _ap_from_rel(rel_row):
    """Step 6: AP over one ranked 0/1 relevance list (fp64).

    ``AP = sum_k (cumsum(rel)[k] / (k+1)) * rel[k] / sum(rel)``; the normaliser is the number
    of relevant items INSIDE the list (upstream convention), 0 if none."""
    rel_row = np.asarray(rel_row, dtype=np.float64)
    n = rel_row.sum()
    if n == 0:
        return 0.0
    cum = np.cumsum(rel_row)
    prec = cum / np.arange(1, rel_row.shape[0] + 1, dtype=np.float64)
    return float((prec * rel_row).sum() / n)


def calculate_mAP(db_codes, db_labels, test_codes, test_labels, R, threshold=0.0,
                  dist_metric="hamming", PRs=None, multiclass=False, landmark_gt=None,
                  db_id=None, test_id=None, remove_first_retrieved=False, tie="stable",
                  return_per_query=False, empty_queries="zero", **_ignored):
    """Oracle for the hot path.  Gallery first, query second.

    ``empty_queries``: "zero" (normative here: AP = 0, kept in the mean) or "skip" (queries without a relevant item in
    their list are left out of the mean -- the other convention found among upstream lineages).

    Definition (normative):
      1. ternary zeroing iff threshold != 0, then sign;
      2. dist = 0.5 * (nbit - q @ d.T);
      3. R_eff = ndb if R == -1 else min(R, ndb); with ``remove_first_retrieved`` the list is
         the first R_eff + 1 ranked items with the rank-0 item dropped
         (``test_hashing.py:105-112``, self-retrieval when the query set is its own gallery);
      4. order = stable sort (distance, then gallery row index);
      5. rel = shares >= 1 positive class;
      6. AP normalised by the relevant items inside the list, fp64; 0 if none;
      7. mAP = mean over ALL queries;
      8. P@k = mean_q hits_k / k, R@k = mean_q hits_k / max(1, total relevant in gallery),
         hits_k = relevant items among the first min(k, len) of the FULL canonical ranking
         (after the optional rank-0 removal) -- independent of R.  With
         ``remove_first_retrieved`` the removed item no longer counts as gallery-relevant.
    ``R`` may be a list (``test_hashing.py:124-128``) -> list of mAPs.
    Returns ``(mAP | [mAP...], recalls, precisions)`` as Python floats / lists."""
    if dist_metric != "hamming":
        raise NotImplementedError(f"dist_metric={dist_metric!r}")
    if landmark_gt is not None or db_id is not None or test_id is not None:
        raise NotImplementedError("GLDv2 landmark ground truth is out of scope")
    PRs = [] if PRs is None else [int(k) for k in PRs]
    r_list = [int(r) for r in R] if isinstance(R, (list, tuple)) else [int(R)]

    ids, _ = ranked_lists(test_codes, db_codes, threshold, tie)
    nq, ndb = ids.shape
    rel_full = relevance_matrix(test_labels, db_labels)           # (nq, ndb) gallery order
    rel_ranked = torch.gather(rel_full, 1, ids).numpy()           # ranked order
    total_rel = rel_full.sum(dim=1).numpy().astype(np.float64)
    if remove_first_retrieved:
        total_rel = total_rel - rel_ranked[:, 0]
        rel_ranked = rel_ranked[:, 1:]
    L = rel_ranked.shape[1]

    aps = np.zeros((len(r_list), nq), dtype=np.float64)
    for ri, r in enumerate(r_list):
        r_eff = L if r == -1 else min(r, L)
        for i in range(nq):
            aps[ri, i] = _ap_from_rel(rel_ranked[i, :r_eff])
    if empty_queries == "skip":
        maps = [float(a[a > 0].mean()) if (a > 0).any() else 0.0 for a in aps]
    elif empty_queries == "zero":
        maps = [float(a.mean()) if nq else 0.0 for a in aps]
    else:
        raise ValueError(f"empty_queries={empty_queries!r}")

    cum = np.cumsum(rel_ranked.astype(np.float64), axis=1) if L else np.zeros((nq, 0))
    recalls, precisions = [], []
    for k in PRs:
        kk = min(k, L)
        hits = cum[:, kk - 1] if kk > 0 else np.zeros(nq)
        precisions.append(float((hits / k).mean()) if nq else 0.0)
        recalls.append(float((hits / np.maximum(1.0, total_rel)).mean()) if nq else 0.0)

    m = maps if isinstance(R, (list, tuple)) else maps[0]
    if return_per_query:
        return m, recalls, precisions, aps
    return m, recalls, precisions


def calculate_mAP_upstream_style(db_codes, db_labels, test_codes, test_labels, R,
                                 threshold=0.0, chunk=32):
    """The structure upstream is recalled to have: 32-row gallery chunk
    GEMMs -> dense (nq, ndb) fp32 -> ``torch.topk(largest=False)`` -> per-query numpy loop.

    Used ONLY as the timed CPU baseline (``bench.py`` ``cpu_baseline`` / ``--impl reference``)
    and to report the tie-order delta.  Tie order is torch.topk's (unspecified)."""
    qs = sign_codes(test_codes, threshold)
    ds = sign_codes(db_codes, threshold)
    nbit = ds.shape[1]
    ql, dl = _pair_labels(test_labels, db_labels)
    dl = np.asarray(dl.cpu().numpy())
    ql = np.array(ql.cpu().numpy(), copy=True)                    # never mutate the caller's
    dist = []
    with torch.no_grad():
        for s in range(0, ds.shape[0], chunk):                     # engine.py:41-54, 64-80 role
            dist.append(0.5 * (nbit - torch.matmul(qs, ds[s:s + chunk].t())))
        dist = torch.cat(dist, 1)
    ndb = ds.shape[0]
    r = ndb if R == -1 else min(int(R), ndb)
    top = torch.topk(dist, r, dim=1, largest=False)[1].numpy()
    apx = []
    for i in range(dist.shape[0]):
        label = ql[i, :]
        label[label == 0] = -1
        imatch = np.sum(np.equal(dl[top[i, :r], :], label), 1) > 0
        rel = np.sum(imatch)
        lx = np.cumsum(imatch)
        px = lx.astype(float) / np.arange(1, r + 1, 1)
        apx.append(np.sum(px * imatch) / rel if rel != 0 else 0.0)
    return float(np.mean(np.array(apx))) if apx else 0.0


def calculate_pr_curve(db_codes, db_labels, test_codes, test_labels, threshold=0.0,
                       dist_metric="hamming", remove_first_retrieved=False, Rs=None, **_ignored):
    """Oracle for the second public symbol (``test_hashing.py:15,152-168``); upstream semantics
    are unpinned, so it is DEFINED here: at each cut-off k in ``Rs`` (default: powers of two up
    to the list length, plus the length itself) report mean_q R@k and mean_q P@k with the same
    hit counting as ``calculate_mAP``.  Returns ``(recalls, precisions, Rs)``."""
    if dist_metric != "hamming":
        raise NotImplementedError(f"dist_metric={dist_metric!r}")
    ndb = torch.as_tensor(db_codes).shape[0]
    L = ndb - (1 if remove_first_retrieved else 0)
    if Rs is None:
        Rs = default_pr_cutoffs(L)
    _, recalls, precisions = calculate_mAP(db_codes, db_labels, test_codes, test_labels, -1,
                                           threshold=threshold, PRs=list(Rs),
                                           remove_first_retrieved=remove_first_retrieved)
    return recalls, precisions, list(Rs)


def default_pr_cutoffs(L):
    out, k = [], 1
    while k < L:
        out.append(k)
        k *= 2
    if L > 0:
        out.append(int(L))
    return out


def get_hamm_dist(codes, centroids, margin=0.0, normalize=False):
    """Per-batch code -> codebook Hamming distance.

    Callers: ``trainers/orthohash.py:362,397,430,465``, ``trainers/dpn.py:30,62``; the in-tree
    statement of the identity is ``get_hd`` at ``trainers/orthohash.py:263-264``:
    ``0.5 * (nbit - a @ b.T)`` (divided by nbit when ``normalize``)."""
    nbit = torch.as_tensor(centroids).shape[1]
    d = hamming_distance_matrix(sign_codes(codes, margin), sign_codes(centroids, 0.0))
    return d / nbit if normalize else d


def zero_mean(db_codes, test_codes):
    """The callers' ``zero_mean_eval`` preprocessing (experiments/train_helper.py:223-226,
    experiments/test_hashing.py:100-103): both sets minus the gallery's column mean.  The mean is accumulated in
    fp64 and rounded to the dtype of the codes (the reference lets torch accumulate in the dtype of the codes; the
    two means agree to ~1e-7 relative, so only an element within that distance of the mean can binarise
    differently)."""
    db_codes, test_codes = torch.as_tensor(db_codes), torch.as_tensor(test_codes)
    mean = db_codes.double().mean(dim=0, keepdim=True).to(db_codes.dtype)
    return db_codes - mean, test_codes - mean.to(test_codes.dtype)
