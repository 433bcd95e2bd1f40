"""CPU oracle on PACKED sign bits: the same normative definition as ``oracle/map_oracle.py``,
restated so that it scales to the large BASELINE shapes (1 M ... 100 M gallery rows) for a handful of queries.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as ``map_oracle``: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s checker / CPU legs may import it).  PARITY UNPINNED for the same
reason: ``utils.hashing.calculate_mAP`` is absent from the reference tree (``experiments/test_hashing.py:15``,
``README.md:11``).  ``tests/test_oracle.py`` pins this file against ``map_oracle`` (dense fp32 matmul + stable sort)
on random and tie-heavy cases, so the two restatements check each other.

Steps (each cites what it follows):
  1. sign: bit k of a row = ``code[k] > 0`` (``models/layers/signhash.py:11``; exact zeros are rejected here --
     ternary codes stay with ``map_oracle``);
  2. Hamming distance = popcount(q XOR g) (inequality-count form, ``trainers/orthohash.py:49``; equals
     ``0.5 * (nbit - q . g)`` of ``trainers/orthohash.py:263-264`` for +-1 codes);
  3. ranking = ascending (distance, gallery row) -- the stable sort of ``map_oracle._stable_order``; only the
     first R items are produced: a distance histogram gives the threshold distance t of the R-th item, the rows
     with distance <= t are taken in row order and stably sorted by distance (identical to a full stable sort
     truncated at R);
  4. relevance = equal class id (single-label form of ``map_oracle.relevance_matrix``), AP as
     ``map_oracle._ap_from_rel``.
"""
from __future__ import annotations

import numpy as np

__all__ = ["pack_sign_bits", "hamming_packed", "topk_packed", "ap_of_ranked"]


def pack_sign_bits(codes):
    """(n, nbit) real-valued array -> (n, ceil(nbit / 64)) uint64, little-endian bit order: bit k%64 of word k//64
    is ``codes[:, k] > 0``.  Raises on exact zeros / NaN (ternary keys are not representable in one plane)."""
    x = np.asarray(codes)
    if x.ndim != 2:
        raise ValueError("codes must be 2-D")
    if np.isnan(x).any():
        raise ValueError("codes contain NaN")
    if (x == 0).any():
        raise ValueError("exact zeros: use map_oracle (ternary keys)")
    n, nbit = x.shape
    words = (nbit + 63) // 64
    pos = np.zeros((n, words * 64), dtype=np.uint8)
    pos[:, :nbit] = x > 0
    return np.packbits(pos, axis=1, bitorder="little").view(np.uint64).reshape(n, words)


def hamming_packed(q_row, g_bits):
    """popcount(q XOR g) of one packed query against every packed gallery row -> int32 (n,)"""
    x = np.bitwise_xor(g_bits, q_row[None, :])
    return np.bitwise_count(x).sum(axis=1, dtype=np.int32)


def topk_packed(q_bits, g_bits, R, nbit, remove_first_retrieved=False):
    """First ``R`` (``-1`` = all) items of the canonical ranking for every query.

    Returns ``(ids int64 (nq, L), dist int32 (nq, L))``, ``L = min(R, ndb) [- 0]`` exactly as
    ``map_oracle.topk_ids``."""
    q_bits, g_bits = np.ascontiguousarray(q_bits), np.ascontiguousarray(g_bits)
    ndb = g_bits.shape[0]
    extra = 1 if remove_first_retrieved else 0
    L = ndb if R == -1 else min(int(R) + extra, ndb)
    ids = np.empty((q_bits.shape[0], L), dtype=np.int64)
    dist = np.empty((q_bits.shape[0], L), dtype=np.int32)
    for i in range(q_bits.shape[0]):
        d = hamming_packed(q_bits[i], g_bits)
        if L == ndb:
            cand = np.arange(ndb, dtype=np.int64)
        else:
            cum = np.cumsum(np.bincount(d, minlength=nbit + 1))
            t = int(np.searchsorted(cum, L))           # smallest distance at which >= L rows are covered
            cand = np.flatnonzero(d <= t)               # ascending row order
        order = np.argsort(d[cand], kind="stable")[:L]  # ties keep the row order
        ids[i] = cand[order]
        dist[i] = d[ids[i]]
    if remove_first_retrieved:
        ids, dist = ids[:, 1:], dist[:, 1:]
    return ids, dist


def ap_of_ranked(ids, q_class, g_class):
    """AP per query from ranked id lists and 1-D class ids (fp64; normalised by the relevant items inside the
    list, 0 if none -- ``map_oracle._ap_from_rel``)."""
    q_class, g_class = np.asarray(q_class), np.asarray(g_class)
    out = np.zeros(ids.shape[0], dtype=np.float64)
    for i in range(ids.shape[0]):
        rel = (g_class[ids[i]] == q_class[i]).astype(np.float64)
        n = rel.sum()
        if n > 0:
            prec = np.cumsum(rel) / np.arange(1, rel.shape[0] + 1, dtype=np.float64)
            out[i] = float((prec * rel).sum() / n)
    return out
