"""Seeded synthetic codes / labels of the BASELINE.json shapes.

Generator contract (identical for the CPU oracle and the GPU path, so both see the same
inputs): ``cb = sign(randn(C, nbit))``; item code = ``cb[label] * flip`` with
``flip = where(rand < p, -1, +1)``; stored as fp32 ``+-1 * (|randn| + 0.1)`` so that the sign
kernel is exercised on real values with no exact zeros.
"""
from __future__ import annotations

import os

import numpy as np
import torch

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (label fixture key, nbit list, flip prob)
DATASET_SHAPES = {
    "cub200": dict(nq=5794, ndb=5994, nclass=200),
    "cars196": dict(nq=8041, ndb=8144, nclass=196),
    "nabirds": dict(nq=24633, ndb=23929, nclass=555),
}


def load_label_fixture(name):
    """Integer label columns of the reference's ``data/<name>/{test,database}.txt`` lists
    (committed under ``tests/golden/labels_<name>.npz`` by ``oracle/make_golden.py``)."""
    path = os.path.join(_GOLDEN, f"labels_{name}.npz")
    z = np.load(path)
    return torch.from_numpy(z["test"].astype(np.int64)), torch.from_numpy(z["database"].astype(np.int64))


def codebook(nclass, nbit, seed=0, device="cpu"):
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    cb = torch.sign(torch.randn(nclass, nbit, generator=g))
    cb[cb == 0] = 1.0
    return cb.to(device)


def codes_from_labels(labels, cb, p=0.15, seed=0, device=None):
    """Real-valued codes clustered around ``cb[label]`` with per-bit flip probability ``p``."""
    device = torch.device(device) if device is not None else labels.device
    g = torch.Generator(device=device).manual_seed(int(seed))
    labels = labels.to(device)
    cb = cb.to(device)
    n, nbit = labels.shape[0], cb.shape[1]
    flip = torch.where(torch.rand(n, nbit, generator=g, device=device) < p, -1.0, 1.0)
    mag = torch.randn(n, nbit, generator=g, device=device).abs() + 0.1
    return (cb[labels] * flip * mag).to(torch.float32)


def one_hot(labels, nclass, dtype=torch.float32):
    out = torch.zeros(labels.shape[0], nclass, dtype=dtype, device=labels.device)
    out[torch.arange(labels.shape[0], device=labels.device), labels] = 1
    return out


def make_dataset_case(name, nbit=64, p=0.15, seed=0, device="cpu"):
    """cfg1-3: label columns from the fixture, clustered codes.  Returns
    ``(db_codes, db_ids, q_codes, q_ids, nclass)`` with 1-D int64 ids."""
    q_ids, d_ids = load_label_fixture(name)
    nclass = DATASET_SHAPES[name]["nclass"]
    cb = codebook(nclass, nbit, seed)
    q = codes_from_labels(q_ids, cb, p, seed * 2 + 1, device)
    d = codes_from_labels(d_ids, cb, p, seed * 2 + 2, device)
    return d, d_ids.to(device), q, q_ids.to(device), nclass


def make_random_case(nq, ndb, nbit, nclass, p=0.30, seed=0, device="cpu", db_chunk=1 << 20, shard=0):
    """cfg4/5-style: labels ``randint(nclass)``; gallery generated in chunks on ``device``.
    ``shard`` selects an independent gallery block (same codebook and queries): rank r of a weak-scaled run
    generates shard r, so the concatenation over ranks is one large gallery without duplicates."""
    device = torch.device(device)
    g = torch.Generator(device=device).manual_seed(int(seed) + 7919)
    cb = codebook(nclass, nbit, seed)
    q_ids = torch.randint(nclass, (nq,), generator=g, device=device)
    if shard:
        g = torch.Generator(device=device).manual_seed(int(seed) + 7919 + 104729 * int(shard))
    d_ids = torch.randint(nclass, (ndb,), generator=g, device=device)
    q = codes_from_labels(q_ids, cb, p, seed * 2 + 1, device)
    d = torch.empty(ndb, nbit, dtype=torch.float32, device=device)
    for s in range(0, ndb, db_chunk):
        e = min(ndb, s + db_chunk)
        d[s:e] = codes_from_labels(d_ids[s:e], cb, p, (seed * 2 + 2) * 1000003 + s + 15485863 * int(shard), device)
    return d, d_ids, q, q_ids, nclass
