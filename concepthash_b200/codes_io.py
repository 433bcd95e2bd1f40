"""Offline inputs of the retrieval evaluation.

The reference writes the outputs of an inference epoch with ``trainer.save_codes`` -> ``io.fast_save`` (torch
pickles): ``outputs/db_best.pth`` / ``outputs/test_best.pth`` hold ``{'codes': (N, nbit) float, 'labels': ...}``
(experiments/train_helper.py:283-284, trainers/base.py:184-188) and ``outputs.pth`` holds
``{'test': test_out, 'db': db_out}`` (experiments/test_hashing.py:172-174).  ``evaluate_dumps`` scores such files
on the B200 path without the model stack, naming its results like ``RetrievalExperiment.evaluation``
(train_helper.py:207-241: one ``mAP{postfix}`` per output key that contains ``"codes"``).

The reference has no packed on-disk format; ``save_packed`` / ``load_packed`` define one for sign codes:
little-endian, header ``CHPK1`` + (n u64, nbit u32, words u32, label kind u32, nclass u32), then ``n x words`` u32
(bit k % 32 of word k / 32 = code[k] > 0 -- the layout of ``ch_pack_sign``), then ``n`` int64 class ids.  A
``PackedCodes`` can be passed to ``calculate_mAP`` / ``retrieve_topk`` wherever real-valued codes are accepted:
the sign/bit-pack kernel is skipped.  (Codes with exact zeros are not representable: they need the real values.)
"""
from __future__ import annotations

import struct

import numpy as np
import torch

MAGIC = b"CHPK1\0\0\0"


class PackedCodes:
    """Sign codes as packed bits: ``bits`` (n, words) int32, ``nbit``.  Duck-types the few tensor attributes the
    evaluator looks at (shape, dim, is_cuda, slicing)."""

    def __init__(self, bits, nbit):
        if bits.dim() != 2 or bits.dtype != torch.int32:
            raise ValueError("bits must be (n, words) int32")
        words = 1 if nbit <= 32 else 2 if nbit <= 64 else 4 if nbit <= 128 else 8
        if nbit <= 0 or nbit > 256 or bits.shape[1] != words:
            raise ValueError(f"{bits.shape[1]} words do not hold {nbit}-bit codes")
        self.bits, self.nbit = bits.contiguous(), int(nbit)

    shape = property(lambda self: (int(self.bits.shape[0]), self.nbit))
    is_cuda = property(lambda self: self.bits.is_cuda)
    device = property(lambda self: self.bits.device)
    dtype = torch.float32

    def dim(self):
        return 2

    def element_size(self):
        return 4

    def detach(self):
        return self

    def to(self, *a, **kw):
        return PackedCodes(self.bits.to(*a, **kw), self.nbit)

    def cpu(self):
        return PackedCodes(self.bits.cpu(), self.nbit)

    def __getitem__(self, idx):
        if not isinstance(idx, slice):
            raise TypeError("PackedCodes supports row slices only")
        return PackedCodes(self.bits[idx], self.nbit)

    @staticmethod
    def from_codes(codes):
        """Host-side writer of the bit layout (bit = code > 0) for tools and tests that have no GPU at hand; raises
        if a sign is 0.  The evaluation path never calls it: ``hashing.pack_codes`` packs with the CUDA kernel."""
        x = torch.as_tensor(codes).detach().cpu()
        if (x == 0).any():
            raise ValueError("codes with exact zeros cannot be stored as packed bits")
        n, nbit = x.shape
        words = 1 if nbit <= 32 else 2 if nbit <= 64 else 4 if nbit <= 128 else 8
        full = np.zeros((n, words * 32), dtype=np.uint8)
        full[:, :nbit] = (x > 0).numpy()
        packed = np.packbits(full, axis=1, bitorder="little").view("<u4").astype(np.uint32)
        return PackedCodes(torch.from_numpy(packed.view(np.int32).copy()), nbit)


def save_packed(path, packed, labels=None, nclass=0):
    """labels: None or 1-D integer class ids"""
    bits = packed.bits.cpu().numpy().view(np.uint32).astype("<u4")
    n, words = bits.shape
    kind = 0 if labels is None else 1
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<QIIII", n, packed.nbit, words, kind, int(nclass)))
        bits.tofile(f)
        if labels is not None:
            lab = torch.as_tensor(labels).cpu()
            if lab.dim() != 1 or lab.shape[0] != n:
                raise ValueError("labels must be 1-D class ids, one per row")
            lab.numpy().astype("<i8").tofile(f)


def load_packed(path):
    """-> (PackedCodes, labels int64 (n,) or None, nclass)"""
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path}: not a CHPK1 file")
        n, nbit, words, kind, nclass = struct.unpack("<QIIII", f.read(24))
        bits = np.fromfile(f, dtype="<u4", count=n * words).reshape(n, words)
        if bits.size != n * words:
            raise ValueError(f"{path}: truncated")
        labels = None
        if kind == 1:
            labels = torch.from_numpy(np.fromfile(f, dtype="<i8", count=n).astype(np.int64))
            if labels.shape[0] != n:
                raise ValueError(f"{path}: truncated labels")
    return PackedCodes(torch.from_numpy(bits.astype(np.uint32).view(np.int32).copy()), nbit), labels, nclass


def load_code_dump(path, part=None, trusted=False):
    """A reference code dump -> ``{'codes...': tensor, 'labels': tensor, ...}``.  ``outputs.pth`` files hold both
    sides; pick one with ``part`` = ``'db'`` / ``'test'``.

    Dumps are plain dicts of tensors (trainers/base.py:184-188), so they load with ``weights_only=True`` -- a
    crafted file cannot execute code.  ``trusted=True`` opts into full unpickling for a legacy dump that holds
    other objects (numpy arrays pickled by ``np.concatenate`` outputs, base.py:303); only for files you wrote."""
    try:
        obj = torch.load(path, map_location="cpu", weights_only=True)
    except Exception as e:
        if not trusted:
            raise ValueError(f"{path}: not loadable as tensors only ({type(e).__name__}: {e}); pass trusted=True "
                             "(--trusted) if this dump is your own and holds non-tensor objects") from e
        obj = torch.load(path, map_location="cpu", weights_only=False)
    if isinstance(obj, dict) and "db" in obj and "test" in obj and "labels" not in obj:
        if part is None:
            raise ValueError(f"{path} holds both sides: pass part='db' or part='test'")
        obj = obj[part]
    if not isinstance(obj, dict) or "labels" not in obj or not any("codes" in k for k in obj):
        raise ValueError(f"{path}: expected a dict with 'labels' and at least one '*codes*' entry")
    return obj


def evaluate_dumps(db, test, R, PRs=(1, 5, 10), threshold=0.0, zero_mean_eval=False, remove_first_retrieved=False,
                   group=None, trusted=False):
    """``db`` / ``test``: paths of code dumps (or already loaded dicts).  Returns ``{'mAP': ..., 'recalls': ...,
    'precisions': ..., 'mAP_<name>': ...}`` exactly as ``RetrievalExperiment.evaluation`` names them."""
    from .hashing import calculate_mAP
    db_out = load_code_dump(db, "db", trusted) if isinstance(db, str) else db
    test_out = load_code_dump(test, "test", trusted) if isinstance(test, str) else test
    names = [k for k in db_out if "codes" in k]                      # train_helper.py:207-214
    res = {}
    for name in names:
        if name not in test_out:
            raise ValueError(f"'{name}' is missing from the query dump")
        postfix = "" if name == "codes" else "_" + name.replace("codes", "").strip("_")
        mAP, recalls, precisions = calculate_mAP(db_out[name], db_out["labels"], test_out[name], test_out["labels"], R,
                                                 threshold=threshold, PRs=list(PRs), zero_mean_eval=zero_mean_eval,
                                                 remove_first_retrieved=remove_first_retrieved, group=group)
        res["mAP" + postfix], res["recalls" + postfix], res["precisions" + postfix] = mAP, recalls, precisions
    return res


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser(description="mAP@R of saved code dumps on the B200 path")
    ap.add_argument("db")
    ap.add_argument("test")
    ap.add_argument("--R", type=int, default=-1)
    ap.add_argument("--PRs", type=int, nargs="*", default=[1, 5, 10])
    ap.add_argument("--threshold", type=float, default=0.0)
    ap.add_argument("--zero-mean-eval", action="store_true")
    ap.add_argument("--trusted", action="store_true", help="allow full unpickling of the dumps (your own files only)")
    a = ap.parse_args()
    print(json.dumps(evaluate_dumps(a.db, a.test, a.R, a.PRs, a.threshold, a.zero_mean_eval, trusted=a.trusted)))
