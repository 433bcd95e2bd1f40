"""Offline inputs: reference code dumps, the packed on-disk format, PackedCodes as evaluator input."""
import numpy as np
import pytest
import torch

from concepthash_b200 import codes_io, synth
from concepthash_b200.evaluator import Evaluator
from oracle import map_oracle as mo
from tests._emu_backend import EmuBackend


def test_packed_roundtrip_and_layout(tmp_path):
    d, dl, _, _, _ = synth.make_random_case(3, 77, 48, 5, p=0.3, seed=1)
    pk = codes_io.PackedCodes.from_codes(d)
    # same bit layout as the sign/bit-pack kernel (emulated): bit k % 32 of word k / 32 = code[k] > 0
    fl = torch.zeros(1, dtype=torch.int32)
    bits, _ = EmuBackend().pack_sign(d, 0.0, fl, False)
    assert torch.equal(bits[:77], pk.bits) and pk.shape == (77, 48)
    path = str(tmp_path / "db.chpk")
    codes_io.save_packed(path, pk, dl, nclass=5)
    pk2, lab2, ncls = codes_io.load_packed(path)
    assert torch.equal(pk2.bits, pk.bits) and pk2.nbit == 48 and torch.equal(lab2, dl) and ncls == 5
    codes_io.save_packed(path, pk)
    assert codes_io.load_packed(path)[1] is None
    with pytest.raises(ValueError):
        codes_io.PackedCodes.from_codes(torch.zeros(2, 8))
    with pytest.raises(ValueError):
        codes_io.PackedCodes(torch.zeros(4, 2, dtype=torch.int32), 16)


def test_code_dump_layouts(tmp_path):
    d, dl, q, ql, _ = synth.make_random_case(4, 30, 16, 3, seed=2)
    db_out, test_out = {"codes": d, "labels": dl, "codes_aux": -d}, {"codes": q, "labels": ql, "codes_aux": -q}
    torch.save(db_out, tmp_path / "db_best.pth")                               # train_helper.py:283
    torch.save({"test": test_out, "db": db_out}, tmp_path / "outputs.pth")     # test_hashing.py:174
    a = codes_io.load_code_dump(str(tmp_path / "db_best.pth"))
    b = codes_io.load_code_dump(str(tmp_path / "outputs.pth"), "db")
    assert torch.equal(a["codes"], d) and torch.equal(b["codes_aux"], -d)
    with pytest.raises(ValueError):
        codes_io.load_code_dump(str(tmp_path / "outputs.pth"))
    torch.save({"x": 1}, tmp_path / "bad.pth")
    with pytest.raises(ValueError):
        codes_io.load_code_dump(str(tmp_path / "bad.pth"))
    # a dump holding arbitrary pickled objects is refused unless the caller vouches for it (no code execution
    # from a crafted file); with trusted=True the legacy path loads it
    import pickle
    import types
    legacy = dict(db_out, labels=dl, extra=types.SimpleNamespace(v=3))
    with open(tmp_path / "legacy.pth", "wb") as f:
        torch.save(legacy, f, pickle_module=pickle)
    with pytest.raises(ValueError, match="trusted"):
        codes_io.load_code_dump(str(tmp_path / "legacy.pth"))
    assert torch.equal(codes_io.load_code_dump(str(tmp_path / "legacy.pth"), trusted=True)["codes"], d)


@pytest.mark.parametrize("tc", [False, True])
def test_packed_codes_as_evaluator_input(tc):
    d, dl, q, ql, _ = synth.make_random_case(9, 300, 64, 4, p=0.3, seed=3)
    be = EmuBackend(rows_per_stripe=64, threads=128, tensor_cores=True) if tc else EmuBackend(rows_per_stripe=64)
    ev = Evaluator(be)
    ev.sample_stride = 0
    pd, pq = codes_io.PackedCodes.from_codes(d), codes_io.PackedCodes.from_codes(q)
    for R in (-1, 20):
        got = ev.evaluate(pd, dl, pq, ql, [R], 0.0, [1, 5], False)
        om, orec, oprec = mo.calculate_mAP(d, dl, q, ql, R, PRs=[1, 5])
        assert np.allclose(got[0], [om], atol=1e-12) and np.allclose(got[1], orec, atol=1e-12)
        assert np.allclose(got[2], oprec, atol=1e-12)
    ids, keys, _ = ev.retrieve(pd[:200], pq, 15)
    oids, odist = mo.topk_ids(q, d[:200], 15)
    assert torch.equal(ids, oids) and torch.equal(keys.float(), odist)
    with pytest.raises(ValueError):
        ev.evaluate(pd, dl, pq, ql, [5], 0.2, [], False)             # a threshold needs magnitudes


@pytest.mark.gpu
def test_evaluate_dumps_on_gpu(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    d, dl, q, ql, ncls = synth.make_random_case(200, 5000, 64, 10, p=0.3, seed=4)
    torch.save({"codes": d, "labels": synth.one_hot(dl, ncls), "codes_neg": -d}, tmp_path / "db_best.pth")
    torch.save({"codes": q, "labels": synth.one_hot(ql, ncls), "codes_neg": -q}, tmp_path / "test_best.pth")
    res = codes_io.evaluate_dumps(str(tmp_path / "db_best.pth"), str(tmp_path / "test_best.pth"), 100)
    om, orec, oprec = mo.calculate_mAP(d, dl, q, ql, 100, PRs=[1, 5, 10])
    assert abs(res["mAP"] - om) < 1e-9 and np.allclose(res["recalls"], orec, atol=1e-9)
    assert abs(res["mAP_neg"] - om) < 1e-9                      # -q vs -d: same Hamming distances
    # packed files: same answer without the real-valued codes
    from concepthash_b200 import hashing
    codes_io.save_packed(str(tmp_path / "db.chpk"), codes_io.PackedCodes.from_codes(d), dl, ncls)
    pk, lab, _ = codes_io.load_packed(str(tmp_path / "db.chpk"))
    m, _, _ = hashing.calculate_mAP(pk, lab, codes_io.PackedCodes.from_codes(q), ql, 100)
    assert abs(m - om) < 1e-9
    pg = hashing.pack_codes(d.cuda())                                # the kernel writes the same bits
    assert torch.equal(pg.bits.cpu(), pk.bits) and pg.nbit == pk.nbit
    ids, dist = hashing.retrieve_topk(codes_io.PackedCodes.from_codes(q), pk.to("cuda"), 50)
    oids, odist = mo.topk_ids(q, d, 50)
    assert torch.equal(ids.cpu(), oids) and torch.equal(dist.cpu(), odist)
