"""GPU parity tests proper: the CUDA path, called through the C-ABI, against the CPU oracle and the
committed golden fixtures.  Bit-exact for packed bits, distances and ranked ids; |mAP delta| <= 1e-9
here (the north star allows 1e-6)."""
import os

import numpy as np
import pytest
import torch

from concepthash_b200 import synth
from oracle import map_oracle as mo

pytestmark = pytest.mark.gpu
TOL = 1e-9
# two exact paths of this library against each other: same ranks, fp64 sums taken in a different order
EPS = 1e-12


def _same(a, b):
    """(maps, recalls, precisions, ap tensor) of two runs agree to EPS"""
    return (np.allclose(a[0], b[0], rtol=0, atol=EPS) and np.allclose(a[1], b[1], rtol=0, atol=EPS) and
            np.allclose(a[2], b[2], rtol=0, atol=EPS) and torch.allclose(a[3], b[3], rtol=0, atol=EPS))


@pytest.fixture(scope="module")
def H():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from concepthash_b200 import hashing
    hashing.get_evaluator()          # raises if the native library is missing: no fallback
    return hashing


def _unpack(bits, nbit):
    b = bits.cpu().numpy().view(np.uint32).view(np.uint8)
    return np.unpackbits(b, axis=1, bitorder="little")[:, :nbit]


# ------------------------------------------------------------------ K1: sign + bit-pack
@pytest.mark.parametrize("nbit", [1, 16, 31, 32, 33, 48, 64, 100, 128, 200, 256])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16, torch.float64])
def test_pack_sign_bits(H, nbit, dtype):
    ev = H.get_evaluator()
    g = torch.Generator().manual_seed(nbit)
    x = torch.randn(257, nbit, generator=g).to(dtype)
    x[3, 0] = 0
    x[200, nbit - 1] = -0.0
    for src in (x, x.cuda()):
        flags = ev.b.zeros((1,), torch.int32)
        bits, nz = ev.b.pack_sign(src, 0.0, flags)
        s = torch.sign(x.float()).numpy()
        assert np.array_equal(_unpack(bits, nbit)[:257], (s > 0).astype(np.uint8))
        assert np.array_equal(_unpack(nz, nbit)[:257], (s != 0).astype(np.uint8))
        assert int(flags.cpu()[0]) == 1
        assert int(bits[257:].abs().sum()) == 0 and int(nz[257:].abs().sum()) == 0   # pad rows are zero
        # padding bits inside the last word are zero too
        words = bits.shape[1]
        full = np.unpackbits(bits.cpu().numpy().view(np.uint32).view(np.uint8), axis=1, bitorder="little")
        assert full[:, nbit:words * 32].sum() == 0


def test_pack_sign_strided_threshold_nan(H):
    ev = H.get_evaluator()
    g = torch.Generator().manual_seed(0)
    base = torch.randn(300, 96, generator=g)
    for view in (base[:, 10:74], base[:, ::2], base[5:205, 3:40], base.t()[:96, :200]):
        for src in (view, view.cuda()):
            flags = ev.b.zeros((1,), torch.int32)
            bits, nz = ev.b.pack_sign(src, 0.5, flags)
            ref = mo.sign_codes(view, 0.5).numpy()
            n, nbit = view.shape
            assert np.array_equal(_unpack(bits, nbit)[:n], (ref > 0).astype(np.uint8))
            assert np.array_equal(_unpack(nz, nbit)[:n], (ref != 0).astype(np.uint8))
    bad = base.clone()
    bad[7, 7] = float("nan")
    with pytest.raises(ValueError, match="NaN"):
        H.calculate_mAP(bad, torch.zeros(300, dtype=torch.long), base[:4], torch.zeros(4, dtype=torch.long), -1)


def test_pack_sign_large_host_chunks(H):
    # > one 64 MB staging chunk from host memory: pipelined H2D path
    ev = H.get_evaluator()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(300_001, 64, generator=g)            # 76.8 MB fp32
    flags = ev.b.zeros((1,), torch.int32)
    bits, _ = ev.b.pack_sign(x.pin_memory(), 0.0, flags)
    bits2, _ = ev.b.pack_sign(x.cuda(), 0.0, flags)
    assert torch.equal(bits, bits2)
    assert np.array_equal(_unpack(bits[:300_001], 64), (x.numpy() > 0).astype(np.uint8))


@pytest.mark.parametrize("nbit", [16, 48, 64, 100, 128, 256])
def test_host_pack_equals_device_pack(H, nbit):
    """PAGEABLE fp32 codes (what trainers/base.py:291-304 hands over) are sign-packed by the host's cores on their
    way into the pinned bounce buffer (host_pack.cpp) -- the bits, the pad rows and the zero / NaN flags must be
    those of the CUDA kernel; more rows than one bounce buffer holds; strided row views (the streamed sample)."""
    ev = H.get_evaluator()
    g = torch.Generator().manual_seed(nbit)
    n = 70_001
    x = torch.randn(n, nbit, generator=g)
    assert not x.is_pinned()
    for poke in (None, "zero", "nan"):
        if poke == "zero":
            x[n // 2, nbit - 1] = 0.0
        if poke == "nan":
            x[n // 3, 0] = float("nan")
        f_host, f_dev = ev.b.zeros((1,), torch.int32), ev.b.zeros((1,), torch.int32)
        host, _ = ev.b.pack_sign(x, 0.0, f_host, want_nz=False)
        dev, _ = ev.b.pack_sign(x.cuda(), 0.0, f_dev, want_nz=False)
        assert torch.equal(host, dev) and int(f_host.cpu()[0]) == int(f_dev.cpu()[0]) == {None: 0, "zero": 1, "nan": 3}[poke]
    y = torch.randn(4000, 4 * nbit, generator=g)
    view = y[:, nbit:2 * nbit]                          # row stride 4 * nbit
    f1, f2 = ev.b.zeros((1,), torch.int32), ev.b.zeros((1,), torch.int32)
    a, _ = ev.b.pack_sign(view, 0.0, f1, want_nz=False)
    b, _ = ev.b.pack_sign(view.cuda(), 0.0, f2, want_nz=False)
    assert torch.equal(a, b)


def test_host_pack_more_rows_than_one_bounce_buffer(H):
    """17 M rows of 16-bit codes = 68 MB of packed words: two passes through the (double-buffered) 64 MB pinned
    bounce buffers, the second host pass overlapping the DMA of the first"""
    ev = H.get_evaluator()
    n = 17_000_000
    x = torch.empty(n, 16).uniform_(-1, 1, generator=torch.Generator().manual_seed(3))
    f1, f2 = ev.b.zeros((1,), torch.int32), ev.b.zeros((1,), torch.int32)
    a, _ = ev.b.pack_sign(x, 0.0, f1, want_nz=False)
    dev = x.cuda()
    del x
    b, _ = ev.b.pack_sign(dev, 0.0, f2, want_nz=False)
    assert torch.equal(a, b)


@pytest.mark.parametrize("ncls", [1, 5, 32, 33, 200, 555])
def test_pack_labels(H, ncls):
    ev = H.get_evaluator()
    g = torch.Generator().manual_seed(ncls)
    ids = torch.randint(ncls, (1000,), generator=g)
    for dtype in (torch.float32, torch.int64, torch.uint8, torch.bool):
        oh = synth.one_hot(ids, ncls).to(dtype)
        for src in (oh, oh.cuda()):
            pid, masks, info = ev.b.pack_labels(src, 0xFFFFFFFF)
            assert torch.equal(pid[:1000].cpu().long(), ids)
            assert info.cpu().tolist()[:3] == [1, int(ids.max()) + 1, 0]
    pid, masks, info = ev.b.pack_labels(ids.cuda(), 0xFFFFFFFE)
    assert masks is None and torch.equal(pid[:1000].cpu().long(), ids)
    multi = (torch.rand(1000, ncls, generator=g) < 0.3).float()
    multi[0] = 0
    pid, masks, info = ev.b.pack_labels(multi, 0xFFFFFFFF)
    got = _unpack(masks, ncls)[:1000]
    assert np.array_equal(got, multi.numpy().astype(np.uint8))
    assert int(info.cpu()[0]) == int(multi.sum(1).max()) and int(info.cpu()[2]) == int((multi.sum(1) == 0).sum())
    assert int(pid[0]) == -1


# ------------------------------------------------------------------ K2: distances, bit-exact
@pytest.mark.parametrize("nbit", [8, 16, 32, 40, 64, 96, 128, 160, 256])
@pytest.mark.parametrize("zeros", [False, True])
def test_hamming_matrix_bit_exact(H, nbit, zeros):
    g = torch.Generator().manual_seed(nbit + zeros)
    q = torch.randn(70, nbit, generator=g)
    d = torch.randn(333, nbit, generator=g)
    if zeros:
        q[torch.rand(q.shape, generator=g) < 0.1] = 0
        d[torch.rand(d.shape, generator=g) < 0.1] = 0
    ref = mo.hamming_distance_matrix(mo.sign_codes(q), mo.sign_codes(d))
    got = H.get_hamm_dist(q.cuda(), d.cuda())
    assert torch.equal(got.cpu(), ref)
    # normalised form: one fp32 division, GPU and CPU division may differ in the last ulp
    assert torch.allclose(H.get_hamm_dist(q, d, normalize=True).cpu(), ref / nbit, rtol=1e-6, atol=0)


# ------------------------------------------------------------------ the drop-in call vs golden fixtures
def _golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "oracle_cases.npz"))
    for name in z["names"]:
        R = z[f"{name}_R"].tolist()
        yield name, z, (R if bool(z[f"{name}_R_is_list"]) else R[0])


@pytest.mark.parametrize("where", ["cpu", "cuda"])
def test_calculate_map_golden(H, golden_dir, where):
    for name, z, R in _golden(golden_dir):
        args = [torch.from_numpy(z[f"{name}_{k}"]).to(where) for k in ("db_codes", "db_labels", "q_codes", "q_labels")]
        snap = [a.clone() for a in args]
        m, rec, prec = H.calculate_mAP(*args, R, threshold=float(z[f"{name}_thr"]), PRs=z[f"{name}_PRs"].tolist(),
                                       remove_first_retrieved=bool(z[f"{name}_rf"]))
        assert isinstance(m, list) == isinstance(R, list)
        m = m if isinstance(m, list) else [m]
        assert all(isinstance(v, float) for v in m + rec + prec)
        assert np.allclose(m, z[f"{name}_mAP"], atol=TOL), name
        assert np.allclose(rec, z[f"{name}_recalls"], atol=TOL), name
        assert np.allclose(prec, z[f"{name}_precisions"], atol=TOL), name
        for a, b in zip(args, snap):
            assert torch.equal(a, b)             # inputs are never mutated


def test_per_query_ap_golden(H, golden_dir):
    ev = H.get_evaluator()
    for name, z, R in _golden(golden_dir):
        args = [torch.from_numpy(z[f"{name}_{k}"]) for k in ("db_codes", "db_labels", "q_codes", "q_labels")]
        r_list = R if isinstance(R, list) else [R]
        _, _, _, ap = ev.evaluate(*args, r_list, float(z[f"{name}_thr"]), z[f"{name}_PRs"].tolist(),
                                  bool(z[f"{name}_rf"]), return_ap=True)
        assert np.allclose(ap.cpu().numpy(), z[f"{name}_aps"], atol=TOL), name


def test_ranked_ids_golden_bit_exact(H, golden_dir):
    for name, z, R in _golden(golden_dir):
        ndb = z[f"{name}_db_codes"].shape[0]
        r = max(x if x > 0 else ndb for x in R) if isinstance(R, list) else R
        ids, dist = H.retrieve_topk(torch.from_numpy(z[f"{name}_q_codes"]), torch.from_numpy(z[f"{name}_db_codes"]),
                                    r, threshold=float(z[f"{name}_thr"]),
                                    remove_first_retrieved=bool(z[f"{name}_rf"]))
        assert np.array_equal(ids.cpu().numpy(), z[f"{name}_ids"].astype(np.int64)), name
        assert np.array_equal((2 * dist).round().cpu().numpy().astype(np.int16), z[f"{name}_dist2"]), name


def test_known_answer(H):
    A, B = [1, 0], [0, 1]
    q = torch.tensor([[1., 1., 1., 1.]])
    d = torch.tensor([[1., 1., 1., 1.], [1., 1., 1., -1.], [1., 1., -1., 1.], [-1., -1., -1., -1.], [1., -1., 1., 1.]])
    ql = torch.tensor([A], dtype=torch.float32)
    dl = torch.tensor([A, B, A, A, B], dtype=torch.float32)
    m, rec, prec = H.calculate_mAP(d, dl, q, ql, -1, PRs=[1, 2])
    assert m == pytest.approx((1 + 2 / 3 + 3 / 5) / 3, abs=1e-15)
    assert prec == pytest.approx([1.0, 0.5]) and rec == pytest.approx([1 / 3, 1 / 3])
    m, _, _ = H.calculate_mAP(d, dl, q, ql, [2, 3, -1, 100])
    assert m == pytest.approx([1.0, (1 + 2 / 3) / 2, (1 + 2 / 3 + 3 / 5) / 3, (1 + 2 / 3 + 3 / 5) / 3])
    m, rec, prec = H.calculate_mAP(d, dl, q, ql, -1, PRs=[2], remove_first_retrieved=True)
    assert m == pytest.approx(0.5) and prec == pytest.approx([0.5]) and rec == pytest.approx([0.5])
    ids, dist = H.retrieve_topk(q, d, -1)
    assert ids.tolist() == [[0, 1, 2, 4, 3]] and dist.tolist() == [[0., 1., 1., 1., 4.]]
    assert H.map_at_r(query_codes=q, db_codes=d, query_labels=ql, db_labels=dl, R=-1) == pytest.approx(m * 0 + (1 + 2 / 3 + 3 / 5) / 3)
    with pytest.raises(NotImplementedError):
        H.calculate_mAP(d, dl, q, ql, -1, dist_metric="euclidean")
    with pytest.raises(ValueError):
        H.calculate_mAP(d, dl, q[:, :3], ql, -1)
    with pytest.raises(ValueError):
        H.calculate_mAP(torch.randn(5, 300), dl, torch.randn(1, 300), ql, -1)      # nbit > 256


# ------------------------------------------------------------------ randomised parity
@pytest.mark.parametrize("seed", range(12))
def test_random_cases_vs_oracle(H, seed):
    rng = np.random.RandomState(seed)
    nbit = int(rng.choice([8, 16, 24, 32, 48, 64, 96, 128, 192, 256]))
    nq, ndb = int(rng.randint(1, 400)), int(rng.randint(1, 3000))
    ncls = int(rng.choice([2, 7, 40, 300]))
    d, dl, q, ql, _ = synth.make_random_case(nq, ndb, nbit, ncls, p=float(rng.choice([0.1, 0.3, 0.5])), seed=seed)
    form = seed % 3
    if form == 0:
        dl, ql = synth.one_hot(dl, ncls), synth.one_hot(ql, ncls)
    elif form == 1:
        g = torch.Generator().manual_seed(seed)
        dl = (synth.one_hot(dl, ncls) + (torch.rand(ndb, ncls, generator=g) < 0.02)).clamp(max=1)
        ql = (synth.one_hot(ql, ncls) + (torch.rand(nq, ncls, generator=g) < 0.02)).clamp(max=1)
    if seed % 4 == 3:
        d[:, ::5] = torch.where(torch.rand(d[:, ::5].shape) < 0.3, 0.0, 1.0) * d[:, ::5]
    R = [-1, 10, 100, 1000, [1, 50, -1]][seed % 5]
    PRs = [[1, 5, 10], [], [3]][seed % 3]
    thr = [0.0, 0.0, 0.3][seed % 3]
    where = "cuda" if seed % 2 else "cpu"
    m, rec, prec = H.calculate_mAP(d.to(where), dl.to(where), q.to(where), ql.to(where), R, threshold=thr, PRs=PRs)
    om, orec, oprec = mo.calculate_mAP(d, dl, q, ql, R, threshold=thr, PRs=PRs)
    assert np.allclose(m, om, atol=TOL) and np.allclose(rec, orec, atol=TOL) and np.allclose(prec, oprec, atol=TOL)
    r = 64 if not isinstance(R, list) and R > 0 else -1
    ids, dist = H.retrieve_topk(q.to(where), d.to(where), r, threshold=thr)
    oids, odist = mo.topk_ids(q, d, r, threshold=thr)
    assert torch.equal(ids.cpu(), oids) and torch.equal(dist.cpu(), odist)


def test_edge_shapes(H):
    for nq, ndb, nbit in [(1, 1, 8), (1, 3, 64), (300, 2, 16), (2, 255, 32), (2, 256, 32), (2, 257, 32), (5, 1025, 128)]:
        d, dl, q, ql, _ = synth.make_random_case(nq, ndb, nbit, 3, seed=nq + ndb)
        for R in (-1, 1, ndb, ndb + 5):
            m, rec, prec = H.calculate_mAP(d, dl, q, ql, R, PRs=[1, 2])
            om, orec, oprec = mo.calculate_mAP(d, dl, q, ql, R, PRs=[1, 2])
            assert np.allclose(m, om, atol=TOL) and np.allclose(rec, orec, atol=TOL) and np.allclose(prec, oprec, atol=TOL)
    # self-retrieval with the query set as its own gallery (test_as_database, test_hashing.py:105-112)
    d, dl, _, _, _ = synth.make_random_case(1, 500, 32, 5, seed=3)
    for R in (-1, 20):
        m, rec, prec = H.calculate_mAP(d, dl, d, dl, R, PRs=[1, 5], remove_first_retrieved=True)
        om, orec, oprec = mo.calculate_mAP(d, dl, d, dl, R, PRs=[1, 5], remove_first_retrieved=True)
        assert np.allclose(m, om, atol=TOL) and np.allclose(rec, orec, atol=TOL) and np.allclose(prec, oprec, atol=TOL)


def test_huge_tie_bucket_and_counter_flush(H):
    """All-identical codes: one bucket holds the whole gallery (> 65535 rows in one stripe), so the
    16-bit shared-memory counters must be flushed mid-stripe and the prefixes must survive it."""
    ev = H.get_evaluator()
    ndb, nq = 150_000, 40
    d = torch.ones(ndb, 32)
    q = torch.ones(nq, 32)
    q[1::2, 0] = -1                      # half the queries at distance 1 from everything
    g = torch.Generator().manual_seed(0)
    dl = torch.randint(50, (ndb,), generator=g)
    ql = torch.randint(50, (nq,), generator=g)
    ev.stripe_rows_override = 75_008     # 2 stripes of > 65535 rows (multiple of 256)
    try:
        for R in (-1, 1000):
            m, rec, prec = H.calculate_mAP(d, dl, q, ql, R, PRs=[1, 10])
            # canonical order = gallery order for every query -> closed form via the oracle on labels only
            rel = (ql[:, None] == dl[None, :]).numpy()
            L = ndb if R == -1 else R
            aps = [mo._ap_from_rel(rel[i, :L]) for i in range(nq)]
            assert abs(m - float(np.mean(aps))) < TOL
            assert np.allclose(prec, [rel[:, :1].mean(), (rel[:, :10].sum(1) / 10).mean()], atol=TOL)
        ids, dist = H.retrieve_topk(q, d, 70_000)
        assert torch.equal(ids.cpu(), torch.arange(70_000).expand(nq, -1))
        assert torch.equal(dist.cpu(), (torch.arange(nq) % 2).float()[:, None].expand(-1, 70_000))
    finally:
        ev.stripe_rows_override = None


# ------------------------------------------------------------------ BASELINE configs at full size
@pytest.mark.parametrize("name,nbit", [("cub200", 64), ("cars196", 16), ("cars196", 32), ("cars196", 64)])
def test_full_size_configs_vs_oracle(H, name, nbit):
    d, dl, q, ql, ncls = synth.make_dataset_case(name, nbit=nbit, p=0.15, seed=0)
    m, rec, prec = H.calculate_mAP(d, synth.one_hot(dl, ncls), q, synth.one_hot(ql, ncls), -1, PRs=[1, 5, 10])
    om, orec, oprec = mo.calculate_mAP(d, dl, q, ql, -1, PRs=[1, 5, 10])
    assert abs(m - om) < TOL and np.allclose(rec, orec, atol=TOL) and np.allclose(prec, oprec, atol=TOL)
    m2, _, _ = H.calculate_mAP(d.cuda(), dl.cuda(), q.cuda(), ql.cuda(), -1)      # 1-D ids, device tensors
    assert m2 == m


def test_nabirds_subset_and_list_path_cross_check(H):
    """cfg3 (NABirds, 24,633 x 23,929, 555 classes): oracle on a 1,500-query subset; on the full query set the
    record path must agree with the independent ranked-list path (ch_ap_from_ranked over retrieve_topk)."""
    d, dl, q, ql, ncls = synth.make_dataset_case("nabirds", nbit=64, p=0.15, seed=0)
    sub = slice(0, 24633, 17)
    m, rec, prec = H.calculate_mAP(d, dl, q[sub], ql[sub], -1, PRs=[1, 5, 10])
    om, orec, oprec = mo.calculate_mAP(d, dl, q[sub], ql[sub], -1, PRs=[1, 5, 10])
    assert abs(m - om) < TOL and np.allclose(rec, orec, atol=TOL) and np.allclose(prec, oprec, atol=TOL)
    ev = H.get_evaluator()
    m_full, _, _, ap = ev.evaluate(d.cuda(), dl.cuda(), q.cuda(), ql.cuda(), [1000], return_ap=True)
    ids, _ = H.retrieve_topk(q.cuda(), d.cuda(), 1000)
    _, gi, _ = ev.b.pack_labels(dl.cuda(), 0xFFFFFFFE)[0], None, None
    qi = ev.b.pack_labels(ql.cuda(), 0xFFFFFFFF)[0]
    gi = ev.b.pack_labels(dl.cuda(), 0xFFFFFFFE)[0]
    cols = ev.b.zeros((q.shape[0], 2), torch.float64)
    ev.b.ap_from_ranked(ids.contiguous(), q.shape[0], 1000, qi, gi, 1, 0, [], cols)
    ap_list = torch.where(cols[:, 1] > 0, cols[:, 0] / cols[:, 1].clamp(min=1), torch.zeros_like(cols[:, 0]))
    assert torch.allclose(ap_list, ap[0], atol=1e-12)
    assert abs(float(ap_list.mean()) - m_full[0]) < 1e-12


def test_cfg4_shape_subset_and_idempotence(H):
    """cfg4 shape (128-bit, 1M gallery, mAP@1000) on a 48-query subset against the oracle, plus
    run-to-run bit reproducibility."""
    d, dl, q, ql, ncls = synth.make_random_case(48, 1_000_000, 128, 101, p=0.30, seed=0, device="cuda")
    m, _, _ = H.calculate_mAP(d, dl, q, ql, 1000)
    m2, _, _ = H.calculate_mAP(d, dl, q, ql, 1000)
    assert m == m2
    ids, dist = H.retrieve_topk(q, d, 1000)
    dc, qc = d.cpu(), q.cpu()
    oids, odist = mo.topk_ids(qc, dc, 1000)
    assert torch.equal(ids.cpu(), oids) and torch.equal(dist.cpu(), odist)
    rel = (ql.cpu()[:, None] == dl.cpu()[oids]).numpy()
    om = float(np.mean([mo._ap_from_rel(r) for r in rel]))
    assert abs(m - om) < TOL


def _check_subset_against_packed_oracle(H, d, dl, q, ql, R, ap_full, nsub=32):
    """AP rows of a full-size run and the ranked ids of the same queries against the packed-bit CPU oracle
    (oracle/packed_oracle.py: popcount + stable (distance, row) order over the WHOLE gallery)."""
    from oracle import packed_oracle as po
    nq, nbit = q.shape
    sub = torch.unique(torch.linspace(0, nq - 1, nsub).round().long())
    gb = po.pack_sign_bits(d.cpu().numpy())
    qb = po.pack_sign_bits(q[sub.to(q.device)].cpu().numpy())
    oids, odist = po.topk_packed(qb, gb, R, nbit)
    oap = po.ap_of_ranked(oids, ql.cpu().numpy()[sub.numpy()], dl.cpu().numpy())
    got = ap_full[0].cpu().numpy()[sub.numpy()]
    assert np.abs(got - oap).max() < TOL, np.abs(got - oap).max()
    ids, dist = H.retrieve_topk(q[sub.to(q.device)], d, R)
    assert np.array_equal(ids.cpu().numpy(), oids)
    assert np.array_equal(dist.cpu().numpy().astype(np.int32), odist)
    return float(oap.mean())


def test_cfg4_full_query_set_vs_oracle(H):
    """BASELINE configs[3] exactly as bench.py times it (128-bit, 25,000 queries x 1,000,000 rows, mAP@1000: 49 query
    groups x several stripes, dense epilogue, two-level sample): per-query AP of the FULL run, on 32 queries spread
    over all query groups, against the CPU oracle; ranked ids of those queries bit-exact."""
    ev = H.get_evaluator()
    d, dl, q, ql, ncls = synth.make_random_case(25_000, 1_000_000, 128, 101, p=0.30, seed=0, device="cuda")
    maps, _, _, ap = ev.evaluate(d, dl, q, ql, [1000], 0.0, [], False, return_ap=True)
    assert ev.stats["mode"] == "topR-sampled" and ev.stats["select_kernel"] == "tcgen05", ev.stats
    assert abs(float(ap[0].mean()) - maps[0]) < 1e-12
    _check_subset_against_packed_oracle(H, d, dl, q, ql, 1000, ap)


def test_cfg5_shard_geometry_vs_oracle(H):
    """BASELINE configs[4] geometry (64-bit, 1000 classes, top-R = 1000; configs/dataset/inat_birds.yaml:4) at a
    size the oracle can check: 20,500 queries (41 query groups, a partial last tile) x 5,100,077 rows (KB = 96
    kernel variant, sparse epilogue, stride-64 sample: candidates are < 0.02 % of the gallery, as in cfg5)."""
    ev = H.get_evaluator()
    d, dl, q, ql, ncls = synth.make_random_case(20_500, 5_100_077, 64, 1000, p=0.30, seed=0, device="cuda")
    maps, rec, prec, ap = ev.evaluate(d, dl, q, ql, [1000], 0.0, [1, 10], False, return_ap=True)
    assert ev.stats["mode"] == "topR-sampled" and ev.stats["select_kernel"] == "tcgen05", ev.stats
    assert not ev.stats.get("select_dense") and ev.stats["sample"]["stride"] == 2 * ev.sample_stride, ev.stats
    _check_subset_against_packed_oracle(H, d, dl, q, ql, 1000, ap)
    # the same call through the drop-in surface: a SPECULATIVE re-evaluation of the shape (one host round trip),
    # bit-identical numbers
    m, rec2, prec2 = H.calculate_mAP(d, dl, q, ql, 1000, PRs=[1, 10])
    assert ev.stats["speculation"] == "hit" and ev.stats["host_syncs"] == 1, ev.stats
    assert m == maps[0] and rec2 == rec and prec2 == prec


def test_speculative_reevaluation_on_gpu(H):
    """Second evaluation of a shape: no mid-flight host round trips, same numbers; new data of the same shape is
    either covered by the hint or detected on the device and repeated -- never wrong.  All modes."""
    ev = H.get_evaluator()
    for nq, ndb, nbit, ncls, R in [(3000, 300_000, 64, 50, 100), (700, 9000, 32, 20, -1), (900, 40_000, 128, 30, 500)]:
        runs = []
        for seed in (5, 5, 6, 7):
            d, dl, q, ql, _ = synth.make_random_case(nq, ndb, nbit, ncls, p=0.30 if seed < 7 else 0.45, seed=seed,
                                                     device="cuda")
            out = ev.evaluate(d, dl, q, ql, [R], 0.0, [1, 10], False, return_ap=True)
            runs.append((out, dict(ev.stats)))
            sub = slice(0, 24)
            om, orec, oprec, oaps = mo.calculate_mAP(d.cpu(), dl.cpu(), q[sub].cpu(), ql[sub].cpu(), R, PRs=[1, 10],
                                                     return_per_query=True)
            assert np.allclose(out[3][0][sub].cpu().numpy(), oaps[0], atol=TOL), (nq, seed, ev.stats)
        assert runs[1][1]["speculation"] == "hit" and runs[1][1]["host_syncs"] == 1, runs[1][1]
        assert _same(runs[0][0], runs[1][0])
        assert all(r[1]["speculation"] in ("hit", "retried", "graph") for r in runs[2:])


def test_cuda_graph_replay_of_small_evaluations(H):
    """Third and later evaluations of a small shape are ONE graph launch: same numbers as the eager path on new data
    of the same shape, in every mode; data that contradicts what the graph assumed (an exact zero -> ternary keys; a
    gallery whose order defeats the sample) is noticed on the device and re-evaluated by the ordinary path."""
    ev = H.get_evaluator()
    assert ev.use_graphs
    for nq, ndb, nbit, ncls, R in [(700, 9000, 32, 20, -1), (3000, 200_000, 64, 50, 100), (900, 40_000, 128, 30, 500),
                                   (5794, 5994, 64, 200, -1)]:
        ev._hints.clear()
        ev._graphs.clear()
        seen = []
        for seed in (1, 2, 3, 4, 5):
            d, dl, q, ql, _ = synth.make_random_case(nq, ndb, nbit, ncls, p=0.30, seed=seed, device="cuda")
            out = ev.evaluate(d, dl, q, ql, [R], 0.0, [1, 10], False)
            seen.append(ev.stats["speculation"])
            ev.use_graphs = False
            try:
                ref = ev.evaluate(d, dl, q, ql, [R], 0.0, [1, 10], False)
                assert ev.stats["speculation"] != "graph"
            finally:
                ev.use_graphs = True
            assert out == ref, (nq, seed, out, ref)
        assert seen[0] == "none" and "graph" in seen[2:], seen
        sub = slice(0, 24)
        om, orec, oprec = mo.calculate_mAP(d.cpu(), dl.cpu(), q[sub].cpu(), ql[sub].cpu(), R, PRs=[1, 10])
        m, rec, prec = H.calculate_mAP(d, dl, q[sub], ql[sub], R, PRs=[1, 10])
        assert abs(m - om) < TOL and np.allclose(rec, orec, atol=TOL) and np.allclose(prec, oprec, atol=TOL)
        # an exact zero in the codes: the graph's status block says "ternary keys" -> ordinary path, right answer
        dz = d.clone()
        dz[7, 3] = 0.0
        out = ev.evaluate(dz, dl, q, ql, [R], 0.0, [1, 10], False)
        assert ev.stats["ternary"] and ev.stats["speculation"] != "graph", ev.stats
        om, orec, oprec = mo.calculate_mAP(dz.cpu(), dl.cpu(), q[sub].cpu(), ql[sub].cpu(), R, PRs=[1, 10])
        m, rec, prec = H.calculate_mAP(dz, dl, q[sub], ql[sub], R, PRs=[1, 10])
        assert abs(m - om) < TOL


def test_query_chunking_on_gpu(H):
    """More slots than the 32-bit slot index allows (forced by a small limit): the query set is evaluated in chunks
    and the result is that of the one-shot run (to the summation order of the final mean)."""
    ev = H.get_evaluator()
    d, dl, q, ql, _ = synth.make_random_case(3000, 300_000, 64, 50, p=0.30, seed=11, device="cuda")
    ref = ev.evaluate(d, dl, q, ql, [100, 1000], 0.0, [1, 10], False, return_ap=True)
    saved = ev.max_slots
    try:
        ev.max_slots = ev.stats["record_slots"] // 3
        ev._hints.clear()
        out = ev.evaluate(d, dl, q, ql, [100, 1000], 0.0, [1, 10], False, return_ap=True)
        assert ev.stats.get("query_chunks", 0) >= 2, ev.stats
    finally:
        ev.max_slots = saved
        ev._hints.clear()
    assert _same(out, ref)


def test_sampled_one_pass_equals_exact_two_pass(H):
    """Top-R with the sample-derived threshold (one full pass) must be bit-identical to the exact two-pass
    path, and must fall back when the gallery order defeats the sample."""
    ev = H.get_evaluator()
    d, dl, q, ql, ncls = synth.make_random_case(3000, 300_000, 64, 50, p=0.30, seed=5, device="cuda")
    res = {}
    default_stride = ev.sample_stride
    for stride in (default_stride, 16, 0):
        ev.sample_stride = stride
        try:
            res[stride] = ev.evaluate(d, dl, q, ql, [100, 1000], 0.0, [1, 5, 10], False, return_ap=True)
            mode = ev.stats["mode"]
        finally:
            ev.sample_stride = default_stride
        assert mode == ("topR-sampled" if stride else "topR")
    assert _same(res[16], res[0]) and _same(res[default_stride], res[0])
    # thresholds through the two-level sample (select pass over the sample itself), forced at this size, and off
    saved = (ev.sample2_min_rows, ev.sample2_min_work, ev.sample_two_level)
    try:
        ev.sample2_min_rows = ev.sample2_min_work = 0
        r2 = ev.evaluate(d, dl, q, ql, [100, 1000], 0.0, [1, 5, 10], False, return_ap=True)
        assert ev.stats["mode"] == "topR-sampled" and "sample2" in ev.stats, ev.stats
        ev.sample_two_level = False
        r1 = ev.evaluate(d, dl, q, ql, [100, 1000], 0.0, [1, 5, 10], False, return_ap=True)
        assert ev.stats["mode"] == "topR-sampled" and "sample2" not in ev.stats, ev.stats
    finally:
        ev.sample2_min_rows, ev.sample2_min_work, ev.sample_two_level = saved
    assert _same(r2, res[0]) and _same(r1, res[0])
    # adversarial: the row sample sees only near duplicates, the rest of the gallery is far away
    n = 400_000
    st = default_stride
    dd = -torch.ones(n, 32, device="cuda")
    dd[0:st * 900:st] = 1.0                     # 900 near rows, all on sampled positions
    qq = torch.ones(64, 32, device="cuda")
    ll = torch.arange(n, device="cuda") % 7
    ql2 = torch.arange(64, device="cuda") % 7
    m, rec, prec = H.calculate_mAP(dd, ll, qq, ql2, 1000, PRs=[1, 10])
    # every query fails the sampled pass: re-ranked by the exact path (per-query repair for few queries, else the
    # whole evaluation)
    assert ev.stats["sample"].get("repaired_queries") == 64 or ev.stats["sample"].get("fallback"), ev.stats
    rel = (ql2[:, None] == ll[None, :]).cpu().numpy()
    order = np.concatenate([np.arange(0, st * 900, st), np.setdiff1d(np.arange(n), np.arange(0, st * 900, st))])[:1000]
    aps = [mo._ap_from_rel(rel[i, order]) for i in range(64)]
    assert abs(m - float(np.mean(aps))) < TOL


@pytest.mark.parametrize("nbit", [32, 64, 128, 256])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_pack_sign_flat_fast_path(H, nbit, dtype):
    """Contiguous device codes without a non-zero plane take pack_sign_flat_kernel: same bits as the general
    kernel, zeros / NaNs still reported in the flags, pad rows zero."""
    ev = H.get_evaluator()
    g = torch.Generator().manual_seed(nbit)
    for n in (31, 32, 1000, 4099):
        x = torch.randn(n, nbit, generator=g).to(dtype).cuda()
        f1, f2 = ev.b.zeros((1,), torch.int32), ev.b.zeros((1,), torch.int32)
        fast, none = ev.b.pack_sign(x, 0.0, f1, want_nz=False)
        slow, _ = ev.b.pack_sign(x, 0.0, f2, want_nz=True)
        assert none is None and torch.equal(fast, slow) and int(f1.cpu()[0]) == 0
        x[n // 2, 3] = 0
        ev.b.pack_sign(x, 0.0, f1, want_nz=False)
        assert int(f1.cpu()[0]) == 1


@pytest.mark.parametrize("nbit,thr", [(16, 0.0), (31, 0.0), (32, 0.0), (48, 0.0), (64, 0.0), (96, 0.0), (100, 0.0),
                                      (127, 0.0), (128, 0.0), (129, 0.0), (160, 0.0), (192, 0.0), (200, 0.0),
                                      (224, 0.0), (255, 0.0), (256, 0.0),
                                      (32, 0.3), (64, 0.3), (100, 0.3), (128, 0.3), (256, 0.3)])
def test_tensor_core_select_equals_popc_select(H, nbit, thr):
    """The tcgen05 (int8 {-1, 0, +1}, UTCIMMA) select pass + candidate-list ranking must give the results of the
    XOR+POPC select pass + record ranking: same AP per query (to fp64 summation order), bit-identical ranked ids --
    and both match the oracle on a query subset.  nbit up to 256 (KB up to 288: 2-stage ring, four threshold
    slots); ``thr`` = configs/val.yaml:12 ``ternary_threshold`` (experiments/test_hashing.py:109): ~16 % of the
    signs become 0, keys are 2 x distance."""
    ev = H.get_evaluator()
    nq, ndb = 700, 260_000 + nbit          # tail tile, several stripes, inactive query lanes in the last tile
    if nbit > 128 or thr:
        ndb = 150_000 + nbit               # (the POPC reference pass of wide / ternary codes is slow)
    d, dl, q, ql, ncls = synth.make_random_case(nq, ndb, nbit, 30, p=0.30, seed=nbit, device="cuda")
    out = {}
    has_pair = ev.b.tc_code_bytes_pair(nbit, bool(thr)) > 0      # two gallery rows per accumulator cell (keys <= 128)
    variants = [(True, False, False), (True, True, False), (False, None, False)]
    if has_pair:
        variants += [(True, False, True), (True, True, True)]
    for tc, dense, pair in variants:
        ev.use_tensor_cores, ev.select_dense_override, ev.paired_rows = tc, dense, pair
        try:
            res = ev.evaluate(d, dl, q, ql, [50, 500], thr, [1, 5, 10], False, return_ap=True)
            kern = ev.stats["select_kernel"]
            assert ev.stats["ternary"] == bool(thr)
            assert not tc or ev.stats["select_rows_per_cell"] == (2 if pair else 1), ev.stats
            ids, keys, _ = ev.retrieve(d, q, 300, thr)
            out[(tc, dense, pair)] = (res, ids, keys)
        finally:
            ev.use_tensor_cores, ev.select_dense_override, ev.paired_rows = True, None, True
        assert kern == ("tcgen05" if tc else "popc")
    rb, ib, kb = out[(False, None, False)]
    for key in variants[:2] + variants[3:]:     # both epilogue variants of the tensor-core kernel, one / two rows per cell
        ra, ia, ka = out[key]
        assert _same(ra, rb), key
        assert torch.equal(ia, ib) and torch.equal(ka, kb), key
    sub = slice(0, 40)
    om, orec, oprec = mo.calculate_mAP(d.cpu(), dl.cpu(), q[sub].cpu(), ql[sub].cpu(), [50, 500], threshold=thr,
                                       PRs=[1, 5, 10])
    m, rec, prec = H.calculate_mAP(d, dl, q[sub], ql[sub], [50, 500], threshold=thr, PRs=[1, 5, 10])
    assert np.allclose(m, om, atol=TOL) and np.allclose(rec, orec, atol=TOL) and np.allclose(prec, oprec, atol=TOL)
    oids, odist = mo.topk_ids(q[sub].cpu(), d.cpu(), 300, threshold=thr)
    assert torch.equal(out[(True, False, False)][1][sub].cpu(), oids)
    assert torch.equal(out[(True, False, False)][2][sub].cpu().float() * (0.5 if thr else 1.0), odist)


@pytest.mark.parametrize("nbit,tern", [(128, False), (64, False), (100, False), (16, False), (7, False), (64, True),
                                       (40, True)])
def test_paired_select_kernel_threshold_edges(H, nbit, tern):
    """The select kernel alone, paired form (two gallery rows per accumulator cell) against brute force, on the
    thresholds where ``thresh - key`` touches the ends of a signed byte: 0, 127, 128 (= every row), beyond the
    largest key; queries of all ones / all minus ones (P = nbit / 0 in the slot constants); gallery rows equal to a
    query and to its complement (keys 0 and the maximum); a partial last tile; padding queries."""
    from concepthash_b200.evaluator import Packed
    ev = H.get_evaluator()
    b = ev.b
    assert b.tc_code_bytes_pair(nbit, tern) > 0
    g = torch.Generator().manual_seed(nbit + tern)
    nq, ndb, rps = 200, 3000, 2048
    qx = torch.randn(nq, nbit, generator=g)
    gx = torch.randn(ndb, nbit, generator=g)
    qx[0], qx[1] = 1.0, -1.0
    gx[5], gx[6], gx[2999], gx[2047], gx[2048] = qx[3], -qx[3], qx[4], -qx[4], qx[0]
    thr = 0.4 if tern else 0.0
    ip = mo.sign_codes(qx, thr) @ mo.sign_codes(gx, thr).T
    keys = (nbit - ip) if tern else (nbit - ip) / 2           # ternary: 2 x distance
    kmax = 2 * nbit if tern else nbit
    th = torch.randint(0, kmax + 2, (nq,), generator=g)
    th[:8] = torch.tensor([0, kmax, 127, 128, 200, 1, kmax - 1, kmax + 1]).clamp(max=1000)
    fl = b.zeros((1,), torch.int32)
    packs = []
    for x in (qx, gx):
        p = Packed()
        p.n, p.nbit = x.shape[0], nbit
        p.bits, p.nz = b.pack_sign(x.cuda(), thr, fl, want_nz=tern)
        p.i8 = None
        packs.append(p)
    qp, gp = packs
    nq_pad, nstripes = 256, 2
    thresh = b.zeros((nq_pad,), torch.int32)
    thresh[:nq] = th.to(torch.int32).cuda()
    off = (torch.arange(nstripes * nq_pad, dtype=torch.int64) * rps).to(torch.int32).view(nstripes, nq_pad).cuda()
    cap = torch.full((nstripes, nq_pad), rps, dtype=torch.int32, device="cuda")
    want = keys <= th[:, None].float()
    try:
        for pair in (True, False):
            for dense in (False, True):
                ev.paired_rows = pair
                gp.i8 = gp.i8b = gp.i8p = qp.i8b = None
                cand = dict(off=off, cap=cap, cnt=b.zeros((nstripes, nq_pad), torch.int32),
                            rows=torch.full((nstripes * nq_pad * rps,), -1, dtype=torch.int32, device="cuda"),
                            err=b.zeros((1,), torch.int32))
                ev._select_tc(qp, gp, (128, nq_pad, nstripes, rps), thresh, cand, dense)
                cnt, rows = cand["cnt"].cpu(), cand["rows"].cpu().view(nstripes, nq_pad, rps)
                assert int(cand["err"].cpu()[0]) == 0
                for qi in range(nq):
                    got = torch.cat([rows[s, qi, :cnt[s, qi]] for s in range(nstripes)])
                    exp = torch.nonzero(want[qi]).flatten().to(torch.int32)
                    assert torch.equal(got, exp), (pair, dense, qi, int(th[qi]), got[:8], exp[:8])
    finally:
        ev.paired_rows = True


def test_streamed_host_gallery_equals_resident(H):
    """A large gallery passed as a HOST tensor is streamed in row blocks behind the select pass; the answer must
    be bit-identical to the device-resident run, and a gallery with exact zeros must fall back cleanly."""
    ev = H.get_evaluator()
    for nbit, nq, ndb in [(64, 1500, 420_000), (128, 900, 300_123)]:
        d, dl, q, ql, ncls = synth.make_random_case(nq, ndb, nbit, 40, p=0.30, seed=nbit + 1, device="cuda")
        ref = ev.evaluate(d, dl, q, ql, [100, 1000], 0.0, [1, 5, 10], False, return_ap=True)
        assert ev.stats["mode"] == "topR-sampled"
        for host in (d.cpu(), d.cpu().pin_memory()):
            out = ev.evaluate(host, dl.cpu(), q.cpu(), ql.cpu(), [100, 1000], 0.0, [1, 5, 10], False, return_ap=True)
            assert ev.stats["mode"] == "topR-sampled-streamed", ev.stats
            assert _same(out, ref)
    # zeros in the gallery -> ternary keys: the streamed attempt must notice and restart on the regular path
    dz = d.cpu().clone()
    dz[::1000, 5] = 0.0
    m, rec, prec = H.calculate_mAP(dz, dl.cpu(), q.cpu(), ql.cpu(), 200, PRs=[1, 10])
    assert ev.stats["ternary"] and not ev.stats["mode"].endswith("streamed")
    sub = slice(0, 30)
    om, orec, oprec = mo.calculate_mAP(dz, dl.cpu(), q.cpu()[sub], ql.cpu()[sub], 200, PRs=[1, 10])
    m2, rec2, prec2 = H.calculate_mAP(dz, dl.cpu(), q.cpu()[sub], ql.cpu()[sub], 200, PRs=[1, 10])
    assert abs(m2 - om) < TOL and np.allclose(rec2, orec, atol=TOL) and np.allclose(prec2, oprec, atol=TOL)


def test_native_gallery_loader(H):
    """csrc/loader.cu: the whole host gallery packed by a native thread into a ring of pinned chunks and copied chunk by
    chunk -- bits, pad rows and flags equal to the pack kernel's; a ring smaller than the gallery (slots are
    refilled); strided rows; waits for row prefixes; and the evaluation paths around it: Python loader (native off),
    column-slice views, a host gallery that is not streamed after all (R = all), NaN."""
    import os
    ev = H.get_evaluator()
    b = ev.b
    side = torch.cuda.Stream()
    g = torch.Generator().manual_seed(5)
    for n, nbit, ring in [(300_123, 128, None), (1_000_003, 64, 1 << 20), (70_001, 48, None), (5, 32, None)]:
        x = torch.randn(n, nbit, generator=g)
        wide = torch.randn(n, nbit + 24, generator=g)
        x[x == 0], wide[wide == 0] = 1.0, 1.0      # (the CPU generator does produce exact zeros, ~6e-8 of its samples)
        for src, poke in ((x, None), (wide[:, 8:8 + nbit], None), (x.pin_memory(), None), (x.clone(), "zero"),
                          (x.clone(), "nan")):
            if poke == "zero":
                src[n // 2, 1] = 0.0
            elif poke == "nan":
                src[n - 1, nbit - 1] = float("nan")
            assert b.host_loader_ok(src)
            bits = torch.full((b.padded_rows(n), b.code_words(nbit)), -1, dtype=torch.int32, device="cuda")
            flags = torch.zeros(1, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            if ring is not None:
                os.environ["CH_LOADER_RING_BYTES"] = str(ring)
            # a second, short job in front (as the evaluator queues the queries in front of the gallery)
            small = x[:min(n, 3000)]
            bits0 = torch.full((b.padded_rows(small.shape[0]), b.code_words(nbit)), -1, dtype=torch.int32, device="cuda")
            flags0 = torch.zeros(1, dtype=torch.int32, device="cuda")
            ids = torch.randint(0, 1 << 40, (max(n // 3, 7),), generator=g)          # raw copy job (label ids)
            ids_dev = torch.zeros_like(ids, device="cuda")
            try:
                ld = b.host_loader_start([(small, bits0, flags0), ("copy", ids, ids_dev), (src, bits, flags)], side)
            finally:
                os.environ.pop("CH_LOADER_RING_BYTES", None)
            cur = torch.cuda.current_stream()
            ld.wait(2, min(n, 1000), cur)
            head = bits[:min(n, 1000)].clone()
            ld.wait(0, small.shape[0], cur, block=poke is None)
            first = bits0.clone()
            ld.wait(1, ids.numel(), cur)
            assert torch.equal(ids_dev.cpu(), ids)
            ld.wait(2, n, cur)
            done = bits.clone()
            fl = ld.join()
            assert ld.join() == fl and fl[0] == 0 and fl[1] == 0     # idempotent
            fl = [fl[0], fl[2]]
            f_dev = torch.zeros(1, dtype=torch.int32, device="cuda")
            ref, _ = b.pack_sign(src.cuda(), 0.0, f_dev, False)
            ref0, _ = b.pack_sign(small.cuda(), 0.0, f_dev.clone(), False)
            assert torch.equal(done, ref) and torch.equal(head, ref[:min(n, 1000)]) and torch.equal(first, ref0)
            assert fl[1] == int(f_dev.cpu()[0]) == int(flags.cpu()[0]) == {None: 0, "zero": 1, "nan": 2}[poke], (n, poke)
            assert int(flags0.cpu()[0]) == 0
    assert not b.host_loader_ok(x.double()) and not b.host_loader_ok(x.cuda()) and not b.host_loader_ok(x.t())
    # a ring of 4 slots, many times: threads wait for slots all the time and depend on thread 0 sending chunks to the
    # very end (it used to leave with the last piece: a rare hang)
    big = torch.randn(400_000, 64, generator=g)
    big[big == 0] = 1.0
    bits = torch.empty((b.padded_rows(big.shape[0]), 2), dtype=torch.int32, device="cuda")
    ref, _ = b.pack_sign(big.cuda(), 0.0, torch.zeros(1, dtype=torch.int32, device="cuda"), False)
    os.environ["CH_LOADER_RING_BYTES"] = str(1 << 20)
    try:
        for rep in range(100):
            ld = b.host_loader_start([(big[:2000], torch.empty((b.padded_rows(2000), 2), dtype=torch.int32, device="cuda"),
                                       None), (big, bits, None)], side)
            ld.wait(1, big.shape[0], torch.cuda.current_stream())
            if rep % 25 == 0:
                assert torch.equal(bits, ref)
            assert ld.join() == [0, 0]
    finally:
        os.environ.pop("CH_LOADER_RING_BYTES", None)
    # ---- the evaluation around it
    d, dl, q, ql, ncls = synth.make_random_case(1200, 330_000, 64, 40, p=0.30, seed=77, device="cuda")
    ref = ev.evaluate(d, dl, q, ql, [100, 1000], 0.0, [1, 5, 10], False, return_ap=True)
    hd, hdl, hq, hql = d.cpu(), dl.cpu(), q.cpu(), ql.cpu()
    wide = torch.randn(330_000, 96)
    wide[:, 16:80] = hd
    for native in (True, False):
        ev.stream_native_loader = native
        try:
            for host in (hd, wide[:, 16:80]):
                out = ev.evaluate(host, hdl, hq, hql, [100, 1000], 0.0, [1, 5, 10], False, return_ap=True)
                assert ev.stats["mode"] == "topR-sampled-streamed", ev.stats
                assert _same(out, ref)
        finally:
            ev.stream_native_loader = True
    # the knobs of the native path that were measured and left off still give the same answer
    for knob, val in (("stream_cand_overlap", True), ("stream_late_labels", True), ("stream_fused_rank", True),
                      ("stream_loader_labels", True), ("stream_chunks_native", 2), ("stream_chunks_native", 5)):
        saved = getattr(ev, knob)
        setattr(ev, knob, val)
        try:
            for rep in range(2):                    # without and with a hint
                out = ev.evaluate(hd, hdl, hq, hql, [100, 1000], 0.0, [1, 5, 10], False, return_ap=True)
                assert ev.stats["mode"] == "topR-sampled-streamed" and _same(out, ref), (knob, rep)
        finally:
            setattr(ev, knob, saved)
    # other host dtypes keep the staged path (Python loader thread) -- before and after a native evaluation
    out = ev.evaluate(hd.half(), hdl, hq.half(), hql, [100, 1000], 0.0, [1, 5, 10], False, return_ap=True)
    assert ev.stats["mode"] == "topR-sampled-streamed" and _same(out, ref)
    out = ev.evaluate(hd, hdl, hq.double(), hql, [100, 1000], 0.0, [1, 5, 10], False, return_ap=True)
    assert ev.stats["mode"] == "topR-sampled-streamed" and _same(out, ref)
    # not streamed after all: the loader's bits are the shard's bits
    ref_all = ev.evaluate(d, dl, q[:200], ql[:200], [-1], 0.0, [1, 10], False)
    out_all = ev.evaluate(hd, hdl, hq[:200], hql[:200], [-1], 0.0, [1, 10], False)
    assert ev.stats["mode"] == "all" and all(np.allclose(a, r, rtol=0, atol=EPS) for a, r in zip(out_all, ref_all))
    hn = hd.clone()
    hn[123_456, 7] = float("nan")
    with pytest.raises(ValueError, match="NaN"):
        ev.evaluate(hn, hdl, hq, hql, [100], 0.0, [], False)
    with pytest.raises(ValueError, match="NaN"):
        ev.evaluate(hn, hdl, hq[:50], hql[:50], [-1], 0.0, [], False)
    assert ev._loader is None
    out = ev.evaluate(hd, hdl, hq, hql, [100, 1000], 0.0, [1, 5, 10], False, return_ap=True)     # and it still works
    assert _same(out, ref)


# ------------------------------------------------------------------ zero_mean_eval fused into K1
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.float16])
def test_zero_mean_eval_fused(H, dtype):
    """`calculate_mAP(..., zero_mean_eval=True)` == the callers' `db -= db.mean(0); test -= db_mean` followed by the
    plain call (train_helper.py:223-226), for device and host inputs, mAP@all and top-R; the column sums are
    deterministic."""
    ev = H.get_evaluator()
    d, dl, q, ql, ncls = synth.make_random_case(300, 6000, 64, 20, p=0.3, seed=3)
    d, q = (d + 0.4).to(dtype), (q + 0.4).to(dtype)
    dz, qz = mo.zero_mean(d, q)
    s1, s2 = ev.b.column_sums(d.cuda()), ev.b.column_sums(d.cuda())
    assert torch.equal(s1, s2)
    assert torch.allclose(s1.cpu(), d.double().sum(0), rtol=1e-12, atol=1e-9)
    for R in (-1, 100):
        om, orec, oprec = mo.calculate_mAP(dz, dl, qz, ql, R, PRs=[1, 5, 10])
        for dev in ("cuda", "cpu"):
            m, rec, prec = H.calculate_mAP(d.to(dev), dl.to(dev), q.to(dev), ql.to(dev), R, PRs=[1, 5, 10],
                                           zero_mean_eval=True)
            assert abs(m - om) < TOL and np.allclose(rec, orec, atol=TOL) and np.allclose(prec, oprec, atol=TOL)
        m0, _, _ = H.calculate_mAP(d.cuda(), dl.cuda(), q.cuda(), ql.cuda(), R)
        assert abs(m0 - m) > 1e-6                     # the offset matters on this data
    # strided bit slice (sub_code_eval, test_hashing.py:87-92) + zero mean + ternary threshold
    ds, qs = d[:, 8:40], q[:, 8:40]
    dzs, qzs = mo.zero_mean(ds, qs)
    om, _, _ = mo.calculate_mAP(dzs, dl, qzs, ql, 50, threshold=0.1)
    m, _, _ = H.calculate_mAP(ds.cuda(), dl.cuda(), qs.cuda(), ql.cuda(), 50, threshold=0.1, zero_mean_eval=True)
    assert abs(m - om) < TOL


def test_tensor_core_path_tiny_shapes(H):
    """one query, a handful of rows, R larger than the gallery: the tensor-core select pass with mostly padding"""
    ev = H.get_evaluator()
    for nq, ndb, nbit, R in [(1, 300, 64, 10), (3, 129, 128, 5), (130, 1000, 32, 50), (5, 257, 16, 400)]:
        d, dl, q, ql, ncls = synth.make_random_case(nq, ndb, nbit, 4, p=0.3, seed=nq + ndb)
        m, rec, prec = H.calculate_mAP(d.cuda(), dl.cuda(), q.cuda(), ql.cuda(), R, PRs=[1, 5])
        om, orec, oprec = mo.calculate_mAP(d, dl, q, ql, R, PRs=[1, 5])
        assert abs(m - om) < TOL and np.allclose(rec, orec, atol=TOL) and np.allclose(prec, oprec, atol=TOL), (nq, ndb)
        ids, dist = H.retrieve_topk(q.cuda(), d.cuda(), min(R, 40))
        oids, odist = mo.topk_ids(q, d, min(R, 40))
        assert torch.equal(ids.cpu(), oids) and torch.equal(dist.cpu(), odist)


# ------------------------------------------------------------------ calculate_pr_curve
def test_pr_curve_matches_oracle(H):
    """second symbol imported at experiments/test_hashing.py:15 (used at :152-168): recall / precision at the
    default power-of-two cut-offs (more than CH_MAX_PR of them: several passes) and at explicit ones, with and
    without remove_first_retrieved, device and host inputs."""
    d, dl, q, ql, ncls = synth.make_random_case(150, 9000, 48, 12, p=0.3, seed=21)
    for rf in (False, True):
        qq, qql = (d[:150].clone(), dl[:150].clone()) if rf else (q, ql)
        orec, oprec, ors = mo.calculate_pr_curve(d, dl, qq, qql, remove_first_retrieved=rf)
        for dev in ("cuda", "cpu"):
            rec, prec, rs = H.calculate_pr_curve(d.to(dev), dl.to(dev), qq.to(dev), qql.to(dev),
                                                 remove_first_retrieved=rf)
            assert rs == ors and np.allclose(rec, orec, atol=TOL) and np.allclose(prec, oprec, atol=TOL)
    cuts = list(range(1, 80, 2))                                  # 40 cut-offs -> two passes of <= 32
    orec, oprec, _ = mo.calculate_pr_curve(d, dl, q, ql, Rs=cuts)
    rec, prec, rs = H.calculate_pr_curve(d.cuda(), dl.cuda(), q.cuda(), ql.cuda(), Rs=cuts)
    assert rs == cuts and np.allclose(rec, orec, atol=TOL) and np.allclose(prec, oprec, atol=TOL)


def test_empty_queries_convention(H):
    """`empty_queries="skip"`: queries whose list holds no relevant item are left out of the mean (the other upstream
    lineage's convention); "zero" is the default.  Sparse labels + small R make such queries common."""
    d, dl, q, ql, _ = synth.make_random_case(300, 4000, 32, 400, p=0.45, seed=5)
    for R in (5, [3, 20], -1):
        z = H.calculate_mAP(d.cuda(), dl.cuda(), q.cuda(), ql.cuda(), R, PRs=[1, 5])
        s = H.calculate_mAP(d.cuda(), dl.cuda(), q.cuda(), ql.cuda(), R, PRs=[1, 5], empty_queries="skip")
        oz = mo.calculate_mAP(d, dl, q, ql, R, PRs=[1, 5])
        os_ = mo.calculate_mAP(d, dl, q, ql, R, PRs=[1, 5], empty_queries="skip")
        assert np.allclose(z[0], oz[0], atol=TOL) and np.allclose(s[0], os_[0], atol=TOL)
        assert np.allclose(s[1], os_[1], atol=TOL) and np.allclose(s[2], os_[2], atol=TOL)
        if R == 5:
            assert s[0] > z[0] + 1e-3            # (the case does hold empty lists)
    with pytest.raises(ValueError):
        H.calculate_mAP(d.cuda(), dl.cuda(), q.cuda(), ql.cuda(), 5, empty_queries="drop")


def test_hinted_key_range_outgrown_on_gpu(H):
    """POPC select pass (``use_tensor_cores=False``; also what wide ternary codes take) evaluated three times on one
    shape: tight clusters twice (thresholds of 0-2 keys; the second run sizes its slabs from the first one's largest
    threshold + 3), then near-random codes (thresholds ~15).  The kernel must clamp the thresholds to the narrowed
    key range (hist.cu), the device check flags the evaluation and it is repeated: the oracle's numbers every time.
    CPU twin: tests/test_evaluator_emu.py::test_hinted_key_range_outgrown_by_the_next_evaluation."""
    ev = H.get_evaluator()
    nq, ndb, nbit, ncls, R = 1234, 260_000, 64, 20, 50
    seen = []
    ev.use_tensor_cores = False
    try:
        for p, seed in ((0.02, 1), (0.02, 1), (0.45, 2), (0.45, 2)):
            d, dl, q, ql, _ = synth.make_random_case(nq, ndb, nbit, ncls, p=p, seed=seed, device="cuda")
            out = ev.evaluate(d, dl, q, ql, [R], 0.0, [1, 10], False, return_ap=True)
            sub = slice(0, 24)
            om, orec, oprec, oaps = mo.calculate_mAP(d.cpu(), dl.cpu(), q[sub].cpu(), ql[sub].cpu(), R, PRs=[1, 10],
                                                     return_per_query=True)
            assert np.allclose(out[3][0][sub].cpu().numpy(), oaps[0], atol=TOL), (p, seed, ev.stats)
            seen.append((ev.stats["mode"], ev.stats["select_kernel"], ev.stats["speculation"],
                         (ev.stats.get("sample") or {}).get("key_limit", -1)))
    finally:
        ev.use_tensor_cores = True
    assert all(s[0].startswith("topR") and s[1] == "popc" for s in seen), seen
    assert seen[1][2] == "hit" and seen[2][2] != "hit", seen
    if all(s[0] == "topR-sampled" for s in seen[:3]):
        assert seen[2][3] > seen[1][3] + 3, seen      # the key range really was outgrown
