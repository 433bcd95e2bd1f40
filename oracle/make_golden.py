"""Generates the committed fixtures under tests/golden/ (run HERE, where /root/reference exists).

    python oracle/make_golden.py

1. ``labels_<dataset>.npz`` -- the integer label columns of the reference's
   ``data/<dataset>/{test,database}.txt`` lists (SURVEY.md §8c "fixtures that exist in the
   reference"): they fix nq / ndb / C of BASELINE configs 1-3 exactly.
2. ``ref_substeps.npz`` -- outputs of the reference's OWN in-tree code for the sub-steps of the
   hot path that do exist in /root/reference (the hot-path function itself is absent, F1):
     * ``models.layers.signhash.sign_hash``        (sign binarisation, signhash.py:11)
     * ``utils.metrics.calculate_accuracy_hamm_dist`` (nearest-codeword accuracy, metrics.py:18-29)
     * ``engine.tensor_to_dataset`` + ``engine.dataloader(…, 32, False, 0, False)``
       (the 32-row gallery chunk iterator the upstream function is built on, engine.py:41-80)
   The oracle is checked against these in tests/test_oracle.py.
3. ``oracle_cases.npz`` -- small seeded inputs with the oracle's outputs (regression pins for the
   GPU parity tests; oracle-generated, i.e. NOT an independent pin).

Nothing here is imported by the product.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def read_labels(path):
    with open(path) as f:
        return np.array([int(line.rsplit(" ", 1)[1]) for line in f if line.strip()], dtype=np.int16)


def labels():
    for name, sub in [("cub200", "cub200_2011"), ("cars196", "cars196"), ("nabirds", "nabirds")]:
        t = read_labels(os.path.join(REF, "data", sub, "test.txt"))
        d = read_labels(os.path.join(REF, "data", sub, "database.txt"))
        np.savez_compressed(os.path.join(OUT, f"labels_{name}.npz"), test=t, database=d)
        print(name, t.shape, d.shape, int(max(t.max(), d.max())) + 1)


def ref_substeps():
    sys.path.insert(0, REF)
    import engine                                    # /root/reference/engine.py
    from models.layers.signhash import sign_hash     # /root/reference/models/layers/signhash.py
    from utils.metrics import calculate_accuracy_hamm_dist  # /root/reference/utils/metrics.py

    g = torch.Generator().manual_seed(1234)
    x = torch.randn(37, 48, generator=g)
    x[3, 5] = 0.0
    x[10, :4] = 0.0
    x[11, 7] = -0.0
    s = sign_hash(x)

    nclass, nbit, b = 23, 32, 64
    cbk = torch.sign(torch.randn(nclass, nbit, generator=g))
    lab = torch.randint(nclass, (b,), generator=g)
    codes = cbk[lab] * torch.where(torch.rand(b, nbit, generator=g) < 0.2, -1.0, 1.0)
    hd = 0.5 * (nbit - codes @ cbk.t())              # identity of trainers/orthohash.py:263-264
    onehot = torch.nn.functional.one_hot(lab, nclass).float()
    acc1 = calculate_accuracy_hamm_dist(hd, onehot)
    acc5 = calculate_accuracy_hamm_dist(hd, onehot, multiclass=True)

    gal = torch.randn(101, 16, generator=g)
    ttd = engine.tensor_to_dataset(gal)
    chunks = [c.clone() for c in engine.dataloader(ttd, 32, False, 0, False)]
    np.savez_compressed(
        os.path.join(OUT, "ref_substeps.npz"),
        sign_in=x.numpy(), sign_out=s.numpy(),
        hd_codes=codes.numpy(), hd_codebook=cbk.numpy(), hd_labels=lab.numpy(),
        hd_acc1=np.float64(acc1.item()), hd_acc5=np.float64(acc5.item()),
        chunk_in=gal.numpy(), chunk_sizes=np.array([c.shape[0] for c in chunks]),
        chunk_cat=torch.cat(chunks, 0).numpy(),
    )
    print("ref_substeps ok", float(acc1), float(acc5), [c.shape[0] for c in chunks])


def oracle_cases():
    sys.path.insert(0, ROOT)
    from oracle import map_oracle as mo
    from concepthash_b200 import synth

    out = {}
    specs = [  # name, nq, ndb, nbit, nclass, p, R, PRs, remove_first, threshold, zero_frac
        ("a", 33, 257, 64, 7, 0.30, -1, [1, 5, 10], False, 0.0, 0.0),
        ("b", 50, 600, 16, 5, 0.30, 40, [1, 5, 10], False, 0.0, 0.0),
        ("c", 64, 300, 32, 11, 0.25, [10, 100, -1], [1, 3], False, 0.0, 0.0),
        ("d", 40, 40, 64, 6, 0.20, -1, [1, 5], True, 0.0, 0.0),
        ("e", 31, 500, 128, 9, 0.35, 77, [10], False, 0.0, 0.0),
        ("f", 29, 333, 48, 8, 0.30, -1, [1, 5, 10], False, 0.5, 0.0),
        ("g", 27, 222, 64, 8, 0.30, 50, [1, 5, 10], False, 0.0, 0.1),
    ]
    for name, nq, ndb, nbit, nclass, p, R, PRs, rf, thr, zf in specs:
        d, dl, q, ql, _ = synth.make_random_case(nq, ndb, nbit, nclass, p, seed=ord(name))
        if rf:
            q, ql = d.clone(), dl.clone()
        if zf > 0:
            g = torch.Generator().manual_seed(99)
            d[torch.rand(d.shape, generator=g) < zf] = 0.0
            q[torch.rand(q.shape, generator=g) < zf] = 0.0
        m, rec, prec, aps = mo.calculate_mAP(d, dl, q, ql, R, threshold=thr, PRs=PRs,
                                             remove_first_retrieved=rf, return_per_query=True)
        ids, dist = mo.topk_ids(q, d, R if not isinstance(R, list) else max(r if r > 0 else ndb for r in R),
                                threshold=thr, remove_first_retrieved=rf)
        out[f"{name}_db_codes"] = d.numpy(); out[f"{name}_db_labels"] = dl.numpy()
        out[f"{name}_q_codes"] = q.numpy(); out[f"{name}_q_labels"] = ql.numpy()
        out[f"{name}_R"] = np.array(R if isinstance(R, list) else [R])
        out[f"{name}_R_is_list"] = np.array(isinstance(R, list))
        out[f"{name}_PRs"] = np.array(PRs); out[f"{name}_rf"] = np.array(rf)
        out[f"{name}_thr"] = np.array(thr)
        out[f"{name}_mAP"] = np.array(m if isinstance(m, list) else [m], dtype=np.float64)
        out[f"{name}_recalls"] = np.array(rec); out[f"{name}_precisions"] = np.array(prec)
        out[f"{name}_aps"] = aps
        out[f"{name}_ids"] = ids.numpy().astype(np.int32)
        out[f"{name}_dist2"] = (2 * dist).round().numpy().astype(np.int16)
        print(name, m, rec, prec)
    out["names"] = np.array([s[0] for s in specs])
    np.savez_compressed(os.path.join(OUT, "oracle_cases.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    labels()
    ref_substeps()
    oracle_cases()
