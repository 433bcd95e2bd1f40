"""e2e wall time of the public call with PAGEABLE host tensors over stream_chunks (CH_PACK_THREADS from the env):
python dev/e2e_sweep.py <workload> [steps]"""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from concepthash_b200 import hashing  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
w, d, dl, q, ql = bench.make_workload(name, "cuda")
hd, hdl, hq, hql = (t.cpu() for t in (d, dl, q, ql))
del d, dl, q, ql
ev = hashing.get_evaluator()
for chunks in (4, 2, 3, 6, 8):
    ev.stream_chunks = chunks
    ev._hints.clear()
    ts = []
    for i in range(steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m = hashing.calculate_mAP(hd, hdl, hq, hql, w["R"])
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print("stream_chunks", chunks, "ms", " ".join("%.2f" % t for t in ts), ev.stats["mode"], ev.stats["geometry"])
