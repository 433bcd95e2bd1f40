"""e2e wall time of the public call with HOST tensors (pageable; PIN=1: pinned) over the knobs of the native-loader
path (blocks per evaluation, list kernels beside the next select), and the raw rate of the loader alone:
python dev/e2e_native_sweep.py <workload> [steps]"""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from concepthash_b200 import hashing  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
PIN = os.environ.get("PIN", "0") == "1"
w, d, dl, q, ql = bench.make_workload(name, "cuda")
hd, hdl, hq, hql = ((t.cpu().pin_memory() if PIN else t.cpu()) for t in (d, dl, q, ql))
del d, dl, q, ql
ev = hashing.get_evaluator()
b = ev.b._b

# ---- the loader alone
side = torch.cuda.Stream(priority=-1)
bits = torch.empty((b.padded_rows(hd.shape[0]), b.code_words(hd.shape[1])), dtype=torch.int32, device="cuda")
flags = torch.zeros(1, dtype=torch.int32, device="cuda")
ts = []
for i in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ld = b.host_loader_start([(hd, bits, flags)], side)
    ld.join()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    ts.append(((t1 - t0) * 1e3, (time.perf_counter() - t0) * 1e3))
print("loader alone (join, +sync) ms:", " ".join("%.2f/%.2f" % t for t in ts),
      " best %.1f GB/s" % (hd.numel() * 4 / min(t[0] for t in ts) / 1e6))


def run(label):
    ev._hints.clear()
    ts = []
    for i in range(steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hashing.calculate_mAP(hd, hdl, hq, hql, w["R"])
        ts.append((time.perf_counter() - t0) * 1e3)
    ts = ts[3:]
    print("%-34s min %.2f  med %.2f  max %.2f   %s %s" % (label, min(ts), sorted(ts)[len(ts) // 2], max(ts),
                                                           ev.stats["mode"], ev.stats["geometry"]))


for native in (True, False):
    ev.stream_native_loader = native
    if not native:
        run("python loader (before)")
        continue
    for labels in (False, True, False, True):
        ev.stream_loader_labels = labels
        run("native loader_labels=%d" % labels)
