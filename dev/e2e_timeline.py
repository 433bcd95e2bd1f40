"""Host + device timeline of ONE public call with pageable (PIN=0, default) or pinned (PIN=1) host tensors: every backend
entry point with its host start / duration and the device start / end of what it enqueued (events on the stream current
at the call).  python dev/e2e_timeline.py <workload>"""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from concepthash_b200 import hashing  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
w, d, dl, q, ql = bench.make_workload(name, "cuda")
PIN = os.environ.get("PIN", "0") == "1"
hd, hdl, hq, hql = ((t.cpu().pin_memory() if PIN else t.cpu()) for t in (d, dl, q, ql))
del d, dl, q, ql
ev = hashing.get_evaluator()
f = lambda: hashing.calculate_mAP(hd, hdl, hq, hql, w["R"])
for _ in range(5):
    f()
torch.cuda.synchronize()
b = ev.b._b        # (ev.b is the pass-through proxy)
log = []
skip = {"on_stream", "empty", "zeros", "full", "padded_rows", "code_words", "tc_code_bytes", "tc_code_bytes_pair",
        "launch_count", "begin", "geometry", "gather_plane_words"}
T0 = [0.0]


def wrap(nm, fn):
    def g(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        out = fn(*a, **k)
        t1 = time.perf_counter()
        e1.record()
        log.append((nm, (t0 - T0[0]) * 1e3, (t1 - t0) * 1e3, e0, e1))
        return out
    return g


for nm in dir(b):
    if nm.startswith("_") or nm in skip:
        continue
    fn = getattr(b, nm)
    if callable(fn) and not isinstance(fn, type):
        try:
            setattr(b, nm, wrap(nm, fn))
        except Exception:
            pass
torch.cuda.synchronize()
base = torch.cuda.Event(enable_timing=True)
base.record()
T0[0] = time.perf_counter()
m = f()
t_end = (time.perf_counter() - T0[0]) * 1e3
torch.cuda.synchronize()
print("step wall ms %.3f  mode %s  mAP %r" % (t_end, ev.stats["mode"], m[0]))
print("%-28s %9s %8s | %9s %9s" % ("entry point", "host t0", "host ms", "dev t0", "dev t1"))
for nm, t0, dt, e0, e1 in log:
    print("%-28s %9.3f %8.3f | %9.3f %9.3f" % (nm, t0, dt, base.elapsed_time(e0), base.elapsed_time(e1)))
