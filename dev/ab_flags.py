"""Same-process A/B of an evaluator attribute on one workload (device-resident inputs, CUDA-event timing, the settings
alternate so that box-to-box and clock differences cancel):
    python dev/ab_flags.py <workload> <attr> <value> <value> ... [--steps N]"""
import ast
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from concepthash_b200 import hashing  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 10
name, attr, values = args[0], args[1], [ast.literal_eval(v) for v in args[2:]]
w, d, dl, q, ql = bench.make_workload(name, "cuda")
ev = hashing.get_evaluator()
res = {repr(v): [] for v in values}
info = {}
for rnd in range(3):
    for v in values:
        setattr(ev, attr, v)
        ev._hints.clear()
        for _ in range(3):
            m = ev.evaluate(d, dl, q, ql, [w["R"]], 0.0, [], False)
        if rnd == 0:
            ev.debug_counts = True
            ev.evaluate(d, dl, q, ql, [w["R"]], 0.0, [], False)
            ev.debug_counts = False
            info[repr(v)] = dict(mAP=m[0][0], candidates=ev.stats.get("candidates"), slots=ev.stats.get("record_slots"),
                                 mode=ev.stats["mode"])
            ev.evaluate(d, dl, q, ql, [w["R"]], 0.0, [], False)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            ev.evaluate(d, dl, q, ql, [w["R"]], 0.0, [], False)
        b.record()
        torch.cuda.synchronize()
        res[repr(v)].append(a.elapsed_time(b) / steps)
for v in values:
    print(name, attr, "=", v, "ms/step", " ".join("%.3f" % t for t in res[repr(v)]), info[repr(v)])
