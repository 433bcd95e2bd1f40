"""Compact view of a bench.py JSON line (dev helper)."""
import json
import sys

line = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])


def brief(x, main=False):
    km = x.get("kernel_ms_detail") or x.get("kernel_ms_per_step") or {}
    top = sorted(km.items(), key=lambda kv: -kv[1])[:12]
    print(f"== {x.get('workload') or x['config']['workload']}  nbit={x.get('nbit') or x['config']['nbit']}  N={x['n_gpus']}")
    print(f"   ms/step {x['ms_per_step']:.4f}  value {x['value']:.4g}  mAP {x.get('mAP') or x['config']['mAP']}"
          f"  launches/step {x['gpu_launches'] / x['steps']:.1f}")
    print("   phase", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in x["phase_ms"].items()})
    print("   kernels", ", ".join(f"{k}={v:.3f}" for k, v in top))
    r = x.get("roofline")
    if r:
        print(f"   roofline {r['kernel'][:40]}: frac {r['frac']:.3f}  ms/launch {r['ms_per_launch']:.3f}"
              f"  issue-frac {r.get('frac_of_measured_mma_issue_peak')}")
    pc = x.get("parity_check")
    if pc:
        print("   parity", {k: pc.get(k) for k in ("queries", "gallery_rows", "max_abs_ap_delta", "ids_equal", "ok")})
    e = x.get("e2e")
    if e:
        print(f"   e2e pageable {e['ms_per_step']:.2f} ms  pinned {e['ms_per_step_pinned_host_tensors']:.2f} ms" +
              (f"  one-hot labels {e['ms_per_step_onehot_labels']:.2f} ms" if "ms_per_step_onehot_labels" in e else ""))
    if x.get("clocks"):
        print("   clocks", x["clocks"])
    if x.get("cpu_baseline"):
        print("   cpu", x["cpu_baseline"]["value"], x["cpu_baseline"]["cores"])


brief(line, True)
for o in line.get("other_workloads", []):
    brief(o)
