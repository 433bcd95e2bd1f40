#!/usr/bin/env python
"""Benchmark of the retrieval-evaluation hot path (BASELINE.json metric: 64-bit Hamming comparisons/s and
mAP@R eval wall-time at 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg5|cub200|cars196|nabirds]
    python bench.py --impl reference ...      # the reference-style CPU implementation, same metric

One "step" = one complete calculate_mAP evaluation (sign/bit-pack of queries + gallery, Hamming passes,
exact top-R selection, label match + AP, means) over one batch of synthetic codes.  The last stdout line
is ONE JSON object (see DESIGN.md "Measurement").
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[3]: 128-bit codes, 25k query x 1M synthetic gallery, mAP@1000 at 1/2/4/8 B200
    "cfg4": dict(nq=25_000, ndb=1_000_000, nbit=128, nclass=101, R=1000, p=0.30, scaling="strong",
                 desc="BASELINE configs[3]: 128-bit, 25,000 queries x 1,000,000 gallery, mAP@1000, 101 classes"),
    # BASELINE.json configs[4] weak-scaled: every GPU owns a 12.5M-row shard (8 GPUs = the 100M gallery)
    "cfg5": dict(nq=100_000, ndb=12_500_000, nbit=64, nclass=1000, R=1000, p=0.30, scaling="weak",
                 desc="BASELINE configs[4] per-GPU shard: 64-bit, 100,000 queries x 12.5M gallery rows per GPU "
                      "(8 GPUs = the 100M gallery), top-R=1000, 1000 classes"),
    # development-size stand-in for cfg5 (same code path, 1/12.5 of the per-GPU shard)
    "cfg5s": dict(nq=100_000, ndb=1_000_000, nbit=64, nclass=1000, R=1000, p=0.30, scaling="strong",
                  desc="cfg5 stand-in: 64-bit, 100,000 queries x 1,000,000 gallery, top-R=1000, 1000 classes"),
    "cub200": dict(dataset="cub200", nbit=64, R=-1, p=0.15, scaling="strong",
                   desc="BASELINE configs[0]: CUB-200-2011 64-bit mAP@all, 5,794 x 5,994, 200 classes"),
    "cars196": dict(dataset="cars196", nbit=64, R=-1, p=0.15, scaling="strong",
                    desc="BASELINE configs[1]: Cars196 64-bit mAP@all, 8,041 x 8,144, 196 classes"),
    "nabirds": dict(dataset="nabirds", nbit=64, R=-1, p=0.15, scaling="strong",
                    desc="BASELINE configs[2]: NABirds 64-bit mAP@all, 24,633 x 23,929, 555 classes"),
}


def make_workload(name, device, nbit_override=None, shard=0):
    from concepthash_b200 import synth
    w = dict(WORKLOADS[name])
    if nbit_override:
        w["nbit"] = nbit_override
    if "dataset" in w:
        d, dl, q, ql, ncls = synth.make_dataset_case(w["dataset"], nbit=w["nbit"], p=w["p"], seed=0, device=device)
        w.update(nq=q.shape[0], ndb=d.shape[0], nclass=ncls)
    else:
        d, dl, q, ql, ncls = synth.make_random_case(w["nq"], w["ndb"], w["nbit"], w["nclass"], p=w["p"], seed=0,
                                                    device=device, shard=shard)
    return w, d, dl, q, ql


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # NVML missing: report that instead of inventing clocks
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml")}
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


def cpu_reference_sample(w, d, dl, q, ql, scale=1.0):
    """Reference-style CPU implementation (oracle/, upstream structure: 32-row chunk GEMMs -> dense dist ->
    torch.topk -> per-query numpy loop) on a bounded sample of the SAME workload, all host threads."""
    from oracle import map_oracle as mo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ndb, nq = d.shape[0], q.shape[0]
    # bounded sample: a query subset against the full (single-GPU) gallery, sized for 10-30 s of CPU work
    if ndb * nq <= 1.2e8:
        sq = nq
    elif w["R"] == -1:
        sq = min(nq, max(64, int(scale * 3.0e7 / ndb)))       # the numpy AP loop dominates: ~2e6 pairs/s
    else:
        sq = min(nq, max(16, int(scale * 4.0e9 / ndb)))       # the chunked GEMM dominates: ~2.5e8 pairs/s on 16 cores
    dc, dlc = d.cpu(), dl.cpu()
    qc, qlc = q[:sq].cpu(), ql[:sq].cpu()
    t0 = time.perf_counter()
    m = mo.calculate_mAP_upstream_style(dc, dlc, qc, qlc, w["R"])
    dt = time.perf_counter() - t0
    return dict(seconds=dt, pairs=float(sq) * ndb, queries=sq, cores=cores, threads=torch.get_num_threads(), mAP=m)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--nbit", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sample-stride", type=int, default=0, help="override the evaluator's row-sample stride")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    unit64 = (args.nbit or wl["nbit"]) / 64.0

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        w, d, dl, q, ql = make_workload(args.workload, "cpu" if "dataset" in wl else
                                        ("cuda" if torch.cuda.is_available() else "cpu"), args.nbit)
        res = None
        times = []
        # each step is a bounded sample (~16 s); with many steps the sample shrinks so that the whole run stays
        # within a few minutes
        scale = min(1.0, 150.0 / (16.0 * max(1, args.warmup + args.steps)))
        for i in range(args.warmup + args.steps):
            res = cpu_reference_sample(w, d, dl, q, ql, scale)
            if i >= args.warmup:
                times.append(res["seconds"])
        ms = 1e3 * float(np.mean(times))
        value = res["pairs"] * unit64 / (ms * 1e-3)
        sample = f"{res['queries']} of {w['nq']} queries x {w['ndb']} gallery rows per step (linear in nq)"
        print(json.dumps({
            "impl": "reference", "metric": "hamming_comparisons_per_sec_64bit", "value": value,
            "unit": "64-bit comparisons/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "nbit": w["nbit"], "R": w["R"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": "64-bit comparisons/s", "cores": res["cores"], "kind": "port",
                             "sample": sample, "threads": res["threads"]},
            "e2e": {"value": value, "unit": "64-bit comparisons/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }))
        return 0

    # ------------------------------------------------------------------ our arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
        group = dist.group.WORLD
    from concepthash_b200 import hashing

    # weak scaling: every rank generates its own gallery block (same queries), no duplicates across ranks
    w, d, dl, q, ql = make_workload(args.workload, device, args.nbit, shard=rank if wl["scaling"] == "weak" else 0)
    ndb_full = d.shape[0]
    if w["scaling"] == "strong" and world > 1:        # row-shard the named gallery over the ranks
        cut = [ndb_full * r // world for r in range(world + 1)]
        d, dl = d[cut[rank]:cut[rank + 1]].contiguous(), dl[cut[rank]:cut[rank + 1]].contiguous()
    total_pairs = float(w["nq"]) * (ndb_full if w["scaling"] == "strong" else ndb_full * world)
    ev = hashing.get_evaluator(device, group)
    if args.sample_stride:
        ev.sample_stride = args.sample_stride
    r_list = [w["R"]]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def run(fn, steps, warmup, sample_clocks=False):
        out = None
        for _ in range(warmup):
            out = fn()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        barrier()
        if sampler:
            sampler.start()
        ev.events = []
        ev.profile = True
        l0 = ev.b.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ev.profile = False
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=device)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item()), out, ev.b.launch_count() - l0, clocks, list(ev.events)

    # value: inputs resident in HBM (fp32 codes + int64 ids as torch CUDA tensors)
    step_dev = lambda: ev.evaluate(d, dl, q, ql, r_list, 0.0, [], False)
    ms, out, launches, clocks, events = run(step_dev, args.steps, args.warmup, sample_clocks=True)
    value = total_pairs * unit64 / (ms * 1e-3)

    # per-kernel times from the CUDA events recorded on the launch stream inside the timed region
    kinds = {}
    for kind, units, a, b in events:
        k = kinds.setdefault(kind, [0.0, 0.0, 0])
        k[0] += a.elapsed_time(b)
        k[1] += units
        k[2] += 1
    words32 = max(1, (w["nbit"] + 31) // 32)
    popc_peak, _ = ev.b.popc_peak()
    peaks, peak_src = measured_peaks()
    # the dominant kernel = the Hamming pass with the largest share of the step
    sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)

    def popc_roofline(kind):
        hk = kinds[kind]
        t_ms = hk[0] / hk[2]
        popc = (hk[1] / hk[2]) * words32 / (t_ms * 1e-3)
        name = {"hist_count": "count pass: every pair histogrammed",
                "hist_count_rec": "count pass + records of relevant pairs (full ranking)",
                "hist_select": "select pass: pairs with key <= threshold counted / matched / recorded"}[kind]
        return {
            "kernel": f"hamming_hist_kernel ({name})", "bound": "int-pipe (POPC)",
            "achieved": popc / 1e9, "peak": popc_peak / 1e9, "unit": "Gpopc32/s", "frac": popc / popc_peak,
            "traffic": ncu_traffic("hamming_hist_kernel"),
            "algorithmic_work": f"pairs x {words32} popc32 per pair per launch",
            "peak_source": "ch_popc_peak micro-benchmark run live on this GPU (MEASURED_PEAKS.json has no "
                           "integer-pipe figure); nominal 148 SM x 16 lanes x f",
            "nominal_peak_at_sampled_clock": 148 * 16 * sm_mhz * 1e6 / 1e9,
            "ms_per_launch": t_ms, "pairs_per_launch": hk[1] / hk[2], "pairs_per_s": (hk[1] / hk[2]) / (t_ms * 1e-3),
            "share_of_step": hk[0] / (ms * args.steps)}

    def ncu_traffic(kernel):
        """DRAM bytes per launch of `kernel` from the committed ncu --set full capture of this workload, or None"""
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                t = json.load(f)
            return t.get(args.workload, {}).get(kernel)
        except Exception:
            return None

    def tensor_roofline(kind):
        hk = kinds[kind]
        t_ms = hk[0] / hk[2]
        pairs = hk[1] / hk[2]
        kb = ev.b.tc_code_bytes(w["nbit"])                    # K bytes the MMA really contracts (codes + threshold block)
        tops = pairs * w["nbit"] * 2 / (t_ms * 1e-3) / 1e12   # ALGORITHMIC: nbit int8 MACs per pair
        peak = 2.0 * peaks.get("bf16_tflops", 1590.0)
        pairs_s = pairs / (t_ms * 1e-3)
        return {
            "kernel": "hamming_select_tc_kernel (select pass: tcgen05.mma kind::i8 -> TMEM, sign-bit epilogue, "
                      "candidate lists)",
            "bound": "tensor", "achieved": tops, "peak": peak, "unit": "TOP/s (int8)", "frac": tops / peak,
            "traffic": ncu_traffic("hamming_select_tc_kernel"),
            "algorithmic_work": f"pairs x {w['nbit']} int8 MACs x 2 per launch (the kernel contracts K = {kb} bytes: "
                                "the codes plus the 32-byte block that carries the per-query threshold)",
            "peak_source": f"2 x bf16_tflops of {peak_src}: int8 dense runs at twice the bf16 rate and no int8 figure "
                           "is measured; nominal dense int8 is 4500 TOP/s",
            "executed_tops_incl_threshold_block": tops * kb / w["nbit"],
            "frac_of_nominal_int8_4500": tops / 4500.0,
            "pairs_per_clk_per_sm": pairs_s / (148 * sm_mhz * 1e6),
            "mma_floor_pairs_per_clk_per_sm": 128.0 * 128.0 / (64.0 * (kb // 32)),
            "popc_kernel_ceiling_pairs_per_s": popc_peak / words32,
            "ms_per_launch": t_ms, "launches_per_step": hk[2] / args.steps, "pairs_per_launch": pairs,
            "pairs_per_s": pairs_s, "share_of_step": hk[0] / (ms * args.steps)}

    ham_kinds = [k for k in kinds if k.startswith("hist")]
    dom = max(ham_kinds, key=lambda k: kinds[k][0])
    roofline = tensor_roofline(dom) if dom == "hist_select_tc" else popc_roofline(dom)
    roofline_other = {k: (tensor_roofline(k) if k == "hist_select_tc" else popc_roofline(k))
                      for k in ham_kinds if k != dom}
    pack = kinds.get("pack_dev")
    roofline_pack = None
    if pack:
        gbs = pack[1] / (pack[0] * 1e-3) / 1e9
        roofline_pack = {"kernel": "pack_bits_kernel (sign + bit-pack)", "bound": "hbm", "achieved": gbs,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "traffic": None,
                         "peak_source": peak_src, "ms_per_step": pack[0] / args.steps}
    kernel_ms = {k: v[0] / args.steps for k, v in kinds.items()}

    # e2e: the public call with HOST (pinned) tensors; H2D of codes + labels and D2H of the result inside
    e2e = None
    if not args.no_e2e:
        hd, hdl, hq, hql = (t.cpu().pin_memory() for t in (d, dl, q, ql))
        step_host = lambda: hashing.calculate_mAP(hd, hdl, hq, hql, w["R"], group=group)
        ms_e, out_e, _, _, ev_e = run(step_host, max(2, args.steps // 2), 2)
        kinds_e = {}
        for kind, units, a, b in ev_e:
            kinds_e[kind] = kinds_e.get(kind, 0.0) + a.elapsed_time(b) / max(2, args.steps // 2)
        h2d = sum(t.numel() * t.element_size() for t in (hd, hdl, hq, hql))
        # the same call with ordinary (pageable) CPU tensors -- what the reference's trainers hand over
        pd, pdl, pq, pql = (t.cpu() for t in (d, dl, q, ql))
        step_page = lambda: hashing.calculate_mAP(pd, pdl, pq, pql, w["R"], group=group)
        ms_p, _, _, _, _ = run(step_page, 2, 1)
        del pd, pdl, pq, pql
        e2e = {"value": total_pairs * unit64 / (ms_e * 1e-3), "unit": "64-bit comparisons/s",
               "ms_per_step_pageable_host_tensors": ms_p,
               "ms_per_step": ms_e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8 * len(r_list) + 64,
               "mAP": out_e[0], "mode": ev.stats.get("mode"), "kernel_ms_per_step": kinds_e}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_sample(w, d, dl, q, ql)
        cpu = {"value": r["pairs"] * unit64 / r["seconds"], "unit": "64-bit comparisons/s", "cores": r["cores"],
               "kind": "port", "threads": r["threads"], "seconds": r["seconds"],
               "sample": f"{r['queries']} of {w['nq']} queries x {w['ndb']} gallery rows (upstream-style "
                         f"oracle/map_oracle.calculate_mAP_upstream_style; linear in nq)"}

    if rank == 0:
        print("evaluator stats:", ev.stats, file=sys.stderr)
        line = {
            "metric": "hamming_comparisons_per_sec_64bit", "value": value, "unit": "64-bit comparisons/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
            "dtype": "s8 x s8 -> s32 (tcgen05 kind::i8 select pass) + u32 xor/popc (sample, count and key passes)"
            if "hist_select_tc" in kinds else "u32 xor/popc",
            "data": "synthetic",
            "config": {"workload": w["desc"], "nq": w["nq"], "ndb_per_gpu": int(d.shape[0]),
                       "ndb_total": int(ndb_full if w["scaling"] == "strong" else ndb_full * world),
                       "nbit": w["nbit"], "R": w["R"], "nclass": w["nclass"],
                       "pairs_per_s": total_pairs / (ms * 1e-3),
                       "mode": ev.stats.get("mode"), "geometry(threads,nq_pad,stripes,rows/stripe)": ev.stats.get("geometry"),
                       "l2_policy": "inputs larger than L2 (fp32 gallery codes >= 512 MB vs 126 MB L2)"
                       if w["ndb"] * w["nbit"] * 4 > 2.0e8 else "small workload: inputs fit L2 (latency-bound case)",
                       "parallelism": f"gallery row-sharded x{world}" if world > 1 else "single GPU",
                       "mAP": out[0][0] if out else None},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "kernel_ms_per_step": kernel_ms,
            "roofline": roofline, "roofline_other_passes": roofline_other, "roofline_pack": roofline_pack,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
