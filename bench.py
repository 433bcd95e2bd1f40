#!/usr/bin/env python
"""Benchmark of the retrieval-evaluation hot path (BASELINE.json metric: 64-bit Hamming comparisons/s and
mAP@R eval wall-time at 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg5|cub200|cars196|nabirds]
    python bench.py --impl reference ...      # the reference-style CPU implementation, same metric

One "step" = one complete calculate_mAP evaluation (sign/bit-pack of queries + gallery, Hamming passes,
exact top-R selection, label match + AP, means) over one batch of synthetic codes.  The last stdout line
is ONE JSON object (see DESIGN.md "Measurement").  Besides the headline workload the line carries

  * ``parity_check``     -- outside the timed region, at every N: per-query AP of the timed configuration and the
                            ranked ids of 32 queries against the CPU oracle (oracle/packed_oracle.py, checker only)
                            over the CONCATENATED gallery of all ranks; a mismatch aborts with a non-zero exit code;
  * ``other_workloads``  -- the other BASELINE configs on the same box: configs[4]'s per-GPU shard (weak-scaled) at
                            every N, and at N = 1 CUB-200 / Cars196-16/32/64 / NABirds, each with its own
                            parity_check;
  * ``phase_ms``         -- kernels / collectives / host share of a step.
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
    # BASELINE.json configs[4] AS NAMED, strong-scaled: the whole 100M-row gallery (= the 8 shards of "cfg5" concatenated,
    # so the mAP must equal the 8-GPU weak-scaled run's to the last bit) split over however many GPUs there are; on few
    # GPUs the candidate slices exceed the 32-bit slot range and calculate_mAP chunks the queries internally
    "cfg5full": dict(nq=100_000, ndb=12_500_000, shards=8, nbit=64, nclass=1000, R=1000, p=0.30, scaling="strong",
                     desc="BASELINE configs[4] as named: 64-bit, 100,000 queries x 100M gallery (strong-scaled: 100M / N "
                          "rows per GPU), top-R=1000, 1000 classes"),
    # development-size stand-in for cfg5 (same code path, 1/12.5 of the per-GPU shard)
    "cfg5s": dict(nq=100_000, ndb=1_000_000, nbit=64, nclass=1000, R=1000, p=0.30, scaling="strong",
                  desc="cfg5 stand-in: 64-bit, 100,000 queries x 1,000,000 gallery, top-R=1000, 1000 classes"),
    "cub200": dict(dataset="cub200", nbit=64, R=-1, p=0.15, scaling="strong",
                   desc="BASELINE configs[0]: CUB-200-2011 64-bit mAP@all, 5,794 x 5,994, 200 classes"),
    "cars196": dict(dataset="cars196", nbit=64, R=-1, p=0.15, scaling="strong",
                    desc="BASELINE configs[1]: Cars196 mAP@all, 8,041 x 8,144, 196 classes"),
    "nabirds": dict(dataset="nabirds", nbit=64, R=-1, p=0.15, scaling="strong",
                    desc="BASELINE configs[2]: NABirds 64-bit mAP@all, 24,633 x 23,929, 555 classes"),
}
# measured by dev/mma_rate_probe.cu on this pool's B200 (profiles/r1i_mma_rate_probe.txt): one UTCIMMA
# M = N = 128, K = 32 per 64.0 clk and SM = 8192 int8 MAC / clk / SM
MMA_ISSUE_MAC_PER_CLK_SM = 8192.0


def make_workload(name, device, nbit_override=None, shard=0):
    from concepthash_b200 import synth
    w = dict(WORKLOADS[name])
    if nbit_override:
        w["nbit"] = nbit_override
    if "dataset" in w:
        d, dl, q, ql, ncls = synth.make_dataset_case(w["dataset"], nbit=w["nbit"], p=w["p"], seed=0, device=device)
        w.update(nq=q.shape[0], ndb=d.shape[0], nclass=ncls)
    elif "shards" in w:
        # the concatenation of the weak-scaled run's per-rank blocks (generated block by block into one allocation)
        ns, per = int(w["shards"]), int(w["ndb"])
        d = torch.empty(ns * per, w["nbit"], dtype=torch.float32, device=device)
        dl = torch.empty(ns * per, dtype=torch.int64, device=device)
        for sh in range(ns):
            db, dlb, q, ql, ncls = synth.make_random_case(w["nq"], per, w["nbit"], w["nclass"], p=w["p"], seed=0,
                                                          device=device, shard=sh)
            d[sh * per:(sh + 1) * per], dl[sh * per:(sh + 1) * per] = db, dlb
            del db, dlb
        w["ndb"] = ns * per
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


# ---------------------------------------------------------------------------------------------------- parity check
def _torch_sign_bytes(codes, chunk=1 << 20):
    """(n, nbit) real codes -> (n, 8 * ceil(nbit / 64)) uint8, little-endian sign bits, by PLAIN torch ops (the
    checker must not depend on the pack kernel it checks).  Also returns whether an exact zero / NaN was seen."""
    n, nbit = codes.shape
    nb = (nbit + 63) // 64 * 8
    out = torch.empty((n, nb), dtype=torch.uint8, device=codes.device)
    w = (2 ** torch.arange(8, device=codes.device)).to(torch.int32)
    bad = False
    for s in range(0, n, chunk):
        x = codes[s:s + chunk]
        bad = bad or bool(((x == 0) | (x != x)).any())
        pos = torch.zeros((x.shape[0], nb * 8), dtype=torch.int32, device=codes.device)
        pos[:, :nbit] = x > 0
        out[s:s + chunk] = (pos.view(-1, nb, 8) * w).sum(-1).to(torch.uint8)
    return out, bad


def parity_check(ev, w, d, dl, q, ql, rank, world, device, nsub=32):
    """The timed configuration against the CPU oracle, outside the timed region.  Every rank runs the collective
    GPU calls; rank 0 runs the oracle over the concatenated gallery."""
    from oracle import packed_oracle as po                # checker only
    R, nbit, nq = w["R"], int(q.shape[1]), int(q.shape[0])
    ndb_total = int(d.shape[0]) * (world if w["scaling"] == "weak" else 1)
    if ndb_total > 20_000_000:
        nsub = min(nsub, 16)                               # the oracle costs ~1 s per query per 100 M rows
    sub = torch.unique(torch.linspace(0, nq - 1, nsub).round().long()).to(device)
    _, _, _, ap = ev.evaluate(d, dl, q, ql, [R], 0.0, [], False, return_ap=True)
    mode = ev.stats.get("mode")
    ids, keys, _ = ev.retrieve(d, q[sub].contiguous(), R)
    bits, bad = _torch_sign_bytes(d)
    lab = dl.to(torch.int32).contiguous()
    if world > 1:
        import torch.distributed as dist
        n_loc = torch.tensor([d.shape[0]], dtype=torch.int64, device=device)
        n_all = [torch.zeros_like(n_loc) for _ in range(world)]
        dist.all_gather(n_all, n_loc)
        n_all = [int(t.item()) for t in n_all]
        nmax = max(n_all)
        pb = torch.zeros((nmax, bits.shape[1]), dtype=torch.uint8, device=device)
        pb[:bits.shape[0]] = bits
        pl = torch.zeros((nmax,), dtype=torch.int32, device=device)
        pl[:lab.shape[0]] = lab
        gb = torch.empty((world * nmax, bits.shape[1]), dtype=torch.uint8, device=device)
        gl = torch.empty((world * nmax,), dtype=torch.int32, device=device)
        dist.all_gather_into_tensor(gb, pb)
        dist.all_gather_into_tensor(gl, pl)
        del pb, pl
        if rank == 0:
            bits = torch.cat([gb[r * nmax:r * nmax + n_all[r]] for r in range(world)])
            lab = torch.cat([gl[r * nmax:r * nmax + n_all[r]] for r in range(world)])
        del gb, gl
    res = None
    if rank == 0:
        t0 = time.perf_counter()
        g64 = bits.cpu().numpy().view(np.uint64)
        qbits, qbad = _torch_sign_bytes(q[sub])
        q64 = qbits.cpu().numpy().view(np.uint64)
        oids, odist = po.topk_packed(q64, g64, R, nbit)
        oap = po.ap_of_ranked(oids, ql[sub].cpu().numpy(), lab.cpu().numpy())
        got_ap = ap[0][sub].cpu().numpy()
        ids_equal = bool(np.array_equal(ids.cpu().numpy(), oids))
        dist_equal = bool(np.array_equal(keys.cpu().numpy().astype(np.int32), odist))
        delta = float(np.abs(got_ap - oap).max())
        res = {"queries": int(sub.numel()), "gallery_rows": int(g64.shape[0]), "R": R,
               "max_abs_ap_delta": delta, "ids_equal": ids_equal, "dist_equal": dist_equal,
               "ok": bool(ids_equal and dist_equal and delta <= 1e-9 and not bad and not qbad),
               "mode": mode, "oracle": "oracle/packed_oracle.py (popcount + stable (distance, row) order)",
               "oracle_seconds": round(time.perf_counter() - t0, 2)}
    del bits, lab
    ok = torch.tensor([1 if (res is None or res["ok"]) else 0], dtype=torch.int32, device=device)
    if world > 1:
        torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
    return res, bool(ok.item())


def tie_order_delta(w, d, dl, q, ql, nsub=512):
    """mAP(upstream's torch.topk tie order) - mAP(canonical stable order) on a query subset, CPU oracle: what a
    maintainer would see move against numbers produced by the old utils.hashing."""
    from oracle import map_oracle as mo
    sub = torch.unique(torch.linspace(0, q.shape[0] - 1, nsub).round().long())
    dc, dlc, qc, qlc = d.cpu(), dl.cpu(), q.cpu()[sub], ql.cpu()[sub]
    a, _, _ = mo.calculate_mAP(dc, dlc, qc, qlc, w["R"], tie="torch_topk")
    b, _, _ = mo.calculate_mAP(dc, dlc, qc, qlc, w["R"], tie="stable")
    return {"queries": int(sub.numel()), "mAP_torch_topk_order": a, "mAP_canonical": b, "delta": a - b}


# ---------------------------------------------------------------------------------------------------- one workload
def measure(ctx, name, nbit, steps, warmup, *, main, e2e=True, cpu=True):
    """Times one workload on this rank set; returns the fields of its JSON object (rank 0) and the parity verdict."""
    from concepthash_b200 import hashing
    rank, world, device, group, ev, peaks, peak_src = (ctx[k] for k in
                                                       ("rank", "world", "device", "group", "ev", "peaks", "peak_src"))
    wl = WORKLOADS[name]
    # weak scaling: every rank generates its own gallery block (same queries), no duplicates across ranks
    w, d, dl, q, ql = make_workload(name, device, nbit, shard=rank if wl["scaling"] == "weak" else 0)
    unit64 = w["nbit"] / 64.0
    ndb_full = d.shape[0]
    if w["scaling"] == "strong" and world > 1:        # row-shard the named gallery over the ranks
        cut = [ndb_full * r // world for r in range(world + 1)]
        d, dl = d[cut[rank]:cut[rank + 1]].contiguous(), dl[cut[rank]:cut[rank + 1]].contiguous()
    total_pairs = float(w["nq"]) * (ndb_full if w["scaling"] == "strong" else ndb_full * world)
    r_list = [w["R"]]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def run(fn, steps, warmup, sample_clocks=False, detail=False, profile=True):
        out = None
        for _ in range(warmup):
            out = fn()
        sampler = ClockSampler(device.index) if sample_clocks else None
        barrier()
        if sampler:
            sampler.start()
        ev.events = []
        ev.profile = profile
        ev.profile_all = detail
        l0 = ev.b.launch_count() + getattr(ev, "graph_launches", 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ev.profile = ev.profile_all = False
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=device)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return (float(ms.item()), out, ev.b.launch_count() + getattr(ev, "graph_launches", 0) - l0, clocks,
                list(ev.events))

    # value: inputs resident in HBM (fp32 codes + int64 ids as torch CUDA tensors)
    thr = float(ctx.get("threshold", 0.0)) if main else 0.0
    step_dev = lambda: ev.evaluate(d, dl, q, ql, r_list, thr, [], False)
    ms, out, launches, clocks, events = run(step_dev, steps, warmup, sample_clocks=True)
    stats = dict(ev.stats)
    # Small evaluations are replayed from a CUDA graph (one launch per evaluation) -- but not while the evaluator
    # brackets its kernels with events, as it does in the loop above.  For them the timed region is run a second time
    # without the brackets: that loop gives ms_per_step, the bracketed one the per-kernel times.
    graph_info = None
    if getattr(ev, "use_graphs", False) and (world == 1 or getattr(ev, "graph_multi_gpu", False)):
        ms_g, out_g, launches_g, clocks_g, _ = run(step_dev, steps, warmup, sample_clocks=True, profile=False)
        if ev.stats.get("speculation") == "graph":
            assert out_g[0] == out[0], (out_g, out)
            graph_info = {"ms_per_step_eager_with_event_brackets": ms, "launches_per_step_captured": launches_g / steps}
            ms, launches, clocks = ms_g, launches_g, clocks_g
    value = total_pairs * unit64 / (ms * 1e-3)

    # per-kernel times from the CUDA events recorded on the launch stream inside the timed region
    kinds, biggest = {}, {}
    for kind, units, a, b in events:
        t = a.elapsed_time(b)
        k = kinds.setdefault(kind, [0.0, 0.0, 0])
        k[0] += t
        k[1] += units
        k[2] += 1
        if kind not in biggest or units > biggest[kind][1]:
            biggest[kind] = [0.0, units, 0]
        if units == biggest[kind][1]:
            biggest[kind][0] += t
            biggest[kind][2] += 1
    words32 = max(1, (w["nbit"] + 31) // 32)
    popc_peak = ctx["popc_peak"]
    sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    sm_count = ev.b.sm_count

    def ncu_traffic(kernel):
        """DRAM bytes per launch of `kernel` from the committed ncu --set full capture of this workload, or None"""
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                t = json.load(f)
            return t.get(name, {}).get(kernel)
        except Exception:
            return None

    def popc_roofline(kind):
        hk = kinds[kind]
        t_ms = hk[0] / hk[2]
        popc = (hk[1] / hk[2]) * words32 / (t_ms * 1e-3)
        what = {"hist_count": "count pass: every pair histogrammed",
                "hist_count_rec": "count pass + records of relevant pairs (full ranking)",
                "hist_select": "select pass: pairs with key <= threshold counted / matched / recorded"}[kind]
        return {
            "kernel": f"hamming_hist_kernel ({what})", "bound": "int-pipe (POPC)",
            "achieved": popc / 1e9, "peak": popc_peak / 1e9, "unit": "Gpopc32/s", "frac": popc / popc_peak,
            "traffic": ncu_traffic("hamming_hist_kernel"),
            "traffic_source": "profiles/ncu_traffic.json (committed ncu --set full capture, not measured in this run)",
            "algorithmic_work": f"pairs x {words32} popc32 per pair per launch",
            "peak_source": "ch_popc_peak micro-benchmark run live on this GPU (MEASURED_PEAKS.json has no "
                           "integer-pipe figure); nominal 148 SM x 16 lanes x f",
            "nominal_peak_at_sampled_clock": sm_count * 16 * sm_mhz * 1e6 / 1e9,
            "ms_per_launch": t_ms, "pairs_per_launch": hk[1] / hk[2], "pairs_per_s": (hk[1] / hk[2]) / (t_ms * 1e-3),
            "share_of_step": hk[0] / (ms * steps)}

    def tensor_roofline(kind):
        hk = kinds[kind]
        t_ms = hk[0] / hk[2]
        pairs = hk[1] / hk[2]
        # K bytes the MMA really contracts per gallery ROW: codes + the block that carries the per-query threshold; in
        # the paired form (two rows per accumulator cell) a plane row of kbp bytes serves two gallery rows
        rows_per_cell = int(stats.get("select_rows_per_cell") or 1)
        if rows_per_cell == 2:
            kbp = ev.b.tc_code_bytes_pair(w["nbit"], bool(stats.get("ternary")))
            kb = kbp / 2.0
            k_note = (f"paired form: one plane row of K = {kbp} bytes = two gallery rows (2 x {w['nbit']} code bytes + the "
                      "block that carries the per-query threshold)")
            floor = 128.0 * 256.0 / (64.0 * (kbp // 32))
        else:
            kb = float(ev.b.tc_code_bytes(w["nbit"], stats.get("select_threshold") == "epilogue"))
            k_note = f"the kernel contracts K = {int(kb)} bytes per row: the codes plus the per-query threshold block"
            floor = 128.0 * 128.0 / (64.0 * (int(kb) // 32))
        tops = pairs * w["nbit"] * 2 / (t_ms * 1e-3) / 1e12   # ALGORITHMIC: nbit int8 MACs per pair
        peak2 = 2.0 * peaks.get("bf16_tflops", 1590.0)
        issue_peak = sm_count * MMA_ISSUE_MAC_PER_CLK_SM * 2 * sm_mhz * 1e6 / 1e12
        pairs_s = pairs / (t_ms * 1e-3)
        return {
            "kernel": "hamming_select_tc_kernel (select pass: tcgen05.mma kind::i8 -> TMEM, sign-bit epilogue, "
                      "candidate lists)",
            "bound": "tensor", "achieved": tops, "peak": issue_peak, "unit": "TOP/s (int8)", "frac": tops / issue_peak,
            "traffic": ncu_traffic("hamming_select_tc_kernel"),
            "traffic_source": "profiles/ncu_traffic.json (committed ncu --set full capture, not measured in this run)",
            "algorithmic_work": f"pairs x {w['nbit']} int8 MACs x 2 per launch ({k_note})",
            "peak_source": "measured int8 MMA issue rate: dev/mma_rate_probe.cu (profiles/r1i_mma_rate_probe.txt), 64.0 clk "
                           "per UTCIMMA M=N=128 K=32 = 8192 MAC/clk/SM, x SMs x the SM clock sampled in this run.  "
                           f"MEASURED_PEAKS.json holds no int8 figure; 2 x its bf16_tflops ({peak_src}) = {peak2:.0f} TOP/s "
                           "is BELOW what this kernel sustains, so it is reported beside, not as the peak; nominal "
                           "dense int8 is 4500 TOP/s",
            "frac_of_2x_measured_bf16_peak": tops / peak2,
            "measured_mma_issue_peak_tops": issue_peak,
            "frac_of_measured_mma_issue_peak": tops / issue_peak,
            "frac_of_measured_mma_issue_peak_executed": tops * kb / w["nbit"] / issue_peak,
            "executed_tops_incl_threshold_block": tops * kb / w["nbit"],
            "frac_of_nominal_int8_4500": tops / 4500.0,
            "gallery_rows_per_accumulator_cell": rows_per_cell,
            "pairs_per_clk_per_sm": pairs_s / (sm_count * sm_mhz * 1e6),
            "mma_floor_pairs_per_clk_per_sm": floor,
            "popc_kernel_ceiling_pairs_per_s": popc_peak / words32,
            "ms_per_launch": t_ms, "launches_per_step": hk[2] / steps, "pairs_per_launch": pairs,
            "pairs_per_s": pairs_s, "share_of_step": hk[0] / (ms * steps)}

    ham_kinds = [k for k in kinds if k.startswith("hist")]
    roofline, roofline_other = None, {}
    if ham_kinds:
        dom = max(ham_kinds, key=lambda k: kinds[k][0])
        roofline = tensor_roofline(dom) if dom == "hist_select_tc" else popc_roofline(dom)
        roofline_other = {k: (tensor_roofline(k) if k == "hist_select_tc" else popc_roofline(k))
                          for k in ham_kinds if k != dom}
    roofline_pack = None
    pack = biggest.get("pack_dev")
    if pack and pack[0] > 0 and hasattr(ev.b, "pack_sign"):
        # K1 on the gallery codes, timed as launches queued BACK TO BACK (10 launches between two events).  Inside a
        # step the launch is the first thing on an idle stream: its event bracket there also spans the host's launch
        # latency (kept beside as ms_per_launch_in_step_bracket) and understates the kernel.
        fl = ev.b.zeros((1,), torch.int32)
        for _ in range(3):
            ev.b.pack_sign(d, thr, fl, bool(thr))
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(10):
            ev.b.pack_sign(d, thr, fl, bool(thr))
        p1.record()
        torch.cuda.synchronize()
        ms_pack = p0.elapsed_time(p1) / 10
        nbytes = d.shape[0] * d.shape[1] * d.element_size() + d.shape[0] * ((d.shape[1] + 31) // 32) * 4 * (2 if thr else 1)
        gbs = nbytes / (ms_pack * 1e-3) / 1e9
        roofline_pack = {"kernel": "pack_sign_flat_kernel / pack_bits_kernel (sign + bit-pack of the gallery codes)",
                         "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": gbs / peaks["hbm_gbs"], "traffic": ncu_traffic("pack_sign_flat_kernel"),
                         "peak_source": peak_src, "ms_per_launch": ms_pack, "bytes_per_launch": nbytes,
                         "timing": "10 launches queued back to back between two CUDA events, outside the timed region",
                         "ms_per_launch_in_step_bracket": pack[0] / pack[2]}
    kernel_ms = {k: v[0] / steps for k, v in kinds.items()}
    # diagnostic pass (NOT the timed region: a bracket around every entry point costs two event records each):
    # where the rest of the step goes
    dsteps = 3
    _, _, _, _, dev_events = run(step_dev, dsteps, 0, detail=True)
    detail_ms = {}
    for kind, units, a, b in dev_events:
        detail_ms[kind] = detail_ms.get(kind, 0.0) + a.elapsed_time(b) / dsteps
    coll_ms = sum(v for k, v in kernel_ms.items() if k.startswith("comm_"))
    kern_ms = sum(v for k, v in detail_ms.items() if not k.startswith("comm_"))
    if graph_info is not None:
        graph_info["note"] = ("the evaluation is replayed from a CUDA graph: ms_per_step is the replay loop (no event "
                              "brackets); kernel_ms_* come from the bracketed eager loop")
    phase_ms = {"kernels": kern_ms, "collectives": coll_ms, "host_and_launch_gaps": max(0.0, ms - kern_ms - coll_ms),
                "host_syncs_per_step": stats.get("host_syncs"), "cuda_graph": graph_info,
                "note": "kernels = every entry point bracketed in a separate diagnostic pass; collectives = NCCL calls "
                        "bracketed in the timed region; the rest of ms_per_step is host time / launch gaps"}

    # e2e: the public call with HOST tensors; H2D of codes + labels and D2H of the result inside.  Led by what the
    # reference's trainers really hand over -- PAGEABLE tensors (trainers/base.py:291-304) -- with pinned beside it.
    e2e_obj, e2e_ok = None, True
    if e2e:
        es, ew = max(2, steps // 2), 2
        pd, pdl, pq, pql = (t.cpu() for t in (d, dl, q, ql))
        step_page = lambda: hashing.calculate_mAP(pd, pdl, pq, pql, w["R"], threshold=thr, group=group)
        # the timed loop runs WITHOUT the per-kernel event brackets (with them the evaluator keeps the gallery blocks on
        # its own thread instead of the loader thread); the per-kernel times come from a second, bracketed loop
        ms_p, out_p, _, _, _ = run(step_page, es, ew, profile=False)
        _, _, _, _, ev_p = run(step_page, es, 1)
        kinds_p = {}
        for kind, units, a, b in ev_p:
            kinds_p[kind] = kinds_p.get(kind, 0.0) + a.elapsed_time(b) / es
        mode_p = ev.stats.get("mode")
        hd, hdl, hq, hql = (t.pin_memory() for t in (pd, pdl, pq, pql))
        del pd, pdl, pq, pql
        step_host = lambda: hashing.calculate_mAP(hd, hdl, hq, hql, w["R"], threshold=thr, group=group)
        ms_e, out_e, _, _, _ = run(step_host, es, ew, profile=False)
        h2d = sum(t.numel() * t.element_size() for t in (hd, hdl, hq, hql))
        # what crosses PCIe: fp32 codes in host memory are sign/bit-packed by the host's cores (csrc/loader.cu,
        # host_pack.cpp) -- 1 bit per code travels, plus the row sample's bits; labels travel as they are
        sample_rows = (ev.stats.get("sample") or {}).get("rows", 0)
        wire = (sum(t.shape[0] for t in (hd, hq)) + sample_rows) * ((w["nbit"] + 31) // 32 * 4) + \
            sum(t.numel() * t.element_size() for t in (hdl, hql))
        del hd, hdl, hq, hql
        e2e_obj = {"value": total_pairs * unit64 / (ms_p * 1e-3), "unit": "64-bit comparisons/s",
                   "ms_per_step": ms_p, "host_memory": "pageable (what trainers/base.py:291-304 hands over)",
                   "ms_per_step_pinned_host_tensors": ms_e,
                   "value_pinned_host_tensors": total_pairs * unit64 / (ms_e * 1e-3),
                   "h2d_bytes_per_step": wire, "host_input_bytes_per_step": h2d,
                   "h2d_note": "fp32 codes are read once by the host's cores (host_input_bytes_per_step) and cross "
                               "PCIe as packed sign bits (h2d_bytes_per_step); pageable and pinned alike",
                   "d2h_bytes_per_step": 8 * len(r_list) + 64,
                   "mAP": out_p[0], "mAP_pinned": out_e[0], "mode": mode_p, "kernel_ms_per_step": kinds_p}
        if "dataset" in wl:
            # the label form the reference's trainers hand over: one-hot (N, C) float32 rows (trainers/base.py:291-304)
            from concepthash_b200 import synth as _synth
            ohd, ohq = _synth.one_hot(dl.cpu(), w["nclass"]), _synth.one_hot(ql.cpu(), w["nclass"])
            pd2, pq2 = d.cpu(), q.cpu()
            step_oh = lambda: hashing.calculate_mAP(pd2, ohd, pq2, ohq, w["R"], threshold=thr, group=group)
            ms_oh, out_oh, _, _, _ = run(step_oh, es, ew, profile=False)
            e2e_obj["ms_per_step_onehot_labels"] = ms_oh
            e2e_obj["labels_note"] = ("ms_per_step: 1-D int64 class ids; ms_per_step_onehot_labels: the one-hot (N, C) "
                                      "float32 label rows of trainers/base.py:291-304, pageable")
            e2e_obj["mAP_onehot"] = out_oh[0]
            del ohd, ohq, pd2, pq2
        # the host paths (native loader, streamed blocks) must return the device-resident answer to the last bit:
        # the candidate lists and the in-order walk do not depend on where the codes came from
        flat = lambda v: [float(x) for x in (v if isinstance(v, (list, tuple)) else [v])]
        e2e_obj["equals_resident"] = bool(out is not None and flat(out_p[0]) == flat(out[0]) == flat(out_e[0]) and
                                          ("mAP_onehot" not in e2e_obj or flat(e2e_obj["mAP_onehot"]) == flat(out[0])))
        e2e_ok = e2e_obj["equals_resident"]

    if thr == 0.0:
        pc, ok = parity_check(ev, w, d, dl, q, ql, rank, world, device)
        ok = ok and e2e_ok
    else:
        pc, ok = {"skipped": "ternary codes (--threshold): the packed-bit oracle is binary; see tests"}, True

    cpu_obj = None
    if cpu and rank == 0 and world == 1:
        r = cpu_reference_sample(w, d, dl, q, ql)
        cpu_obj = {"value": r["pairs"] * unit64 / r["seconds"], "unit": "64-bit comparisons/s", "cores": r["cores"],
                   "kind": "port", "threads": r["threads"], "seconds": r["seconds"],
                   "sample": f"{r['queries']} of {w['nq']} queries x {w['ndb']} gallery rows (upstream-style "
                             f"oracle/map_oracle.calculate_mAP_upstream_style; linear in nq)"}
    tie = None
    if rank == 0 and world == 1 and "dataset" in wl:
        tie = tie_order_delta(w, d, dl, q, ql)

    config = {"workload": w["desc"], "nq": w["nq"], "ndb_per_gpu": int(d.shape[0]),
              "ndb_total": int(ndb_full if w["scaling"] == "strong" else ndb_full * world),
              "nbit": w["nbit"], "R": w["R"], "nclass": w["nclass"], "ternary_threshold": thr,
              "pairs_per_s": total_pairs / (ms * 1e-3),
              "mode": stats.get("mode"), "geometry(threads,nq_pad,stripes,rows/stripe)": stats.get("geometry"),
              "l2_policy": "inputs larger than L2 (fp32 gallery codes >= 512 MB vs 126 MB L2)"
              if w["ndb"] * w["nbit"] * 4 > 2.0e8 else "small workload: inputs fit L2 (latency-bound case)",
              "parallelism": f"gallery row-sharded x{world}" if world > 1 else "single GPU",
              "mAP": out[0][0] if out else None}
    dtype = ("s8 x s8 -> s32 (tcgen05 kind::i8 select pass) + u32 xor/popc (sample, count and key passes)"
             if "hist_select_tc" in kinds else "u32 xor/popc")
    if rank == 0:
        print(f"[{name}/{w['nbit']}] evaluator stats:", stats, file=sys.stderr)
    del d, dl, q, ql
    if hasattr(ev, "release_graphs"):
        ev.release_graphs()                  # (a captured graph owns its arenas: give them back before the next workload)
    torch.cuda.empty_cache()
    if main:
        return {
            "metric": "hamming_comparisons_per_sec_64bit", "value": value, "unit": "64-bit comparisons/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": dtype,
            "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e_obj, "gpu_launches": launches,
            "kernel_ms_per_step": kernel_ms, "kernel_ms_detail": detail_ms, "phase_ms": phase_ms,
            "parity_check": pc, "roofline": roofline, "roofline_other_passes": roofline_other, "roofline_pack": roofline_pack,
            "cpu_baseline": cpu_obj, "tie_order_delta": tie,
        }, ok
    slim = None
    if roofline is not None:
        slim = {k: roofline[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "ms_per_launch",
                                         "share_of_step") if k in roofline}
        if "frac_of_measured_mma_issue_peak" in roofline:
            slim["frac_of_measured_mma_issue_peak"] = roofline["frac_of_measured_mma_issue_peak"]
    return {
        "workload": w["desc"], "nbit": w["nbit"], "R": w["R"], "scaling": w["scaling"], "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms, "value": value, "unit": "64-bit comparisons/s",
        "mAP": config["mAP"], "mode": stats.get("mode"), "ndb_per_gpu": config["ndb_per_gpu"],
        "ndb_total": config["ndb_total"], "gpu_launches": launches, "kernel_ms_per_step": kernel_ms,
        "kernel_ms_detail": detail_ms, "phase_ms": phase_ms, "clocks": clocks, "roofline": slim, "parity_check": pc, "tie_order_delta": tie,
        "e2e": e2e_obj,
    }, ok


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
    ap.add_argument("--no-others", action="store_true", help="skip the other_workloads block")
    ap.add_argument("--sample-stride", type=int, default=0, help="override the evaluator's row-sample stride")
    ap.add_argument("--threshold", type=float, default=0.0,
                    help="ternary_threshold (configs/val.yaml:12): |code| < threshold -> 0; the packed-bit oracle of "
                         "parity_check handles binary codes only, so the check is skipped (tests cover ternary)")
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

    ev = hashing.get_evaluator(device, group)
    if args.sample_stride:
        ev.sample_stride = args.sample_stride
    if world > 1 and os.environ.get("CH_GRAPH_MULTI_GPU", "1") == "1":
        # small per-rank problems (cfg4 over 8 GPUs): the launch sequence incl. its NCCL collectives is replayed from a
        # CUDA graph; the graphs are released before the process group goes away (evaluator.release_graphs)
        ev.graph_multi_gpu = True
    peaks, peak_src = measured_peaks()
    popc_peak, _ = ev.b.popc_peak()
    ctx = dict(rank=rank, world=world, device=device, group=group, ev=ev, peaks=peaks, peak_src=peak_src,
               popc_peak=popc_peak, threshold=args.threshold)

    try:
        return _measure_all(args, ctx, ev, rank, world, wl)
    finally:
        ev.release_graphs()
        if world > 1:
            torch.distributed.destroy_process_group()


def _measure_all(args, ctx, ev, rank, world, wl):
    line, ok_all = measure(ctx, args.workload, args.nbit, args.steps, args.warmup, main=True,
                           e2e=not args.no_e2e, cpu=not args.no_cpu_baseline)
    others = []
    if not args.no_others:
        plan = []
        if args.workload != "cfg5":
            plan.append(("cfg5", 0, 5, 3))              # configs[4]: weak-scaled shard, at every N
        if args.workload != "cfg5full":
            plan.append(("cfg5full", 0, 3, 3))          # configs[4] as named: the whole 100M gallery, strong-scaled
        if world == 1:
            plan += [(n, b, 10, 3) for n, b in (("cub200", 64), ("cars196", 16), ("cars196", 32), ("cars196", 64),
                                                ("nabirds", 64)) if not (n == args.workload and b == wl["nbit"])]
        for name, nbit, steps, warm in plan:
            # (the dataset shapes are what the reference's trainers evaluate, from pageable CPU tensors: their e2e too)
            obj, ok = measure(ctx, name, nbit, steps, warm, main=False, e2e="dataset" in WORKLOADS[name], cpu=False)
            others.append(obj)
            ok_all = ok_all and ok
    if rank == 0:
        line["other_workloads"] = others
        line["parity_ok"] = ok_all
        print(json.dumps(line))
    if not ok_all:
        if rank == 0:
            print("PARITY CHECK FAILED (see parity_check in the JSON line)", file=sys.stderr)
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
