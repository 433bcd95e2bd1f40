"""Turns `ncu -i X.ncu-rep --page raw --csv` output and a launch-list CSV into the text summaries kept here.
usage: python profiles/summarize_ncu.py raw <raw.csv> | launches <launches.csv>"""
import collections
import csv
import sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput', 'dram__cycles_active', 'sm__throughput', 'warps_active', 'registers_per_thread',
        'occupancy_limit', 'smsp__inst_executed.sum', 'issue_active', 'pipe_alu_cycles', 'pipe_xu', 'pipe_fma_cycles',
        'pipe_tensor', 'cycles_elapsed.avg.per_second', 'bank_conflicts', 'shared_mem_per_block', 'grid_size',
        'block_size', 'waves_per', 'issue_stalled', 'sm__inst_executed_pipe', 'l1tex__data_pipe_lsu_wavefronts_mem_shared']


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    keys = [k for k in hdr if any(s in k for s in KEEP)]
    for r in rows[2:]:
        print('----')
        for w in keys:
            i = hdr.index(w)
            print(w, '=', r[i], units[i])


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[h]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        v *= {'us': 1e-3, 'ns': 1e-6, 's': 1e3, 'ms': 1.0}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki].split('(')[0][:90], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches (cold-cache, serialised: compare SHARES)")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{a[1]:10.3f} ms {100 * a[1] / tot:5.1f}%  n={a[0]:4d}  {k}")


if __name__ == "__main__":
    {"raw": raw, "launches": launches}[sys.argv[1]](sys.argv[2])
