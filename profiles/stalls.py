"""Per-instruction stall summary of an ncu report: python profiles/stalls.py X.ncu-rep [min_pct]
(reads `ncu -i X --page source --csv`; needs -lineinfo builds and --import-source on captures)"""
import csv
import subprocess
import sys


def main(path, min_pct=0.8):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[start - 1][1] if start else "")
    hdr = rows[start]
    ix = {k: i for i, k in enumerate(hdr)}
    data = [r for r in rows[start + 1:] if len(r) == len(hdr)]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    agg = {}
    print("total samples", tot, "instructions", len(data))
    for n, r in enumerate(data):
        s = int(r[ix["# Samples"]])
        for k in stall_cols:
            agg[k] = agg.get(k, 0) + int(r[ix[k]])
        if s > tot * min_pct / 100:
            st = {k: int(r[ix[k]]) for k in stall_cols if int(r[ix[k]]) > 0}
            top = sorted(st.items(), key=lambda x: -x[1])[:3]
            print(f"{n:5d} {r[ix['Source']].strip()[:58]:58s} {s:7d} {100 * s / tot:5.1f}%  x{r[ix['Instructions Executed']]:>9s}  {top}")
    print("all:", [(k, f"{100 * v / tot:.1f}%") for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]])


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.8)
