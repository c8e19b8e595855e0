"""Aggregate the per-instruction warp-stall samples of an `ncu --page source --csv` export.

    python tools/ncu_stalls.py gpurun_out/r2_ncu_rb0_source.csv [top_n]

Prints, per kernel in the file, the share of every stall reason over all sampled instructions and the top-N
instructions by samples with their dominant reason.
"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 15
rows = list(csv.reader(open(path, newline="")))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) >= len(hdr) - 2:
                body.append(rows[j])
            j += 1
        col = {h: k for k, h in enumerate(hdr)}
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = defaultdict(float)
        samples = 0.0
        per = []
        for r in body:
            try:
                n = float(r[col["# Samples"]])
            except (ValueError, IndexError):
                continue
            samples += n
            best = ("", 0.0)
            for h in stall_cols:
                try:
                    v = float(r[col[h]])
                except (ValueError, IndexError):
                    v = 0.0
                tot[h] += v
                if v > best[1]:
                    best = (h, v)
            per.append((n, r[col["Source"]].strip(), best[0], r[col["Instructions Executed"]]))
        print(f"== {name[:110]}  ({int(samples)} samples, {len(body)} instructions)")
        for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
            print(f"   {h:28s} {100 * v / max(samples, 1):5.1f} %")
        for n, src, why, ex in sorted(per, key=lambda t: -t[0])[:top_n]:
            print(f"   {100 * n / max(samples, 1):5.1f} %  {src[:70]:70s} {why:22s} exec {ex}")
        i = j
    else:
        i += 1
