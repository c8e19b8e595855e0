"""Key metrics per launch from an ncu report: python tools/ncu_table.py report.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
cols = [("Kernel Name", "kernel", 34), ("gpu__time_duration.sum", "us", 8), ("dram__bytes_read.sum", "rdMB", 7),
        ("dram__bytes_write.sum", "wrMB", 7), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 7),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 7), ("smsp__inst_executed.sum", "Minst", 7),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%", 6),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 6),
        ("lts__t_sectors_srcunit_tex_op_read.sum", "l2rdMB", 8), ("launch__grid_size", "grid", 6)]
idx = [hdr.index(c[0]) if c[0] in hdr else None for c in cols]
print(" ".join(f"{c[1]:>{c[2]}}" for c in cols))
for d in data:
    out = []
    for (name, short, w), i in zip(cols, idx):
        v = d[i] if i is not None else ""
        if short == "kernel": v = v.replace("void ", "").replace("unnamed>::", "").replace("wt::<", "")[:w]
        elif short == "Minst": v = f"{float(v) / 1e6:.1f}"
        elif short == "l2rdMB": v = f"{float(v) * 32 / 1e6:.0f}"
        elif v:
            try: v = f"{float(v):.1f}"
            except ValueError: pass
        out.append(f"{v:>{w}}")
    print(" ".join(out))
