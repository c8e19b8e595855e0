"""Key metrics per launch from `ncu --page raw --csv` exports (tools/ncu_targets.sh): python tools/ncu_table.py a_raw.csv [b_raw.csv ...]"""
import csv, io, sys
def table(path):
    rows = list(csv.reader(open(path)))
    # find header row starting with "ID"
    hi = next(i for i,r in enumerate(rows) if r and r[0]=="ID")
    hdr, data = rows[hi], rows[hi+2:]
    cols = [("Kernel Name","kernel",44),("gpu__time_duration.sum","us",8),("dram__bytes_read.sum","rdMB",8),("dram__bytes_write.sum","wrMB",8),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","tensor%",7),("smsp__issue_active.avg.pct_of_peak_sustained_active","issue%",7),
            ("smsp__inst_executed.sum","Minst",7),("lts__throughput.avg.pct_of_peak_sustained_elapsed","lts%",6),("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","dram%",6),
            ("lts__t_sectors_srcunit_tex_op_read.sum","l2rdMB",8),("launch__grid_size","grid",6),("launch__registers_per_thread","regs",5)]
    units = rows[hi+1]
    idx=[hdr.index(c[0]) if c[0] in hdr else None for c in cols]
    print(" ".join(f"{c[1]:>{c[2]}}" for c in cols))
    for d in data:
        if len(d) < len(hdr)-2: continue
        out=[]
        for (name,short,w),i in zip(cols,idx):
            v=d[i] if i is not None else ""
            u=units[i] if i is not None else ""
            if short=="kernel": v=v.replace("void ","").replace("wt::<unnamed>::","").replace("(int)","").replace("(bool)","")[:w]
            elif short=="Minst": v=f"{float(v)/1e6:.1f}"
            elif short=="l2rdMB": v=f"{float(v)*32/1e6:.0f}"
            elif short in("rdMB","wrMB"):
                f=float(v); m={"Gbyte":1e3,"Mbyte":1,"Kbyte":1e-3,"byte":1e-6}.get(u,1); v=f"{f*m:.1f}"
            elif v:
                try: v=f"{float(v):.1f}"
                except ValueError: pass
            out.append(f"{v:>{w}}")
        print(" ".join(out))
for p in sys.argv[1:]: table(p)
