"""Decode repeatability check (run on a B200): same codes decoded repeatedly, after other work, and in child processes."""
import os, subprocess, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.gpu_util import native_model, Taps
NAMES = ["dec_embed", "dec_pos0", "dec_cnx0", "dec_final", "dec_headlin"]
m = native_model("small320", 2)
g = torch.Generator().manual_seed(5)
codes = torch.randint(0, 4096, (1, 3, 75), generator=g).cuda()
bw = torch.tensor([0]).cuda()
tp = Taps(m, NAMES)
def run():
    a = m.decode(m.codes_to_features(codes), bandwidth_id=bw)
    torch.cuda.synchronize()
    out = {n: tp.bufs[n][:3 * 75 * 1300].clone() for n in NAMES}
    out["audio"] = a.clone()
    return out
ref = run()
nrep = int(os.environ.get("NREP", "150"))
bad = 0
for i in range(nrep):
    o = run()
    for n in NAMES + ["audio"]:
        if not torch.equal(o[n], ref[n]):
            d = (o[n] - ref[n]).abs()
            idx = (d > 0).nonzero().flatten()
            print("rep", i, n, "count", idx.numel(), "max", d.max().item(), "idx range", idx.min().item(), idx.max().item(), flush=True)
            bad += 1
            break
print("in-process repeats differing:", bad, "of", nrep, flush=True)
if len(sys.argv) > 1:
    torch.save({k: v.cpu() for k, v in ref.items()}, sys.argv[1]); sys.exit(0)
R = {k: v.cpu() for k, v in ref.items()}
for c in range(6):
    e = dict(os.environ); e["NREP"] = "10"
    subprocess.run([sys.executable, __file__, "/tmp/d2.pt"], env=e, check=True, stdout=subprocess.DEVNULL)
    o = torch.load("/tmp/d2.pt")
    msg = "equal"
    for n in NAMES + ["audio"]:
        if not torch.equal(o[n], R[n]):
            d = (o[n] - R[n]).abs(); idx = (d > 0).nonzero().flatten()
            cols = sorted(set((idx % 768).tolist())) if n == "dec_embed" else []
            msg = f"{n} count {idx.numel()} max {d.max().item():.3e} cols {cols[:3]}..{cols[-3:]} ncols {len(cols)}"
            break
    print("child", c, msg, flush=True)
