"""Decode repeatability check (run on a B200): same codes decoded repeatedly, after other work, and in child processes."""
import os, subprocess, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.gpu_util import native_model, Taps
from wavtokenizer_b200 import spec
NAMES = ["dec_embed", "dec_pos0", "dec_pos1", "dec_pos2", "dec_pos4", "dec_norm", "dec_cnx0", "dec_cnx1", "dec_cnx5", "dec_cnx11", "dec_final", "dec_headlin"]
def run(m, codes, bw):
    tp = Taps(m, NAMES)
    a = m.decode(m.codes_to_features(codes), bandwidth_id=bw)
    out = {n: tp.get(n) for n in NAMES}
    tp.close()
    out["audio"] = a.cpu()
    return out
def cmp(tag, A, B):
    for n in NAMES + ["audio"]:
        x, y = A[n], B[n]
        if torch.equal(x, y):
            continue
        d = (x - y).abs()
        idx = (d > 0).nonzero()
        print(tag, n, "DIFF count", idx.shape[0], "of", d.numel(), "max", d.max().item(), "ref max", x.abs().max().item(), "first", idx[:3].tolist(), "last", idx[-3:].tolist())
        return
    print(tag, "all equal")
m = native_model("small320", 2)
g = torch.Generator().manual_seed(5)
codes = torch.randint(0, 4096, (1, 3, 75), generator=g).cuda()
bw = torch.tensor([0]).cuda()
if len(sys.argv) > 1:
    torch.save(run(m, codes, bw), sys.argv[1]); sys.exit(0)
A = run(m, codes, bw)
A2 = run(m, codes, bw)
cmp("same-process repeat", A, A2)
wav = spec.synthetic_audio(16, 48000, seed=9).cuda()
f, c = m.encode_infer(wav, bandwidth_id=bw)
_ = m.decode(m.codes_to_features(c), bandwidth_id=bw)
B = run(m, codes, bw)
cmp("after dirtying the arena", A, B)
for env in ({}, {"WT_MEM_V1": "1"}, {"WT_LSTM_PUBLISH": "0", "WT_LSTM_KEEP_C": "0"}):
    e = dict(os.environ); e.update(env)
    subprocess.run([sys.executable, __file__, "/tmp/diag_child.pt"], env=e, check=True)
    cmp(f"child {env}", A, torch.load("/tmp/diag_child.pt"))
    cmp(f"child {env} vs dirty", B, torch.load("/tmp/diag_child.pt"))
