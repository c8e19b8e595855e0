"""Per-step critical path of the persistent LSTM kernel (CTA 0, steps 4..7)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.gpu_util import native_model
from wavtokenizer_b200 import _native, spec
lib = _native.lib()
m = native_model("small320", 2)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
wav = spec.synthetic_audio(B, 72000, seed=1).cuda()
bw = torch.tensor([0]).cuda()
m.encode_infer(wav, bandwidth_id=bw)
torch.cuda.synchronize()
dbg = torch.zeros(148 * 64 + 64, dtype=torch.int64, device="cuda")  # generic GEMMs stamp 64 slots per CTA
lib.wt_debug_timeline(dbg.data_ptr())
m.encode_infer(wav, bandwidth_id=bw)
torch.cuda.synchronize()
lib.wt_debug_timeline(None)
d = dbg[148 * 64:].view(8, 8).cpu()  # last layer's launch overwrote the first
names = ["published(t-1) seen", "last TMA issued", "first kb landed", "last kb landed", "acc ready", "cell done", "published", "all warps stored h"]
for i in range(4):
    base = int(d[i][0])
    print(f"step {4+i}: t0={base}", {n: int(d[i][k]) - base for k, n in enumerate(names)})
print("step period:", int(d[1][0] - d[0][0]), int(d[2][0] - d[1][0]), int(d[3][0] - d[2][0]))
