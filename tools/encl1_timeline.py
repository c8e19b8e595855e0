"""Per-tile critical path of the fused encoder level 0 -> 1 kernel (CTA 0, its tiles 2..5). Needs a WT_TIMELINE=1 build."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.gpu_util import native_model
from wavtokenizer_b200 import _native, spec
lib = _native.lib()
m = native_model("small320", 2)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
wav = spec.synthetic_audio(B, 72000, seed=1).cuda()
bw = torch.tensor([0]).cuda()
m.encode_infer(wav, bandwidth_id=bw)
torch.cuda.synchronize()
dbg = torch.zeros(148 * 64 + 64 + 64, dtype=torch.int64, device="cuda")
lib.wt_debug_timeline(dbg.data_ptr())
m.encode_infer(wav, bandwidth_id=bw)
torch.cuda.synchronize()
lib.wt_debug_timeline(None)
d = dbg[148 * 64 + 64:].view(4, 16).cpu()
names = ["A0 TMA issued", "GEMM1 issued", "GEMM2 issued", "GEMM3 issued", "ep1 starts", "ep2 starts", "ep3 starts", "ep3 done"]
for i in range(4):
    base = int(d[i][1])
    print(f"tile {4+i}: t0={base}", {n: int(d[i][k]) - base for k, n in enumerate(names)})
print("tile period (GEMM1 issue to GEMM1 issue):", [int(d[i + 1][1] - d[i][1]) for i in range(3)])
