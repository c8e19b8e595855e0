"""Single tcgen05 GEMM launches for ncu (--set full): pick a shape with --case."""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wavtokenizer_b200 import _native
lib = _native.lib()
CASES = {
    "pw1": dict(rows=57600, Cin=768, taps=1, N=2304, passes=1, act=1, split=True),     # ConvNeXt pwconv1 + GELU, B=256
    "pw2": dict(rows=57600, Cin=2304, taps=1, N=768, passes=1, act=0, split=False),
    "k3": dict(rows=57600, Cin=768, taps=3, N=768, passes=3, act=0, split=False),      # ResnetBlock conv
    "down0": dict(rows=16 * 36001, Cin=128, taps=1, N=64, passes=3, act=0, split=True),
}
ap = argparse.ArgumentParser(); ap.add_argument("--case", default="pw1"); a = ap.parse_args()
c = CASES[a.case]
dev = "cuda:0"
A = torch.randn(c["rows"], c["Cin"], device=dev) * 0.5
W = torch.randn(c["N"], c["taps"] * c["Cin"], device=dev) * 0.05
b = torch.randn(c["N"], device=dev)
out = torch.empty(c["rows"], c["N"], device=dev)
osp = torch.empty(c["rows"], c["N"], device=dev) if c["split"] else None
for _ in range(2):
    _native.check(lib.wt_test_tap_gemm(0, A.data_ptr(), c["rows"], c["Cin"], c["taps"], W.data_ptr(), c["N"], b.data_ptr(),
                                       None, None, c["act"], c["passes"], None if c["split"] else out.data_ptr(),
                                       osp.data_ptr() if c["split"] else None, None))
torch.cuda.synchronize()
print("ok", a.case)
