"""Where does a narrow tcgen05 GEMM spend its time? Same launch with and without epilogue stores."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wavtokenizer_b200 import _native
lib = _native.lib()
dev = "cuda:0"


def t(rows, Cin, N, passes, mode):
    A = torch.randn(rows, Cin, device=dev) * 0.5
    W = torch.randn(N, Cin, device=dev) * 0.05
    b = torch.randn(N, device=dev)
    out = torch.empty(rows, N, device=dev)
    osp = torch.empty(rows, N, device=dev)
    args = dict(f32=(out.data_ptr(), None), split=(None, osp.data_ptr()), none=(None, None))[mode]
    # the hook converts A/W to planes and syncs on every call: time the GEMM alone with the timeline stamps
    dbg = torch.zeros(148 * 64, dtype=torch.int64, device=dev)
    lib.wt_debug_timeline(dbg.data_ptr())
    _native.check(lib.wt_test_tap_gemm(0, A.data_ptr(), rows, Cin, 1, W.data_ptr(), N, b.data_ptr(), None, None, 0, passes,
                                       args[0], args[1], None))
    lib.wt_debug_timeline(None)
    d = dbg.view(148, 64).cpu()
    return int(d[:, 43].max())  # cycles until the slowest CTA's epilogue finished its last tile


for rows, Cin, N in [(4 * 72002, 128, 16), (4 * 72002, 128, 32), (4 * 36001, 128, 64), (8 * 9001, 512, 128)]:
    tiles = (rows + 127) // 128
    res = {m: t(rows, Cin, N, 3, m) for m in ("none", "f32", "split")}
    ncta = int(os.environ.get('WT_TC_GRID', '148'))
    per = {m: round(v / (tiles / ncta)) for m, v in res.items()}
    print(f"rows {rows} K {Cin} N {N}: cycles/tile/SM  no-store {per['none']}  fp32 {per['f32']}  split planes {per['split']}")
