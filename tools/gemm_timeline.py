"""Critical-path timeline of one tcgen05 GEMM launch (clock64 stamps per role), for shapes of interest."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wavtokenizer_b200 import _native  # noqa: E402

lib = _native.lib()


def run(rows, Cin, taps, N, passes, act=0, label="", split=False):
    dev = "cuda:0"
    A = torch.randn(rows, Cin, device=dev) * 0.5
    W = torch.randn(N, taps * Cin, device=dev) * 0.05
    b = torch.randn(N, device=dev)
    out = torch.empty(rows, N, device=dev)
    dbg = torch.zeros(148 * 64, dtype=torch.int64, device=dev)
    osplit = torch.empty(rows, N, device=dev) if split else None
    for rep in range(3):
        if rep == 2:
            lib.wt_debug_timeline(dbg.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _native.check(lib.wt_test_tap_gemm(0, A.data_ptr(), rows, Cin, taps, W.data_ptr(), N, b.data_ptr(), None, None,
                                           act, passes, None if split else out.data_ptr(),
                                           osplit.data_ptr() if split else None, None))
    lib.wt_debug_timeline(None)
    torch.cuda.synchronize()
    d = dbg.view(148, 64).cpu()
    c = d[0]
    nkb = taps * Cin // 64
    print(f"== {label} rows {rows} Cin {Cin} taps {taps} N {N} passes {passes}: CTA0 cycles")
    print("  prologue", int(c[0]), "| issue kb:", [int(x) for x in c[1:1 + min(nkb, 16)]])
    print("  landed kb:", [int(x) for x in c[17:17 + min(nkb, 16)]])
    print("  acc ready", int(c[40]), "epi done", int(c[41]), "end", int(c[42]))
    print("  tile4 detail: producer issue kb0/kb1", int(c[33]), int(c[34]), "| landed kb0/kb1", int(c[35]), int(c[36]),
          "| epi first chunk math done", int(c[38]))
    for ti in range(5):
        print(f"   tile {ti}: mma start {int(c[44 + 4 * ti])} issued {int(c[45 + 4 * ti])} acc ready {int(c[46 + 4 * ti])} "
              f"epi done {int(c[47 + 4 * ti])}")


run(128 * 148 * 8, 128, 1, 32, 3, label="K=128 N=32 split out", split=True)
run(128 * 148 * 8, 128, 1, 32, 3, label="K=128 N=32 fp32 out", split=False)
