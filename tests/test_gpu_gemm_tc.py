"""-m gpu: the tcgen05 tap-GEMM alone (C-ABI test hook) against an fp64 torch contraction."""
import pytest
import torch

from tests import helpers
from wavtokenizer_b200 import _native

pytestmark = pytest.mark.gpu


def run(rows, Cin, taps, N, passes, act=0, bias=True, gamma=False, res=False, split=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(rows, Cin, generator=g) * 0.5
    W = torch.randn(N, taps * Cin, generator=g) * 0.05
    b = torch.randn(N, generator=g) * 0.1 if bias else None
    gm = torch.randn(N, generator=g) * 0.1 + 0.1 if gamma else None
    r = torch.randn(rows, N, generator=g) if res else None
    dev = "cuda:0"
    out = torch.full((rows, N), float("nan"), device=dev)
    out_split = torch.zeros(rows, N, device=dev) if split else None
    p = lambda t: t.to(dev).contiguous().data_ptr() if t is not None else None
    keep = [t.to(dev).contiguous() if t is not None else None for t in (A, W, b, gm, r)]
    ptr = [t.data_ptr() if t is not None else None for t in keep]
    _native.check(_native.lib().wt_test_tap_gemm(0, ptr[0], rows, Cin, taps, ptr[1], N, ptr[2], ptr[3], ptr[4], act,
                                                 passes, out.data_ptr(), out_split.data_ptr() if split else None,
                                                 None))
    # fp64 reference of the same tap contraction (zero rows outside [0, rows))
    c = (taps - 1) // 2
    Ad = torch.zeros(rows + taps - 1, Cin, dtype=torch.float64)
    Ad[c:c + rows] = A.double()
    ref = torch.zeros(rows, N, dtype=torch.float64)
    for j in range(taps):
        ref += Ad[j:j + rows] @ W.double()[:, j * Cin:(j + 1) * Cin].t()
    if b is not None:
        ref += b.double()
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    if gm is not None:
        ref *= gm.double()
    if r is not None:
        ref += r.double()
    return ref, out.cpu(), (out_split.cpu() if split else None)


@pytest.mark.parametrize("rows,Cin,taps,N", [(128, 64, 1, 128), (300, 768, 1, 768), (1000, 768, 1, 2304),
                                             (517, 512, 7, 768), (260, 768, 3, 768), (130, 768, 1, 1284),
                                             (257, 1344, 1, 1280)])
def test_three_pass_matches_fp64(rows, Cin, taps, N):
    ref, out, _ = run(rows, Cin, taps, N, passes=3)
    assert bool(torch.isfinite(out).all())
    assert helpers.snr_db(ref, out) > 90, helpers.snr_db(ref, out)


def test_single_pass_is_fp16_accurate():
    ref, out, _ = run(700, 768, 1, 2304, passes=1)
    snr = helpers.snr_db(ref, out)
    assert 55 < snr < 90, snr


def test_epilogue_variants():
    ref, out, split = run(515, 768, 1, 2304, passes=3, act=1, split=True)
    assert helpers.snr_db(ref, out) > 90
    assert helpers.snr_db(ref, split) > 90          # hi + lo planes carry ~22 bits
    ref, out, _ = run(515, 2304, 1, 768, passes=3, gamma=True, res=True)
    assert helpers.snr_db(ref, out) > 90
    ref, out, _ = run(40000, 768, 1, 768, passes=3, res=True, seed=3)   # > 148 tiles: persistent loop + TMEM ping-pong
    assert helpers.snr_db(ref, out) > 90


@pytest.mark.parametrize("N,Cin,taps", [(16, 64, 1), (32, 128, 1), (64, 128, 3), (128, 64, 1)])
def test_narrow_tiles_epilogue_groups(N, Cin, taps):
    """Narrow accumulators: four TMEM stages, four epilogue groups; > 4 x 148 tiles wraps every ring several times."""
    rows = 128 * 148 * 5 + 77
    ref, out, split = run(rows, Cin, taps, N, passes=3, split=True, seed=N)
    assert bool(torch.isfinite(out).all())
    assert helpers.snr_db(ref, out) > 90
    assert helpers.snr_db(ref, split) > 90
