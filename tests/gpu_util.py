"""Helpers for the -m gpu parity tests: build the native model and read debug taps."""
from __future__ import annotations

import ctypes
from typing import Dict, Iterable

import torch

from wavtokenizer_b200 import WavTokenizer, _native
from tests import helpers


def native_model(tag: str, plan: int = 0) -> WavTokenizer:
    cfg, sd = helpers.model(tag)
    m = WavTokenizer(cfg)
    m.load_state_dict(sd)
    m = m.to("cuda:0")
    m.set_plan(plan)  # the library default is plan 2
    return m


class Taps:
    """Request named stage outputs (channels-last [B, T, C]) from the next encode/decode."""

    def __init__(self, model: WavTokenizer, names: Iterable[str], capacity: int = 8 << 20):
        self.model = model
        self.lib = _native.lib()
        self.h = model.native().ptr
        self.bufs: Dict[str, torch.Tensor] = {}
        for n in names:
            buf = torch.zeros(capacity, dtype=torch.float32, device=model.device)
            self.bufs[n] = buf
            _native.check(self.lib.wt_tap_request(self.h, n.encode(), buf.data_ptr(), capacity))

    def get(self, name: str) -> torch.Tensor:
        """Stage output as [B, C, T] (reference layout) on the CPU."""
        B, T, C = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        _native.check(self.lib.wt_tap_shape(self.h, name.encode(), ctypes.byref(B), ctypes.byref(T), ctypes.byref(C)))
        n = B.value * T.value * C.value
        if n == 0:
            raise KeyError(f"stage {name} did not run")
        torch.cuda.synchronize()
        return self.bufs[name][:n].view(B.value, T.value, C.value).permute(0, 2, 1).contiguous().cpu()

    def close(self) -> None:
        _native.check(self.lib.wt_tap_clear(self.h))


def golden_sub(t: torch.Tensor, n: int = 384) -> torch.Tensor:
    """Same deterministic strided subsample oracle/make_golden.py applied to the reference taps."""
    f = t.reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].to(torch.float32)
