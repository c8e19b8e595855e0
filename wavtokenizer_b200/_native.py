"""ctypes binding of the C ABI in include/wavtok_b200.h (csrc/libwavtok_b200.so).

There is no fallback: if the CUDA library is missing or fails to load, importing a symbol
from here raises, and so does every product entry point.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, Iterable, Optional

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libwavtok_b200.so")

WT_OK, WT_ERR_VALUE, WT_ERR_INDEX, WT_ERR_TYPE, WT_ERR_RUNTIME = 0, 1, 2, 3, 4
_ERR = {WT_ERR_VALUE: ValueError, WT_ERR_INDEX: IndexError, WT_ERR_TYPE: TypeError, WT_ERR_RUNTIME: RuntimeError}

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


class WtConfig(ctypes.Structure):
    _fields_ = [("strides", ctypes.c_int32 * 4), ("n_filters", ctypes.c_int32), ("dimension", ctypes.c_int32),
                ("lstm_layers", ctypes.c_int32), ("vq_bins", ctypes.c_int32), ("num_quantizers", ctypes.c_int32),
                ("dim", ctypes.c_int32), ("intermediate_dim", ctypes.c_int32), ("num_layers", ctypes.c_int32),
                ("adanorm_num_embeddings", ctypes.c_int32), ("n_fft", ctypes.c_int32), ("hop_length", ctypes.c_int32)]


class WtTensor(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("data", ctypes.c_void_p), ("numel", ctypes.c_int64)]


# every symbol include/wavtok_b200.h declares: name -> (restype, argtypes)
_P, _I32, _I64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
SYMBOLS = {
    "wt_create": (ctypes.c_int, [ctypes.POINTER(WtConfig), ctypes.POINTER(WtTensor), _I32, _I32, ctypes.POINTER(_P)]),
    "wt_destroy": (ctypes.c_int, [_P]),
    "wt_frames_for": (_I32, [_P, _I32]),
    "wt_workspace_bytes": (_I64, [_P, _I32, _I32]),
    "wt_reserve": (ctypes.c_int, [_P, _I32, _I32]),
    "wt_encode": (ctypes.c_int, [_P, _P, _I32, _I32, _P, _P, _P]),
    "wt_encode_ragged": (ctypes.c_int, [_P, _P, ctypes.POINTER(_I32), _I32, _P, _P, _P]),
    "wt_encoder_forward": (ctypes.c_int, [_P, _P, _I32, _I32, _P, _P]),
    "wt_seanet_decoder": (ctypes.c_int, [_P, _P, _I32, _I32, _P, _P]),
    "wt_codes_to_features": (ctypes.c_int, [_P, _P, _I32, _I32, _I32, _P, _P]),
    "wt_decode": (ctypes.c_int, [_P, _P, _I32, _I32, _I32, _P, _P]),
    "wt_decode_ragged": (ctypes.c_int, [_P, _P, ctypes.POINTER(_I32), _I32, _I32, _P, _P]),
    "wt_vq": (ctypes.c_int, [_P, _P, _I64, _P, _P, _P]),
    "wt_encode_decode_host": (ctypes.c_int, [_P, _P, _I32, _I32, _I32, _P, _P, _P]),
    "wt_tap_request": (ctypes.c_int, [_P, ctypes.c_char_p, _P, _I64]),
    "wt_tap_shape": (ctypes.c_int, [_P, ctypes.c_char_p, ctypes.POINTER(_I32), ctypes.POINTER(_I32),
                                    ctypes.POINTER(_I32)]),
    "wt_tap_clear": (ctypes.c_int, [_P]),
    "wt_launch_count": (_I64, [_P]),
    "wt_set_plan": (ctypes.c_int, [_P, _I32]),
    "wt_timing_enable": (ctypes.c_int, [_P, _I32]),
    "wt_timing_read": (ctypes.c_int, [_P, _I32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_I64)]),
    "wt_convert_audio_length": (_I64, [_I64, _I64, _I64]),
    "wt_convert_audio": (ctypes.c_int, [_I32, _P, _I64, _I32, _I64, _I64, _I64, _I32, _P, _P]),
    "wt_save_audio_pcm16": (ctypes.c_int, [_I32, _P, _I64, _I64, _I32, _P, _P, _P, _P]),
    "wt_timing_read_kernel": (ctypes.c_int, [_P, _I32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_I64),
                                             ctypes.POINTER(ctypes.c_double)]),
    "wt_timing_read_kernel_bytes": (ctypes.c_int, [_P, _I32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_I64),
                                                   ctypes.POINTER(ctypes.c_double)]),
    "wt_check_errors": (ctypes.c_int, [_P]),
    "wt_test_tap_gemm": (ctypes.c_int, [_I32, _P, _I32, _I32, _I32, _P, _I32, _P, _P, _P, _I32, _I32, _P, _P, _P]),
    "wt_debug_timeline": (ctypes.c_int, [_P]),
    "wt_last_error": (ctypes.c_char_p, []),
    "wt_version": (ctypes.c_char_p, []),
}

_lib: Optional[ctypes.CDLL] = None


def sources() -> Iterable[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into csrc/libwavtok_b200.so (nvcc cross-compiles without a GPU).

    Each translation unit is compiled to its own object (in parallel, rebuilt only when it or a header is newer),
    then the objects are linked into the shared library."""
    from concurrent.futures import ThreadPoolExecutor
    srcs = list(sources())
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "wavtok_b200.h"))
    hdr_m = max(os.path.getmtime(d) for d in hdrs)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f != "-shared"]
    if os.environ.get("WT_TIMELINE"):  # instrumented build for tools/gemm_timeline.py (per-CTA clock64 stamps)
        cflags.append("-DWT_TIMELINE=1")
    cflags += os.environ.get("WT_EXTRA_NVCC_FLAGS", "").split()  # experiments: compile-time tunables (-DWT_...=n)

    def compile_one(src: str):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_m):
            return obj, None
        cmd = [nvcc] + cflags + ["-c", "-o", obj, src]
        if verbose:
            print(" ".join(cmd))
        out = subprocess.run(cmd, capture_output=True, text=True)
        return obj, (out.stdout + out.stderr if out.returncode != 0 else None)

    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 1)) as ex:
        res = list(ex.map(compile_one, srcs))
    errs = [e for _, e in res if e]
    if errs:
        raise RuntimeError("nvcc failed:\n" + "\n".join(errs))
    objs = [o for o, _ in res]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(o) for o in objs):
        return LIB_PATH
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs + ["-lcuda"]
    if verbose:
        print(" ".join(cmd))
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + out.stdout + out.stderr)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """Load the CUDA library; raises if it is absent (no CPU / eager fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"wavtok_b200: CUDA library {LIB_PATH} is not built; run `python -c 'import __graft_entry__ as g; "
                "g.build()'` (there is no CPU fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status: int) -> None:
    if status != WT_OK:
        msg = lib().wt_last_error().decode("utf-8", "replace")
        raise _ERR.get(status, RuntimeError)(msg)


def make_config(cfg) -> WtConfig:
    c = WtConfig()
    for i, s in enumerate(cfg.strides):
        c.strides[i] = int(s)
    c.n_filters, c.dimension, c.lstm_layers = cfg.n_filters, cfg.dimension, cfg.lstm_layers
    c.vq_bins, c.num_quantizers = cfg.vq_bins, cfg.num_quantizers
    c.dim, c.intermediate_dim, c.num_layers = cfg.dim, cfg.intermediate_dim, cfg.num_layers
    c.adanorm_num_embeddings, c.n_fft, c.hop_length = cfg.adanorm_num_embeddings, cfg.n_fft, cfg.hop_length
    return c


class Handle:
    """Owns one wt_handle (prepared weights + workspace) on one CUDA device."""

    def __init__(self, cfg, state: Dict[str, "object"], device_index: int):
        import torch
        L = lib()
        keep = []
        arr = (WtTensor * len(state))()
        for i, (name, t) in enumerate(state.items()):
            h = t.detach().to(device="cpu", dtype=torch.float32).contiguous()
            keep.append(h)
            arr[i].name = name.encode()
            arr[i].data = h.data_ptr()
            arr[i].numel = h.numel()
        out = _P()
        c = make_config(cfg)
        check(L.wt_create(ctypes.byref(c), arr, len(state), int(device_index), ctypes.byref(out)))
        self._h = out
        self.device_index = int(device_index)

    @property
    def ptr(self):
        if self._h is None:
            raise RuntimeError("wt_handle already destroyed")
        return self._h

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            lib().wt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
