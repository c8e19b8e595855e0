"""Drop-in for the reference inference API (reference decoder/pretrained.py:32-239).

``WavTokenizer`` keeps the reference's constructor class-methods, method names, argument
meaning, tensor layouts, state-dict key names and error behaviour, and runs every method on
the hand-written sm_100a library through the C ABI (include/wavtok_b200.h). PyTorch is used
only for device memory, streams and the nn.Module parameter container. There is no eager
or CPU fallback: calling a compute method without a CUDA device or without the built
library raises.
"""
from __future__ import annotations

import ctypes
from typing import Any, Dict, List, Optional, Tuple

import torch
from torch import nn

from . import _native, ragged, spec
from .spec import ModelConfig


class _Node(nn.Module):
    """Plain container used to reproduce the reference's dotted state-dict names."""


class _Encoder(_Node):
    """``model.feature_extractor.encodec.encoder`` — callable on wav [B, 1, T] like the
    reference SEANetEncoder (reference encoder/modules/seanet.py:143-144; used directly by
    extract_features.py:46)."""

    def __init__(self, root: "WavTokenizer"):
        super().__init__()
        object.__setattr__(self, "_root", root)

    @torch.inference_mode()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, C, T = x.shape  # ValueError on a wrong rank, like conv.py:196
        if C != 1:
            raise RuntimeError(f"expected input[{B}, {C}, {T}] to have 1 channels")
        return self._root._encoder_forward(x.reshape(B, T))


class _SeanetDecoder(nn.Module):
    """``model.feature_extractor.encodec.decoder`` — callable on z [B, 512, L] like the reference SEANetDecoder
    (reference encoder/modules/seanet.py:189-238). Its weights are the
    ``feature_extractor.encodec.decoder.*`` tensors of a checkpoint; they are optional and do not appear in
    ``state_dict()`` (the hot path never reads them). Without them every call raises."""

    def __init__(self, root: "WavTokenizer"):
        super().__init__()
        object.__setattr__(self, "_root", root)

    @torch.inference_mode()
    def forward(self, z: torch.Tensor) -> torch.Tensor:
        root = self._root
        if not root._seanet_dec:
            raise RuntimeError("this model holds no feature_extractor.encodec.decoder weights (load a checkpoint that has them)")
        B, C, L = z.shape
        if C != root.cfg.dimension:
            raise RuntimeError(f"expected input[{B}, {C}, {L}] to have {root.cfg.dimension} channels")
        if z.dtype != torch.float32:
            raise RuntimeError(f"Input type ({z.dtype}) and weight type (torch.float32) should be the same")
        root._check_input(z, "z")
        z = z.contiguous()
        out = torch.empty(B, 1, L * root.cfg.hop_length, dtype=torch.float32, device=z.device)
        with torch.cuda.device(z.device):
            _native.check(_native.lib().wt_seanet_decoder(root.native().ptr, z.data_ptr(), B, L, out.data_ptr(), root._stream()))
        return out


class _VQLayer(_Node):
    @property
    def codebook(self) -> torch.Tensor:  # reference core_vq.py:286-288
        return self._codebook.embed


class WavTokenizer(nn.Module):
    """B200-native WavTokenizer (inference only)."""

    def __init__(self, cfg: ModelConfig, config_path: Optional[str] = None):
        super().__init__()
        self.cfg = cfg
        self.config_path = config_path
        self._handle: Optional[_native.Handle] = None
        self._kinds: Dict[str, str] = {}
        self._seanet_dec: Dict[str, torch.Tensor] = {}  # optional SEANet-decoder weights
        init = spec.synthetic_state_dict(cfg, seed=0)  # a fresh (random-init) model, like from_hparams0802
        for name, (shape, kind) in spec.state_spec(cfg).items():
            self._register(name, init[name], kind)
        # attributes the reference exposes and callers read (decoder/pretrained.py:230)
        q = self.feature_extractor.encodec.quantizer
        q.bins = cfg.vq_bins
        q.n_q = cfg.num_quantizers
        q.dimension = cfg.dimension
        self.feature_extractor.bandwidths = list(cfg.bandwidths)
        self.feature_extractor.frame_rate = 25  # "not use" (feature_extractors.py:68)
        self.feature_extractor.encodec.add_module("decoder", _SeanetDecoder(self))
        self.eval()

    # ------------------------------------------------------------------ module tree
    def _register(self, name: str, value: torch.Tensor, kind: str) -> None:
        parts = name.split(".")
        mod: nn.Module = self
        path: List[str] = []
        for p in parts[:-1]:
            path.append(p)
            if p not in mod._modules:
                joined = ".".join(path)
                if joined == "feature_extractor.encodec.encoder":
                    child: nn.Module = _Encoder(self)
                elif joined == "feature_extractor.encodec.quantizer.vq.layers":
                    child = nn.ModuleList()
                elif joined.startswith("feature_extractor.encodec.quantizer.vq.layers.") and len(path) == 6:
                    child = _VQLayer()
                else:
                    child = _Node()
                mod.add_module(p, child)
            mod = mod._modules[p]
        if kind == "param":
            mod.register_parameter(parts[-1], nn.Parameter(value.clone(), requires_grad=False))
        else:
            mod.register_buffer(parts[-1], value.clone())
        self._kinds[name] = kind

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_hparams0802(cls, config_path: str) -> "WavTokenizer":
        """Build from a reference-style YAML (reference decoder/pretrained.py:81-92)."""
        return cls(spec.load_config(config_path), config_path)

    @classmethod
    def from_pretrained0802(cls, config_path: str, model_path: str) -> "WavTokenizer":
        """Load a Lightning checkpoint (reference decoder/pretrained.py:95-114): keep the
        ``backbone.`` / ``head.`` / ``feature_extractor.`` keys of ``['state_dict']``."""
        model = cls.from_hparams0802(config_path)
        state_dict_raw = torch.load(model_path, map_location="cpu")["state_dict"]
        state_dict = {k: v for k, v in state_dict_raw.items()
                      if k.startswith(("backbone.", "head.", "feature_extractor."))}
        model.load_state_dict(state_dict)
        model.eval()
        return model

    @classmethod
    def from_pretrained0911(cls, config_path: str, model_folder_path: str) -> "WavTokenizer":
        """Average the three best checkpoints of a folder (reference decoder/pretrained.py:117-156): only files whose
        name starts with ``vocos_`` count, they are ranked by the validation loss embedded in the file name
        (characters ``[-11:-5]``, compared as strings), every file whose loss string is among the three smallest is
        loaded, and their ``backbone.`` / ``head.`` / ``feature_extractor.`` tensors are averaged."""
        import os
        model = cls.from_hparams0802(config_path)
        names = [f for f in os.listdir(model_folder_path) if f.startswith("vocos_")]
        if not names:
            raise FileNotFoundError(f"no 'vocos_*' checkpoint in {model_folder_path} (reference pretrained.py:126-129)")
        best = sorted(f[-11:-5] for f in names)[:3]
        dicts = []
        for f in names:
            if f[-11:-5] not in best:
                continue
            raw = torch.load(model_folder_path + "/" + f, map_location="cpu")["state_dict"]
            dicts.append({k: v for k, v in raw.items() if k.startswith(("backbone.", "head.", "feature_extractor."))})
        avg = {}
        for k in dicts[0]:
            acc = dicts[0][k].clone()  # same accumulation order as the reference (file order of os.listdir)
            for d in dicts[1:]:
                acc += d[k]
            avg[k] = acc / len(dicts)
        model.load_state_dict(avg)
        model.eval()
        return model

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Strict on the hot-path keys. The SEANet-decoder keys every reference checkpoint carries
        (feature_extractors.py:76-79) are optional: a COMPLETE set is kept for ``feature_extractor.encodec.decoder``
        (SURVEY.md section 8(f) row 4), anything else under that prefix is accepted and ignored as before."""
        filtered = {k: v for k, v in state_dict.items() if not k.startswith(spec.UNUSED_PREFIX)}
        want = spec.seanet_decoder_spec(self.cfg)
        have = {k: v for k, v in state_dict.items() if k in want}
        if len(have) == len(want) and all(tuple(have[k].shape) == tuple(s) for k, s in want.items()):
            self._seanet_dec = {k: have[k].detach().to(torch.float32).cpu().clone() for k in want}
        else:
            self._seanet_dec = {}
        out = super().load_state_dict(filtered, strict=strict, assign=assign)
        self._invalidate()
        return out

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self._invalidate()
        return out

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("wavtokenizer_b200 implements the inference path only (SURVEY.md section 8)")
        return super().train(False)

    def _invalidate(self) -> None:
        if getattr(self, "_handle", None) is not None:
            self._handle.close()
        self._handle = None
        self.__dict__.pop("_plan", None)          # a new handle starts on the library's default plan
        self.__dict__.pop("_ragged_lanes", None)  # replicas hold copies of the old weights

    # ------------------------------------------------------------------ native plumbing
    @property
    def device(self) -> torch.device:
        return self.head.istft.window.device

    def native(self) -> _native.Handle:
        """The wt_handle for the current weights/device (built on first use)."""
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("wavtokenizer_b200 has no CPU path: move the model to a CUDA device "
                               "(model.to('cuda')) before calling it")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if self._handle is None or self._handle.device_index != idx:
            self._invalidate()
            inited = self.feature_extractor.encodec.quantizer.vq.layers._modules["0"]._codebook.inited
            if float(inited.sum()) == 0.0:
                raise RuntimeError(
                    "codebook is not initialised (inited == 0): the reference would run k-means inside infer "
                    "(core_vq.py:140-151); load a checkpoint or install a codebook first")
            tensors = {k: v for k, v in self.state_dict().items()}
            tensors.update(self._seanet_dec)
            self._handle = _native.Handle(self.cfg, tensors, idx)
        return self._handle

    def _stream(self) -> ctypes.c_void_p:
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check_input(self, t: torch.Tensor, what: str) -> None:
        if t.device != self.device:
            raise RuntimeError(f"Expected all tensors to be on the same device, but {what} is on {t.device} "
                               f"and the model on {self.device}")

    def set_plan(self, plan: int) -> None:
        _native.check(_native.lib().wt_set_plan(self.native().ptr, int(plan)))
        self.__dict__["_plan"] = int(plan)
        for lane in self.__dict__.get("_ragged_lanes", []):
            lane[2].set_plan(plan)

    def _bandwidth_index(self, kwargs: Dict[str, Any], need_list_semantics: bool) -> int:
        if "bandwidth_id" not in kwargs or kwargs["bandwidth_id"] is None:
            if need_list_semantics:
                raise TypeError("infer() missing 1 required positional argument: 'bandwidth_id'")
            raise AssertionError("bandwidth_id is required (adanorm_num_embeddings is set)")  # models.py:227
        bw = kwargs["bandwidth_id"]
        if isinstance(bw, torch.Tensor):
            if bw.numel() != 1:
                if need_list_semantics:  # list[tensor] needs a single-element integer tensor
                    raise TypeError("only integer tensors of a single element can be converted to an index")
                raise RuntimeError("bandwidth_id must hold one id for the whole batch (decoder/modules.py:81-86)")
            if bw.dtype.is_floating_point or bw.dtype == torch.bool:
                raise TypeError("only integer tensors of a single element can be converted to an index")
            i = int(bw.reshape(-1)[0].item())
        else:
            i = int(bw)
        n = len(self.cfg.bandwidths) if need_list_semantics else self.cfg.adanorm_num_embeddings
        if need_list_semantics and -n <= i < 0:
            i += n  # Python list indexing accepts negatives
        if not 0 <= i < n:
            raise IndexError("list index out of range" if need_list_semantics else "index out of range in self")
        return i

    def _check_audio(self, audio_input: torch.Tensor) -> Tuple[int, int]:
        if audio_input.dim() != 2:
            raise ValueError(f"not enough values to unpack (expected audio [B, T], got {tuple(audio_input.shape)})"
                             if audio_input.dim() < 2 else
                             f"too many values to unpack (expected audio [B, T], got {tuple(audio_input.shape)})")
        if audio_input.dtype != torch.float32:
            raise RuntimeError(f"Input type ({audio_input.dtype}) and weight type (torch.float32) should be the same")
        self._check_input(audio_input, "audio_input")
        B, T = audio_input.shape
        if T < 1:
            raise RuntimeError("audio must hold at least one sample")
        return B, T

    # ------------------------------------------------------------------ public API
    @torch.inference_mode()
    def encode_infer(self, audio_input: torch.Tensor, **kwargs: Any) -> Tuple[torch.Tensor, torch.Tensor]:
        """audio [B, T] -> (features [B, 512, L] f32, codes [1, B, L] int64)
        (reference decoder/pretrained.py:186-189, feature_extractors.py:131-142)."""
        h = self.native()
        B, T = self._check_audio(audio_input)
        self._bandwidth_index(kwargs, need_list_semantics=True)  # looked up, then ignored (vq.py:126-137)
        wav = audio_input.contiguous()
        L = self.cfg.frames_for(T)
        feats = torch.empty(B, self.cfg.dimension, L, dtype=torch.float32, device=wav.device)
        codes = torch.empty(1, B, L, dtype=torch.int64, device=wav.device)
        with torch.cuda.device(wav.device):
            _native.check(_native.lib().wt_encode(h.ptr, wav.data_ptr(), B, T, feats.data_ptr(), codes.data_ptr(),
                                                  self._stream()))
        return feats, codes

    @torch.inference_mode()
    def encode(self, audio_input: torch.Tensor, **kwargs: Any) -> Tuple[torch.Tensor, torch.Tensor]:
        """Eval-mode ``feature_extractor.forward`` (reference decoder/pretrained.py:179-182). With one
        codebook it quantises exactly like ``infer`` (vq.py:84-113 picks n_q <= num_quantizers)."""
        if self.cfg.num_quantizers != 1:
            raise NotImplementedError("encode() with several codebooks is the training-time quantiser path")
        return self.encode_infer(audio_input, **kwargs)

    @torch.inference_mode()
    def _encoder_forward(self, wav: torch.Tensor) -> torch.Tensor:
        h = self.native()
        B, T = self._check_audio(wav)
        wav = wav.contiguous()
        z = torch.empty(B, self.cfg.dimension, self.cfg.frames_for(T), dtype=torch.float32, device=wav.device)
        with torch.cuda.device(wav.device):
            _native.check(_native.lib().wt_encoder_forward(h.ptr, wav.data_ptr(), B, T, z.data_ptr(), self._stream()))
        return z

    @torch.inference_mode()
    def codes_to_features(self, codes: torch.Tensor) -> torch.Tensor:
        """codes [K, L] or [K, B, L] -> features [B, 512, L] (reference decoder/pretrained.py:209-239)."""
        h = self.native()
        if codes.dim() == 2:
            codes = codes.unsqueeze(1)
        if codes.dim() != 3:
            raise RuntimeError(f"codes must be [K, L] or [K, B, L], got {tuple(codes.shape)}")
        if codes.dtype.is_floating_point or codes.dtype == torch.bool:
            raise RuntimeError("Expected tensor for argument #1 'indices' to have one of the following scalar "
                               f"types: Long, Int; but got {codes.dtype} instead")
        self._check_input(codes, "codes")
        K, B, L = codes.shape
        if K > self.cfg.num_quantizers:
            raise IndexError("index out of range in self")
        c64 = codes.to(torch.int64).contiguous()
        feats = torch.empty(B, self.cfg.dimension, L, dtype=torch.float32, device=codes.device)
        with torch.cuda.device(codes.device):
            _native.check(_native.lib().wt_codes_to_features(h.ptr, c64.data_ptr(), K, B, L, feats.data_ptr(),
                                                             self._stream()))
        return feats

    @torch.inference_mode()
    def decode(self, features_input: torch.Tensor, **kwargs: Any) -> torch.Tensor:
        """features [B, 512, L] -> audio [B, L*hop] (reference decoder/pretrained.py:192-207)."""
        h = self.native()
        bw = self._bandwidth_index(kwargs, need_list_semantics=False)
        if features_input.dim() != 3 or features_input.shape[1] != self.cfg.input_channels:
            raise RuntimeError(f"expected features [B, {self.cfg.input_channels}, L], got {tuple(features_input.shape)}")
        if features_input.dtype != torch.float32:
            raise RuntimeError(f"Input type ({features_input.dtype}) and weight type (torch.float32) should be the same")
        self._check_input(features_input, "features_input")
        feats = features_input.contiguous()
        B, _, L = feats.shape
        audio = torch.empty(B, L * self.cfg.hop_length, dtype=torch.float32, device=feats.device)
        with torch.cuda.device(feats.device):
            _native.check(_native.lib().wt_decode(h.ptr, feats.data_ptr(), B, L, bw, audio.data_ptr(), self._stream()))
        return audio

    @torch.inference_mode()
    def forward(self, audio_input: torch.Tensor, **kwargs: Any) -> torch.Tensor:
        """Copy-synthesis audio -> audio (reference decoder/pretrained.py:159-175)."""
        features, _ = self.encode(audio_input, **kwargs)
        return self.decode(features, **kwargs)

    # ------------------------------------------------------------------ beyond the reference API
    @torch.inference_mode()
    def vq(self, frames: torch.Tensor, return_quantized: bool = True):
        """EuclideanCodebook.quantize/dequantize on row-major frames [N, 512]
        (reference encoder/quantization/core_vq.py:175-190) — BASELINE.json's VQ-only sweep."""
        h = self.native()
        if frames.dim() != 2 or frames.shape[1] != self.cfg.dimension or frames.dtype != torch.float32:
            raise ValueError(f"expected float32 frames [N, {self.cfg.dimension}]")
        self._check_input(frames, "frames")
        x = frames.contiguous()
        N = x.shape[0]
        codes = torch.empty(N, dtype=torch.int64, device=x.device)
        quant = torch.empty_like(x) if return_quantized else None
        with torch.cuda.device(x.device):
            _native.check(_native.lib().wt_vq(h.ptr, x.data_ptr(), N, codes.data_ptr(),
                                              quant.data_ptr() if quant is not None else None, self._stream()))
        return codes, quant

    def _lanes(self, n: int):
        """``n`` (model, stream) lanes for ragged input: lane 0 is this model on the caller's stream, the others are
        replicas with their own native handle, workspace and stream (a handle serves one stream at a time, SURVEY.md
        8(b) threading contract)."""
        lanes = self.__dict__.setdefault("_ragged_lanes", [])
        dev = self.device
        if lanes and (lanes[0][0] != dev or lanes[0][1] != id(self._handle)):
            lanes.clear()  # device or weights changed: rebuild the replicas
        while len(lanes) < n - 1:
            m = WavTokenizer(self.cfg, self.config_path)
            m.load_state_dict(self.state_dict())
            m = m.to(dev)
            m.set_plan(self.plan())
            lanes.append((dev, id(self._handle), m, torch.cuda.Stream(device=dev)))
        models = [self] + [lane[2] for lane in lanes[:n - 1]]
        streams = [torch.cuda.current_stream(dev)] + [lane[3] for lane in lanes[:n - 1]]
        return models, streams

    def plan(self) -> int:
        return int(self.__dict__.get("_plan", 2))

    @torch.inference_mode()
    def encode_infer_ragged(self, clips, max_bucket: int = 0, streams: int = 1, batched: Optional[bool] = None,
                            **kwargs: Any):
        """Clips of DIFFERENT lengths ([T_i] or [1, T_i] float32) -> per clip, in input order, exactly what the
        reference's one-file-at-a-time loop returns (infer.py:44-54): ``(features [1, 512, L_i], codes [1, 1, L_i])``.
        Nothing is padded to a common length.

        Default (``batched`` None/True, every clip long enough for the tensor-core encoder layout): ONE C-ABI call
        (``wt_encode_ragged``): the conv front runs per run of equal-length clips and the LSTM - a latency chain whose
        cost does not depend on the batch - runs once for all clips, as do the last conv and the VQ.
        ``batched=False`` (or very short clips): equal-length clips are stacked and share one ``wt_encode`` call per
        length bucket (``ragged.length_buckets``); ``streams`` > 1 then runs the buckets on that many CUDA streams (one
        model replica each) so that small buckets, which cannot fill the GPU alone, overlap."""
        flat = []
        for i, c in enumerate(clips):
            if c.dim() == 2 and c.shape[0] == 1:
                c = c[0]
            if c.dim() != 1:
                raise ValueError(f"clip {i}: expected [T] or [1, T], got {tuple(c.shape)}")
            flat.append(c)
        if batched is None:
            batched = streams <= 1
        if batched and len(flat) > 1 and self.plan() >= 1 and all(self._ragged_ok(int(c.numel())) for c in flat):
            return self._encode_ragged_native(flat, **kwargs)

        def make(model):
            def fn(batch):
                feats, codes = model.encode_infer(batch, **kwargs)
                return (feats, 0), (codes, 1)
            return fn
        if streams > 1 and len(flat) > 1:
            self.native()
            models, sts = self._lanes(int(streams))
            out = ragged.run_bucketed(flat, [make(m) for m in models], max_bucket, sts)
        else:
            out = ragged.run_bucketed(flat, make(self), max_bucket)
        return [(f.unsqueeze(0), c.unsqueeze(1)) for f, c in out]

    def _ragged_ok(self, T: int) -> bool:
        """Same rule as the library's tensor-core encoder layout: every level's reflect halo stays inside the clip."""
        Tc = T
        for s in self.cfg.strides:
            Tn = -(-Tc // s)
            hr = s // 2 + (Tn * s - Tc)
            if Tc < 4 or Tc <= hr + 1 or Tc <= s:
                return False
            Tc = Tn
        return Tc >= 4

    def _encode_ragged_native(self, flat, **kwargs: Any):
        h = self.native()
        self._bandwidth_index(kwargs, need_list_semantics=True)  # looked up, then ignored (vq.py:126-137)
        for c in flat:
            if c.dtype != torch.float32:
                raise RuntimeError(f"Input type ({c.dtype}) and weight type (torch.float32) should be the same")
            self._check_input(c, "audio_input")
        # clips of equal length next to each other (one conv-front pass per run), longest first
        order = sorted(range(len(flat)), key=lambda i: (-int(flat[i].numel()), i))
        lens = [int(flat[i].numel()) for i in order]
        wav = torch.cat([flat[i] for i in order]).contiguous()
        Ls = [self.cfg.frames_for(n) for n in lens]
        D = self.cfg.dimension
        feats = torch.empty(sum(Ls) * D, dtype=torch.float32, device=wav.device)
        codes = torch.empty(sum(Ls), dtype=torch.int64, device=wav.device)
        arr = (ctypes.c_int32 * len(lens))(*lens)
        with torch.cuda.device(wav.device):
            _native.check(_native.lib().wt_encode_ragged(h.ptr, wav.data_ptr(), arr, len(lens), feats.data_ptr(),
                                                         codes.data_ptr(), self._stream()))
        out: List[Any] = [None] * len(flat)
        off = 0
        for i, L in zip(order, Ls):
            out[i] = (feats[off * D:(off + L) * D].view(1, D, L), codes[off:off + L].view(1, 1, L))
            off += L
        return out

    @torch.inference_mode()
    def decode_ragged(self, features, max_bucket: int = 0, streams: int = 1, batched: Optional[bool] = None,
                      **kwargs: Any):
        """Features of different lengths ([512, L_i] or [1, 512, L_i]) -> [audio [1, L_i * hop]] in input order,
        each equal to a batch-of-one ``decode`` (reference decoder/pretrained.py:192-207).

        Default (``batched`` None/True): ONE C-ABI call (``wt_decode_ragged``): all clips share one padded row space
        whose GEMMs run at batch size, while GroupNorm, attention, the depthwise convs and the overlap-add read every
        clip's own length. ``batched=False``: one ``wt_decode`` per length bucket, optionally on ``streams`` streams."""
        flat = []
        for i, f in enumerate(features):
            if f.dim() == 3 and f.shape[0] == 1:
                f = f[0]
            if f.dim() != 2:
                raise ValueError(f"features {i}: expected [C, L] or [1, C, L], got {tuple(f.shape)}")
            flat.append(f)
        if batched is None:
            batched = streams <= 1
        if batched and len(flat) > 1 and self.plan() >= 1:
            return self._decode_ragged_native(flat, **kwargs)

        def make(model):
            return lambda batch: ((model.decode(batch, **kwargs), 0),)
        if streams > 1 and len(flat) > 1:
            self.native()
            models, sts = self._lanes(int(streams))
            out = ragged.run_bucketed(flat, [make(m) for m in models], max_bucket, sts)
        else:
            out = ragged.run_bucketed(flat, make(self), max_bucket)
        return [a.unsqueeze(0) for (a,) in out]

    def _decode_ragged_native(self, flat, **kwargs: Any):
        h = self.native()
        bw = self._bandwidth_index(kwargs, need_list_semantics=False)
        C = self.cfg.input_channels
        for f in flat:
            if f.shape[0] != C:
                raise RuntimeError(f"expected features [{C}, L], got {tuple(f.shape)}")
            if f.dtype != torch.float32:
                raise RuntimeError(f"Input type ({f.dtype}) and weight type (torch.float32) should be the same")
            self._check_input(f, "features_input")
        order = sorted(range(len(flat)), key=lambda i: (-int(flat[i].shape[1]), i))  # similar lengths share a chunk
        Ls = [int(flat[i].shape[1]) for i in order]
        packed = torch.cat([flat[i].contiguous().reshape(-1) for i in order])
        hop = self.cfg.hop_length
        audio = torch.empty(sum(Ls) * hop, dtype=torch.float32, device=packed.device)
        arr = (ctypes.c_int32 * len(Ls))(*Ls)
        with torch.cuda.device(packed.device):
            _native.check(_native.lib().wt_decode_ragged(h.ptr, packed.data_ptr(), arr, len(Ls), bw, audio.data_ptr(),
                                                         self._stream()))
        out: List[Any] = [None] * len(flat)
        off = 0
        for i, L in zip(order, Ls):
            out[i] = audio[off * hop:(off + L) * hop].view(1, L * hop)
            off += L
        return out

    def encode_decode_host(self, wav_host: torch.Tensor, bandwidth_id: int = 0):
        """Whole hot path on HOST tensors (pinned recommended): H2D, encode, decode, D2H, sync.
        The returned tensors are reused by the next call with the same shapes (clone to keep them)."""
        h = self.native()
        if wav_host.device.type != "cpu" or wav_host.dtype != torch.float32 or wav_host.dim() != 2:
            raise ValueError("expected a float32 CPU tensor [B, T]")
        wav_host = wav_host.contiguous()
        B, T = wav_host.shape
        L = self.cfg.frames_for(T)
        pin = wav_host.is_pinned()
        # pinned result buffers are cached per shape: cudaHostAlloc of ~75 MB costs more than the whole step
        key = (B, L, pin)
        cache = self.__dict__.setdefault("_host_out", {})
        if key not in cache:
            cache.clear()
            cache[key] = (torch.empty(1, B, L, dtype=torch.int64, pin_memory=pin),
                          torch.empty(B, L * self.cfg.hop_length, dtype=torch.float32, pin_memory=pin))
        codes, audio = cache[key]
        with torch.cuda.device(self.device):
            _native.check(_native.lib().wt_encode_decode_host(h.ptr, wav_host.data_ptr(), B, T, int(bandwidth_id),
                                                              codes.data_ptr(), audio.data_ptr(), self._stream()))
        return codes, audio

    def check_errors(self) -> None:
        """Wait for pending device-side checks and raise what they found: ``IndexError`` for a code >= vq_bins given to
        ``codes_to_features`` (reference decoder/pretrained.py:236 -> embedding lookup). That call does not
        synchronise the stream to find out; without this method the error is raised by a later call on the model."""
        if self._handle is not None:
            with torch.cuda.device(self.device):
                _native.check(_native.lib().wt_check_errors(self._handle.ptr))

    def reserve(self, B: int, T: int) -> None:
        _native.check(_native.lib().wt_reserve(self.native().ptr, int(B), int(T)))

    def launch_count(self) -> int:
        return int(_native.lib().wt_launch_count(self.native().ptr))
