"""Benchmark of the WavTokenizer hot path (encode_infer -> VQ -> codes_to_features -> decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--plan P]

One "step" = one pass of the hot path over one batch of synthetic 24 kHz clips. The workload is
BASELINE.json configs[1]: WavTokenizer-small-320 (frame75), 256 x 3 s clips per GPU (weak scaling:
each rank owns its own 256 clips; the only collective is the all-gather of codes). Prints ONE JSON line.

  value        audio-seconds per second, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e          same metric through the host-buffer C-ABI entry (pinned host wav in, codes+audio out)
  roofline     dominant contraction category, timed live with CUDA events on its stream
  cpu_baseline the CPU oracle port (oracle/) on a bounded sample of the same workload, host cores
  parity_sample  code match % / waveform SNR of the measured path against the oracle on that CPU sample
  other_configs  BASELINE.json's VQ-only sweep (1e6 frames) and decode-only detokenisation (256 x 10 s streams)
  next_rows      SURVEY.md 8(f) row 1: convert_audio and the save_audio limiter + PCM16 back-end, HBM roofline each

`--impl reference` times the reference algorithm's CPU port (oracle/, torch CPU ops, all host threads)
on a bounded sample of the same workload: /root/reference is pure Python and does not exist on the GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from wavtokenizer_b200 import spec  # noqa: E402

CONFIG = "wavtokenizer_smalldata_frame75_3s_nq1_code4096_dim512_kmeans200_attn.yaml"
WORKLOAD = "WavTokenizer-small-320-24k-4096 encode_infer+codes_to_features+decode, 256 x 3 s clips per GPU"
CLIPS_PER_GPU = 256
T = 72000
SR = 24000
METRIC = "encode+decode audio-sec/sec (24 kHz)"
UNIT = "audio-s/s"
CPU_SAMPLE_CLIPS = 4
# algorithmic FLOPs per frame of the ConvNeXt pointwise GEMMs: 12 x 2 x (2*768*2304)
# dram__bytes_read.sum + dram__bytes_write.sum per launch, one `ncu --set full` capture of the kernel's modal launch shape
# in this workload (decoder k3 conv 768 -> 768 over 128 clips; ConvNeXt GEMM-1): profiles/r01_final_summary.md
DOMINANT_TRAFFIC = {"tap_gemm_tc_kernel<256, 3>": 249.5e6, "tap_gemm_tc_kernel<256, 1>": 135.0e6}
CATS = ["enc_conv", "lstm", "vq", "dec_conv", "pwconv", "head_idft", "attention", "memory_bound"]


def cfg_path() -> str:
    return os.path.join(ROOT, "wavtokenizer_b200", "configs", CONFIG)


def algorithmic_flops(cfg, L: int) -> dict:
    """Per 3 s clip, per category (2*MACs; SURVEY.md Appendix A formulas)."""
    D, H = cfg.dim, cfg.intermediate_dim
    enc = 0.0
    Tc, C = T, cfg.n_filters
    enc += 2.0 * Tc * C * 7
    for s in cfg.strides:
        enc += 2.0 * Tc * (3 * C * (C // 2) + (C // 2) * C + C * C)
        Tn = -(-Tc // s)
        enc += 2.0 * Tn * (2 * s * C) * (2 * C)
        Tc, C = Tn, 2 * C
    enc += 2.0 * L * 7 * C * cfg.dimension
    lstm = cfg.lstm_layers * L * 16.0 * C * C
    vq = 2.0 * L * cfg.vq_bins * cfg.dimension
    dec_conv = 2.0 * L * (7 * cfg.dimension * D + 8 * 3 * D * D + 4 * D * D)
    pw = cfg.num_layers * 2.0 * L * 2 * D * H
    head = 2.0 * L * D * (cfg.n_fft + 2)
    attn = 4.0 * L * L * D
    return {"enc_conv": enc, "lstm": lstm, "vq": vq, "dec_conv": dec_conv, "pwconv": pw, "head_idft": head,
            "attention": attn, "memory_bound": 0.0}


def make_state(cfg, seed: int, encode_frames):
    """Random-init weights + a codebook of sampled encoder frames."""
    sd = spec.synthetic_state_dict(cfg, seed)
    spec.install_codebook(sd, torch.zeros(cfg.vq_bins, cfg.dimension))  # placeholder so the encoder can run
    cal = spec.synthetic_audio(4, T, seed=7)
    z = encode_frames(sd, cal)  # [4, 512, L]
    frames = z.permute(0, 2, 1).reshape(-1, cfg.dimension).float().cpu()
    g = torch.Generator().manual_seed(5)
    base = frames[torch.randperm(frames.shape[0], generator=g)[:512]].to(torch.bfloat16)
    spec.install_codebook(sd, spec.expand_codebook(base, cfg.vq_bins, seed=5))
    return sd


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md clocks line). NVML (nvidia_ml_py) is
    polled every 10 ms from a thread; if NVML cannot be loaded the same fields are read through nvidia-smi."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.rows, self._stop = index, [], threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            phys = int(ids[index]) if index < len(ids) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        bits = [getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8), getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)]
        return [str(sm), str(self.max_sm)] + ["Active" if r & b else "Not Active" for b in bits]

    def _run(self):
        while not self._stop.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                    parts = [p.strip() for p in out.strip().split(",")]
                    if len(parts) == 6:
                        self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.01 if self.nvml is not None else 0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.thread.join(timeout=6)

    def summary(self) -> dict:
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def oracle_step(sd, cfg, wav, bw):
    from oracle import wavtok_oracle as O  # CPU baseline leg only
    with torch.inference_mode():
        feats, codes = O.encode_infer(sd, cfg, wav, library_lstm=True)
        audio = O.decode(sd, cfg, O.codes_to_features(sd, cfg, codes), bw)
    return codes, audio


def time_cpu(sd, cfg, clips: int, steps: int, warmup: int) -> dict:
    torch.set_num_threads(os.cpu_count() or 1)
    wav = spec.synthetic_audio(clips, T, seed=100)
    bw = torch.tensor([0])
    for _ in range(warmup):
        oracle_step(sd, cfg, wav, bw)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        oracle_step(sd, cfg, wav, bw)
        ts.append(time.perf_counter() - t0)
    sec = sum(ts) / len(ts)
    return {"value": clips * T / SR / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{clips} x 3 s clips per step of the same workload, {steps} steps after {warmup} warm-up, "
                      f"oracle/wavtok_oracle.py (torch CPU fp32 ops, nn.LSTM), {sec * 1e3:.0f} ms/step",
            "ms_per_step": sec * 1e3}


def run_reference(args, rank: int) -> None:
    if rank != 0:
        return
    from oracle import wavtok_oracle as O
    cfg = spec.load_config(cfg_path())
    torch.set_num_threads(os.cpu_count() or 1)

    def enc(sd, cal):
        with torch.inference_mode():
            return O.seanet_encoder(sd, cfg, cal.unsqueeze(1), library_lstm=True)
    sd = make_state(cfg, 1, enc)
    r = time_cpu(sd, cfg, CPU_SAMPLE_CLIPS, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample_clips_per_step": CPU_SAMPLE_CLIPS, "device": "host CPU"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_native(args, rank: int, local_rank: int, world: int) -> None:
    import torch.distributed as dist
    from wavtokenizer_b200 import WavTokenizer, _native
    from wavtokenizer_b200.shard import gather_codes

    assert torch.cuda.is_available(), "bench.py (native arm) needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout (one JSON line only)
        dist.init_process_group("nccl", device_id=dev)
    cfg = spec.load_config(cfg_path())
    L = cfg.frames_for(T)

    model = WavTokenizer(cfg)

    def enc(sd, cal):
        model.load_state_dict(sd)
        m = model.to(dev)
        z = m._encoder_forward(cal.to(dev))
        torch.cuda.synchronize()
        return z
    sd = make_state(cfg, 1, enc)
    model.load_state_dict(sd)
    model = model.to(dev)
    model.set_plan(args.plan)
    B = CLIPS_PER_GPU
    model.reserve(B, T)
    lib, hptr = _native.lib(), model.native().ptr

    wav_host = spec.synthetic_audio(B, T, seed=1000 + rank).pin_memory()
    wav = wav_host.to(dev)
    bw = torch.tensor([0], device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step():
        feats, codes = model.encode_infer(wav, bandwidth_id=bw)
        audio = model.decode(model.codes_to_features(codes), bandwidth_id=bw)
        allc = gather_codes(codes, B * world) if world > 1 else codes
        return allc, audio

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = lib.wt_launch_count(hptr)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:
        torch.cuda.synchronize()
        for a, b in evs:
            flush.zero_()  # evict L2 between timed iterations (outside the event pair)
            a.record()
            step()
            b.record()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = (lib.wt_launch_count(hptr) - launches0) // args.steps
    ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    audio_s = B * world * T / SR
    value = audio_s / (ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI entry (pinned host in / out) ----
    for _ in range(2):
        model.encode_decode_host(wav_host, 0)
    e2e_t = []
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        codes_h, audio_h = model.encode_decode_host(wav_host, 0)
        e2e_t.append(time.perf_counter() - t0)
    e2e_ms = 1e3 * sum(e2e_t) / len(e2e_t)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e = {"value": audio_s / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * T * 4,
           "d2h_bytes_per_step": B * L * 8 + B * L * cfg.hop_length * 4, "ms_per_step": e2e_ms,
           "api": "wt_encode_decode_host (WavTokenizer.encode_decode_host), pinned host buffers"}

    # ---- per-category breakdown, timed live with CUDA events on the launching stream ----
    import ctypes
    _native.check(lib.wt_timing_enable(hptr, 1))
    step()
    torch.cuda.synchronize()
    flops = algorithmic_flops(cfg, L)
    breakdown = {}
    for i, name in enumerate(CATS):
        t_ms, n = ctypes.c_double(), ctypes.c_int64()
        _native.check(lib.wt_timing_read(hptr, i, ctypes.byref(t_ms), ctypes.byref(n)))
        breakdown[name] = {"ms": round(t_ms.value, 3), "launches": n.value,
                           "algorithmic_tflops": round(flops[name] * B / 1e12, 4),
                           "tflops_per_s": round(flops[name] * B / max(t_ms.value, 1e-9) / 1e9, 2)}
    breakdown["lstm"]["note"] = ("the two LSTM layers run as a wavefront on two streams: ms sums intervals that overlap "
                                 "in time (wall share of the step is about half of it)")
    # ---- same record per tcgen05 GEMM kernel variant (kern = BN * 10 + passes) ----
    kernels = {}
    for bn in (16, 32, 64, 128, 256):
        for passes in (3, 1):
            t_ms, n, fl = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double()
            _native.check(lib.wt_timing_read_kernel(hptr, bn * 10 + passes, ctypes.byref(t_ms), ctypes.byref(n),
                                                    ctypes.byref(fl)))
            if n.value:
                kernels[f"tap_gemm_tc_kernel<{bn}, {passes}>"] = {
                    "ms": round(t_ms.value, 3), "launches": n.value, "passes": passes,
                    "algorithmic_tflops": round(fl.value / 1e12, 4),
                    "tflops_per_s": round(fl.value / max(t_ms.value, 1e-9) / 1e9, 2)}
    _native.check(lib.wt_timing_enable(hptr, 0))
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback (B200_PROFILING.md sustained)"
    # dominant kernel = the kernel (by name) with the largest share of the step
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    d = kernels[dom]
    achieved = d["tflops_per_s"]
    roofline = {"bound": "tensor", "kernel": dom, "achieved": round(achieved, 2), "peak": peak_tf, "unit": "TFLOP/s",
                "frac": round(achieved / peak_tf, 4), "traffic": DOMINANT_TRAFFIC.get(dom), "peak_source": peak_src,
                "launches_per_step": d["launches"], "avg_launch_ms": round(d["ms"] / max(d["launches"], 1), 4),
                "share_of_step": round(d["ms"] / ms, 3),
                "executed_frac": round(achieved * d["passes"] / peak_tf, 4),
                "note": "achieved = algorithmic FLOPs (2*M*N*K per launch, split-precision passes NOT counted) of all "
                        "launches of this kernel in one step / their summed CUDA-event time on the launching stream; "
                        "executed_frac counts the 3 split-fp16 MMA passes the parity bar requires (SURVEY.md App. D); "
                        "traffic = dram read+write bytes per launch of the modal launch shape, ncu --set full "
                        "(profiles/r01_final_summary.md)"}

    # ---- the convert_audio front-end (44.1 kHz stereo -> 24 kHz mono, 171 taps) ----
    from wavtokenizer_b200 import convert_audio
    sr_in = 44100
    t_in = T * sr_in // SR
    raw = torch.randn(B, 2, t_in, device=dev).clamp_(-1, 1)
    for _ in range(3):
        mono = convert_audio(raw, sr_in, SR, 1)
    ca = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ca:
        flush.zero_()
        a.record()
        mono = convert_audio(raw, sr_in, SR, 1)
        b.record()
    torch.cuda.synchronize()
    ca_ms = sum(a.elapsed_time(b) for a, b in ca) / args.steps
    ca_bytes = raw.numel() * 4 + mono.numel() * 4
    hbm_peak = float(peaks.get("hbm_gbs", 6500.0))
    next_rows = {"convert_audio": {
        "workload": f"{B} clips x 3 s, {sr_in} Hz stereo -> {SR} Hz mono (polyphase sinc, 171 taps x 80 phases)",
        "ms": round(ca_ms, 3), "audio_s_per_s": round(B * T / SR / (ca_ms * 1e-3), 1),
        "roofline": {"bound": "hbm", "achieved": round(ca_bytes / (ca_ms * 1e-3) / 1e9, 1), "peak": hbm_peak,
                     "unit": "GB/s", "frac": round(ca_bytes / (ca_ms * 1e-3) / 1e9 / hbm_peak, 4),
                     "algorithmic_bytes": ca_bytes,
                     "note": "compulsory bytes (stereo in + mono out) / CUDA-event time; this ratio is 171 FMA per output "
                             "sample, so the kernel is fp32-FMA bound before it is HBM bound: "
                             f"{round(B * T * 171 / (ca_ms * 1e-3) / 1e12, 2)} TFMA/s achieved"}}}
    del raw, mono

    def timed(fn, n):
        for _ in range(2):
            fn()
        evp = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b in evp:
            flush.zero_()
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evp) / n

    # ---- BASELINE.json configs[4]: VQ-only sweep (fused distance GEMM + argmin), 1e6 frames per GPU ----
    n_vq = 1_000_000
    cb = sd[spec.CODEBOOK_PREFIX + "0._codebook.embed"].to(dev)
    g = torch.Generator(device=dev).manual_seed(11 + rank)
    xs = cb[torch.randint(0, cb.shape[0], (n_vq,), device=dev, generator=g)]
    xs = xs + 2e-3 * torch.randn(n_vq, cfg.dimension, device=dev, generator=g)
    vq_ms = timed(lambda: model.vq(xs, return_quantized=False), args.steps)
    vq_fl = 2.0 * n_vq * cfg.vq_bins * cfg.dimension
    other = {"vq_only": {
        "workload": f"{n_vq} frames x {cfg.dimension} fp32 vs {cfg.vq_bins} x {cfg.dimension} codebook per GPU, codes only",
        "ms": round(vq_ms, 3), "frames_per_s": round(n_vq * world / (vq_ms * 1e-3), 1),
        "roofline": {"bound": "tensor", "achieved": round(vq_fl / (vq_ms * 1e-3) / 1e12, 2), "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": round(vq_fl / (vq_ms * 1e-3) / 1e12 / peak_tf, 4),
                     "executed_frac": round(3 * vq_fl / (vq_ms * 1e-3) / 1e12 / peak_tf, 4),
                     "note": "4.194 algorithmic MFLOP per frame (SURVEY.md 8(d)); includes the fp32 -> split-fp16 "
                             "plane conversion of the frames (2048 B read + 2048 B written per frame)"}}}
    del xs
    # ---- BASELINE.json configs[3]: decode-only detokenisation, 10 s random token streams (bounded: 256 per GPU) ----
    n_st, l_st = 256, 750
    rc = torch.randint(0, cfg.vq_bins, (1, n_st, l_st), device=dev, generator=g)
    dec_ms = timed(lambda: model.decode(model.codes_to_features(rc), bandwidth_id=bw), args.steps)
    other["decode_only"] = {
        "workload": f"codes_to_features + decode of {n_st} random token streams x 10 s (L = {l_st}) per GPU",
        "ms": round(dec_ms, 3), "audio_s_per_s": round(n_st * world * l_st * cfg.hop_length / SR / (dec_ms * 1e-3), 1),
        "tokens_per_s": round(n_st * world * l_st / (dec_ms * 1e-3), 1)}
    del rc
    # ---- SURVEY.md 8(f) row 1, second half: save_audio limiter + PCM_S 16 of the decoded batch ----
    from wavtokenizer_b200 import pcm16
    dec_audio = torch.randn(B, T, device=dev, generator=g) * 0.5
    for mode in ("clamp", "rescale"):
        p_ms = timed(lambda: pcm16(dec_audio, mode), args.steps)
        p_bytes = dec_audio.numel() * (6 if mode == "clamp" else 10)
        next_rows[f"save_audio_pcm16_{mode}"] = {
            "workload": f"{B} clips x 3 s fp32 -> int16, limiter '{mode}' (encoder/utils.py:95-103)",
            "ms": round(p_ms, 4), "audio_s_per_s": round(B * T / SR / (p_ms * 1e-3), 1),
            "roofline": {"bound": "hbm", "achieved": round(p_bytes / (p_ms * 1e-3) / 1e9, 1), "peak": hbm_peak,
                         "unit": "GB/s", "frac": round(p_bytes / (p_ms * 1e-3) / 1e9 / hbm_peak, 4),
                         "algorithmic_bytes": p_bytes,
                         "note": "4 B read + 2 B written per sample" + ("" if mode == "clamp" else
                                 " + 4 B for the peak pass; two launches + a memset")}}
    del dec_audio
    # ---- ragged input, 64 clips of 64 DISTINCT lengths (1 - 5 s), one bucket per clip ----
    n_rag = 64
    rag = [torch.randn(SR + i * (4 * SR // n_rag) + 7 * i, device=dev, generator=g).clamp_(-1, 1) for i in range(n_rag)]
    rag_s = sum(x.numel() for x in rag) / SR

    def rag_step(k):
        enc = model.encode_infer_ragged(rag, streams=k, bandwidth_id=bw)
        return model.decode_ragged([f for f, _ in enc], streams=k, bandwidth_id=bw)
    ragged_row = {"workload": f"{n_rag} clips of {n_rag} distinct lengths, 1 - 5 s ({rag_s:.0f} audio-s): one batch-of-one "
                              "bucket per clip, encode_infer_ragged + decode_ragged"}
    for k in (1, 4):
        r_ms = timed(lambda: rag_step(k), max(2, args.steps // 2))
        ragged_row[f"streams_{k}"] = {"ms": round(r_ms, 2), "audio_s_per_s": round(rag_s / (r_ms * 1e-3), 1)}
    next_rows["ragged_batches"] = ragged_row
    del rag

    if rank == 0:
        cpu = None
        parity = None
        if world == 1 and not args.no_cpu:
            cpu_r = time_cpu(sd, cfg, CPU_SAMPLE_CLIPS, 2, 1)
            cpu = {k: cpu_r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            # BASELINE.json metric (iii): code match % of the measured path against the oracle on the same CPU sample
            from oracle import wavtok_oracle as O  # checker only
            wav_s = spec.synthetic_audio(CPU_SAMPLE_CLIPS, T, seed=100)
            feats_n, codes_n = model.encode_infer(wav_s.to(dev), bandwidth_id=bw)
            audio_n = model.decode(feats_n, bandwidth_id=bw)
            with torch.inference_mode():
                z = O.seanet_encoder(sd, cfg, wav_s.unsqueeze(1), library_lstm=True)
                _, c_ref = O.vq_infer(sd, z)
                a_ref = O.decode(sd, cfg, O.codes_to_features(sd, cfg, codes_n.cpu()), torch.tensor([0]))
            rep = O.vq_tie_report(z.permute(0, 2, 1).reshape(-1, cfg.dimension), cb.cpu(), codes_n.cpu(), c_ref)
            err = (a_ref - audio_n.cpu()).double()
            parity = {"sample": f"{CPU_SAMPLE_CLIPS} x 3 s clips ({CPU_SAMPLE_CLIPS * L} frames), plan {args.plan} vs "
                                "oracle/wavtok_oracle.py fp32",
                      "code_match_pct": round(float(rep["match_pct"]), 4),
                      "near_tie_mismatches": int(rep["near_ties"]),
                      "hard_mismatches": int(rep["hard_mismatches"]),
                      "waveform_snr_db": round(float(10 * torch.log10(a_ref.double().pow(2).sum() / err.pow(2).sum())), 2),
                      "waveform_max_abs_err": float(err.abs().max())}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32" if args.plan == 0 else "split-f16x2 operands, f32 accumulate",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "clips_per_gpu": B, "samples_per_clip": T, "frames_per_clip": L,
                           "plan": args.plan, "l2": "256 MiB flush between timed iterations; activations per step "
                           "(> 10 GB) exceed the 126 MB L2", "sharding": f"by clip, {world} rank(s), all-gather of codes"},
                "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "clocks": clocks.summary(), "breakdown": breakdown, "kernels": kernels, "next_rows": next_rows,
                "other_configs": other, "parity_sample": parity,
                "algorithmic_gflop_per_audio_s": round(sum(flops.values()) / 3 / 1e9, 3)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--plan", type=int, default=int(os.environ.get("WT_PLAN", "2")))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_native(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
