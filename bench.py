"""Benchmark of the WavTokenizer hot path (encode_infer -> VQ -> codes_to_features -> decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--plan P]
                    [--config small320|small600|medium] [--clips-per-gpu B] [--quick]

One "step" = one pass of the hot path over one batch of synthetic 24 kHz clips. The default workload is
BASELINE.json configs[1]: WavTokenizer-small-320 (frame75), 256 x 3 s clips per GPU (weak scaling: each rank owns its
own clips; the only collective is the all-gather of codes). Prints ONE JSON line.

  value          audio-seconds per second, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e            same metric through the host-buffer C-ABI entry (pinned host wav in, codes + audio out)
  roofline       the kernel with the largest share of the step (`rooflines`: every kernel above 2 % of the step)
  cpu_baseline   the UNMODIFIED reference (oracle/_ref, copied by oracle/fetch_ref.py) on a bounded sample, host cores
  parity_sample  code match % / waveform SNR of the measured path against the oracle on a CPU-sized sample
  other_configs  the same measurement on small-600 (the configuration north_star's target is quoted on) and on the
                 medium model at 1024 clips per rank (configs[2]); the VQ-only sweep at 1e7 frames per GPU
                 (configs[4]); decode-only detokenisation of 512 x 10 s token streams per GPU (configs[3])
  next_rows      SURVEY.md 8(f): convert_audio, the save_audio limiter + PCM16 back-end, ragged batches, SEANet decoder

`--impl reference` drives the unmodified reference's own public API (decoder/pretrained.py:186-239) on the host CPU
with all host threads, on a bounded sample of the same workload (falls back to the oracle port, kind "port", when
oracle/_ref is absent).
"""
from __future__ import annotations

import argparse
import ctypes
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

CONFIGS = {
    # name: (yaml under wavtokenizer_b200/configs, label, default clips per GPU)
    "small320": ("wavtokenizer_smalldata_frame75_3s_nq1_code4096_dim512_kmeans200_attn.yaml",
                 "WavTokenizer-small-320-24k-4096", 256),
    "small600": ("wavtokenizer_smalldata_frame40_3s_nq1_code4096_dim512_kmeans200_attn.yaml",
                 "WavTokenizer-small-600-24k-4096", 256),
    "medium": ("wavtokenizer_mediumdata_music_audio_frame75_3s_nq1_code4096_dim512_kmeans200_attn.yaml",
               "WavTokenizer-medium-music-audio-320-24k-4096", 1024),
}
T = 72000
SR = 24000
METRIC = "encode+decode audio-sec/sec (24 kHz)"
UNIT = "audio-s/s"
CPU_SAMPLE_CLIPS = 8
PARITY_CLIPS = 4
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the modal launch shape, one `ncu --set full` capture each with
# the pass sizes of the device-resident step (128-clip encoder passes, one 256-clip decoder pass):
# profiles/r02_ncu_table_traffic_final2.txt (tools/ncu_round2_traffic.sh) and r02_ncu_table_mem_final2.txt
# (tools/ncu_round2_mem.sh). <256, 3>: decoder k3 conv 768 -> 768 (197 MB read + 147 MB written; algorithmic 366 MB);
# <256, 1>: mean of ConvNeXt GEMM-1 (93 + 219 MB) and GEMM-2 (455 + 143 MB), 12 launches each; <128, 3>: level-1 strided
# conv; LSTM: one time segment (same shape as before: profiles/r02_ncu_table_final.txt).
TRAFFIC = {"tap_gemm_tc_kernel<256, 3>": 343.8e6, "tap_gemm_tc_kernel<256, 1>": 454.9e6,
           "tap_gemm_tc_kernel<128, 3>": 2.326e9, "groupnorm_kernel": 312.3e6, "dwconv_ln_kernel": 239.9e6,
           "lstm_persistent_kernel": 70.2e6, "enc_l0_tc_kernel": 1.163e9, "enc_l1_fused_kernel": 2.326e9}
CATS = ["enc_conv", "lstm", "vq", "dec_conv", "pwconv", "head_idft", "attention", "memory_bound"]
NAMED_KERNELS = {1: "lstm_persistent_kernel", 2: "resblock0_fused_kernel", 3: "groupnorm_kernel", 4: "dwconv_ln_kernel",
                 5: "layernorm_kernel", 6: "spectral_kernel", 7: "overlap_add_kernel", 8: "softmax_planes_kernel",
                 9: "vt_planes_kernel", 10: "features_to_rows_kernel", 11: "codes_to_features_kernel",
                 12: "lstm_skip_elu_pad_kernel", 14: "enc_l1_fused_kernel", 15: "enc_l0_tc_kernel"}


def cfg_path(name: str) -> str:
    return os.path.join(ROOT, "wavtokenizer_b200", "configs", CONFIGS[name][0])


def workload(name: str, clips: int) -> str:
    return f"{CONFIGS[name][1]} encode_infer+codes_to_features+decode, {clips} x 3 s clips per GPU"


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


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (oracle/_ref) or, when it is absent, the oracle port
# ---------------------------------------------------------------------------------------------------------------------
def cpu_runner(name: str, sd, cfg):
    """(step(wav, bw) -> (codes, audio), kind, description): the reference's own public API when it is available."""
    try:
        from oracle import fetch_ref  # baseline leg only
        ref = fetch_ref.load_reference(cfg_path(name), sd)

        def step(wav, bw):
            with torch.inference_mode():
                feats, codes = ref.encode_infer(wav, bandwidth_id=bw)          # decoder/pretrained.py:186-189
                audio = ref.decode(ref.codes_to_features(codes), bandwidth_id=bw)  # :209-239, :192-207
            return codes, audio
        return step, "reference", ("unmodified reference (decoder.pretrained.WavTokenizer from oracle/_ref: "
                                   "encode_infer + codes_to_features + decode, torch CPU fp32)")
    except ImportError:
        from oracle import wavtok_oracle as O  # CPU baseline leg only

        def step(wav, bw):
            with torch.inference_mode():
                feats, codes = O.encode_infer(sd, cfg, wav, library_lstm=True)
                audio = O.decode(sd, cfg, O.codes_to_features(sd, cfg, codes), bw)
            return codes, audio
        return step, "port", "oracle/wavtok_oracle.py (torch CPU fp32 ops, nn.LSTM); oracle/_ref is absent"


def time_cpu(name: str, sd, cfg, clips: int, steps: int, warmup: int) -> dict:
    torch.set_num_threads(os.cpu_count() or 1)
    step, kind, what = cpu_runner(name, sd, cfg)
    wav = spec.synthetic_audio(clips, T, seed=100)
    bw = torch.tensor([0])
    for _ in range(warmup):
        step(wav, bw)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step(wav, bw)
        ts.append(time.perf_counter() - t0)
    sec = sum(ts) / len(ts)
    return {"value": clips * T / SR / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{clips} x 3 s clips per step of the same workload, {steps} steps after {warmup} warm-up, "
                      f"{what}, {sec * 1e3:.0f} ms/step",
            "ms_per_step": sec * 1e3}


def cpu_state(name: str):
    """Seeded weights + codebook for the CPU arm (encoder frames from the oracle port: no GPU needed)."""
    from oracle import wavtok_oracle as O
    cfg = spec.load_config(cfg_path(name))

    def enc(sd, cal):
        with torch.inference_mode():
            return O.seanet_encoder(sd, cfg, cal.unsqueeze(1), library_lstm=True)
    return cfg, make_state(cfg, 1, enc)


def run_reference(args, rank: int) -> None:
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    cfg, sd = cpu_state(args.config)
    r = time_cpu(args.config, sd, cfg, CPU_SAMPLE_CLIPS, args.steps, args.warmup)
    clips = args.clips_per_gpu or CONFIGS[args.config][2]
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload(args.config, clips), "sample_clips_per_step": CPU_SAMPLE_CLIPS,
                       "device": "host CPU"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# native arm
# ---------------------------------------------------------------------------------------------------------------------
def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def build_model(name: str, dev, plan: int):
    from wavtokenizer_b200 import WavTokenizer
    cfg = spec.load_config(cfg_path(name))
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
    model.set_plan(plan)
    return cfg, sd, model


def event_ms(fn, n: int, flush, warm: int = 2) -> float:
    for _ in range(warm):
        fn()
    evp = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evp:
        flush.zero_()  # evict L2 between timed iterations (outside the event pair)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evp) / n


def kernel_records(lib, hptr, step_ms: float, peaks: dict) -> tuple:
    """Per-kernel record of ONE step timed with CUDA events around every launch on its own stream, and the roofline of
    every kernel that takes more than 2 % of the step."""
    from wavtokenizer_b200 import _native
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    hbm = float(peaks.get("hbm_gbs", 6500.0))
    src_tf = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback (B200_PROFILING.md sustained)"
    src_bw = "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback (B200_PROFILING.md)"
    kernels, roofs = {}, {}
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
    for kid, kname in NAMED_KERNELS.items():
        t_ms, n, fl, by = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double(), ctypes.c_double()
        _native.check(lib.wt_timing_read_kernel(hptr, kid, ctypes.byref(t_ms), ctypes.byref(n), ctypes.byref(fl)))
        _native.check(lib.wt_timing_read_kernel_bytes(hptr, kid, ctypes.byref(t_ms), ctypes.byref(n), ctypes.byref(by)))
        if n.value:
            kernels[kname] = {"ms": round(t_ms.value, 3), "launches": n.value,
                              "algorithmic_tflops": round(fl.value / 1e12, 4), "algorithmic_gb": round(by.value / 1e9, 4),
                              "tflops_per_s": round(fl.value / max(t_ms.value, 1e-9) / 1e9, 2),
                              "gb_per_s": round(by.value / max(t_ms.value, 1e-9) / 1e6, 1)}
    for kname, d in kernels.items():
        if d["ms"] < 0.02 * step_ms:
            continue
        base = {"kernel": kname, "launches_per_step": d["launches"], "avg_launch_ms": round(d["ms"] / d["launches"], 4),
                "share_of_step": round(d["ms"] / step_ms, 3), "traffic": TRAFFIC.get(kname)}
        if kname.startswith("tap_gemm_tc_kernel") or kname == "lstm_persistent_kernel":
            passes = d.get("passes", 3)
            r = {"bound": "tensor", "achieved": d["tflops_per_s"], "peak": peak_tf, "unit": "TFLOP/s",
                 "frac": round(d["tflops_per_s"] / peak_tf, 4), "executed_frac": round(d["tflops_per_s"] * passes / peak_tf, 4),
                 "peak_source": src_tf}
            if kname == "lstm_persistent_kernel":
                r["note"] = ("recurrence of one LSTM layer (16*D*D FLOP per clip and step): a chain of per-step latencies "
                             "(h hand-over between the 32 CTAs of a batch tile), not a throughput-bound kernel; the two "
                             "layers run as a wavefront on two streams, so their summed ms overlap in time")
        else:
            r = {"bound": "hbm", "achieved": d["gb_per_s"], "peak": hbm, "unit": "GB/s",
                 "frac": round(d["gb_per_s"] / hbm, 4), "peak_source": src_bw}
            if kname == "enc_l1_fused_kernel":
                r["note"] = ("level-0 strided conv + level-1 ResBlock in one tcgen05 kernel: 256 B of ELU(y0) planes in + "
                             "256 B of ELU(y1) planes out per level-1 position, intermediates in TMEM / shared memory; "
                             f"{d['tflops_per_s']} algorithmic TFLOP/s (3-pass split fp16)")
            if kname == "enc_l0_tc_kernel":
                r["note"] = ("conv0 + ResBlock 0: 4 B in + 128 B of split-fp16 planes out per sample; k3 / 1x1 products as "
                             "A-from-TMEM tcgen05 MMAs, conv0 + composed shortcut (448 FMA per sample) and the ELU / split "
                             f"epilogues on the CUDA cores, which bound it: {d['tflops_per_s']} algorithmic TFLOP/s")
            if kname == "resblock0_fused_kernel":
                r["note"] = ("4 B in + 128 B of split-fp16 planes out per sample; the kernel is fp32-FMA bound before it "
                             f"is HBM bound: {d['tflops_per_s']} algorithmic TFLOP/s on the CUDA cores")
        roofs[kname] = {**base, **r}
    return kernels, roofs


def measure(name: str, B: int, args, dev, rank: int, local_rank: int, world: int, flush, peaks: dict, full: bool) -> dict:
    """Device-resident value (+ e2e, breakdown, per-kernel rooflines when `full`) of one configuration."""
    import torch.distributed as dist
    from wavtokenizer_b200 import _native
    from wavtokenizer_b200.shard import gather_codes

    cfg, sd, model = build_model(name, dev, args.plan)
    L = cfg.frames_for(T)
    model.reserve(B, T)
    lib, hptr = _native.lib(), model.native().ptr
    wav_host = spec.synthetic_audio(B, T, seed=1000 + rank).pin_memory()
    wav = wav_host.to(dev)
    bw = torch.tensor([0], device=dev)

    def step():
        feats, codes = model.encode_infer(wav, bandwidth_id=bw)
        audio = model.decode(model.codes_to_features(codes), bandwidth_id=bw)
        allc = gather_codes(codes, B * world, bins=cfg.vq_bins) if world > 1 else codes
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
    out = {"name": name, "cfg": cfg, "sd": sd, "model": model, "L": L, "B": B, "ms": ms, "value": audio_s / (ms * 1e-3),
           "launches": int(launches), "clocks": clocks.summary()}

    # ---- end to end through the host-buffer C-ABI entry (pinned host in / out) ----
    for _ in range(2):
        model.encode_decode_host(wav_host, 0)
    e2e_t = []
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.encode_decode_host(wav_host, 0)
        e2e_t.append(time.perf_counter() - t0)
    e2e_ms = 1e3 * sum(e2e_t) / len(e2e_t)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    out["e2e"] = {"value": audio_s / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * T * 4,
                  "d2h_bytes_per_step": B * L * 8 + B * L * cfg.hop_length * 4, "ms_per_step": e2e_ms,
                  "api": "wt_encode_decode_host (WavTokenizer.encode_decode_host), pinned host buffers"}

    # ---- per-category and per-kernel record of one step, timed live with CUDA events on the launching stream ----
    _native.check(lib.wt_timing_enable(hptr, 1))
    step()
    torch.cuda.synchronize()
    flops = algorithmic_flops(cfg, L)
    breakdown = {}
    for i, cat in enumerate(CATS):
        t_ms, n = ctypes.c_double(), ctypes.c_int64()
        _native.check(lib.wt_timing_read(hptr, i, ctypes.byref(t_ms), ctypes.byref(n)))
        breakdown[cat] = {"ms": round(t_ms.value, 3), "launches": n.value,
                          "algorithmic_tflops": round(flops[cat] * B / 1e12, 4),
                          "tflops_per_s": round(flops[cat] * B / max(t_ms.value, 1e-9) / 1e9, 2)}
    breakdown["lstm"]["note"] = ("the two LSTM layers run as a wavefront on two streams: ms sums intervals that overlap "
                                 "in time (wall share of the step is about half of it)")
    kernels, roofs = kernel_records(lib, hptr, ms, peaks)
    _native.check(lib.wt_timing_enable(hptr, 0))
    out.update(breakdown=breakdown, kernels=kernels, rooflines=roofs, flops=flops)
    dom = max(roofs, key=lambda k: roofs[k]["share_of_step"])
    out["roofline"] = dict(roofs[dom])
    out["roofline"]["note"] = (
        "dominant kernel = largest summed CUDA-event time in one step. achieved = algorithmic FLOPs (2*M*N*K per launch, "
        "split-precision passes NOT counted) of all its launches in the step / their summed CUDA-event time on the "
        "launching stream; executed_frac counts the 3 split-fp16 MMA passes the parity bar requires (SURVEY.md App. D); "
        "traffic = dram read+write bytes per launch of the modal launch shape, ncu --set full (profiles/)")
    if not full:
        for k in ("kernels",):
            out.pop(k)
    return out


def compact(m: dict, world: int) -> dict:
    """An `other_configs` row: the same quantities as the headline line, for another model configuration."""
    total = sum(m["flops"].values())
    return {"workload": workload(m["name"], m["B"]), "value": round(m["value"], 1), "unit": UNIT,
            "ms_per_step": round(m["ms"], 3), "e2e": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in m["e2e"].items()},
            "gpu_launches": m["launches"], "frames_per_clip": m["L"],
            "algorithmic_gflop_per_audio_s": round(total / 3 / 1e9, 3),
            "algorithmic_tflops_per_s": round(total * m["B"] * world / (m["ms"] * 1e-3) / 1e12, 1),
            "breakdown_ms": {k: v["ms"] for k, v in m["breakdown"].items()},
            "roofline": m["roofline"],
            "rooflines": {k: {kk: v[kk] for kk in ("bound", "achieved", "peak", "unit", "frac", "share_of_step")}
                          for k, v in m["rooflines"].items()},
            "clocks": m["clocks"]}


def parity_sample(m: dict, dev, plan: int) -> dict:
    """BASELINE.json metric (iii): code match % of the measured path against the oracle on a CPU-sized sample."""
    from oracle import wavtok_oracle as O  # checker only
    cfg, sd, model, L = m["cfg"], m["sd"], m["model"], m["L"]
    bw = torch.tensor([0], device=dev)
    cb = sd[spec.CODEBOOK_PREFIX + "0._codebook.embed"]
    wav_s = spec.synthetic_audio(PARITY_CLIPS, T, seed=100)
    feats_n, codes_n = model.encode_infer(wav_s.to(dev), bandwidth_id=bw)
    audio_n = model.decode(feats_n, bandwidth_id=bw)
    with torch.inference_mode():
        z = O.seanet_encoder(sd, cfg, wav_s.unsqueeze(1), library_lstm=True)
        _, c_ref = O.vq_infer(sd, z)
        a_ref = O.decode(sd, cfg, O.codes_to_features(sd, cfg, codes_n.cpu()), torch.tensor([0]))
    rep = O.vq_tie_report(z.permute(0, 2, 1).reshape(-1, cfg.dimension), cb.cpu(), codes_n.cpu(), c_ref)
    err = (a_ref - audio_n.cpu()).double()
    return {"sample": f"{PARITY_CLIPS} x 3 s clips ({PARITY_CLIPS * L} frames), plan {plan} vs oracle/wavtok_oracle.py fp32",
            "code_match_pct": round(float(rep["match_pct"]), 4),
            "mismatches": int(rep["mismatches"]),
            "near_tie_mismatches": int(rep["near_ties"]),
            "near_ties_rel_distance": int(rep["near_ties_rel_distance"]),
            "hard_mismatches": int(rep["hard_mismatches"]),
            "worst_rel_gap": float(rep["worst_rel_gap"]),
            "worst_rel_gap_distance": float(rep["worst_rel_gap_distance"]),
            "native_closer_in_fp64": int(rep["a_closer_fp64"]),
            "near_tie_definition": "gap < 1e-5 * (|x|^2 + |c|^2) in fp64 (term magnitude, SURVEY.md section 7); "
                                   "near_ties_rel_distance uses north_star's literal gap < 1e-5 * distance",
            "waveform_snr_db": round(float(10 * torch.log10(a_ref.double().pow(2).sum() / err.pow(2).sum())), 2),
            "waveform_max_abs_err": float(err.abs().max())}


def extras(m: dict, args, dev, rank: int, world: int, flush, peaks: dict) -> tuple:
    """BASELINE.json configs[3] / [4] and SURVEY.md 8(f) rows on the headline model."""
    from wavtokenizer_b200 import convert_audio, pcm16
    cfg, sd, model, B = m["cfg"], m["sd"], m["model"], m["B"]
    bw = torch.tensor([0], device=dev)
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    hbm_peak = float(peaks.get("hbm_gbs", 6500.0))
    n_ev = max(2, min(args.steps, 5))
    g = torch.Generator(device=dev).manual_seed(11 + rank)
    other, next_rows = {}, {}

    # ---- configs[4]: VQ-only sweep (fused distance GEMM + argmin), 1e7 frames per GPU (1e8 at N = 8 with 1.25e7) ----
    n_vq = 10_000_000
    cb = sd[spec.CODEBOOK_PREFIX + "0._codebook.embed"].to(dev)
    xs = torch.empty(n_vq, cfg.dimension, device=dev)
    for i in range(0, n_vq, 1_000_000):  # frames near codebook rows, generated in slices (20.5 GB in total)
        sl = xs[i:i + 1_000_000]
        sl.copy_(cb[torch.randint(0, cb.shape[0], (sl.shape[0],), device=dev, generator=g)])
        sl.add_(torch.randn(sl.shape, device=dev, generator=g), alpha=2e-3)
    codes_out = torch.empty(n_vq, dtype=torch.int64, device=dev)
    from wavtokenizer_b200 import _native
    lib, hptr = _native.lib(), model.native().ptr

    def vq_all():
        with torch.cuda.device(dev):
            _native.check(lib.wt_vq(hptr, xs.data_ptr(), n_vq, codes_out.data_ptr(), None,
                                    ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    vq_ms = event_ms(vq_all, 3, flush, warm=1)
    vq_fl = 2.0 * n_vq * cfg.vq_bins * cfg.dimension
    other["vq_only"] = {
        "workload": f"{n_vq} frames x {cfg.dimension} fp32 vs {cfg.vq_bins} x {cfg.dimension} codebook per GPU, codes "
                    "only (wt_vq streams the frames in passes of 2^18)",
        "ms": round(vq_ms, 3), "frames_per_s": round(n_vq * world / (vq_ms * 1e-3), 1),
        "roofline": {"bound": "tensor", "achieved": round(vq_fl / (vq_ms * 1e-3) / 1e12, 2), "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": round(vq_fl / (vq_ms * 1e-3) / 1e12 / peak_tf, 4),
                     "executed_frac": round(3 * vq_fl / (vq_ms * 1e-3) / 1e12 / peak_tf, 4),
                     "note": "4.194 algorithmic MFLOP per frame (SURVEY.md 8(d)); includes the fp32 -> split-fp16 "
                             "plane conversion of the frames (2048 B read + 2048 B written per frame)"}}
    del xs, codes_out
    torch.cuda.empty_cache()
    # ---- configs[3]: decode-only detokenisation, 10 s random token streams, 512 per GPU (4096 at N = 8) ----
    n_st, l_st = 512, 750
    rc = torch.randint(0, cfg.vq_bins, (1, n_st, l_st), device=dev, generator=g)
    dec_ms = event_ms(lambda: model.decode(model.codes_to_features(rc), bandwidth_id=bw), 3, flush, warm=1)
    dec_fl = n_st * l_st * (125.6e6 + 3072.0 * l_st)
    other["decode_only"] = {
        "workload": f"codes_to_features + decode of {n_st} random token streams x 10 s (L = {l_st}) per GPU",
        "ms": round(dec_ms, 3), "audio_s_per_s": round(n_st * world * l_st * cfg.hop_length / SR / (dec_ms * 1e-3), 1),
        "tokens_per_s": round(n_st * world * l_st / (dec_ms * 1e-3), 1),
        "algorithmic_tflops_per_s": round(dec_fl / (dec_ms * 1e-3) / 1e12, 1),
        "note": "125.6 MFLOP per frame + 3072*L attention FLOP per frame (SURVEY.md 8(d))"}
    del rc
    torch.cuda.empty_cache()

    # ---- the convert_audio front-end (44.1 kHz stereo -> 24 kHz mono, 171 taps) ----
    sr_in = 44100
    t_in = T * sr_in // SR
    raw = torch.randn(B, 2, t_in, device=dev).clamp_(-1, 1)
    ca_ms = event_ms(lambda: convert_audio(raw, sr_in, SR, 1), n_ev, flush)
    mono = convert_audio(raw, sr_in, SR, 1)
    ca_bytes = raw.numel() * 4 + mono.numel() * 4
    next_rows["convert_audio"] = {
        "workload": f"{B} clips x 3 s, {sr_in} Hz stereo -> {SR} Hz mono (polyphase sinc, 171 taps x 80 phases)",
        "ms": round(ca_ms, 3), "audio_s_per_s": round(B * T / SR / (ca_ms * 1e-3), 1),
        "roofline": {"bound": "hbm", "achieved": round(ca_bytes / (ca_ms * 1e-3) / 1e9, 1), "peak": hbm_peak,
                     "unit": "GB/s", "frac": round(ca_bytes / (ca_ms * 1e-3) / 1e9 / hbm_peak, 4),
                     "algorithmic_bytes": ca_bytes,
                     "note": "compulsory bytes (stereo in + mono out) / CUDA-event time; this ratio is 171 FMA per output "
                             "sample, so the kernel is fp32-FMA bound before it is HBM bound: "
                             f"{round(B * T * 171 / (ca_ms * 1e-3) / 1e12, 2)} TFMA/s achieved"}}
    del raw, mono
    # ---- SURVEY.md 8(f) row 1, second half: save_audio limiter + PCM_S 16 of the decoded batch ----
    dec_audio = torch.randn(B, T, device=dev, generator=g) * 0.5
    for mode in ("clamp", "rescale"):
        p_ms = event_ms(lambda: pcm16(dec_audio, mode), n_ev, flush)
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

    def rag_step(k, batched):
        enc = model.encode_infer_ragged(rag, streams=k, batched=batched, bandwidth_id=bw)
        return model.decode_ragged([f for f, _ in enc], streams=k, bandwidth_id=bw)
    ragged_row = {"workload": f"{n_rag} clips of {n_rag} distinct lengths, 1 - 5 s ({rag_s:.0f} audio-s), "
                              "encode_infer_ragged + decode_ragged; 'buckets': one batch-of-one call per clip on k streams; "
                              "'batched': the ragged C-ABI entries (the default of encode_infer_ragged / decode_ragged)"}
    for k in (1, 4):
        r_ms = event_ms(lambda: rag_step(k, False), 2, flush, warm=1)
        ragged_row[f"buckets_streams_{k}"] = {"ms": round(r_ms, 2), "audio_s_per_s": round(rag_s / (r_ms * 1e-3), 1)}
    e_ms = event_ms(lambda: model.encode_infer_ragged(rag, batched=True, bandwidth_id=bw), 2, flush, warm=1)
    b_ms = event_ms(lambda: model.encode_infer_ragged(rag, batched=False, bandwidth_id=bw), 2, flush, warm=1)
    ragged_row["encode_only"] = {"batched_ms": round(e_ms, 2), "buckets_ms": round(b_ms, 2),
                                 "batched_audio_s_per_s": round(rag_s / (e_ms * 1e-3), 1)}

    def rag_batched():
        enc = model.encode_infer_ragged(rag, batched=True, bandwidth_id=bw)
        return model.decode_ragged([f for f, _ in enc], batched=True, bandwidth_id=bw)
    r_ms = event_ms(rag_batched, 3, flush, warm=1)
    ragged_row["batched"] = {"ms": round(r_ms, 2), "audio_s_per_s": round(rag_s / (r_ms * 1e-3), 1),
                             "api": "wt_encode_ragged + wt_decode_ragged: one LSTM recurrence and one padded decoder row "
                                    "space for all clips, per-clip lengths in the kernels that look across rows"}
    next_rows["ragged_batches"] = ragged_row
    del rag
    # ---- the SEANet decoder (feature_extractor.encodec.decoder), fp32 CUDA-core kernels ----
    try:
        from wavtokenizer_b200.pretrained import WavTokenizer
        m4 = WavTokenizer(cfg)
        full = dict(m4.state_dict())
        full.update(sd)
        full.update(spec.synthetic_seanet_decoder(cfg, 7))
        m4.load_state_dict(full)
        m4 = m4.to(dev)
        n4, l4 = 16, cfg.frames_for(T)
        z4 = 0.03 * torch.randn(n4, cfg.dimension, l4, device=dev, generator=g)
        s_ms = event_ms(lambda: m4.feature_extractor.encodec.decoder(z4), 2, flush, warm=1)
        next_rows["seanet_decoder"] = {
            "workload": f"feature_extractor.encodec.decoder on {n4} x 3 s latents (L = {l4}), random-init weights",
            "ms": round(s_ms, 3), "audio_s_per_s": round(n4 * l4 * cfg.hop_length / SR / (s_ms * 1e-3), 1),
            "note": "next to the hot path, not on it: fp32 CUDA-core tap GEMMs, one launch per LSTM step"}
        del m4, z4
    except Exception as e:  # the row is informational
        next_rows["seanet_decoder"] = {"error": str(e)[:200]}
    torch.cuda.empty_cache()
    return other, next_rows


def run_native(args, rank: int, local_rank: int, world: int) -> None:
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py (native arm) needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout (one JSON line only)
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    B = args.clips_per_gpu or CONFIGS[args.config][2]
    head = measure(args.config, B, args, dev, rank, local_rank, world, flush, peaks, full=True)
    other, next_rows = ({}, {})
    if not args.quick:
        other, next_rows = extras(head, args, dev, rank, world, flush, peaks)
    parity = cpu = None
    if rank == 0 and not args.no_cpu:
        parity = parity_sample(head, dev, args.plan)  # rank 0 at every N: a CPU-sized sample, outside every timed region
        if world == 1:
            cpu_r = time_cpu(args.config, head["sd"], head["cfg"], CPU_SAMPLE_CLIPS, 5, 2)
            cpu = {k: cpu_r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    # ---- the other model configurations BASELINE.json names, same measurement (own roofline each) ----
    if not args.quick and args.config == "small320" and not args.clips_per_gpu:
        head["model"] = None  # frees the handle (weights + workspace) before the next model is built
        torch.cuda.empty_cache()
        for name in ("small600", "medium"):
            sub_args = argparse.Namespace(**{**vars(args), "steps": max(3, min(args.steps, 5))})
            m = measure(name, CONFIGS[name][2], sub_args, dev, rank, local_rank, world, flush, peaks, full=False)
            row = compact(m, world)
            if name == "small600":
                row["note"] = ("the configuration north_star's 1e5 audio-s/s target is quoted on: 6.95 algorithmic GFLOP "
                               "per audio-second (SURVEY.md 8(d)), L = 120 frames per 3 s clip")
            else:
                row["note"] = ("BASELINE.json configs[2] at its per-rank size: 8192 clips sharded by clip over 8 GPUs = "
                               "1024 clips per rank (the model section of the YAML equals small-320's)")
            other[f"{name}_{CONFIGS[name][2]}x3s"] = row
            m["model"] = None
            del m
            torch.cuda.empty_cache()
    if rank == 0:
        cfg, L, ms = head["cfg"], head["L"], head["ms"]
        flops = head["flops"]
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32" if args.plan == 0 else "split-f16x2 operands, f32 accumulate",
                "data": "synthetic",
                "config": {"workload": workload(args.config, B), "clips_per_gpu": B, "samples_per_clip": T,
                           "frames_per_clip": L, "plan": args.plan,
                           "l2": "256 MiB flush between timed iterations; activations per step (> 10 GB) exceed the 126 MB L2",
                           "sharding": f"by clip, {world} rank(s), all-gather of codes"},
                "e2e": head["e2e"], "gpu_launches": head["launches"], "roofline": head["roofline"], "cpu_baseline": cpu,
                "clocks": head["clocks"], "breakdown": head["breakdown"], "kernels": head.get("kernels"),
                "rooflines": head["rooflines"], "next_rows": next_rows, "other_configs": other, "parity_sample": parity,
                "algorithmic_gflop_per_audio_s": round(sum(flops.values()) / 3 / 1e9, 3)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--plan", type=int, default=int(os.environ.get("WT_PLAN", "2")))
    ap.add_argument("--config", default="small320", choices=sorted(CONFIGS))
    ap.add_argument("--clips-per-gpu", type=int, default=0, help="0: the configuration's default (256; medium 1024)")
    ap.add_argument("--quick", action="store_true", help="headline measurement only (no other_configs / next_rows)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and parity_sample legs")
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
