"""Multi-GPU equality on real devices: the sharded encode + NCCL all-gather of the codes equals the single-GPU result
bit for bit. Skipped when fewer than two GPUs are visible; the host-side logic of the
same path is covered on CPU by tests/test_shard.py (gloo, world size 2)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world: int, n_items: int) -> str:
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mgpu_equal_worker.py"), str(n_items)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT,
                         env=dict(os.environ, NCCL_DEBUG="WARN"))
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    return out.stdout


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_codes_equal_single_gpu_world2():
    out = _run(2, 37)
    assert "equal=True" in out, out


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 8, reason="needs eight GPUs")
def test_sharded_codes_equal_single_gpu_world8():
    out = _run(8, 203)
    assert "equal=True" in out, out


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_second_device_in_one_process():
    """One process, two handles on two devices (the header's "one handle per GPU" contract): the per-device launch
    state (shared-memory opt-in, SM count, tensor-map cache) must not leak from cuda:0 to cuda:1, and the caller's
    current device must survive the calls."""
    from tests import helpers
    from wavtokenizer_b200 import WavTokenizer, spec
    cfg, sd = helpers.model("small320")
    wav = spec.synthetic_audio(3, 24000, seed=5)
    outs = []
    torch.cuda.set_device(0)
    for d in (0, 1):
        m = WavTokenizer(cfg)
        m.load_state_dict(sd)
        m = m.to(f"cuda:{d}")
        bw = torch.tensor([1], device=f"cuda:{d}")
        f, c = m.encode_infer(wav.to(f"cuda:{d}"), bandwidth_id=bw)
        a = m.decode(m.codes_to_features(c), bandwidth_id=bw)
        torch.cuda.synchronize(d)
        assert torch.cuda.current_device() == 0
        outs.append((c.cpu(), a.cpu()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
