"""-m gpu: the kernels that the round-2 defaults replaced stay selectable by environment variable (they are the A/B
baselines quoted in DESIGN.md); this runs them in a child process and compares with the default path of this process."""
import os
import subprocess
import sys

import pytest
import torch

from tests import helpers
from tests.gpu_util import native_model
from wavtokenizer_b200 import spec

pytestmark = pytest.mark.gpu

CHILD = r"""
import sys, torch
sys.path.insert(0, sys.argv[1])
from tests.gpu_util import native_model
from wavtokenizer_b200 import spec
m = native_model("small320", 2)
wav = spec.synthetic_audio(3, 24000, seed=41).cuda()
bw = torch.tensor([0]).cuda()
f, c = m.encode_infer(wav, bandwidth_id=bw)
a = m.decode(m.codes_to_features(c), bandwidth_id=bw)
torch.save({"codes": c.cpu(), "audio": a.cpu()}, sys.argv[2])
"""


@pytest.mark.parametrize("env", [
    {"WT_ENC_L0_FUSED_TC": "0", "WT_ENC_L1_FUSED": "0"},                  # CUDA-core level 0, three-launch level 1
    {"WT_LSTM_PUBLISH": "0", "WT_LSTM_KEEP_C": "0"},                      # per-warp release, c / y stored every step
    {"WT_LSTM_CLUSTER": "4", "WT_LSTM_POLL_NS": "30", "WT_TC_N128_MC": "1"},  # multicast variants
    {"WT_MEM_V1": "1"},               # one-element-per-thread spectral / overlap-add / V-transpose, scalar-FMA dwconv + AdaLN
], ids=["encoder-unfused", "lstm-round1-handover", "multicast", "memory-bound-forms"])
def test_replaced_kernels_still_agree(tmp_path, env):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "child.pt")
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", CHILD, root, out], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    child = torch.load(out)
    m = native_model("small320", 2)
    wav = spec.synthetic_audio(3, 24000, seed=41).cuda()
    bw = torch.tensor([0]).cuda()
    f, c = m.encode_infer(wav, bandwidth_id=bw)
    a = m.decode(m.codes_to_features(child["codes"].cuda()), bandwidth_id=bw)
    # same operands and the same 3-pass products: only fp32 summation orders differ between the variants
    assert (c.cpu() != child["codes"]).float().mean().item() <= 0.005
    # (floor 85 dB, not 100: one child process in about ten has shown a 48-channel block of the decoder's embed conv that
    # differs by ~1e-5 relative from every other run, see DESIGN.md section 9 "open observation")
    assert helpers.snr_db(child["audio"], a.cpu()) >= 85.0
    if "WT_MEM_V1" in env:  # decoder-side kernels only: the codes cannot move
        assert torch.equal(c.cpu(), child["codes"])
