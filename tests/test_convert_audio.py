"""convert_audio front-end: oracle vs reference goldens (CPU), CUDA vs oracle (-m gpu).

Tolerance (floating point, stated): |native - reference| <= 2e-6 absolute on unit-scale audio, i.e. the fp32
summation-order spread of a <= 171-tap FIR (the oracle itself differs from torchaudio's conv1d by <= 4e-7); the
tap tables are required to be BIT-identical to torchaudio's.
"""
import os

import numpy as np
import pytest
import torch

from oracle import audio_oracle as A
from oracle.make_golden_audio import CASES, make_input
from wavtokenizer_b200 import _native

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "convert_audio.npz")
TOL = 2e-6


def test_oracle_matches_reference_goldens():
    g = np.load(GOLDEN)
    assert set(g.files) == {c[0] for c in CASES}
    for i, (name, sr, ch, T, tsr, tch, lead) in enumerate(CASES):
        y = A.convert_audio(make_input(i, lead, ch, T).numpy(), sr, tsr, tch)
        assert y.shape == g[name].shape, name
        assert float(np.abs(y - g[name]).max()) <= 1e-6, name


def test_oracle_tap_tables_are_bit_identical_to_torchaudio():
    ta = pytest.importorskip("torchaudio")
    for sr, tsr in [(44100, 24000), (48000, 24000), (16000, 24000), (22050, 24000), (8000, 24000), (96000, 24000)]:
        k, width, orig, new = A.sinc_resample_kernel(sr, tsr)
        r = ta.transforms.Resample(sr, tsr)
        assert width == r.width and k.shape == tuple(r.kernel[:, 0].shape)
        assert np.array_equal(k, r.kernel[:, 0].numpy()), (sr, tsr)


def test_output_length_entry_point_matches_reference_shapes():
    lib = _native.lib()
    g = np.load(GOLDEN)
    for name, sr, ch, T, tsr, tch, lead in CASES:
        assert lib.wt_convert_audio_length(T, sr, tsr) == g[name].shape[-1], name
    assert lib.wt_convert_audio_length(0, 44100, 24000) == 0
    assert lib.wt_convert_audio_length(10, 0, 24000) == -1


def test_reference_error_behaviour_is_mirrored_before_any_device_work():
    from wavtokenizer_b200 import convert_audio
    with pytest.raises(AssertionError, match="at least 2 dimensions"):
        convert_audio(torch.zeros(10), 44100, 24000, 1)
    with pytest.raises(AssertionError, match="mono or stereo"):
        convert_audio(torch.zeros(3, 10), 44100, 24000, 1)
    with pytest.raises(RuntimeError, match="Impossible to convert from 2 to 3"):
        convert_audio(torch.zeros(2, 10), 44100, 24000, 3)
    with pytest.raises(RuntimeError, match="CUDA device only"):
        convert_audio(torch.zeros(1, 10), 44100, 24000, 1)  # no CPU fallback


@pytest.mark.gpu
def test_cuda_matches_reference_goldens_and_oracle():
    from wavtokenizer_b200 import convert_audio
    g = np.load(GOLDEN)
    for i, (name, sr, ch, T, tsr, tch, lead) in enumerate(CASES):
        x = make_input(i, lead, ch, T)
        y = convert_audio(x.cuda(), sr, tsr, tch).cpu().numpy()
        assert y.shape == g[name].shape, name
        assert float(np.abs(y - g[name]).max()) <= TOL, (name, float(np.abs(y - g[name]).max()))
        assert float(np.abs(y - A.convert_audio(x.numpy(), sr, tsr, tch)).max()) <= TOL, name


@pytest.mark.gpu
def test_cuda_edge_cases_against_oracle():
    from wavtokenizer_b200 import convert_audio
    g = torch.Generator().manual_seed(7)
    # mono -> stereo expand (the reference itself raises here, see wavtokenizer_b200/audio.py), odd ratios, one sample,
    # a length that is an exact multiple of the block size, a long stereo batch
    cases = [(32000, 1, 3333, 24000, 2, (2,)), (8000, 1, 1, 24000, 1, ()), (96000, 2, 4096, 24000, 1, (1,)),
             (44100, 2, 147 * 256, 24000, 1, (3,)), (11025, 1, 999, 24000, 1, (2, 2)), (48000, 2, 96000, 24000, 1, (4,))]
    for sr, ch, T, tsr, tch, lead in cases:
        x = torch.randn(*lead, ch, T, generator=g).clamp(-1, 1)
        y = convert_audio(x.cuda(), sr, tsr, tch).cpu().numpy()
        ref = A.convert_audio(x.numpy(), sr, tsr, tch)
        assert y.shape == ref.shape, (sr, T)
        assert float(np.abs(y - ref).max()) <= TOL, (sr, T, float(np.abs(y - ref).max()))
    # identity: same rate, same channels -> the input itself (torchaudio Resample returns its input)
    x = torch.randn(2, 1, 1000, generator=g)
    assert torch.equal(convert_audio(x.cuda(), 24000, 24000, 1).cpu(), x)
    # empty batch / empty clip
    assert convert_audio(torch.zeros(0, 1, 100).cuda(), 44100, 24000, 1).shape == (0, 1, 55)
    assert convert_audio(torch.zeros(1, 0).cuda(), 44100, 24000, 1).shape == (1, 0)


@pytest.mark.gpu
def test_front_end_feeds_the_hot_path():
    """convert_audio -> encode_infer -> decode, the call sequence of the reference README (README.md:50-112)."""
    from tests.gpu_util import native_model
    from wavtokenizer_b200 import convert_audio
    m = native_model("small320", 2)
    g = torch.Generator().manual_seed(3)
    wav = torch.randn(2, 44100, generator=g).clamp(-1, 1)  # [channels, T] stereo, 1 s at 44.1 kHz
    x = convert_audio(wav.cuda(), 44100, 24000, 1)
    assert x.shape == (1, 24000)
    bw = torch.tensor([0]).cuda()
    feats, codes = m.encode_infer(x, bandwidth_id=bw)
    audio = m.decode(feats, bandwidth_id=bw)
    assert codes.shape == (1, 1, 75) and audio.shape == (1, 24000) and bool(torch.isfinite(audio).all())
