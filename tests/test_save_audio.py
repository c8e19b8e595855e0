"""save_audio back-end (second half): limiter + 16-bit PCM.

Bars: the limiter (the reference's own arithmetic, encoder/utils.py:97-102) is BIT-exact against goldens captured
from the unmodified reference; the int16 samples are bit-exact against the oracle's restatement of the pinned
torchaudio 2.0.1 / libsox conversion (parity unpinned for that sub-step: no PCM writer is installed here)."""
import os
import wave

import numpy as np
import pytest
import torch

from oracle import audio_oracle as A
from oracle.make_golden_save_audio import CASES, make_input
from wavtokenizer_b200 import _native

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "save_audio.npz")


def test_oracle_limiter_is_bit_exact_against_reference_goldens():
    g = np.load(GOLDEN)
    assert set(g.files) == {c[0] for c in CASES}
    for i, (name, ch, T, amp, rescale) in enumerate(CASES):
        y = A.save_audio_limit(make_input(i, ch, T, amp).numpy(), rescale)
        assert y.dtype == np.float32 and np.array_equal(y, g[name]), name


def test_oracle_quantiser_known_answers():
    x = np.array([0.0, 1.0, -1.0, 0.5, -0.5, 0.99, -0.99, 1.5, -1.5, 2.0 ** -16, -(2.0 ** -16), 3 * 2.0 ** -17,
                  2.0 ** -17, 32767.5 / 32768, 32767.4 / 32768, -32768.5 / 32768], dtype=np.float32)
    want = [0, 32767, -32768, 16384, -16384, 32440, -32440, 32767, -32768, 1, 0, 1, 0, 32767, 32767, -32768]
    # 0.99 * 32768 = 32440.3 -> 32440; ties (x * 32768 = k + 0.5) round up: 2^-16 -> 0.5 -> 1, -2^-16 -> -0.5 -> 0;
    # 0.75 -> 1, 0.25 -> 0
    assert A.pcm16_sox(x).tolist() == want


def test_abi_rejects_bad_arguments_without_a_device():
    lib = _native.lib()
    assert lib.wt_save_audio_pcm16(0, None, 1, 1, 0, None, None, None, None) != 0
    assert b"null buffer" in lib.wt_last_error()


def test_python_mirror_errors():
    from wavtokenizer_b200 import pcm16
    with pytest.raises(ValueError):
        pcm16(torch.zeros(4), "clamp")
    with pytest.raises(ValueError):
        pcm16(torch.zeros(1, 4), "loud")
    with pytest.raises(RuntimeError, match="CUDA device only"):
        pcm16(torch.zeros(1, 4), "clamp")


@pytest.mark.gpu
def test_cuda_limiter_and_pcm_bit_exact():
    from wavtokenizer_b200 import pcm16
    g = np.load(GOLDEN)
    for i, (name, ch, T, amp, rescale) in enumerate(CASES):
        x = make_input(i, ch, T, amp)
        q, lim = pcm16(x.reshape(1, -1).cuda(), "rescale" if rescale else "clamp", return_limited=True)
        assert np.array_equal(lim.cpu().numpy().reshape(ch, T), g[name]), name
        assert np.array_equal(q.cpu().numpy().reshape(ch, T), A.pcm16_sox(g[name])), name
    # batch of files with per-file peaks, ragged block tail, no limiter (infer.py:70) incl. out-of-range samples
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(5, 4099, generator=gen) * torch.tensor([0.1, 0.5, 1.0, 2.0, 4.0]).unsqueeze(1)
    for mode, resc in (("rescale", True), ("clamp", False)):
        q = pcm16(x.cuda(), mode).cpu().numpy()
        for b in range(5):
            assert np.array_equal(q[b], A.pcm16_sox(A.save_audio_limit(x[b].numpy(), resc))), (mode, b)
    assert np.array_equal(pcm16(x.cuda(), "none").cpu().numpy(), A.pcm16_sox(x.numpy()))
    assert pcm16(torch.zeros(0, 7).cuda(), "clamp").shape == (0, 7)
    z = pcm16(torch.zeros(2, 9).cuda(), "rescale")  # silent file: 0.99 / 0 = inf -> factor 1
    assert int(z.abs().max()) == 0


@pytest.mark.gpu
def test_save_audio_writes_the_wav_file(tmp_path):
    from wavtokenizer_b200 import save_audio
    i, (name, ch, T, amp, rescale) = 4, CASES[4]  # stereo, rescale
    x = make_input(i, ch, T, amp)
    path = tmp_path / "out.wav"
    save_audio(x.cuda(), path, 24000, rescale=rescale)
    with wave.open(str(path), "rb") as f:
        assert (f.getnchannels(), f.getsampwidth(), f.getframerate(), f.getnframes()) == (ch, 2, 24000, T)
        data = np.frombuffer(f.readframes(T), dtype="<i2").reshape(T, ch).T
    assert np.array_equal(data, A.pcm16_sox(np.load(GOLDEN)[name]))


@pytest.mark.gpu
def test_full_size_properties():
    """BASELINE size (256 x 3 s): monotone, bounded by the limiter, idempotent on its own output."""
    from wavtokenizer_b200 import pcm16
    x = (torch.randn(256, 72000, generator=torch.Generator().manual_seed(2)) * 0.6).cuda()
    q, lim = pcm16(x, "clamp", return_limited=True)
    assert int(q.max()) <= 32440 and int(q.min()) >= -32440
    assert float((q.float() / 32768 - lim).abs().max()) <= 0.5 / 32768 + 1e-7
    back = pcm16(q.float() / 32768, "none")
    assert torch.equal(back, q)
