"""`from encoder.utils import convert_audio, save_audio` (reference encoder/utils.py:79-103) -> the CUDA kernels behind
wt_convert_audio / wt_save_audio_pcm16."""
from wavtokenizer_b200.audio import convert_audio, save_audio  # noqa: F401

__all__ = ["convert_audio", "save_audio"]
