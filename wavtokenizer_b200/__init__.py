"""wavtokenizer_b200 — B200-native (sm_100a) WavTokenizer inference path behind the reference API."""
from .audio import convert_audio, pcm16, save_audio  # noqa: F401
from .pretrained import WavTokenizer  # noqa: F401
from .spec import ModelConfig, load_config  # noqa: F401

__all__ = ["WavTokenizer", "ModelConfig", "load_config", "convert_audio", "pcm16", "save_audio"]
