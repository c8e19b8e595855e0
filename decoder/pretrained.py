"""`from decoder.pretrained import WavTokenizer` (reference decoder/pretrained.py:32) -> the sm_100a implementation."""
from wavtokenizer_b200.pretrained import WavTokenizer  # noqa: F401

__all__ = ["WavTokenizer"]
