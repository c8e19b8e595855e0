"""Import-path shim: `from encoder.utils import convert_audio` (reference README.md:52, infer.py:5) resolves to the
CUDA front-end / back-end in `wavtokenizer_b200.audio`. Nothing is implemented here."""
