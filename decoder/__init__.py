"""Import-path shim: lets the reference's scripts (`from decoder.pretrained import WavTokenizer`, reference README.md:50-112,
infer.py:5-6) run unedited on the B200-native implementation in `wavtokenizer_b200`. Nothing is implemented here."""
