"""Make the UNMODIFIED reference importable on the GPU box: copy its hot-path packages into oracle/_ref/.

    python oracle/fetch_ref.py            # needs /root/reference (the build container)

/root/reference is a pure-Python research repo with no setup.py / pyproject.toml, so there is nothing to pip-install;
the packages the path imports (decoder/, encoder/, configs/ -- decoder/pretrained.py:4-10) are copied verbatim into
the git-IGNORED directory oracle/_ref/ (listed in .gitignore, not in .gpurunignore: like a built .so it travels to
the GPU box but never enters history). Nothing in the product package imports it; the only users are
`bench.py --impl reference`, bench.py's `cpu_baseline` leg and tests/test_oracle_golden.py (checker side).
TEST / BASELINE INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.environ.get("WT_REFERENCE_DIR", "/root/reference")
REF_DST = os.path.join(ROOT, "oracle", "_ref")
PACKAGES = ("decoder", "encoder", "configs")


def fetch(force: bool = False) -> str | None:
    """Copy the reference packages; returns the destination, or None when the reference checkout is absent."""
    if not os.path.isdir(REF_SRC):
        return REF_DST if available() else None
    for name in PACKAGES:
        src, dst = os.path.join(REF_SRC, name), os.path.join(REF_DST, name)
        if not os.path.isdir(src):
            raise FileNotFoundError(src)
        if os.path.isdir(dst):
            if not force:
                continue
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return REF_DST


def available() -> bool:
    return all(os.path.isdir(os.path.join(REF_DST, p)) for p in PACKAGES)


def reference_root() -> str:
    """Directory holding the unmodified reference packages: /root/reference when mounted, else oracle/_ref."""
    if os.path.isdir(os.path.join(REF_SRC, "decoder")):
        return REF_SRC
    if available():
        return REF_DST
    raise ImportError("the reference is not available: run `python oracle/fetch_ref.py` in the build container")


def _ours(k: str) -> bool:
    return k in ("decoder", "encoder") or k.startswith(("decoder.", "encoder."))


def load_reference(config_path: str, state_dict=None):
    """The unmodified reference model: `decoder.pretrained.WavTokenizer.from_hparams0802(config_path)` (reference
    decoder/pretrained.py:81-92), then `state_dict` merged over its own and loaded through its own
    `load_state_dict`, in eval mode.

    This repo ships drop-in shim packages that are also named decoder/ and encoder/ (INTEGRATION.md), and the
    reference resolves its YAML `class_path`s with `__import__` at construction time (pretrained.py:26): the
    reference's packages are therefore installed under those names only while the model is being built (the shims are
    stashed and put back), after which its methods need no further imports. Both can so live in one process."""
    base = reference_root()
    stash = {k: sys.modules.pop(k) for k in [k for k in sys.modules if _ours(k)]}
    sys.path.insert(0, base)
    try:
        from decoder.pretrained import WavTokenizer as Ref  # noqa: WPS433
        ref = Ref.from_hparams0802(config_path).eval()
        if state_dict is not None:
            full = dict(ref.state_dict())
            full.update(state_dict)
            ref.load_state_dict(full)
    finally:
        sys.path.remove(base)
        for k in [k for k in sys.modules if _ours(k)]:
            del sys.modules[k]
        sys.modules.update(stash)
    return ref


if __name__ == "__main__":
    dst = fetch(force="--force" in sys.argv)
    print(f"reference packages in {dst}" if dst else f"{REF_SRC} not found; nothing copied")
