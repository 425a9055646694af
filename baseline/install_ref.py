"""Install the UNMODIFIED reference package into baseline/_ref (git-ignored, travels to the GPU box).

    python baseline/install_ref.py [--src /root/reference]

The reference (jacquelm/AcademiCodec) is a plain source tree without setup.py / pyproject, so a
`pip install` has nothing to build: installing it means copying its `academicodec/` package.  Nothing
under baseline/_ref is ever committed (.gitignore) and nothing in academicodec_b200/ imports it; it is
used by
  * bench.py --impl reference   (the reference's own modules timed on the host CPU), and
  * tests/test_gpu_models.py    (the reference's SoundStream / HiFi-Codec models with the quantizer swapped).
`load()` puts baseline/_ref on sys.path (stubbing matplotlib, which the reference's utils.py imports and the
image lacks) and returns the imported `academicodec` package, or None when the install is absent.
"""
from __future__ import annotations

import argparse
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def install(src: str = "/root/reference") -> str | None:
    pkg = os.path.join(src, "academicodec")
    if not os.path.isdir(pkg):
        return None
    out = os.path.join(DEST, "academicodec")
    if os.path.isdir(out):
        shutil.rmtree(out)
    shutil.copytree(pkg, out, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.ipynb"))
    return out


def available() -> bool:
    return os.path.isfile(os.path.join(DEST, "academicodec", "quantization", "core_vq.py"))


def load():
    """import academicodec from baseline/_ref; None if it is not installed."""
    if not available():
        return None
    for name in ("matplotlib", "matplotlib.pylab", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                stub = types.ModuleType(name)
                stub.use = lambda *a, **k: None
                sys.modules[name] = stub
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    import academicodec  # noqa: F401  (the reference package)
    return academicodec


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    print(install(a.src) or "reference source tree not found")
