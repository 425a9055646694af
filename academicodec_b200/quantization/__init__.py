"""Drop-in for `academicodec.quantization` (reference quantization/__init__.py:7-8)."""
from .vq import QuantizedResult, ResidualVectorQuantizer  # noqa: F401
