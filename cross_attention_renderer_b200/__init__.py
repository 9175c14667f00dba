"""B200-native per-ray rendering path of the cross-attention renderer.

Drop-in for ``CrossAttentionRenderer.forward`` of the reference
(reference models.py:190-626): same constructor, ``state_dict`` keys and
``out_dict``; the arithmetic runs in hand-written sm_100a CUDA kernels behind
a C-ABI library (include/car_b200.h).  There is no CPU fallback: importing
``models`` works anywhere, but rendering raises if the CUDA library or a GPU is
missing.
"""
from . import params, synthetic  # noqa: F401

__all__ = ["params", "synthetic"]
__version__ = "0.1.0"
