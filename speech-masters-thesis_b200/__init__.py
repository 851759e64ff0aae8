"""vqb200: B200-native (sm_100a) vector-quantisation bottleneck, drop-in for
``models/vqvae/bottleneck.py`` of vliu15/speech-masters-thesis.  See DESIGN.md / INTEGRATION.md."""
from . import _lib, dist
from .quantizer import (Bottleneck, BottleneckBlock, NoBottleneck, NoBottleneckBlock, assign, decode_nct,
                        gather_rows)
from .drop_in import patch_reference

__all__ = ["Bottleneck", "BottleneckBlock", "NoBottleneck", "NoBottleneckBlock", "assign", "decode_nct",
           "gather_rows", "patch_reference", "_lib"]
