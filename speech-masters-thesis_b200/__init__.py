"""vqb200: B200-native (sm_100a) vector-quantisation bottleneck, drop-in for
``models/vqvae/bottleneck.py`` of vliu15/speech-masters-thesis.  See DESIGN.md / INTEGRATION.md."""
from . import _lib, dist
from .quantizer import (Bottleneck, BottleneckBlock, GroupedBottleneck, NoBottleneck, NoBottleneckBlock, assign,
                        assign_grouped, decode_nct, gather_rows, invalidate_prepared)
from .drop_in import patch_reference
from .codes import CodeShard, CodeShardWriter, HostEncoder

__all__ = ["Bottleneck", "BottleneckBlock", "GroupedBottleneck", "NoBottleneck", "NoBottleneckBlock", "assign", "assign_grouped",
           "decode_nct", "gather_rows", "invalidate_prepared", "patch_reference", "CodeShard", "CodeShardWriter", "HostEncoder", "_lib"]
