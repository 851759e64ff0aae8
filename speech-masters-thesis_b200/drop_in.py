"""Install the B200 quantiser into the reference code base without editing it.

    import vqb200; vqb200.patch_reference()        # before `from models.vqvae.vqvae import VQVAE`

After this, ``models.vqvae.bottleneck.{BottleneckBlock, Bottleneck, NoBottleneck, NoBottleneckBlock}`` are
the classes of this package, so ``models/vqvae/vqvae.py:74-79``, ``scripts/generate_vq_dataset.py:69,75`` and
``models/transformer_lm/transformer_lm.py:96-103`` pick them up unchanged.  State dict keys are identical
(``bottleneck.level_blocks.<i>.k``), so existing checkpoints load.  ``models.vqtts.bottleneck.Bottleneck`` (the grouped,
phoneme-conditioned quantiser; vqtts/bottleneck.py:7-77) becomes ``GroupedBottleneck`` when that module is importable.
"""
import importlib
import sys


def patch_reference(module_name: str = "models.vqvae.bottleneck"):
    from . import quantizer
    mod = sys.modules.get(module_name) or importlib.import_module(module_name)
    for name in ("BottleneckBlock", "Bottleneck", "NoBottleneckBlock", "NoBottleneck"):
        setattr(mod, name, getattr(quantizer, name))
    # modules that did `from models.vqvae.bottleneck import Bottleneck, NoBottleneck` before the patch
    for other in list(sys.modules.values()):
        if other is None or other is mod or not getattr(other, "__name__", "").startswith("models."):
            continue
        for name in ("BottleneckBlock", "Bottleneck", "NoBottleneckBlock", "NoBottleneck"):
            if hasattr(other, name):
                setattr(other, name, getattr(quantizer, name))
    # the grouped quantiser lives in its own module under the same class name `Bottleneck`
    if module_name == "models.vqvae.bottleneck":
        try:
            tts = sys.modules.get("models.vqtts.bottleneck") or importlib.import_module("models.vqtts.bottleneck")
        except ImportError:
            tts = None
        if tts is not None:
            tts.BottleneckBlock = quantizer.BottleneckBlock
            tts.Bottleneck = quantizer.GroupedBottleneck
    return mod
