"""Stage the UNMODIFIED reference modules of the hot path (and its two callers' layers) under ``baseline/_ref/``.

TEST INFRASTRUCTURE ONLY (same rules as vq_oracle.py: only tests/, smoke() and bench.py's CPU legs use it).

The reference is pure Python, so "building" it is copying the few modules the path needs from where they lie under
``/root/reference`` into the git-ignored (but NOT gpurun-ignored) ``baseline/_ref/``, so that the files travel to the GPU
box like a built ``.so`` does.  Nothing is ever committed from there.  ``/root/reference`` only exists in the authoring
container: on the GPU box this script is a no-op and the staged copy (if any) is used as is.

What it enables on the box:
  * ``cpu_baseline.kind == "reference"``: bench.py times ``models.vqvae.bottleneck.BottleneckBlock`` itself;
  * the drop-in test: ``vqb200.patch_reference()`` against the real ``models.vqvae.vqvae.VQVAE`` / ``Encoder`` / ``Decoder``;
  * the grouped quantiser oracle: ``models.vqtts.bottleneck.Bottleneck``.
"""
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(os.path.dirname(HERE), "baseline", "_ref")

FILES = [
    "models/__init__.py", "models/base.py",
    "models/vqvae/__init__.py", "models/vqvae/bottleneck.py", "models/vqvae/encdec.py", "models/vqvae/conv.py",
    "models/vqvae/resnet.py", "models/vqvae/vqvae.py", "models/vqvae/losses.py",
    "models/vqtts/__init__.py", "models/vqtts/bottleneck.py",
    "models/glow_tts/__init__.py", "models/glow_tts/submodules.py",
    "datasets/__init__.py", "datasets/transforms.py",
    "utils/__init__.py", "utils/torch_utils.py",
    "configs/models/vqvae.yaml",
]


def stage(verbose=False):
    """Copy FILES from /root/reference when it exists.  Returns True when baseline/_ref holds the modules afterwards."""
    if os.path.isdir(REF_SRC):
        for rel in FILES:
            src, dst = os.path.join(REF_SRC, rel), os.path.join(REF_DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if not os.path.exists(dst) or open(src, "rb").read() != open(dst, "rb").read():
                shutil.copyfile(src, dst)
                if verbose:
                    print("staged", rel)
    return available()


def available():
    return os.path.exists(os.path.join(REF_DST, "models", "vqvae", "bottleneck.py"))


def _librosa_stub():
    """``datasets/transforms.py`` imports librosa (absent from this image) for two helpers used by ``STFT.__init__``;
    they are restated here from their documented behaviour.  Anything else raises."""
    import numpy as np

    def pad_center(data, size, axis=-1, **kwargs):
        n = data.shape[axis]
        lpad = int((size - n) // 2)
        widths = [(0, 0)] * data.ndim
        widths[axis] = (lpad, int(size - n - lpad))
        return np.pad(data, widths, mode=kwargs.get("mode", "constant"))

    def tiny(x):
        x = np.asarray(x)
        dtype = x.dtype if np.issubdtype(x.dtype, np.floating) else np.dtype(np.float32)
        return np.finfo(dtype).tiny

    lib = types.ModuleType("librosa")
    util = types.ModuleType("librosa.util")
    util.pad_center, util.tiny = pad_center, tiny
    filters = types.ModuleType("librosa.filters")

    def _missing(*a, **k):
        raise RuntimeError("librosa is not installed in this image (stub)")

    filters.mel = filters.window_sumsquare = _missing
    lib.util, lib.filters = util, filters
    return {"librosa": lib, "librosa.util": util, "librosa.filters": filters}


def activate():
    """Put baseline/_ref first on sys.path (and the librosa stub in sys.modules when librosa is missing).  Returns False
    when nothing is staged.  Importing ``models.vqvae.bottleneck`` afterwards gives the unmodified reference."""
    stage()
    if not available():
        return False
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    try:
        import librosa  # noqa: F401
    except ImportError:
        for name, mod in _librosa_stub().items():
            sys.modules.setdefault(name, mod)
    return True


def default_vqvae_config():
    """configs/models/vqvae.yaml as the attribute-style object ``VQVAE.__init__`` reads (vqvae.py:15-96)."""
    import yaml
    from types import SimpleNamespace

    def ns(d):
        return SimpleNamespace(**{k: ns(v) if isinstance(v, dict) else v for k, v in d.items()})

    with open(os.path.join(REF_DST, "configs", "models", "vqvae.yaml")) as f:
        return ns(yaml.safe_load(f))


if __name__ == "__main__":
    print("reference staged:" if stage(verbose=True) else "reference not available:", REF_DST)
