"""CPU-only checks: the C-ABI library loads and exports every symbol include/vqb200.h declares, the Python
module mirrors the reference's surface, and nothing silently falls back when there is no GPU."""
import ctypes
import os
import re
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    import vqb200
    return vqb200._lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vqb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vq_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    import vqb200
    names = header_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/vqb200.h but not exported"
        assert name in vqb200._lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(vqb200._lib.SIGNATURES) == set(names)
    assert lib.vq_version() == 100


def test_workspace_query_and_error_channel(lib):
    assert lib.vq_workspace_bytes(16, 1736, 512, 128) > 512 * 128 * 2
    assert lib.vq_workspace_bytes(-1, 5, 512, 128) == 0
    # no GPU here: a compute call must fail loudly, never "succeed"
    if not torch.cuda.is_available():
        buf = (ctypes.c_float * 64)()
        idx = (ctypes.c_int64 * 8)()
        ws = ctypes.create_string_buffer(1 << 20)
        addr = (ctypes.addressof(ws) + 255) & ~255
        rc = lib.vq_assign(ctypes.addressof(buf), 1, 4, 4, ctypes.addressof(buf), 4, ctypes.addressof(idx), None, None,
                           addr, 1 << 19, 0, None)
        assert rc != 0 and b"B200" in lib.vq_last_error()


def test_module_surface_matches_reference():
    import vqb200
    blk = vqb200.BottleneckBlock(512, 128, 0.99, 1.0)
    assert list(blk.state_dict().keys()) == ["k"] and blk.k.shape == (512, 128) and blk.k.dtype == torch.float32
    assert blk.init is False and blk.k_sum is None and blk.k_elem is None
    for name in ("reset_k", "_tile", "init_k", "restore_k", "update_k", "preprocess", "postprocess", "quantize",
                 "dequantize", "encode", "decode", "forward"):
        assert callable(getattr(blk, name))
    wrap = vqb200.Bottleneck(512, 128, 0.99, 1, 1.0)
    assert list(wrap.state_dict().keys()) == ["level_blocks.0.k"] and wrap.levels == 1
    nb = vqb200.NoBottleneck(2)
    xs = [torch.zeros(1, 2, 3), torch.zeros(1, 2, 3)]
    zs, xq, losses, mets = nb(xs, [None, None])
    assert zs is xs and len(losses) == 2 and set(mets[0]) == {"entropy", "usage", "used_curr", "pn", "dk"}
    assert vqb200.NoBottleneckBlock()(xs[0], None) == (xs[0], xs[0], 0, {})
    blk.k.normal_()
    blk.restore_k(num_tokens=1024.0, threshold=2.0)
    assert blk.init and blk.threshold == 2.0 and torch.allclose(blk.k_elem, torch.full((512,), 2.0))
    assert torch.allclose(blk.k_sum, blk.k * 2.0)


def test_layout_helpers_match_oracle():
    import vqb200
    from oracle import vq_oracle as O
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 6, 5, generator=g)
    mask = (torch.arange(5).view(1, 1, 5) < torch.tensor([5, 3]).view(2, 1, 1)).float()
    blk = vqb200.BottleneckBlock(4, 6, 0.99, 1.0)
    rows, pn, mcol = blk.preprocess(x, mask)
    r2, m2, valid = O.flatten_nct(x, mask)
    assert torch.equal(rows, r2) and torch.equal(mcol, m2)
    assert torch.allclose(pn, O.prenorm(r2, valid))
    x_l, x_d, m3 = blk.postprocess(torch.arange(10), rows, (2, 5), mcol)
    assert torch.equal(x_d, x) and torch.equal(m3, mask) and x_l.shape == (2, 5)
    # 2*emb_width input: halves are summed (bottleneck.py:105-113)
    blk3 = vqb200.BottleneckBlock(4, 3, 0.99, 1.0)
    rows3, _, _ = blk3.preprocess(x, mask)
    assert torch.equal(rows3, r2[:, :3] + r2[:, 3:])


def test_cpu_tensors_fail_loudly():
    import vqb200
    blk = vqb200.BottleneckBlock(8, 4, 0.99, 1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        blk(torch.zeros(1, 4, 3), torch.ones(1, 1, 3), update_k=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        blk.decode(torch.zeros(1, 3, dtype=torch.int64))
    from importlib import import_module
    q = import_module(vqb200.__name__ + ".quantizer")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        q.restart_rows_device(torch.zeros(1, 4, 3), None, 8, torch.zeros(1, dtype=torch.int64))
    fast = vqb200.BottleneckBlock(8, 4, 0.99, 1.0, rng_parity=False)      # the device-RNG mode has no CPU path either
    fast.train()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fast(torch.zeros(1, 4, 3), torch.ones(1, 1, 3), update_k=True)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "speech-masters-thesis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "/root/reference" not in text, f


def test_patch_reference_swaps_classes():
    import vqb200
    fake = types.ModuleType("models.vqvae.bottleneck")
    fake.BottleneckBlock = fake.Bottleneck = fake.NoBottleneck = fake.NoBottleneckBlock = object
    user = types.ModuleType("models.vqvae.vqvae")
    user.Bottleneck = user.NoBottleneck = object
    sys.modules["models.vqvae.bottleneck"], sys.modules["models.vqvae.vqvae"] = fake, user
    try:
        vqb200.patch_reference()
        assert fake.BottleneckBlock is vqb200.BottleneckBlock and user.Bottleneck is vqb200.Bottleneck
        assert user.NoBottleneck is vqb200.NoBottleneck
    finally:
        del sys.modules["models.vqvae.bottleneck"], sys.modules["models.vqvae.vqvae"]
