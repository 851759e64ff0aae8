"""The drop-in inside the REAL reference code: ``vqb200.patch_reference()`` against the unmodified
``models.vqvae.vqvae.VQVAE`` (vqvae.py:74-79,98-132), the ``ConvenientVQVAE.encode_and_quantize`` flow of
scripts/generate_vq_dataset.py:61-70 and BASELINE.json configs[2] (bf16 autocast training step).

The reference modules come from the git-ignored staging copy ``baseline/_ref`` (oracle/stage_ref.py writes it in the
authoring container; it travels to the GPU box with the snapshot).  Tests skip when it is absent."""
import copy
import importlib
import sys

import pytest
import torch

from oracle import stage_ref

REF_OK = stage_ref.activate()
needs_ref = pytest.mark.skipif(not REF_OK, reason="baseline/_ref is not staged (run oracle/stage_ref.py in the authoring container)")
REF_CLASSES = ("BottleneckBlock", "Bottleneck", "NoBottleneckBlock", "NoBottleneck")


def _fresh_reference_modules():
    """(bottleneck module, vqvae module) of the UNPATCHED reference."""
    for name in ("models.vqvae.vqvae", "models.vqvae.bottleneck"):
        sys.modules.pop(name, None)
    bott = importlib.import_module("models.vqvae.bottleneck")
    vqv = importlib.import_module("models.vqvae.vqvae")
    return bott, vqv


def _build_pair(seed=0):
    """The reference VQVAE and the same model with the B200 quantiser patched in, same weights."""
    import vqb200
    bott, vqv = _fresh_reference_modules()
    torch.manual_seed(seed)
    ref = vqv.VQVAE(stage_ref.default_vqvae_config())
    assert type(ref.bottleneck).__module__ == "models.vqvae.bottleneck"
    saved = {n: getattr(bott, n) for n in REF_CLASSES}
    try:
        vqb200.patch_reference()
        assert vqv.Bottleneck is vqb200.Bottleneck and bott.BottleneckBlock is vqb200.BottleneckBlock
        ours = vqv.VQVAE(stage_ref.default_vqvae_config())
    finally:
        for n, c in saved.items():
            setattr(bott, n, c)
            setattr(vqv, n, c) if hasattr(vqv, n) else None
    assert isinstance(ours.bottleneck, vqb200.Bottleneck)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())           # checkpoints load unchanged
    ours.load_state_dict(copy.deepcopy(ref.state_dict()))
    return ref, ours


@needs_ref
def test_patch_reference_inside_the_real_vqvae_cpu():
    """No GPU needed: construction, class identity, state-dict keys; a CPU forward must fail loudly (no fallback)."""
    ref, ours = _build_pair()
    assert "bottleneck.level_blocks.0.k" in ours.state_dict()
    x = torch.randn(1, 1, 128 * 8) * 0.1
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ours(x, torch.tensor([128 * 8]))


@needs_ref
def test_patch_reference_covers_the_grouped_tts_quantiser():
    """models/vqtts/bottleneck.py::Bottleneck (same class name, other module) becomes GroupedBottleneck."""
    import vqb200
    bott, vqv = _fresh_reference_modules()
    tts = importlib.import_module("models.vqtts.bottleneck")
    saved = {n: getattr(bott, n) for n in REF_CLASSES}
    saved_tts = (tts.Bottleneck, tts.BottleneckBlock)
    try:
        vqb200.patch_reference()
        assert tts.Bottleneck is vqb200.GroupedBottleneck and tts.BottleneckBlock is vqb200.BottleneckBlock
        assert vqv.Bottleneck is vqb200.Bottleneck                                    # the VQ-VAE wrapper is not clobbered
        blk = tts.Bottleneck(n_vocab=5, l_bins=8, emb_width=16, mu=0.99, threshold=1.0)
        assert blk.k.shape == (40, 16) and list(blk.state_dict().keys()) == ["k"]
    finally:
        tts.Bottleneck, tts.BottleneckBlock = saved_tts
        for n, c in saved.items():
            setattr(bott, n, c)
            setattr(vqv, n, c) if hasattr(vqv, n) else None


def _seed_codebooks(ref, ours, dev, gen):
    K, D = ref.bottleneck.level_blocks[0].k.shape
    code = torch.randn(K, D, generator=gen) * 0.05
    for m in (ref, ours):
        blk = m.bottleneck.level_blocks[0]
        blk.k = code.clone().to(dev)
        blk.k_sum = (code.clone() * 2).to(dev)
        blk.k_elem = torch.full((K,), 2.0, device=dev)
        blk.init = True
    return code


def _waveforms(gen, n=4, frames=(64, 48, 33, 20)):
    lengths = torch.tensor(frames[:n]) * 128
    x = torch.randn(n, 1, int(lengths.max()), generator=gen) * 0.1
    x = x * (torch.arange(x.shape[2]).view(1, 1, -1) < lengths.view(-1, 1, 1))
    return x, lengths


@pytest.mark.gpu
@needs_ref
def test_real_vqvae_training_step_fp32():
    """vqvae.py:98-132 end to end on the GPU: same batch, same weights, reference quantiser (stock PyTorch CUDA ops) vs
    the B200 quantiser.  Indices equal, losses / metrics / codebook state / encoder gradients agree."""
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(3)
    ref, ours = _build_pair()
    ref, ours = ref.to(dev).train(), ours.to(dev).train()
    _seed_codebooks(ref, ours, dev, gen)
    x, lengths = _waveforms(gen)
    seen = {}
    for name, m in (("ref", ref), ("ours", ours)):
        blk = m.bottleneck.level_blocks[0]
        hook = blk.register_forward_hook(lambda mod, args, out, name=name: seen.__setitem__(name, out))
        torch.manual_seed(11)
        out, metrics = m(x.to(dev), lengths.to(dev))
        out["loss"].backward()
        hook.remove()
        seen[name + "_out"], seen[name + "_metrics"] = out, metrics
    assert torch.equal(seen["ref"][0], seen["ours"][0])                              # indices
    assert torch.equal(seen["ref"][1], seen["ours"][1])                              # straight-through latents, bit for bit
    for key in ("loss", "loss_recon", "loss_stft", "loss_commit"):
        a, b = float(seen["ref_out"][key]), float(seen["ours_out"][key])
        assert abs(a - b) <= 1e-5 * abs(a) + 1e-8, (key, a, b)
    for key in ("fit", "entropy", "usage", "dk", "used_curr"):
        a, b = float(seen["ref_metrics"][key]), float(seen["ours_metrics"][key])
        assert abs(a - b) <= 1e-4 * abs(a) + 1e-7, (key, a, b)
    rb, ob = ref.bottleneck.level_blocks[0], ours.bottleneck.level_blocks[0]
    assert torch.allclose(rb.k, ob.k, rtol=1e-5, atol=1e-6)
    assert torch.allclose(rb.k_sum, ob.k_sum, rtol=1e-5, atol=1e-6) and torch.allclose(rb.k_elem, ob.k_elem, rtol=1e-5, atol=1e-6)
    worst = 0.0
    for (n1, p1), (n2, p2) in zip(ref.named_parameters(), ours.named_parameters()):
        assert n1 == n2 and (p1.grad is None) == (p2.grad is None)
        if p1.grad is not None:
            scale = float(p1.grad.abs().max()) + 1e-12
            worst = max(worst, float((p1.grad - p2.grad).abs().max()) / scale)
    assert worst < 1e-3, worst                 # cuDNN/cuBLAS reductions are not bit-reproducible between two module instances


@pytest.mark.gpu
@needs_ref
def test_generate_script_flow_and_decode():
    """scripts/generate_vq_dataset.py:61-77 restated (the script itself needs omegaconf): encoder -> ``encode`` ->
    ``decode`` -> decoder through ``level_blocks[-1]``, reference vs patched model."""
    from models.glow_tts import submodules
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(4)
    ref, ours = _build_pair(seed=1)
    ref, ours = ref.to(dev).eval(), ours.to(dev).eval()
    _seed_codebooks(ref, ours, dev, gen)
    x, lengths = _waveforms(gen)
    outs = []
    with torch.no_grad():
        for m in (ref, ours):
            x_mask = torch.unsqueeze(submodules.sequence_mask(lengths.to(dev), x.size(2)), 1).to(x.dtype)
            q, q_mask = m.encoders[-1](x.to(dev), x_mask)
            z = m.bottleneck.level_blocks[-1].encode(q, q_mask)
            xd = m.bottleneck.level_blocks[-1].decode(z)
            y, _ = m.decoders[-1]([xd], [q_mask], all_levels=False)
            outs.append((z.cpu(), xd.cpu(), y.cpu(), q_mask.sum(-1).long().cpu()))
    assert outs[0][0].dtype == torch.int64 and torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
    assert torch.allclose(outs[0][2], outs[1][2], rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
@needs_ref
def test_bf16_autocast_training_step():
    """BASELINE.json configs[2]: the training step under bf16 autocast.  The B200 quantiser takes the bf16 latents as they
    are (no ``x.float()`` copy) and computes exact FP32-semantics indices; the parity target is the reference quantiser run
    in FP32 on the same bf16 latents (``autocast(enabled=False)``, ``x.float()``), as SURVEY.md section 5 defines it."""
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(5)
    ref, ours = _build_pair(seed=2)
    ref, ours = ref.to(dev).train(), ours.to(dev).train()
    _seed_codebooks(ref, ours, dev, gen)
    rblk = ref.bottleneck.level_blocks[0]
    ref_forward = rblk.forward

    def fp32_reference_block(xq, mask, update_k=True):
        with torch.autocast("cuda", enabled=False):
            return ref_forward(xq.float(), mask.float(), update_k=update_k)

    rblk.forward = fp32_reference_block
    x, lengths = _waveforms(gen)
    seen = {}
    for name, m in (("ref", ref), ("ours", ours)):
        blk = m.bottleneck.level_blocks[0]
        hook = blk.register_forward_hook(lambda mod, args, out, name=name: seen.__setitem__(name, (args, out)))
        torch.manual_seed(12)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out, metrics = m(x.to(dev), lengths.to(dev))
        out["loss"].backward()
        hook.remove()
        seen[name + "_out"] = out
    (r_args, r_out), (o_args, o_out) = seen["ref"], seen["ours"]
    assert o_args[0].dtype == torch.bfloat16                                   # the encoder hands over bf16 latents
    assert torch.equal(r_args[0], o_args[0])
    assert torch.equal(r_out[0], o_out[0])                                     # indices
    assert torch.equal(r_out[1].float(), o_out[1].float())                     # x_q values
    for key in ("loss_commit", "loss_recon"):
        a, b = float(seen["ref_out"][key]), float(seen["ours_out"][key])
        assert abs(a - b) <= 2e-3 * abs(a) + 1e-7, (key, a, b)
    rb, ob = ref.bottleneck.level_blocks[0], ours.bottleneck.level_blocks[0]
    assert torch.allclose(rb.k_sum, ob.k_sum, rtol=1e-5, atol=1e-6) and torch.allclose(rb.k, ob.k, rtol=1e-5, atol=1e-6)
