"""GPU parity tests (run on the B200 box with ``-m gpu``): the CUDA path, called through the module and
through the raw C ABI, against (a) the golden vectors the unmodified reference produced and (b) the CPU
oracle on seeded random inputs; plus size-independent properties at BASELINE.json sizes.

Bars: indices bit-exact except near-ties (fp64 gap <= eps = 2^-18 (||x||^2 + max||e||^2), counted);
straight-through latents bit-exact wherever the index matches; scalars rel 1e-5; codebook state rel 1e-5;
gradients rel 1e-5 (abs 1e-7)."""
import ctypes
import glob
import os

import numpy as np
import pytest
import torch

from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TRAIN_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "train_*.npz"))) + ["eval_k128_d64"]
ALGOS = ["simt", "auto"]
SAFE_GAP = 1e-3


@pytest.fixture(scope="module")
def vq():
    import __graft_entry__ as ge
    ge.build()
    import vqb200
    assert torch.cuda.is_available(), "GPU tests need a GPU"
    assert vqb200._lib.load().vq_device_supported() == 1, "not a compute-capability-10.x device"
    return vqb200


DEV = torch.device("cuda:0")


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def T(a, dev=None):
    t = torch.from_numpy(np.asarray(a))
    return t.to(dev) if dev is not None else t


def close(a, b, rtol=1e-5, atol=1e-7):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else a
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else b
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


def seeded_block(vq, g, algo):
    K, D = g["k0"].shape
    blk = vq.BottleneckBlock(K, D, float(g["mu"]), float(g["threshold"]), algo=algo).to(DEV)
    blk.k = T(g["k0"], DEV).clone()
    blk.k_sum = T(g["k_sum0"], DEV).clone()
    blk.k_elem = T(g["k_elem0"], DEV).clone()
    blk.init = True
    return blk


def check_indices(rows, k, want, got, gap=None):
    """Exact where the fp32 gap is comfortable; every other disagreement must be a near-tie."""
    want, got = torch.as_tensor(want).reshape(-1), torch.as_tensor(got).reshape(-1).cpu()
    if gap is not None:
        safe = torch.from_numpy(gap > SAFE_GAP)
        assert torch.equal(want[safe], got[safe])
    rep = O.audit_indices(rows, k, want, got)
    assert rep["errors"] == 0, rep
    return want == got


# ------------------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("name", TRAIN_CASES)
def test_forward_against_reference_golden(vq, name, algo):
    g = load(name)
    blk = seeded_block(vq, g, algo)
    blk.train()
    x = T(g["x"], DEV).clone().requires_grad_(True)
    mask = T(g["mask"], DEV)
    torch.manual_seed(int(g["rng_seed"]))
    x_l, x_q, commit, metrics = blk(x, mask, update_k=bool(g["update_k"]))
    assert x_l.dtype == torch.int64 and tuple(x_l.shape) == g["x_l"].shape and tuple(x_q.shape) == g["x"].shape
    rows, _, _ = O.flatten_nct(T(g["x"]), T(g["mask"]))
    same = check_indices(rows, T(g["k0"]), g["x_l"], x_l, g["gap"])
    n, d, t = g["x"].shape
    same_nct = same.view(n, 1, t).expand(n, d, t)
    assert torch.equal(x_q.detach().cpu()[same_nct], T(g["x_q"])[same_nct])         # bit-exact latents
    ((T(g["grad_w"], DEV) * x_q).sum() + float(g["grad_commit"]) * commit).backward()
    assert set(metrics) == {k[7:] for k in g if k.startswith("metric_")}
    if same.all():
        close(commit, g["commit"])
        close(metrics["fit"], g["metric_fit"])
        close(x.grad, g["grad_x"], rtol=1e-5, atol=1e-7)
        if bool(g["update_k"]):
            close(blk.k_sum, g["k_sum1"], rtol=1e-5, atol=1e-6)
            close(blk.k_elem, g["k_elem1"], rtol=1e-5, atol=1e-6)
            assert metrics["used_curr"].dtype == torch.int64 and metrics["used_curr"].dim() == 0
            assert int(metrics["used_curr"]) == int(g["metric_used_curr"])
            assert float(metrics["usage"]) == float(g["metric_usage"])
            close(metrics["entropy"], g["metric_entropy"])
            if int(g["mask"].sum()) >= g["k0"].shape[0]:   # restart rows replay the reference's CPU randperm exactly
                close(blk.k, g["k1"], rtol=1e-5, atol=1e-6)
                close(metrics["dk"], g["metric_dk"])
            else:                         # fewer rows than codes: randn_like runs on another device's RNG
                alive = T(g["k_elem1"]) >= float(g["threshold"])
                close(blk.k.cpu()[alive], T(g["k1"])[alive], rtol=1e-5, atol=1e-6)
                valid_rows = rows[T(g["mask"]).reshape(-1) != 0]
                dead = blk.k.cpu()[~alive]
                dmin = torch.cdist(dead, valid_rows).min(dim=1).values
                assert float(dmin.max()) < 0.01 * 6          # jittered copies of batch rows (std 0.01/sqrt(D) per dim)
    for v in metrics.values():
        assert v.dim() == 0


def test_init_k_replays_reference_rng(vq):
    g = load("init_k64_d32")
    K, D = g["k1"].shape
    blk = vq.BottleneckBlock(K, D, 0.99, 1.0).to(DEV)
    blk.train()
    torch.manual_seed(int(g["rng_seed"]))
    x_l, x_q, commit, metrics = blk(T(g["x"], DEV), T(g["mask"], DEV), update_k=True)
    assert blk.init
    assert torch.equal(x_l.cpu(), T(g["x_l"]))
    assert torch.equal(x_q.cpu(), T(g["x_q"]))
    close(blk.k, g["k1"], rtol=1e-5, atol=1e-6)
    close(blk.k_elem, g["k_elem1"], rtol=1e-5, atol=1e-6)
    close(commit, g["commit"])
    for key in ("fit", "entropy", "usage", "dk"):
        close(metrics[key], g["metric_" + key])


def test_init_k_with_fewer_rows_than_codes(vq):
    g = load("init_tile_k128_d8")
    K, D = g["k1"].shape
    blk = vq.BottleneckBlock(K, D, 0.99, 1.0).to(DEV)
    blk.train()
    x, mask = T(g["x"], DEV), T(g["mask"], DEV)
    x_l, x_q, commit, metrics = blk(x, mask, update_k=True)
    rows, _, valid = O.flatten_nct(T(g["x"]), T(g["mask"]))
    assert int(valid.sum()) < K and blk.k.shape == (K, D)
    # every code is a jittered copy of a valid batch row, so every valid row quantises (almost) onto itself
    assert float(commit) < 1e-3 and int(metrics["used_curr"]) == int(g["metric_used_curr"])


def test_encode_decode_and_eval_wrapper(vq):
    g = load("encode_k512_d128")
    K, D = g["k0"].shape
    wrap = vq.Bottleneck(K, D, 0.99, 1, 1.0).to(DEV)
    wrap.load_state_dict({"level_blocks.0.k": T(g["k0"])})
    wrap.eval()
    blk = wrap.level_blocks[0]
    x, mask = T(g["x"], DEV), T(g["mask"], DEV)
    rows, _, _ = O.flatten_nct(T(g["x"]), T(g["mask"]))
    z = blk.encode(x, mask)
    check_indices(rows, T(g["k0"]), g["z"], z, g["gap"])
    assert torch.equal(blk.decode(T(g["z"], DEV)).cpu(), T(g["x_dec"]))
    assert torch.equal(wrap.decode([T(g["z"], DEV)])[0].cpu(), T(g["x_dec"]))
    xr = x.clone().requires_grad_(True)
    zs, xqs, commits, mets = wrap([xr], [mask])
    assert len(mets) == int(g["wrap_n_metrics"]) == 0 and not xqs[0].requires_grad
    same = check_indices(rows, T(g["k0"]), g["wrap_z"], zs[0], g["gap"])
    if same.all():
        assert torch.equal(xqs[0].cpu(), T(g["wrap_xq"]))
        close(commits[0], g["wrap_commit"])
    assert not blk.init            # eval never initialises the codebook


# ------------------------------------------------------------------------------ random inputs vs the oracle
SHAPES = [  # (N, D, T, K, clustered)
    (4, 128, 300, 512, False),
    (3, 128, 257, 512, True),
    (2, 64, 130, 1024, False),
    (2, 48, 37, 40, False),       # odd T, D not a multiple of 16, K < 128
    (1, 256, 512, 256, True),
    (5, 16, 1, 8, False),         # T == 1
    (2, 512, 96, 640, False),
    (1, 8, 2000, 3, True),
    (2, 192, 64, 300, True),      # two depth slices in the asynchronous K2 kernel, the second one partial
]


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("shape", SHAPES)
def test_training_step_against_oracle(vq, shape, algo):
    n, d, t, K, clustered = shape
    gen = torch.Generator().manual_seed(hash(shape) % 10007)
    code = torch.randn(K, d, generator=gen)
    lengths = torch.randint(max(1, t // 3), t + 1, (n,), generator=gen)
    lengths[0] = t
    x, mask = O.synthetic_batch(lengths, d, gen, codebook=code if clustered else None)
    st = O.CodebookState(K, d, 0.99, 1.0, code.clone(), code.clone() * 3, torch.full((K,), 3.0), True)
    blk = vq.BottleneckBlock(K, d, 0.99, 1.0, algo=algo).to(DEV)
    blk.k, blk.k_sum, blk.k_elem, blk.init = code.to(DEV), (code * 3).to(DEV), torch.full((K,), 3.0, device=DEV), True
    blk.train()
    xg = x.to(DEV).requires_grad_(True)
    torch.manual_seed(17)
    x_l, x_q, commit, metrics = blk(xg, mask.to(DEV), update_k=True)
    w = torch.randn(x.shape, generator=gen)
    ((w.to(DEV) * x_q).sum() + 0.3 * commit).backward()
    xo = x.clone().requires_grad_(True)
    torch.manual_seed(17)
    o_l, o_q, o_commit, o_metrics = O.forward(st, xo, mask, update_k=True)
    ((w * o_q).sum() + 0.3 * o_commit).backward()
    rows, _, _ = O.flatten_nct(x, mask)
    same = check_indices(rows, code, o_l, x_l)
    same_nct = same.view(n, 1, t).expand(n, d, t)
    assert torch.equal(x_q.detach().cpu()[same_nct], o_q.detach()[same_nct])
    close(metrics["fit"], o_metrics["fit"], rtol=2e-5)
    if same.all():
        close(commit, o_commit)
        close(xg.grad, xo.grad, rtol=1e-5, atol=1e-7)
        close(blk.k_sum, st.k_sum, rtol=1e-5, atol=1e-5)
        close(blk.k_elem, st.k_elem, rtol=1e-6, atol=1e-6)
        assert int(metrics["used_curr"]) == int(o_metrics["used_curr"])
        assert float(metrics["usage"]) == float(o_metrics["usage"])
        close(metrics["entropy"], o_metrics["entropy"])
        if int(mask.sum()) >= K:
            close(blk.k, st.k, rtol=1e-5, atol=1e-5)
            close(metrics["dk"], o_metrics["dk"], rtol=1e-4)


def test_row_major_entry_points(vq):
    gen = torch.Generator().manual_seed(5)
    K, D, M = 96, 24, 700
    code, rows = torch.randn(K, D, generator=gen), torch.randn(M, D, generator=gen)
    blk = vq.BottleneckBlock(K, D, 0.99, 1.0).to(DEV)
    blk.k, blk.k_sum, blk.k_elem, blk.init = code.to(DEV), code.to(DEV).clone(), torch.ones(K, device=DEV), True
    x_l, fit = blk.quantize(rows.to(DEV))
    o_l, o_fit, _ = O.assign(rows, code)
    check_indices(rows, code, o_l, x_l)
    close(fit, o_fit, rtol=2e-5)
    _, fit_m = blk.quantize(rows.to(DEV), torch.ones(M, 1, device=DEV))
    close(fit_m, O.assign(rows, code, torch.ones(M, 1))[1], rtol=2e-5)
    assert torch.equal(blk.dequantize(o_l.to(DEV)).cpu(), O.gather(o_l, code))
    st = O.CodebookState(K, D, 0.99, 1.0, code.clone(), code.clone(), torch.ones(K), True)
    torch.manual_seed(3)
    m = blk.update_k(rows.to(DEV), o_l.to(DEV))
    torch.manual_seed(3)
    om = O.update_codebook(st, rows, o_l)
    close(blk.k, st.k, rtol=1e-5, atol=1e-5)
    close(blk.k_sum, st.k_sum, rtol=1e-5, atol=1e-5)
    for key in ("entropy", "usage", "dk"):
        close(m[key], om[key], rtol=1e-4)


def test_exact_ties_pick_lowest_index(vq):
    gen = torch.Generator().manual_seed(9)
    K, D = 256, 128
    code = torch.randn(K, D, generator=gen)
    code[1::2] = code[0::2]
    x = code[torch.randint(0, K, (3, 200), generator=gen)].permute(0, 2, 1).contiguous()   # rows ARE codes
    for algo in ALGOS:
        idx, _ = vq.assign(x.to(DEV), code.to(DEV), algo=algo)
        assert bool((idx % 2 == 0).all()), algo


def test_laplace_smoothing_is_opt_in(vq):
    gen = torch.Generator().manual_seed(2)
    K, D = 64, 16
    code = torch.randn(K, D, generator=gen)
    lengths = torch.tensor([150, 99])
    x, mask = O.synthetic_batch(lengths, D, gen, codebook=code)
    outs = []
    for eps in (0.0, 1e-2):
        blk = vq.BottleneckBlock(K, D, 0.99, 0.0, laplace_eps=eps).to(DEV)
        blk.k, blk.k_sum, blk.k_elem, blk.init = code.to(DEV), code.to(DEV).clone(), torch.ones(K, device=DEV), True
        blk.train()
        torch.manual_seed(0)
        blk(x.to(DEV), mask.to(DEV))
        outs.append(blk.k.cpu())
    assert not torch.equal(outs[0], outs[1]) and torch.allclose(outs[0], outs[1], rtol=0.05, atol=0.05)


# ------------------------------------------------------------------------------ properties at full size
def test_full_size_properties(vq):
    """BASELINE.json configs[0]/[1] sizes (K=512, D=128, LJSpeech-length batch): properties that need no oracle
    pass over the data: encode->decode->encode is idempotent, decoded rows are codebook rows, per-code counts
    sum to the number of valid frames, per-code sums add up to the column sums of the valid frames, fit
    equals sum(min_d)/K, and a 4k-row sample agrees with the oracle."""
    gen = torch.Generator().manual_seed(0)
    K, D, n = 512, 128, 64
    lengths = O.ljspeech_like_lengths(n, gen)
    code = torch.randn(K, D, generator=gen)
    x, mask = O.synthetic_batch(lengths, D, gen, codebook=code)
    xd, md, kd = x.to(DEV), mask.to(DEV), code.to(DEV)
    blk = vq.BottleneckBlock(K, D, 0.99, 1.0).to(DEV)
    blk.k, blk.k_sum, blk.k_elem, blk.init = kd.clone(), kd.clone(), torch.ones(K, device=DEV), True
    z = blk.encode(xd, md)
    dec = blk.decode(z)
    assert torch.equal(dec, kd[z].permute(0, 2, 1))
    assert torch.equal(blk.encode(dec, md), z)                                   # idempotent
    blk.train()
    torch.manual_seed(1)
    x_l, x_q, commit, metrics = blk(xd, md, update_k=True)
    assert torch.equal(x_l, z)
    _, min_d = vq.assign(xd, kd, want_min_d=True)
    close(metrics["fit"], min_d.double().sum() / K, rtol=1e-5)
    # statistics via the raw ABI
    lib = vq._lib.load()
    stats = torch.zeros(K * D + K, device=DEV)
    rc = lib.vq_ema_accumulate(xd.data_ptr(), z.data_ptr(), md.data_ptr(), n, D, x.shape[2], K, stats.data_ptr(),
                               None, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.vq_last_error()
    assert float(stats[K * D:].sum()) == float(mask.sum())
    col = (xd * md).double().sum(dim=(0, 2))
    close(stats[:K * D].view(K, D).double().sum(0), col, rtol=1e-4, atol=1e-2)
    assert torch.equal(stats[K * D:].cpu(), torch.bincount(z[md[:, 0] != 0].cpu(), minlength=K).float())
    # oracle on a slice
    rows, _, _ = O.flatten_nct(x[:4], mask[:4])
    o_l, _, _ = O.assign(rows, code)
    check_indices(rows, code, o_l, z[:4])
    # masked latents are zero exactly where the mask is
    assert float((x_q * (1 - md)).abs().max()) == 0.0


def test_large_codebook_paths(vq):
    """K beyond the shared-memory EMA slab and beyond one code tile (sweep corner K=8192, D=64)."""
    gen = torch.Generator().manual_seed(4)
    K, D = 8192, 64
    code = torch.randn(K, D, generator=gen)
    lengths = torch.tensor([700, 512, 300])
    x, mask = O.synthetic_batch(lengths, D, gen, codebook=code)
    rows, _, valid = O.flatten_nct(x, mask)
    o_l, _, _ = O.assign(rows, code)
    for algo in ALGOS:
        idx, _ = vq.assign(x.to(DEV), code.to(DEV), algo=algo)
        check_indices(rows, code, o_l, idx)
    s_sum, s_elem = O.local_statistics(rows[valid], o_l[valid], K)
    lib = vq._lib.load()
    stats = torch.zeros(K * D + K, device=DEV)
    z, xd, md = o_l.view(3, -1).to(DEV), x.to(DEV), mask.to(DEV)      # keep the device tensors alive across the call
    rc = lib.vq_ema_accumulate(xd.data_ptr(), z.data_ptr(), md.data_ptr(), 3, D, x.shape[2], K,
                               stats.data_ptr(), None, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    close(stats[:K * D].view(K, D), s_sum, rtol=1e-5, atol=1e-5)
    assert torch.equal(stats[K * D:].cpu(), s_elem)


# ------------------------------------------------------------------------------ raw C ABI
def test_raw_abi_error_behaviour_and_host_path(vq):
    lib = vq._lib.load()
    gen = torch.Generator().manual_seed(6)
    K, D, n, t = 512, 128, 6, 400
    code = torch.randn(K, D, generator=gen)
    x = torch.randn(n, D, t, generator=gen)
    xd, kd = x.to(DEV), code.to(DEV)
    idx = torch.empty(n, t, dtype=torch.int64, device=DEV)
    ws = torch.empty(int(lib.vq_workspace_bytes(n, t, K, D)), dtype=torch.uint8, device=DEV)
    s = torch.cuda.current_stream().cuda_stream
    assert lib.vq_assign(xd.data_ptr(), n, D, t, kd.data_ptr(), K, idx.data_ptr(), None, None, ws.data_ptr(), 1024, 0, s) != 0
    assert b"workspace too small" in lib.vq_last_error()
    assert lib.vq_assign(xd.data_ptr(), n, D, t, kd.data_ptr(), K, idx.data_ptr(), None, None, ws.data_ptr(), ws.numel(), 7, s) != 0
    assert lib.vq_assign(xd.data_ptr(), n, D, t, kd.data_ptr(), K, idx.data_ptr(), None, None, ws.data_ptr(), ws.numel(), 0, s) == 0
    rows, _, _ = O.flatten_nct(x, torch.ones(n, 1, t))
    o_l, _, o_min = O.assign(rows, code)
    check_indices(rows, code, o_l, idx)
    # host-buffer path: pageable numpy in, numpy out
    ctx = lib.vq_host_ctx_create(0, n * t, K, D)
    assert ctx, lib.vq_last_error()
    try:
        kh = code.numpy()
        assert lib.vq_host_ctx_set_codebook(ctx, kh.ctypes.data) == 0
        xh = x.numpy()
        out = np.empty((n, t), dtype=np.int64)
        total = ctypes.c_double(0.0)
        assert lib.vq_encode_host(ctx, xh.ctypes.data, n, t, out.ctypes.data, ctypes.byref(total)) == 0, lib.vq_last_error()
        check_indices(rows, code, o_l, torch.from_numpy(out))
        assert abs(total.value - float(o_min.double().sum())) <= 2e-5 * abs(total.value)
        assert lib.vq_encode_host(ctx, xh.ctypes.data, n + 1, t, out.ctypes.data, None) != 0     # larger than the context
    finally:
        lib.vq_host_ctx_destroy(ctx)


def test_prepared_codebook_cache_follows_the_codebook(vq):
    """The module prepares a frozen codebook once (VQ_ALGO_PREPARED afterwards); in-place edits, a new tensor object or a
    second module sharing the workspace must all trigger a fresh preparation."""
    gen = torch.Generator().manual_seed(21)
    K, D = 256, 64
    x = torch.randn(3, D, 128, generator=gen)
    mask = torch.ones(3, 1, 128)
    rows, _, _ = O.flatten_nct(x, mask)
    xd, md = x.to(DEV), mask.to(DEV)

    def oracle_idx(code):
        return O.assign(rows, code)[0]

    a = vq.BottleneckBlock(K, D, 0.99, 1.0).to(DEV)
    code_a = torch.randn(K, D, generator=gen)
    a.k = code_a.to(DEV)
    for _ in range(3):                                   # 2nd and 3rd call reuse the prepared operands
        check_indices(rows, code_a, oracle_idx(code_a), a.encode(xd, md))
    a.k.mul_(-1.0)                                       # in-place edit: version counter moves
    check_indices(rows, -code_a, oracle_idx(-code_a), a.encode(xd, md))
    b = vq.BottleneckBlock(K, D, 0.99, 1.0).to(DEV)      # another module, same device workspace
    code_b = torch.randn(K, D, generator=gen)
    b.k = code_b.to(DEV)
    check_indices(rows, code_b, oracle_idx(code_b), b.encode(xd, md))
    check_indices(rows, -code_a, oracle_idx(-code_a), a.encode(xd, md))     # a's tag is stale now: prepared again
    a.k = code_b.to(DEV) * 0.5                           # rebinding: a new tensor object
    check_indices(rows, code_b * 0.5, oracle_idx(code_b * 0.5), a.encode(xd, md))
    # raw ABI: the flag on a prepared workspace gives the same indices as a full call
    lib = vq._lib.load()
    kd = (code_b * 0.5).to(DEV)
    ws = torch.empty(int(lib.vq_workspace_bytes(3, 128, K, D)), dtype=torch.uint8, device=DEV)
    i0 = torch.empty(3, 128, dtype=torch.int64, device=DEV)
    i1 = torch.empty_like(i0)
    s = torch.cuda.current_stream().cuda_stream
    assert lib.vq_assign(xd.data_ptr(), 3, D, 128, kd.data_ptr(), K, i0.data_ptr(), None, None, ws.data_ptr(), ws.numel(), 0, s) == 0
    assert lib.vq_assign(xd.data_ptr(), 3, D, 128, kd.data_ptr(), K, i1.data_ptr(), None, None, ws.data_ptr(), ws.numel(), 256, s) == 0
    assert torch.equal(i0, i1)


def test_device_side_restart_rows(vq):
    """rng_parity=False: no host sync; restart rows are valid batch rows (or jittered copies when there are too few),
    everything that does not depend on the random stream still matches the oracle."""
    gen = torch.Generator().manual_seed(33)
    K, D = 64, 16
    code = torch.randn(K, D, generator=gen)
    lengths = torch.tensor([120, 77, 10])
    x, mask = O.synthetic_batch(lengths, D, gen, codebook=code)
    rows, _, valid = O.flatten_nct(x, mask)
    blk = vq.BottleneckBlock(K, D, 0.99, 50.0, rng_parity=False).to(DEV)        # threshold 50: every code is re-seeded
    blk.k, blk.k_sum, blk.k_elem, blk.init = code.to(DEV), code.to(DEV).clone(), torch.ones(K, device=DEV), True
    blk.train()
    x_l, x_q, commit, metrics = blk(x.to(DEV), mask.to(DEV))
    st = O.CodebookState(K, D, 0.99, 50.0, code.clone(), code.clone(), torch.ones(K), True)
    o_l, o_q, o_commit, o_m = O.forward(st, x, mask, update_k=True, k_rand=torch.zeros(K, D))
    check_indices(rows, code, o_l, x_l)
    close(commit, o_commit)
    close(blk.k_sum, st.k_sum, rtol=1e-5, atol=1e-5)
    assert float(metrics["usage"]) == 0.0
    dmin = (blk.k.cpu()[:, None, :] - rows[valid][None, :, :]).abs().amax(dim=2).min(dim=1).values
    assert float(dmin.max()) == 0.0                                             # every new code IS a valid batch row
    fresh = vq.BottleneckBlock(K, D, 0.99, 1.0, rng_parity=False).to(DEV)       # init path with fewer rows than codes
    fresh.train()
    tiny_x, tiny_m = x[:, :, :8].contiguous(), mask[:, :, :8].contiguous()
    fresh(tiny_x.to(DEV), tiny_m.to(DEV))
    assert fresh.init and fresh.k.shape == (K, D) and bool(torch.isfinite(fresh.k).all())


def test_three_accumulator_stage_mode(vq, monkeypatch):
    """The two accumulator layouts of the tcgen05 kernel -- VQ_K1_STAGES=3 (three N = 128 stages, constant operand of the
    folded k-step read from shared memory through a stride-0 descriptor; default at D = 128) and VQ_K1_STAGES=2 (two stages
    filled by N = 256 MMAs) -- must give the same indices as each other and as the oracle."""
    gen = torch.Generator().manual_seed(5)
    K, D = 512, 128
    code = torch.randn(K, D, generator=gen)
    lengths = torch.tensor([512, 300, 256, 131])
    x, mask = O.synthetic_batch(lengths, D, gen, codebook=code)
    x = x[:, :, :512].contiguous()
    mask = mask[:, :, :512].contiguous()
    rows, _, _ = O.flatten_nct(x, mask)
    o_l = O.assign(rows, code)[0]
    xd, kd = x.to(DEV), code.to(DEV)
    monkeypatch.setenv("VQ_K1_STAGES", "2")
    base = vq.assign(xd, kd, algo="tc")[0].cpu()
    monkeypatch.setenv("VQ_K1_STAGES", "3")
    three = vq.assign(xd, kd, algo="tc")[0].cpu()
    monkeypatch.delenv("VQ_K1_STAGES")
    assert torch.equal(base, three)
    check_indices(rows, code, o_l, three)


def test_restart_rows_device_entry_point(vq):
    """vq_restart_rows_device: every drawn row is a valid frame of the batch; all valid frames are reachable; fewer valid
    frames than codes -> jittered copies; no valid frame -> zeros; a NULL mask means every frame is valid."""
    from importlib import import_module
    q = import_module(vq.__name__ + ".quantizer")
    gen = torch.Generator().manual_seed(3)
    n, d, t, K = 5, 24, 96, 160
    x = torch.randn(n, d, t, generator=gen)
    lengths = torch.tensor([96, 50, 1, 0, 33])
    mask = (torch.arange(t)[None, :] < lengths[:, None]).float().view(n, 1, t)
    mask[0, 0, 10:20] = 0.0                                   # holes: the mask is not a prefix mask
    rows, _, valid = O.flatten_nct(x, mask)
    vrows = rows[valid]                                       # 170 valid frames >= K: no jitter, draws are exact copies
    counter = torch.zeros(1, dtype=torch.int64, device=DEV)
    hits = torch.zeros(vrows.shape[0])
    outs = []
    for _ in range(24):
        out = q.restart_rows_device(x.to(DEV), mask.to(DEV), K, counter).cpu()
        match = (out[:, None, :] == vrows[None, :, :]).all(dim=2)    # [K, M]
        assert bool(match.any(dim=1).all())                          # every draw is a valid frame, bit for bit
        hits += match.float().sum(dim=0)
        outs.append(out)
    assert int(counter) == 24
    assert float(hits.min()) > 0                                     # 3840 draws over 170 frames: each one is hit
    assert float(hits.max()) < 3.0 * 24 * K / vrows.shape[0]         # roughly uniform
    assert not torch.equal(outs[0], outs[1])                         # the counter moved the stream
    few = torch.zeros_like(mask)
    few[1, 0, :3] = 1.0                                              # 3 valid frames < 64 codes: jittered copies
    j = q.restart_rows_device(x.to(DEV), few.to(DEV), 64, counter).cpu()
    base = O.flatten_nct(x, few)[0][O.flatten_nct(x, few)[2]]
    dist = (j[:, None, :] - base[None, :, :]).abs().amax(dim=2).min(dim=1).values
    assert float(dist.max()) < 0.05 and float(dist.min()) > 0.0
    none = q.restart_rows_device(x.to(DEV), torch.zeros_like(mask).to(DEV), 8, counter).cpu()
    assert float(none.abs().max()) == 0.0
    allv = q.restart_rows_device(x.to(DEV), None, 256, counter).cpu()
    assert bool((allv[:, None, :] == rows[None, :, :]).all(dim=2).any(dim=1).all())


def test_ema_accumulate_many_tiles_per_block(vq):
    """More than 256 valid 64-frame tiles per thread block: the accumulate kernel compacts its valid tiles into a
    shared-memory list of 256 entries and must walk through several lists (incl. a ragged last one and padded tiles)."""
    lib = vq._lib.load()
    gen = torch.Generator().manual_seed(11)
    n, d, t, K = 304, 128, 4096, 64                       # 304 * 64 tiles = 19456 > 74 blocks * 256
    x = torch.randn(n, d, t, generator=gen, dtype=torch.float32).to(DEV)
    idx = torch.randint(0, K, (n, t), generator=gen).to(DEV)
    lengths = torch.randint(1, t + 1, (n,), generator=gen)
    lengths[::7] = t
    lengths[3::11] = 0                                    # whole utterances of padding
    mask = (torch.arange(t)[None, :] < lengths[:, None]).float().to(DEV)
    stats = torch.zeros(K * d + K, device=DEV)
    scratch = torch.empty(n * ((t + 63) // 64), dtype=torch.uint8, device=DEV)
    rc = lib.vq_ema_accumulate(x.data_ptr(), idx.data_ptr(), mask.data_ptr(), n, d, t, K, stats.data_ptr(), scratch.data_ptr(),
                               torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.vq_last_error()
    sel = mask.reshape(-1) != 0
    flat = x.permute(0, 2, 1).reshape(-1, d)
    ref = torch.zeros(K, d, device=DEV, dtype=torch.float64).index_add_(0, idx.reshape(-1)[sel], flat[sel].double())
    cnt = torch.bincount(idx.reshape(-1)[sel], minlength=K).float()
    assert torch.equal(stats[K * d:], cnt)
    close(stats[:K * d].view(K, d), ref.float(), rtol=2e-4, atol=2e-2)     # sums of ~10^4 N(0,1) terms in FP32


def test_training_forward_in_a_cuda_graph(vq):
    """rng_parity=False makes the training forward free of host synchronisation: it can be captured once and replayed;
    every replay updates the codebook like an eager call on the same input would."""
    gen = torch.Generator().manual_seed(21)
    K, D = 128, 64
    code = torch.randn(K, D, generator=gen)
    lengths = torch.tensor([256, 200, 131, 64])
    x, mask = O.synthetic_batch(lengths, D, gen, codebook=code)
    xd, md = x.to(DEV), mask.to(DEV)

    def fresh():
        blk = vq.BottleneckBlock(K, D, 0.99, 1.0, rng_parity=False).to(DEV)
        blk.k, blk.k_sum, blk.k_elem, blk.init = code.to(DEV).clone(), code.to(DEV).clone(), torch.ones(K, device=DEV), True
        blk.train()
        return blk

    eager = fresh()
    with torch.no_grad():
        e_l, e_q, e_commit, e_metrics = eager(xd, md, update_k=True)
    blk = fresh()
    static_k = blk.k
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        warm = fresh()
        for _ in range(2):
            warm(xd, md, update_k=True)                   # allocates the cached workspaces outside the capture
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(graph):
        g_l, g_q, g_commit, g_metrics = blk(xd, md, update_k=True)
    blk.k = static_k                                      # replay reads the codebook it was captured with
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(g_l.cpu(), e_l.cpu())
    assert torch.equal(g_q.cpu(), e_q.cpu())
    close(g_commit, e_commit)
    assert float(g_metrics["usage"]) == float(e_metrics["usage"])
    assert int(g_metrics["used_curr"]) == int(e_metrics["used_curr"])
    close(g_metrics["entropy"], e_metrics["entropy"])
