"""GPU parity at the sizes and corners round 1 left unproven (run with ``-m gpu``):

* every row of the bench batch (256 LJSpeech-like utterances, 444 k rows) against the oracle, on clustered AND i.i.d.
  Gaussian latents -- the late tiles of the persistent tcgen05 kernel are checked row for row, not by properties;
* the corners of the codebook sweep (K up to 65 536, D 64..512) against the ORACLE (not the repo's own SIMT kernel);
* the shapes the advisor found (K = 3584 / 4096 with algo='auto', D = 512 with K <= 128);
* the one collective of the path on real NCCL: a 2-rank training forward vs the single-process oracle.

Bars as in test_gpu_parity.py: indices bit-exact except near-ties (fp64 gap <= 2^-18 (||x||^2 + max||e||^2), counted)."""
import os
import socket

import pytest
import torch

from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def vq():
    import __graft_entry__ as ge
    ge.build()
    import vqb200
    assert torch.cuda.is_available(), "GPU tests need a GPU"
    return vqb200


def audit(x, code, idx, chunk=8192):
    """All rows of the NCT batch x against the chunked oracle; returns the audit report (errors must be 0)."""
    n, d, t = x.shape
    rows = x.permute(0, 2, 1).reshape(-1, d)
    o_l, _ = O.assign_chunked(rows, code, chunk)
    rep = O.audit_indices(rows, code, o_l, idx.cpu().reshape(-1))
    assert rep["rows"] == n * t
    assert rep["errors"] == 0, rep
    assert rep["match"] >= 0.9999, rep
    return rep


@pytest.mark.parametrize("clustered", [True, False], ids=["clustered", "gaussian"])
def test_every_row_of_the_bench_batch(vq, clustered):
    gen = torch.Generator().manual_seed(0)
    K, D, n = 512, 128, 256
    code = torch.randn(K, D, generator=gen)
    lengths = O.ljspeech_like_lengths(n, gen)
    x, mask = O.synthetic_batch(lengths, D, gen, codebook=code if clustered else None)
    scalars = torch.zeros(16, dtype=torch.float64, device=DEV)
    idx, _ = vq.assign(x.to(DEV), code.to(DEV), algo="tc", scalars=scalars)
    rep = audit(x, code, idx)
    unsafe = float(scalars[vq._lib.S_UNSAFE_ROWS])
    assert (unsafe == 0) if clustered else (0 < unsafe < 0.05 * rep["rows"])    # the exact re-scan really ran on Gaussian data
    # ... and the SIMT kernel on the same batch
    idx_s, _ = vq.assign(x.to(DEV), code.to(DEV), algo="simt")
    audit(x, code, idx_s)


SWEEP_CORNERS = [  # (K, D, rows)
    (65536, 64, 32768),
    (65536, 512, 32768),
    (8192, 128, 32768),
    (512, 512, 32768),
    (4096, 128, 8192),       # algo='auto' used to fail with "shared memory budget exceeded" for K in 3457..4096
    (3584, 128, 8192),
    (128, 512, 8192),        # D > 448 with a resident codebook: the folded step's constant slice does not fit TMEM
    (2048, 256, 16384),
]


@pytest.mark.parametrize("clustered", [True, False], ids=["clustered", "gaussian"])
@pytest.mark.parametrize("corner", SWEEP_CORNERS, ids=lambda c: f"K{c[0]}_D{c[1]}")
def test_sweep_corner_against_oracle(vq, corner, clustered):
    K, D, rows = corner
    gen = torch.Generator().manual_seed(K + D)
    code = torch.randn(K, D, generator=gen)
    n, t = 8, rows // 8
    lengths = torch.full((n,), t)
    x, _ = O.synthetic_batch(lengths, D, gen, codebook=code if clustered else None)
    for algo in ("auto", "tc"):
        idx, _ = vq.assign(x.to(DEV), code.to(DEV), algo=algo)
        audit(x, code, idx, chunk=1024 if K > 8192 else 8192)


@pytest.mark.parametrize("shape", [(6, 128, 1733, 512), (3, 128, 131, 512), (2, 64, 1001, 1024), (1, 256, 37, 8192), (4, 128, 2, 512)],
                         ids=lambda s: f"N{s[0]}_D{s[1]}_T{s[2]}_K{s[3]}")
@pytest.mark.parametrize("clustered", [True, False], ids=["clustered", "gaussian"])
def test_tcgen05_route_for_t_not_a_multiple_of_4(vq, shape, clustered):
    """T % 4 != 0 (VQTTS / vqlatent crops) used to drop the whole call to the CUDA-core kernel (TMA needs 16-byte global
    strides); the tcgen05 kernel now loads such tensors with cp.async.  algo='tc' must accept them and stay exact; a misaligned
    view (offset by one float) takes the same route."""
    n, D, t, K = shape
    gen = torch.Generator().manual_seed(t * 7 + K)
    code = torch.randn(K, D, generator=gen)
    lengths = torch.randint(max(1, t // 2), t + 1, (n,), generator=gen)
    lengths[0] = t
    x, _ = O.synthetic_batch(lengths, D, gen, codebook=code if clustered else None)
    idx, _ = vq.assign(x.to(DEV), code.to(DEV), algo="tc")
    audit(x, code, idx)
    if t % 4 == 1:                                            # an aligned T with a misaligned base pointer
        flat = torch.empty(n * D * (t + 3) + 1, device=DEV)
        view = flat[1:1 + n * D * (t + 3)].view(n, D, t + 3)
        xp = torch.nn.functional.pad(x, (0, 3))
        view.copy_(xp.to(DEV))
        assert view.data_ptr() % 16 != 0
        idx2, _ = vq.assign(view, code.to(DEV), algo="tc")
        audit(xp, code, idx2)


@pytest.mark.parametrize("shape", [(8, 128, 1024, 512, True), (8, 128, 1024, 512, False), (3, 64, 136, 1024, False), (2, 256, 72, 8192, True),
                                   (2, 128, 37, 512, False)],
                         ids=lambda s: f"N{s[0]}_D{s[1]}_T{s[2]}_K{s[3]}_{'c' if s[4] else 'g'}")
def test_bf16_latents_without_upcast(vq, shape):
    """vq_assign_bf16: bf16 latents straight into K1 (tcgen05 route when T % 8 == 0, the exact CUDA-core kernel otherwise) must
    give the indices of the oracle run on x.float() -- the reference semantics for a bf16-autocast encoder."""
    n, D, t, K, clustered = shape
    gen = torch.Generator().manual_seed(t + K)
    code = torch.randn(K, D, generator=gen)
    lengths = torch.full((n,), t)
    x, mask = O.synthetic_batch(lengths, D, gen, codebook=code if clustered else None)
    xb = x.to(torch.bfloat16)
    for algo in ("auto", "simt"):
        idx, min_d = vq.assign(xb.to(DEV), code.to(DEV), algo=algo, want_min_d=True)
        rep = audit(xb.float(), code, idx)
        rows = xb.float().permute(0, 2, 1).reshape(-1, D)
        o_l, o_min = O.assign_chunked(rows, code)
        if rep["mismatches"] == 0:
            assert torch.allclose(min_d.cpu().reshape(-1), o_min, rtol=2e-5, atol=2e-4)
    # the module's encode takes bf16 as it is
    blk = vq.BottleneckBlock(K, D, 0.99, 1.0).to(DEV)
    blk.k, blk.init = code.to(DEV), True
    z = blk.encode(xb.to(DEV), mask.to(DEV))
    audit(xb.float(), code, z)


def test_second_device_gets_its_shared_memory_opt_in(vq):
    """cudaFuncSetAttribute is per device: the same process must be able to run the kernels on cuda:1 after cuda:0."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    gen = torch.Generator().manual_seed(1)
    K, D = 512, 128
    code = torch.randn(K, D, generator=gen)
    x, mask = O.synthetic_batch(torch.tensor([256, 128]), D, gen, codebook=code)
    outs = []
    for dev in (torch.device("cuda:0"), torch.device("cuda:1")):
        blk = vq.BottleneckBlock(K, D, 0.99, 1.0).to(dev)
        blk.k, blk.k_sum, blk.k_elem, blk.init = code.to(dev), code.to(dev).clone(), torch.ones(K, device=dev), True
        blk.train()
        torch.manual_seed(0)
        x_l, x_q, commit, metrics = blk(x.to(dev), mask.to(dev))
        outs.append((x_l.cpu(), x_q.cpu(), blk.k.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.allclose(outs[0][2], outs[1][2], rtol=1e-6, atol=1e-7)


def test_two_streams_do_not_share_a_workspace(vq):
    """Two streams running K1 concurrently on adversarial data (both use the fallback worklist) must both be exact."""
    gen = torch.Generator().manual_seed(8)
    K, D = 512, 128
    code = torch.randn(K, D, generator=gen)
    xs = [O.synthetic_batch(torch.full((16,), 1024), D, gen)[0] for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    kd = code.to(DEV)
    xd = [x.to(DEV) for x in xs]
    torch.cuda.synchronize()
    outs = [None, None]
    for _ in range(3):
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                outs[i] = vq.assign(xd[i], kd, algo="tc")[0]
    torch.cuda.synchronize()
    for i in range(2):
        audit(xs[i], code, outs[i])


@pytest.mark.parametrize("shape", [(64, 128, 512, True), (5, 96, 300, False), (3, 32, 40, True), (2, 128, 512, False), (3, 200, 96, True)],
                         ids=lambda s: f"N{s[0]}_D{s[1]}_K{s[2]}")
def test_fused_forward_ema_matches_separate_kernels(vq, shape):
    """vq_gather_st_fwd_ema (K2 + K3a in one pass over x, per-code sums in a shared-memory slab) against the two separate kernels
    and against the oracle's one-hot GEMM statistics: x_q bit-equal, counts exact, sums 1e-5."""
    n, D, K, clustered = shape
    gen = torch.Generator().manual_seed(n * 1000 + D)
    code = torch.randn(K, D, generator=gen)
    lengths = O.ljspeech_like_lengths(n, gen) // 4 * 4
    lengths[-1] = 8                                               # a nearly empty utterance: many fully padded tiles
    x, mask = O.synthetic_batch(lengths, D, gen, codebook=code if clustered else None)
    xd, md, kd = x.to(DEV), mask.to(DEV), code.to(DEV)
    outs = []
    for fuse in (True, False):
        blk = vq.BottleneckBlock(K, D, 0.99, 1.0).to(DEV)
        blk.fuse_ema = fuse
        blk.k, blk.k_sum, blk.k_elem, blk.init = kd.clone(), kd.clone() * 2, torch.full((K,), 2.0, device=DEV), True
        blk.train()
        assert blk._fuse_ema_ok(xd) == fuse
        torch.manual_seed(3)
        x_l, x_q, commit, metrics = blk(xd, md, update_k=True)
        outs.append((x_l.cpu(), x_q.cpu(), float(commit), blk.k.cpu(), blk.k_sum.cpu(), blk.k_elem.cpu(),
                     {k: float(v) for k, v in metrics.items()}))
    a, b = outs
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and a[2] == b[2]
    assert torch.equal(a[5], b[5])                                                   # counts: exact
    assert torch.allclose(a[4], b[4], rtol=1e-5, atol=1e-5) and torch.allclose(a[3], b[3], rtol=1e-5, atol=1e-5)
    for key in a[6]:
        assert abs(a[6][key] - b[6][key]) <= 1e-5 * abs(b[6][key]) + 1e-6, key
    # and against the oracle statistics (bottleneck.py:64-68) through the raw entry point
    rows, _, valid = O.flatten_nct(x, mask)
    s_sum, s_elem = O.local_statistics(rows[valid], a[0].reshape(-1)[valid], K)
    lib = vq._lib.load()
    stats = torch.zeros(K * D + K, device=DEV)
    x_q = torch.empty_like(xd)
    scalars = torch.zeros(16, dtype=torch.float64, device=DEV)
    res = torch.zeros(8, device=DEV)
    idx = a[0].to(DEV)
    rc = lib.vq_gather_st_fwd_ema(xd.data_ptr(), idx.data_ptr(), md.data_ptr(), kd.data_ptr(), n, D, x.shape[2], K, x_q.data_ptr(),
                                  scalars.data_ptr(), res.data_ptr(), stats.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.vq_last_error()
    assert torch.equal(stats[K * D:].cpu(), s_elem)
    assert torch.allclose(stats[:K * D].view(K, D).cpu(), s_sum, rtol=1e-5, atol=2e-5)
    assert torch.equal(x_q.cpu(), a[1])
    # shapes the fused kernel does not take are refused, not mangled
    assert lib.vq_gather_st_fwd_ema_supported(128, 1024, 513) == 0 and lib.vq_gather_st_fwd_ema_supported(513, 1024, 512) == 0
    assert lib.vq_gather_st_fwd_ema_supported(128, 1023, 512) == 0 and lib.vq_gather_st_fwd_ema_supported(128, 1024, 512) == 1


def test_laplace_smoothing_formula(vq):
    """Opt-in smoothing: k = k_sum / ((k_elem + eps) / (n + K eps) * n) with n = sum of the UPDATED k_elem."""
    gen = torch.Generator().manual_seed(2)
    K, D, eps = 64, 16, 0.5
    code = torch.randn(K, D, generator=gen)
    x, mask = O.synthetic_batch(torch.tensor([150, 99]), D, gen, codebook=code)
    blk = vq.BottleneckBlock(K, D, 0.99, 0.0, laplace_eps=eps).to(DEV)
    blk.k, blk.k_sum, blk.k_elem, blk.init = code.to(DEV), code.to(DEV).clone(), torch.ones(K, device=DEV), True
    blk.train()
    torch.manual_seed(0)
    blk(x.to(DEV), mask.to(DEV))
    st = O.CodebookState(K, D, 0.99, 0.0, code.clone(), code.clone(), torch.ones(K), True)
    torch.manual_seed(0)
    O.forward(st, x, mask, update_k=True)                         # k_sum / k_elem after the step (no smoothing in the reference)
    n = st.k_elem.double().sum()
    smoothed = (st.k_elem.double() + eps) / (n + K * eps) * n
    want = (st.k_sum.double() / smoothed.view(K, 1)).float()
    assert torch.allclose(blk.k_elem.cpu(), st.k_elem, rtol=1e-6, atol=1e-7)      # stored unsmoothed
    assert torch.allclose(blk.k.cpu(), want, rtol=1e-5, atol=1e-6)
    assert not torch.allclose(blk.k.cpu(), st.k, rtol=1e-3, atol=1e-4)            # and it really differs from the reference


# ------------------------------------------------------------------------------ the collective, on real NCCL
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, out, rng_parity, p2p):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["VQ_P2P"] = "1" if p2p else "0"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        import vqb200
        dev = torch.device("cuda", rank)
        x, mask, code = _nccl_batch()
        a, b = vqb200.dist.shard_range(x.shape[0], world, rank)
        K, D = code.shape
        blk = vqb200.BottleneckBlock(K, D, 0.99, 1.0, rng_parity=rng_parity).to(dev)
        blk.k, blk.k_sum, blk.k_elem, blk.init = code.to(dev), (code * 2).to(dev), torch.full((K,), 2.0, device=dev), True
        blk.train()
        torch.manual_seed(100 + rank)
        res = []
        for _ in range(2):                                    # two steps: the second runs on the all-reduced codebook
            x_l, x_q, commit, metrics = blk(x[a:b].to(dev), mask[a:b].to(dev), update_k=True)
            res.append(dict(x_l=x_l.cpu(), commit=commit.cpu(), metrics={k: v.cpu() for k, v in metrics.items()}))
        torch.cuda.synchronize()
        torch.save(dict(steps=res, k=blk.k.cpu(), k_sum=blk.k_sum.cpu(), k_elem=blk.k_elem.cpu(), used_p2p=bool(getattr(blk, "_peer", None))),
                   f"{out}.{rank}")
    finally:
        dist.destroy_process_group()


def _nccl_batch():
    gen = torch.Generator().manual_seed(77)
    K, D = 512, 128
    code = torch.randn(K, D, generator=gen)
    lengths = O.ljspeech_like_lengths(8, gen) // 2 // 4 * 4
    x, mask = O.synthetic_batch(lengths, D, gen, codebook=code)
    return x, mask, code


@pytest.mark.parametrize("p2p", [True, False], ids=["nvlink_peer_memory", "nccl_allreduce"])
@pytest.mark.parametrize("rng_parity", [True, False], ids=["rng_parity", "device_rng"])
def test_two_rank_training_forward_over_nccl(vq, tmp_path, rng_parity, p2p):
    """bottleneck.py:72-75 on hardware: 2 ranks, utterance-sharded batch; the statistics are exchanged over NVLink peer
    memory (csrc/k3_p2p.cuh) or by ONE NCCL all-reduce overlapped with K2.  Every rank must end with bit-identical
    k / k_sum / k_elem, equal (1e-5) to the single-process oracle on the whole batch (every code stays above the revival
    threshold, so the restart rows -- which come from rank 0's shard -- are unused)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "nccl")
    mp.spawn(_nccl_worker, args=(2, _free_port(), out, rng_parity, p2p), nprocs=2, join=True)
    x, mask, code = _nccl_batch()
    K, D = code.shape
    st = O.CodebookState(K, D, 0.99, 1.0, code.clone(), code.clone() * 2, torch.full((K,), 2.0), True)
    got = [torch.load(f"{out}.{r}") for r in range(2)]
    assert got[0]["used_p2p"] == got[1]["used_p2p"] and (got[0]["used_p2p"] or not p2p or os.environ.get("VQ_ALLOW_NO_P2P"))
    for name in ("k", "k_sum", "k_elem"):
        assert torch.equal(got[0][name], got[1][name]), name                    # replicas stay bit-identical
    rows, _, _ = O.flatten_nct(x, mask)
    for step in range(2):
        k_before = st.k.clone()
        o_l, _, _, o_m = O.forward(st, x, mask, update_k=True, k_rand=torch.zeros(K, D))
        both = torch.cat([got[0]["steps"][step]["x_l"], got[1]["steps"][step]["x_l"]], 0)
        rep = O.audit_indices(rows, k_before, o_l.reshape(-1), both.reshape(-1))
        assert rep["errors"] == 0, rep
        for r in range(2):
            m = got[r]["steps"][step]["metrics"]
            assert int(m["used_curr"]) == int(o_m["used_curr"]) and float(m["usage"]) == float(o_m["usage"])
            assert abs(float(m["entropy"]) - float(o_m["entropy"])) <= 1e-5 * abs(float(o_m["entropy"]))
            assert abs(float(m["dk"]) - float(o_m["dk"])) <= 1e-4 * abs(float(o_m["dk"])) + 1e-7
    assert float(st.k_elem.min()) >= 1.0                                         # no revival happened: k_rand was unused
    assert torch.allclose(got[0]["k_elem"], st.k_elem, rtol=1e-5, atol=1e-6)
    assert torch.allclose(got[0]["k_sum"], st.k_sum, rtol=1e-5, atol=1e-5)
    assert torch.allclose(got[0]["k"], st.k, rtol=1e-5, atol=1e-5)


def _revival_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        import vqb200
        dev = torch.device("cuda", rank)
        x, mask, code = _nccl_batch()
        a, b = vqb200.dist.shard_range(x.shape[0], world, rank)
        K, D = code.shape
        blk = vqb200.BottleneckBlock(K, D, 0.99, 1000.0).to(dev)          # threshold 1000: every code is re-seeded from k_rand
        blk.k, blk.k_sum, blk.k_elem, blk.init = code.to(dev), code.to(dev).clone(), torch.ones(K, device=dev), True
        blk.train()
        torch.manual_seed(7 + rank)                                       # different RNG streams: only rank 0's rows may win
        blk(x[a:b].to(dev), mask[a:b].to(dev), update_k=True)
        torch.cuda.synchronize()
        torch.save(dict(k=blk.k.cpu(), lo=a, hi=b), f"{out}.{rank}")
    finally:
        dist.destroy_process_group()


def test_two_rank_revival_uses_rank0_rows(vq, tmp_path):
    """bottleneck.py:73: the restart rows are rank 0's, on every rank."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "revive")
    mp.spawn(_revival_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = [torch.load(f"{out}.{r}") for r in range(2)]
    assert torch.equal(got[0]["k"], got[1]["k"])
    x, mask, code = _nccl_batch()
    rows0, _, valid0 = O.flatten_nct(x[got[0]["lo"]:got[0]["hi"]], mask[got[0]["lo"]:got[0]["hi"]])
    pool = rows0[valid0]
    dmin = (got[0]["k"][:64, None, :] - pool[None, :, :]).abs().amax(dim=2).min(dim=1).values
    assert float(dmin.max()) == 0.0                                       # every new code is a valid frame of RANK 0's shard
