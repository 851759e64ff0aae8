"""GPU parity of the phoneme-conditioned quantiser (models/vqtts/bottleneck.py): goldens produced by the
unmodified reference (b = 1, the only batch size it runs with) and the oracle restatement at the TTS config's codebook
shape (n_vocab = 149, l_bins = 512, D = 128 -> K = 76 288)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "grouped_*.npz")))


@pytest.fixture(scope="module")
def vq():
    import __graft_entry__ as ge
    ge.build()
    import vqb200
    return vqb200


def T(a, dev=None):
    t = torch.from_numpy(np.asarray(a))
    return t.to(dev) if dev is not None else t


def close(a, b, rtol=1e-5, atol=1e-7):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else a
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else b
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", CASES)
def test_grouped_forward_against_reference_golden(vq, name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    n_vocab, l_bins, training = int(g["n_vocab"]), int(g["l_bins"]), bool(g["training"])
    D = g["k0"].shape[1]
    blk = vq.GroupedBottleneck(n_vocab, l_bins, D, float(g["mu"]), float(g["threshold"])).to(DEV)
    blk.k, blk.k_sum, blk.k_elem, blk.init = T(g["k0"], DEV).clone(), T(g["k_sum0"], DEV).clone(), T(g["k_elem0"], DEV).clone(), True
    blk.train(training)
    y = T(g["y_enc"], DEV).clone().requires_grad_(True)
    torch.manual_seed(int(g["rng_seed"]))
    q_rel, y_d, commit, metrics = blk(y, T(g["x_id"], DEV), T(g["attn"], DEV), update_k=training)
    assert q_rel.dtype == torch.int64 and torch.equal(q_rel.cpu(), T(g["q_rel"]))
    assert torch.equal(y_d.detach().cpu(), T(g["y_d"]))                              # bit-exact straight-through latents
    close(commit, g["commit"])
    ((T(g["grad_w"], DEV) * y_d).sum() + float(g["grad_commit"]) * commit).backward()
    close(y.grad, g["grad_y"], rtol=1e-5, atol=1e-7)
    assert set(metrics) == {k[7:] for k in g if k.startswith("metric_")}
    close(metrics["fit"], g["metric_fit"], rtol=2e-5)
    if training:
        close(blk.k_sum, g["k_sum1"], rtol=1e-5, atol=1e-6)
        close(blk.k_elem, g["k_elem1"], rtol=1e-5, atol=1e-6)
        assert int(metrics["used_curr"]) == int(g["metric_used_curr"]) and float(metrics["usage"]) == float(g["metric_usage"])
        close(metrics["entropy"], g["metric_entropy"])
        alive = T(g["k_elem1"]) >= float(g["threshold"])                             # revived codes get randn jitter from another device's RNG
        close(blk.k.cpu()[alive], T(g["k1"])[alive], rtol=1e-5, atol=1e-6)


def _tts_batch(gen, b, tx, ty, n_vocab, l_bins, c, code):
    x_lens = torch.randint(tx // 2, tx + 1, (b,), generator=gen)
    y_lens = torch.randint(ty // 2, ty + 1, (b,), generator=gen)
    y_lens[0], x_lens[0] = ty, tx
    x_id = torch.randint(0, n_vocab, (b, tx), generator=gen)
    attn = torch.zeros(b, tx, ty)
    for i in range(b):
        cuts = torch.sort(torch.randperm(int(y_lens[i]) - 1, generator=gen)[:int(x_lens[i]) - 1] + 1).values.tolist()
        bounds = [0] + cuts + [int(y_lens[i])]
        for j in range(int(x_lens[i])):
            attn[i, j, bounds[j]:bounds[j + 1]] = 1.0
    tok = O.align_tokens(x_id, attn)
    rel = torch.randint(0, l_bins, (b, ty), generator=gen)
    y = code[tok * l_bins + rel].permute(0, 2, 1).contiguous() + 0.4 * torch.randn(b, c, ty, generator=gen)
    return y, x_id, attn


def test_grouped_tts_shape_against_oracle(vq):
    """n_vocab = 149, l_bins = 512, D = 128 (configs/models/vqtts.yaml scale): K = 76 288, batch of 3 utterances."""
    gen = torch.Generator().manual_seed(7)
    n_vocab, l_bins, c = 149, 512, 128
    K = n_vocab * l_bins
    code = torch.randn(K, c, generator=gen)
    y, x_id, attn = _tts_batch(gen, 3, 20, 160, n_vocab, l_bins, c, code)
    st = O.CodebookState(K, c, 0.99, 1.0, code.clone(), code.clone() * 3, torch.full((K,), 3.0), True)
    blk = vq.GroupedBottleneck(n_vocab, l_bins, c, 0.99, 1.0).to(DEV)
    blk.k, blk.k_sum, blk.k_elem, blk.init = code.to(DEV), (code * 3).to(DEV), torch.full((K,), 3.0, device=DEV), True
    blk.train()
    yg = y.to(DEV).requires_grad_(True)
    torch.manual_seed(5)
    q_rel, y_d, commit, metrics = blk(yg, x_id.to(DEV), attn.to(DEV))
    yo = y.clone().requires_grad_(True)
    torch.manual_seed(5)
    o_rel, o_d, o_commit, o_m = O.grouped_forward(st, yo, x_id, attn, n_vocab, l_bins, training=True)
    # audit the indices inside each frame's group with the fp64 near-tie rule
    tok = O.align_tokens(x_id, attn).reshape(-1)
    rows = y.permute(0, 2, 1).reshape(-1, c)
    rep = O.audit_indices(rows, code, (tok * l_bins + o_rel.reshape(-1)), (tok * l_bins + q_rel.cpu().reshape(-1)))
    assert rep["errors"] == 0, rep
    if rep["mismatches"] == 0:
        assert torch.equal(y_d.detach().cpu(), o_d.detach().contiguous())
        close(commit, o_commit)
        close(metrics["fit"], o_m["fit"], rtol=2e-5)
        close(blk.k_sum, st.k_sum, rtol=1e-5, atol=1e-5)
        close(blk.k_elem, st.k_elem, rtol=1e-6, atol=1e-6)
        assert int(metrics["used_curr"]) == int(o_m["used_curr"]) and float(metrics["usage"]) == float(o_m["usage"])


def test_grouped_raw_entry_point(vq):
    """vq_assign_grouped through the C ABI: relative + absolute indices, winning distances, odd shapes (T % 4 != 0, D % 4 != 0)."""
    lib = vq._lib.load()
    gen = torch.Generator().manual_seed(9)
    n_vocab, l_bins, c, n, t = 6, 40, 30, 3, 37
    K = n_vocab * l_bins
    code = torch.randn(K, c, generator=gen)
    x = torch.randn(n, c, t, generator=gen)
    tok = torch.randint(0, n_vocab, (n, t), generator=gen)
    xd, kd, td = x.to(DEV), code.to(DEV), tok.to(DEV)
    q_rel = torch.empty(n, t, dtype=torch.int64, device=DEV)
    q_abs = torch.empty_like(q_rel)
    min_d = torch.empty(n, t, device=DEV)
    scalars = torch.zeros(16, dtype=torch.float64, device=DEV)
    ws = torch.empty(int(lib.vq_workspace_bytes(n, t, K, c)), dtype=torch.uint8, device=DEV)
    rc = lib.vq_assign_grouped(xd.data_ptr(), n, c, t, kd.data_ptr(), n_vocab, l_bins, td.data_ptr(), q_rel.data_ptr(), q_abs.data_ptr(),
                               min_d.data_ptr(), scalars.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.vq_last_error()
    rows = x.permute(0, 2, 1).reshape(-1, c)
    kg = code.view(n_vocab, l_bins, c)[tok.reshape(-1)]
    dist = (rows.unsqueeze(1) ** 2).sum(-1) - 2 * torch.bmm(rows.unsqueeze(1), kg.transpose(1, 2)).squeeze(1) + (kg ** 2).sum(-1)
    o_min, o_rel = dist.min(-1)
    assert torch.equal(q_rel.cpu().reshape(-1), o_rel)
    assert torch.equal(q_abs.cpu().reshape(-1), tok.reshape(-1) * l_bins + o_rel)
    close(min_d.cpu().reshape(-1), o_min, rtol=1e-5, atol=1e-5)
    close(float(scalars[0]), float(o_min.double().sum()), rtol=1e-5)


def test_host_encoder_and_code_shard_round_trip(vq, tmp_path):
    """latents on the host -> ragged uint16 codes (vq_encode_host_u16) -> one binary shard -> what
    ``dump_batch_to_pickle`` would have pickled (q[:ql] of every utterance, scripts/generate_vq_dataset.py:86-87), on the
    golden encode case produced by the unmodified reference plus a multi-chunk LJSpeech-like batch vs the oracle."""
    z = np.load(os.path.join(GOLDEN, "encode_k512_d128.npz"))
    g = {k: z[k] for k in z.files}
    ql = g["mask"].sum(axis=(1, 2)).astype(np.int32)
    enc = vq.HostEncoder(0, 300 * 1800, g["k0"])
    codes, lengths = enc.encode(g["x"], ql)
    want = np.concatenate([g["z"][i, :ql[i]] for i in range(len(ql))])
    gap = g["gap"].reshape(g["z"].shape)
    safe = np.concatenate([gap[i, :ql[i]] for i in range(len(ql))]) > 1e-3
    assert codes.dtype == np.uint16 and codes.size == int(ql.sum()) and np.array_equal(codes[safe], want[safe])
    w = vq.CodeShardWriter(str(tmp_path / "val.vqb2"), vocab_size=512)
    w.append_batch(codes, lengths)
    w.close()
    shard = vq.CodeShard(str(tmp_path / "val.vqb2"))
    for i in range(len(ql)):
        assert np.array_equal(np.asarray(shard[i]["q"])[gap[i, :ql[i]] > 1e-3], g["z"][i, :ql[i]][gap[i, :ql[i]] > 1e-3])
    # a batch large enough for several chunks, filled in place in the pinned staging view
    gen = torch.Generator().manual_seed(3)
    code = T(g["k0"])
    lens = O.ljspeech_like_lengths(150, gen)
    x, mask = O.synthetic_batch(lens, 128, gen, codebook=code)
    stage = enc.x_staging(*[x.shape[0], x.shape[2]])
    stage[...] = x.numpy()
    codes, lengths = enc.encode(stage, lens.numpy().astype(np.int32))
    rows = x.permute(0, 2, 1).reshape(-1, 128)
    o_l, _ = O.assign_chunked(rows, code)
    o_l = o_l.view(x.shape[0], x.shape[2])
    want = torch.cat([o_l[i, :int(lens[i])] for i in range(len(lens))])
    valid_rows = torch.cat([rows.view(x.shape[0], x.shape[2], 128)[i, :int(lens[i])] for i in range(len(lens))])
    rep = O.audit_indices(valid_rows, code, want, torch.from_numpy(codes.astype(np.int64)))
    assert rep["rows"] == int(lens.sum()) and rep["errors"] == 0, rep
    enc.close()
