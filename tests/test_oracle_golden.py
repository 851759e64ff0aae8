"""The CPU oracle (oracle/vq_oracle.py) against the vectors the unmodified reference produced
(tests/golden/*.npz, written by tests/golden/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import vq_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TRAIN_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "train_*.npz"))) + ["eval_k128_d64"]
INIT_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "init_*.npz")))
SAFE_GAP = 1e-3   # rows whose two best fp32 distances differ by more than this must match bit-exactly


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def T(a):
    return torch.from_numpy(np.asarray(a))


def state_from(g):
    K, D = g["k0"].shape
    return O.CodebookState(K, D, float(g["mu"]), float(g["threshold"]), T(g["k0"]).clone(),
                           T(g["k_sum0"]).clone(), T(g["k_elem0"]).clone(), True)


def assert_indices(g, got, want):
    got, want = got.reshape(-1), want.reshape(-1)
    safe = g["gap"] > SAFE_GAP if "gap" in g else np.ones_like(want, bool)
    assert np.array_equal(got[safe], want[safe])
    return got == want


def close(a, b, rtol=1e-6, atol=1e-7):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_forward_matches_reference(name):
    torch.set_num_threads(1)
    g = load(name)
    st = state_from(g)
    x = T(g["x"]).clone().requires_grad_(True)
    mask = T(g["mask"])
    torch.manual_seed(int(g["rng_seed"]))
    x_l, x_q, commit, metrics = O.forward(st, x, mask, update_k=bool(g["update_k"]))
    same = assert_indices(g, x_l.numpy(), g["x_l"])
    assert x_l.dtype == torch.int64 and tuple(x_l.shape) == g["x_l"].shape
    if same.all():
        assert np.array_equal(x_q.detach().numpy(), g["x_q"])            # bit-exact straight-through latents
        close(commit.item(), g["commit"])
        ((T(g["grad_w"]) * x_q).sum() + float(g["grad_commit"]) * commit).backward()
        close(x.grad.numpy(), g["grad_x"], rtol=1e-6, atol=1e-8)
        # closed-form backward used by the CUDA path
        gx = O.backward_wrt_x(x.detach(), mask, x_l, T(g["k0"]), T(g["grad_w"]), float(g["grad_commit"]))
        close(gx.numpy(), g["grad_x"], rtol=1e-5, atol=1e-7)
        if bool(g["update_k"]):
            close(st.k.numpy(), g["k1"], rtol=1e-6, atol=1e-7)
            close(st.k_sum.numpy(), g["k_sum1"], rtol=1e-6, atol=1e-7)
            close(st.k_elem.numpy(), g["k_elem1"], rtol=1e-6, atol=1e-7)
    want_keys = {k[7:] for k in g if k.startswith("metric_")}
    assert set(metrics) == want_keys
    for key in want_keys:
        if same.all():
            close(metrics[key].item(), g["metric_" + key], rtol=1e-5)
    if "used_curr" in metrics:
        assert metrics["used_curr"].dtype == torch.int64


def test_faithful_fit_equals_identity():
    g = load("train_k512_d128")
    st = state_from(g)
    rows, mcol, _ = O.flatten_nct(T(g["x"]), T(g["mask"]))
    _, fit_a, _ = O.assign(rows, st.k, mcol, faithful_fit=True)
    _, fit_b, _ = O.assign(rows, st.k, mcol, faithful_fit=False)
    close(fit_a.item(), fit_b.item(), rtol=1e-5)
    close(fit_a.item(), g["metric_fit"], rtol=1e-6)


@pytest.mark.parametrize("name", INIT_CASES)
def test_init_path_replays_rng(name):
    torch.set_num_threads(1)
    g = load(name)
    K, D = g["k1"].shape
    st = O.CodebookState(K, D, float(g["mu"]), float(g["threshold"]))
    torch.manual_seed(int(g["rng_seed"]))
    x_l, x_q, commit, metrics = O.forward(st, T(g["x"]), T(g["mask"]), update_k=True)
    assert np.array_equal(x_l.numpy(), g["x_l"])
    assert np.array_equal(x_q.numpy(), g["x_q"])
    close(st.k.numpy(), g["k1"], rtol=1e-6, atol=1e-7)
    close(st.k_elem.numpy(), g["k_elem1"], rtol=1e-6, atol=1e-7)
    close(commit.item(), g["commit"])
    for key in ("fit", "entropy", "used_curr", "usage", "dk"):
        close(metrics[key].item(), g["metric_" + key], rtol=1e-5)


def test_encode_decode_and_ties():
    g = load("encode_k512_d128")
    K, D = g["k0"].shape
    st = O.CodebookState(K, D, k=T(g["k0"]), init=True)
    z = O.encode(st, T(g["x"]), T(g["mask"]))
    assert_indices(g, z.numpy(), g["z"])
    assert np.array_equal(O.decode(st, T(g["z"])).numpy(), g["x_dec"])
    # duplicated codes: the lowest index wins, as torch.min does in the reference
    t = load("train_ties_k32_d8")
    assert (t["x_l"] % 2 == 0).all()


def test_audit_flags_real_errors_and_accepts_near_ties():
    gen = torch.Generator().manual_seed(0)
    k = torch.randn(64, 16, generator=gen)
    rows = torch.randn(500, 16, generator=gen)
    idx, _, _ = O.assign(rows, k)
    rep = O.audit_indices(rows, k, idx, idx)
    assert rep["mismatches"] == 0 and rep["match"] == 1.0
    bad = idx.clone()
    bad[:5] = (bad[:5] + 1) % 64
    rep = O.audit_indices(rows, k, idx, bad)
    assert rep["mismatches"] == 5 and rep["errors"] == 5
    k2 = k.clone()
    k2[1] = k2[0]                                       # exact tie between code 0 and 1
    idx2, _, _ = O.assign(rows, k2)
    swapped = torch.where(idx2 == 0, torch.ones_like(idx2), idx2)
    rep = O.audit_indices(rows, k2, idx2, swapped)
    assert rep["errors"] == 0 and rep["near_ties"] == rep["mismatches"]


def test_synthetic_lengths_look_like_ljspeech():
    gen = torch.Generator().manual_seed(0)
    lens = O.ljspeech_like_lengths(13100, gen)
    assert int(lens.min()) >= 188 and int(lens.max()) <= 1736 and (lens % 4 == 0).all()
    assert 1050 < float(lens.float().mean()) < 1200
    x, mask = O.synthetic_batch(lens[:3], 8, gen)
    assert x.shape == (3, 8, int(lens[:3].max())) and mask.shape == (3, 1, x.shape[2])
    assert torch.all(x[0, :, int(lens[0]):] == 0.25)


GROUPED_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "grouped_*.npz")))


@pytest.mark.parametrize("name", GROUPED_CASES)
def test_grouped_forward_matches_reference(name):
    """models/vqtts/bottleneck.py (phoneme-conditioned quantiser) restated in oracle.grouped_forward."""
    torch.set_num_threads(1)
    g = load(name)
    st = state_from(g)
    y = T(g["y_enc"]).clone().requires_grad_(True)
    training = bool(g["training"])
    torch.manual_seed(int(g["rng_seed"]))
    q_rel, y_d, commit, metrics = O.grouped_forward(st, y, T(g["x_id"]), T(g["attn"]), int(g["n_vocab"]), int(g["l_bins"]),
                                                    training=training, update_k=training)
    assert q_rel.dtype == torch.int64 and np.array_equal(q_rel.numpy(), g["q_rel"])
    assert np.array_equal(y_d.detach().contiguous().numpy(), g["y_d"])
    close(commit.item(), g["commit"])
    ((T(g["grad_w"]) * y_d).sum() + float(g["grad_commit"]) * commit).backward()
    close(y.grad.numpy(), g["grad_y"], rtol=1e-6, atol=1e-8)
    assert set(metrics) == {k[7:] for k in g if k.startswith("metric_")}
    for key, val in metrics.items():
        close(float(val), g["metric_" + key], rtol=1e-6, atol=1e-7)
    if training:
        close(st.k.numpy(), g["k1"], rtol=1e-6, atol=1e-7)
        close(st.k_sum.numpy(), g["k_sum1"], rtol=1e-6, atol=1e-7)
        close(st.k_elem.numpy(), g["k_elem1"], rtol=1e-6, atol=1e-7)
