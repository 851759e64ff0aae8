"""K2 / K3 kernels alone on the bench workload (256 LJSpeech-like utterances, K=512, D=128): CUDA-event time of REPS
back-to-back calls (inputs 228 MB > L2, so every call streams from HBM).  `K23_ONLY=k3a` etc. restricts the run (ncu)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
import vqb200  # noqa: E402
from oracle import vq_oracle as O  # noqa: E402

lib = vqb200._lib.load()
dev = torch.device("cuda:0")
K, D, N_UTT, REPS = 512, 128, 256, int(os.environ.get("K23_REPS", "20"))
only = os.environ.get("K23_ONLY", "")


def main():
    gen = torch.Generator().manual_seed(0)
    code = torch.randn(K, D, generator=gen)
    lengths = O.ljspeech_like_lengths(N_UTT, gen)
    x, mask = O.synthetic_batch(lengths, D, gen, codebook=code)
    n, d, t = x.shape
    rows, valid = n * t, int(mask.sum())
    xd, kd, md = x.to(dev), code.to(dev), mask.to(dev)
    stream = torch.cuda.current_stream().cuda_stream
    idx = torch.empty(n, t, dtype=torch.int64, device=dev)
    ws = torch.empty(int(lib.vq_workspace_bytes(n, t, K, d)), dtype=torch.uint8, device=dev)
    assert lib.vq_assign(xd.data_ptr(), n, d, t, kd.data_ptr(), K, idx.data_ptr(), None, None, ws.data_ptr(), ws.numel(), 0, stream) == 0
    x_q = torch.empty_like(xd)
    g_in = torch.randn_like(xd)
    scalars = torch.zeros(16, dtype=torch.float64, device=dev)
    res = torch.zeros(8, device=dev)
    stats = torch.zeros(K * D + K, device=dev)
    g_commit = torch.ones((), device=dev)
    scratch = torch.empty(n * ((t + 63) // 64), dtype=torch.uint8, device=dev)
    k_sum, k_elem, k_new = kd.clone(), torch.ones(K, device=dev), torch.empty_like(kd)
    calls = {
        "k2_fwd": (lambda: lib.vq_gather_st_fwd(xd.data_ptr(), idx.data_ptr(), md.data_ptr(), kd.data_ptr(), n, d, t, K, x_q.data_ptr(),
                                                scalars.data_ptr(), res.data_ptr(), stream), rows * (8 * D + 12)),
        "k2_fused": (lambda: lib.vq_gather_st_fwd_ema(xd.data_ptr(), idx.data_ptr(), md.data_ptr(), kd.data_ptr(), n, d, t, K, x_q.data_ptr(),
                                                      scalars.data_ptr(), res.data_ptr(), stats.data_ptr(), stream), rows * (8 * D + 12) + 4 * K * (D + 1)),
        "k2_bwd": (lambda: lib.vq_gather_st_bwd(xd.data_ptr(), idx.data_ptr(), md.data_ptr(), kd.data_ptr(), g_in.data_ptr(),
                                                g_commit.data_ptr(), scalars.data_ptr(), n, d, t, K, x_q.data_ptr(), stream), rows * (12 * D + 12)),
        "k2_dec": (lambda: lib.vq_decode(idx.data_ptr(), kd.data_ptr(), n, d, t, K, x_q.data_ptr(), stream), rows * (4 * D + 8)),
        "k3a": (lambda: lib.vq_ema_accumulate(xd.data_ptr(), idx.data_ptr(), md.data_ptr(), n, d, t, K, stats.data_ptr(), scratch.data_ptr(), stream),
                valid * (4 * D + 8) + rows * 4 + 4 * K * (D + 1)),
        "k3b": (lambda: lib.vq_ema_finalize(stats.data_ptr(), kd.data_ptr(), kd.data_ptr(), k_new.data_ptr(), k_sum.data_ptr(), k_elem.data_ptr(),
                                            K, D, 0.99, 1.0, 0.0, scalars.data_ptr(), res.data_ptr(), None, stream), 4 * K * (6 * D + 3)),
    }
    out = {"rows": rows, "valid_frames": valid, "reps": REPS}
    for name, (fn, nbytes) in calls.items():
        if only and name not in only.split(","):
            continue
        for _ in range(3):
            assert fn() == 0, lib.vq_last_error()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(REPS):
                fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / REPS)
        out[name] = {"ms": round(best, 5), "GBps": round(nbytes / best / 1e6, 1), "frac_hbm_6544": round(nbytes / best / 1e6 / 6544.0, 3)}
    # K3a correctness smoke: statistics against a torch scatter-add
    stats.zero_()
    calls["k3a"][0]()
    flat = xd.permute(0, 2, 1).reshape(-1, d)
    sel = md.reshape(-1) != 0
    ref = torch.zeros(K, d, device=dev).index_add_(0, idx.reshape(-1)[sel], flat[sel])
    cnt = torch.bincount(idx.reshape(-1)[sel], minlength=K).float()
    out["k3a_check"] = {"max_rel_err_sums": float(((stats[:K * d].view(K, d) - ref).abs().max() / ref.abs().max()).cpu()),
                        "counts_equal": bool(torch.equal(stats[K * d:], cnt))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
