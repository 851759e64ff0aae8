"""Statistics of the exact re-scan's candidate masks on an adversarial (i.i.d. Gaussian) batch, and timing of the kernel."""
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


def a256(v):
    return (v + 255) // 256 * 256


def main(K=512, D=128, n_utt=256):
    gen = torch.Generator().manual_seed(0)
    code = torch.randn(K, D, generator=gen)
    lengths = O.ljspeech_like_lengths(n_utt, gen)
    x, mask = O.synthetic_batch(lengths, D, gen, codebook=None)
    n, d, t = x.shape
    rows = n * t
    xd, kd = x.to(dev), code.to(dev)
    idx = torch.empty(n, t, dtype=torch.int64, device=dev)
    ws = torch.zeros(int(lib.vq_workspace_bytes(n, t, K, d)), dtype=torch.uint8, device=dev)
    scalars = torch.zeros(16, dtype=torch.float64, device=dev)
    rc = lib.vq_assign(xd.data_ptr(), n, d, t, kd.data_ptr(), K, idx.data_ptr(), None, scalars.data_ptr(), ws.data_ptr(), ws.numel(), 2,
                       torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.vq_last_error()
    torch.cuda.synchronize()
    n_list = int(scalars[3].item())
    Kp, Dp = (K + 127) // 128 * 128, (D + 63) // 64 * 64
    off = 256 + 3 * a256(Kp * 4) + a256(Kp * Dp * 2)
    rows_list = ws[off:off + 4 * n_list].view(torch.int32).cpu()
    off2 = off + a256(rows * 4)
    masks = ws[off2:off2 + 4 * n_list].view(torch.int32).cpu().to(torch.int64) & 0xFFFFFFFF
    pop = torch.zeros(n_list, dtype=torch.int64)
    for b in range(32):
        pop += (masks >> b) & 1
    hist = torch.bincount(pop, minlength=33).tolist()
    print(json.dumps({"n_list": n_list, "rows": rows, "chains_flagged_hist": hist, "mean_chains": float(pop.float().mean()),
                      "full_masks": int((masks == 0xFFFFFFFF).sum())}))
    os.makedirs("gpurun_out", exist_ok=True)


if __name__ == "__main__":
    main()
