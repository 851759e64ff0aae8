"""Per-tile timeline of thread block 0 of the tcgen05 assignment kernel (clock64 probes, see vq_assign_debug)."""
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
EV = ["x_issue", "mma_start", "mma_issued", "conv_start", "conv_x_ready", "conv_end", "fin_wait", "fin_go", "scan_start", "scan_end"]


def main(n_utt=256, K=512, D=128):
    gen = torch.Generator().manual_seed(0)
    code = torch.randn(K, D, generator=gen)
    lengths = O.ljspeech_like_lengths(n_utt, gen)
    x, mask = O.synthetic_batch(lengths, D, gen, codebook=code)
    n, d, t = x.shape
    xd, kd = x.to(dev), code.to(dev)
    idx = torch.empty(n, t, dtype=torch.int64, device=dev)
    ws = torch.empty(int(lib.vq_workspace_bytes(n, t, K, d)), dtype=torch.uint8, device=dev)
    tiles = 64 if os.environ.get("TL_PROBE") else 32      # 64 enables the issuer-side completion probe (it serialises N=128 batches)
    trace = torch.zeros(16, tiles, dtype=torch.int64, device=dev)
    for _ in range(3):
        rc = lib.vq_assign_debug(xd.data_ptr(), n, d, t, kd.data_ptr(), K, idx.data_ptr(), None, None, ws.data_ptr(), ws.numel(),
                                 torch.cuda.current_stream().cuda_stream, trace.data_ptr(), tiles)
        assert rc == 0, lib.vq_last_error()
    torch.cuda.synchronize()
    tr = trace.cpu()
    t0 = int(tr[tr > 0].min())
    rows = []
    print("tile " + " ".join(f"{e:>12s}" for e in EV))
    for i in range(tiles):
        if int(tr[3, i]) == 0:
            break
        vals = [int(tr[e, i]) - t0 if int(tr[e, i]) else -1 for e in range(10)]
        rows.append(vals)
        print(f"{i:4d} " + " ".join(f"{v:12d}" for v in vals))
    print("per code-tile probes, local tiles 8..15: mma_go(after acc_empty)  mma_commit  scan_go(after acc_full)  scan_released  scan_done  mma_done(observer warp / issuer probe)")
    fine = []
    for q in range(32):
        vals = [int(tr[e, q]) - t0 if int(tr[e, q]) else -1 for e in range(10, 16)]
        fine.append(vals)
        print(f"tile {8 + q // 4} nt {q % 4}: " + " ".join(f"{v:10d}" for v in vals))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"events": EV, "rows": rows, "fine": fine}, open("gpurun_out/tc_timeline.json", "w"))


if __name__ == "__main__":
    main()
