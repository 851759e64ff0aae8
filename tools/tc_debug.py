"""Diagnostic for the tcgen05 assignment kernel: shortlist quality, bound validity, agreement with the exact kernel."""
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


def run(n, d, t, K, clustered, seed=0):
    gen = torch.Generator().manual_seed(seed)
    code = torch.randn(K, d, generator=gen)
    lengths = torch.full((n,), t)
    x, mask = O.synthetic_batch(lengths, d, gen, codebook=code if clustered else None)
    xd, kd = x.to(dev), code.to(dev)
    idx = torch.full((n, t), -1, dtype=torch.int64, device=dev)
    dbg = torch.zeros(n * t, 4, device=dev)
    sc = torch.zeros(16, dtype=torch.float64, device=dev)
    ws = torch.empty(int(lib.vq_workspace_bytes(n, t, K, d)), dtype=torch.uint8, device=dev)
    rc = lib.vq_assign_debug(xd.data_ptr(), n, d, t, kd.data_ptr(), K, idx.data_ptr(), dbg.data_ptr(), sc.data_ptr(),
                             ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream, None, 0)
    if rc:
        return {"error": lib.vq_last_error().decode()}
    torch.cuda.synchronize()
    ref, _ = vqb200.assign(xd, kd, algo="simt")
    torch.cuda.synchronize()
    s1, bound, g1, err = dbg.unbind(1)
    rows = x.permute(0, 2, 1).reshape(-1, d)
    audit = O.audit_indices(rows, code, ref.cpu().reshape(-1), idx.cpu().reshape(-1))
    # true best approx-vs-exact deviation: exact score of the shortlisted code minus its approximate score
    dev_abs = (s1 - g1).abs()
    out = {"shape": [n, d, t, K], "clustered": clustered, "unsafe_rows": float(sc[3]), "rows": n * t,
           "mismatch_vs_simt": audit["mismatches"], "errors": audit["errors"],
           "max_abs_s1_minus_g1": float(dev_abs.max()), "median_err_bound": float(err.median()),
           "max_ratio_dev_over_err": float((dev_abs / err).max()), "sum_min_d_tc": float(sc[0]),
           "g1_sample": [float(v) for v in g1[:4]], "s1_sample": [float(v) for v in s1[:4]], "bound_sample": [float(v) for v in bound[:4]]}
    return out


if __name__ == "__main__":
    res = []
    for cfg in [(2, 128, 256, 512, False), (2, 128, 256, 512, True), (1, 64, 128, 128, False), (3, 128, 1000, 512, True),
                (1, 256, 512, 1024, False), (1, 512, 256, 2048, True), (64, 128, 1736, 512, True)]:
        try:
            r = run(*cfg)
        except Exception as e:  # noqa: BLE001
            r = {"shape": cfg, "exception": repr(e)}
        print(json.dumps(r), flush=True)
        res.append(r)
        if "exception" in r or "error" in r:
            break
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/tc_debug.json", "w"), indent=1)
