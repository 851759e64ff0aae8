"""BASELINE.json configs[3]: codebook sweep K = 512 .. 65536 (powers of two) x D = 64 .. 512 -- the 32 grid points of
SURVEY.md 8(d) config 4 -- on one B200 at 262 144 frames: K1 time, distance-GEMM TFLOP/s against the measured tensor
peak, frames sent to the exact re-scan, and index agreement with the ORACLE (oracle/vq_oracle.py, chunked) on the first
4096 rows of every point, mismatches classified in fp64.  Writes gpurun_out/sweep.json (clustered latents) or
gpurun_out/sweep_gaussian.json (--gaussian).  --quick: the 9 corner / edge points only."""
import ctypes
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
PEAK = 1647.2
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops"])
except (OSError, ValueError, KeyError):
    pass


def run(K, D, n=128, t=2048, clustered=True, check_rows=4096):
    gen = torch.Generator(device=dev).manual_seed(K * 1000 + D)
    code = torch.randn(K, D, generator=gen, device=dev)
    if clustered:
        j = torch.randint(0, K, (n, t), generator=gen, device=dev)
        x = code[j].permute(0, 2, 1).contiguous() + 0.5 * torch.randn(n, D, t, generator=gen, device=dev)
        del j
    else:
        x = torch.randn(n, D, t, generator=gen, device=dev)
    idx = torch.empty(n, t, dtype=torch.int64, device=dev)
    ws = torch.empty(int(lib.vq_workspace_bytes(n, t, K, D)), dtype=torch.uint8, device=dev)
    sc = torch.zeros(16, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def call(scalars=None):
        rc = lib.vq_assign(x.data_ptr(), n, D, t, code.data_ptr(), K, idx.data_ptr(), None, scalars, ws.data_ptr(), ws.numel(), 0, stream)
        assert rc == 0, lib.vq_last_error()

    for _ in range(3):
        call()
    torch.cuda.synchronize()
    lib.vq_profile_enable(1)
    for _ in range(5):
        call()
    prof = (ctypes.c_float * 4)()
    assert lib.vq_profile_read(prof) == 0
    lib.vq_profile_enable(0)
    call(sc.data_ptr())
    torch.cuda.synchronize()
    unsafe = float(sc[3])
    # agreement with the oracle on the first utterances (fp32 matmul + min on the host, chunked; disagreements audited in fp64)
    nn = max(1, check_rows // t)
    rows = x[:nn].permute(0, 2, 1).reshape(-1, D).cpu()
    code_cpu = code.cpu()
    ref, _ = O.assign_chunked(rows, code_cpu, chunk=1024 if K > 8192 else 4096)
    audit = O.audit_indices(rows, code_cpu, ref, idx[:nn].cpu().reshape(-1))
    flops = 2.0 * n * t * K * D
    main_ms, total_ms = float(prof[1]), float(prof[0] + prof[1] + prof[2])
    return {"K": K, "D": D, "rows": n * t, "clustered": clustered, "prepare_ms": float(prof[0]), "assign_main_ms": main_ms,
            "fallback_ms": float(prof[2]), "tflops_main": flops / (main_ms * 1e-3) / 1e12, "frac_of_peak_main": flops / (main_ms * 1e-3) / 1e12 / PEAK,
            "tflops_step": flops / (total_ms * 1e-3) / 1e12, "frames_per_s_step": n * t / (total_ms * 1e-3),
            "unsafe_rows": unsafe, "unsafe_frac": unsafe / (n * t), "check": {k: audit[k] for k in ("rows", "mismatches", "near_ties", "errors")}}


if __name__ == "__main__":
    out = []
    shapes = [(K, D) for K in (512, 1024, 2048, 4096, 8192, 16384, 32768, 65536) for D in (64, 128, 256, 512)]
    if "--quick" in sys.argv:
        shapes = [(512, 64), (512, 128), (512, 256), (512, 512), (2048, 128), (8192, 128), (8192, 256), (65536, 64), (65536, 512)]
    for K, D in shapes:
        try:
            r = run(K, D, clustered="--gaussian" not in sys.argv)
        except Exception as e:  # noqa: BLE001
            r = {"K": K, "D": D, "exception": repr(e)}
        print(json.dumps(r), flush=True)
        out.append(r)
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    name = "sweep_gaussian.json" if "--gaussian" in sys.argv else "sweep.json"
    json.dump({"peak_tflops": PEAK, "results": out}, open(os.path.join("gpurun_out", name), "w"), indent=1)
