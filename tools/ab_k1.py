"""A/B timing of vq_assign between two builds of the library on the same box (ctypes only, no package import):
    python tools/ab_k1.py <libA.so> <libB.so> [gaussian]"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vq_oracle as O  # noqa: E402  (synthetic inputs only)

vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int


def bind(path):
    lib = ctypes.CDLL(path)
    lib.vq_workspace_bytes.restype = ctypes.c_size_t
    lib.vq_workspace_bytes.argtypes = [i64, i64, ci, ci]
    lib.vq_assign.argtypes = [vp, i64, i64, i64, vp, ci, vp, vp, vp, vp, ctypes.c_size_t, ci, vp]
    lib.vq_last_error.restype = ctypes.c_char_p
    return lib


def main():
    gaussian = "gaussian" in sys.argv
    libs = [a for a in sys.argv[1:] if a.endswith(".so")]
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(0)
    K, D = 512, 128
    code = torch.randn(K, D, generator=gen)
    lengths = O.ljspeech_like_lengths(256, gen)
    x, _ = O.synthetic_batch(lengths, D, gen, codebook=None if gaussian else code)
    n, d, t = x.shape
    xd, kd = x.to(dev), code.to(dev)
    idx = torch.empty(n, t, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    out = {}
    bound = [(p, bind(p)) for p in libs]
    for rep in range(3):
        for path, lib in bound:
            ws = torch.empty(int(lib.vq_workspace_bytes(n, t, K, D)), dtype=torch.uint8, device=dev)
            flag = 0
            for _ in range(5):
                assert lib.vq_assign(xd.data_ptr(), n, d, t, kd.data_ptr(), K, idx.data_ptr(), None, None, ws.data_ptr(), ws.numel(), flag, stream) == 0, lib.vq_last_error()
                flag = 256
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(100):
                lib.vq_assign(xd.data_ptr(), n, d, t, kd.data_ptr(), K, idx.data_ptr(), None, None, ws.data_ptr(), ws.numel(), 256, stream)
            b.record()
            torch.cuda.synchronize()
            out.setdefault(os.path.basename(path), []).append(round(a.elapsed_time(b) / 100, 5))
    print(json.dumps({"gaussian": gaussian, "ms_per_step": out}))


if __name__ == "__main__":
    main()
