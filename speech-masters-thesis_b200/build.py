"""Builds libvqb200.so (sm_100a only) in-tree with nvcc.  `python -m ... build.py` or build.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "vq_api.cu")
OUT_DIR = os.path.join(HERE, "lib")
OUT = os.path.join(OUT_DIR, "libvqb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "--use_fast_math=false",
    "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in sorted(os.listdir(d))] + [os.path.join(HERE, "..", "include", "vqb200.h")]


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(s) <= t for s in sources())


def build(force=False, verbose=False):
    """Compile csrc/vq_api.cu -> lib/libvqb200.so.  Returns the path.  Raises on failure."""
    if not force and up_to_date():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    flags = [f for f in FLAGS if f != "--use_fast_math=false"]
    exp = os.environ.get("VQ_EXPERIMENT")
    if exp:
        flags = flags + [f"-DVQ_EXPERIMENT={int(exp)}"]
    for knob in ("VQ_K3_WARPS", "VQ_K3_DW", "VQ_K3_STAGES", "VQ_K3_OCC"):       # build-time experiment knobs of K3a
        if os.environ.get(knob):
            flags = flags + [f"-D{knob}={int(os.environ[knob])}"]
    cmd = [NVCC] + flags + ["-o", OUT, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(OUT_DIR, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-6000:])
    if verbose:
        print(log)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
