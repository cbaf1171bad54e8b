"""Build libwvb.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m wavpackdecoder_b200.build [--force]

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwvb.so")
SOURCES = ["wvb_cuda.cu", "wvb_index.cpp"]
HEADERS = ["wvb_pcm.cuh", "wvb_checksum.cuh", "wvb_dsf.cuh", "wvb_dsd.cuh", "wvb_dsd_core.cuh", "wvb_grid.h", "wvb_md5.cuh", "wvb_plan.h", "wv_tables.h", os.path.join("..", "..", "include", "wvb.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC,-fwrapv,-O2",
    "-cudart", "static", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, out=None, defines=()):
    """out / defines: an experimental build next to the shipped one (select it at run time with WVB_LIB=<path>)."""
    if out is None and not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", out or LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(HERE, "build.log" if out is None else os.path.basename(out) + ".log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libwvb.so")
    if verbose:
        print(log)
    return out or LIB


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    build(force="--force" in sys.argv, verbose=True, out=outs[0] if outs else None, defines=defs)
