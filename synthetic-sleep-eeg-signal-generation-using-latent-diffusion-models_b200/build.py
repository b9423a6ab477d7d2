"""Build the eegldm CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python <package>/build.py            # -> <package>/eegldm/libeegldm.so

The shared object is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "eegldm", "libeegldm.so")
SOURCES = ["engine.cu", "kernels_simt.cu", "conv_tc.cu", "attn_tc.cu", "train_kernels.cu", "spectral.cu", "psd.cu", "disc.cu", "train_tc.cu", "unet_train_kernels.cu"]
OPTIONAL_SOURCES = []
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def _sources():
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    srcs += [os.path.join(CSRC, s) for s in OPTIONAL_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    return srcs


def _stamp(srcs):
    h = hashlib.sha256()
    deps = list(srcs) + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h", ".inc"))]
    deps.append(os.path.join(ROOT, "include", "eegldm.h"))
    for p in deps:
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources()
    stamp_path = OUT + ".stamp"
    stamp = _stamp(srcs)
    if not force and os.path.exists(OUT) and os.path.exists(stamp_path) and open(stamp_path).read() == stamp:
        return OUT
    nvcc = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for s in srcs:  # compile translation units in parallel
        o = os.path.join(HERE, "build", os.path.basename(s) + ".o")
        objs.append(o)
        cmd = [nvcc] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", s, "-o", o]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", OUT] + objs
    link += ["-ldl"]   # cuFFT is dlopen'ed at first use (spectral.cu)
    subprocess.run(link, check=True)
    with open(stamp_path, "w") as f:
        f.write(stamp)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
