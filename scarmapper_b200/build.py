"""
build.py — compile libscar_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

    python -m scarmapper_b200.build [--force]

Output: scarmapper_b200/libscar_b200.so (git-ignored; travels to the GPU box with gpurun).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libscar_b200.so")
SOURCES = ["scar_api.cu", "scar_pack.cu", "scar_demux.cu", "scar_window.cu", "scar_table.cu", "scar_fastq.cu", "scar_fastq_dev.cu",
           "scar_homeology.cu", "scar_rawdata.cu", "scar_alignment.cu", "scar_frequency.cu", "scar_ingest.cu"]
LINK_LIBS = ["-lz", "-lpthread"]          # zlib: gzip FASTQ inflated natively (scar_ingest.cu)
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-cudart", "static"]


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d += [os.path.join(HERE, "..", "include", f) for f in ("scar_b200.h", "scar_records.h")]
    return d


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(p) for p in _deps()):
        return LIB
    # an object is rebuilt when its source or any header is newer; the sources compile side by side
    headers = [p for p in _deps() if not p.endswith((".cu", ".o"))]
    newest_header = max(os.path.getmtime(p) for p in headers)
    extra = os.environ.get("SCAR_NVCC_EXTRA", "").split()
    jobs, objs = [], []
    for src in SOURCES:
        path, obj = os.path.join(CSRC, src), os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or extra or verbose or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(path), newest_header):
            jobs.append([nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj])
    procs = [subprocess.Popen(cmd) for cmd in jobs]
    failed = [cmd for cmd, p in zip(jobs, procs) if p.wait() != 0]
    if failed:
        raise subprocess.CalledProcessError(1, failed[0])
    subprocess.check_call([nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-o", LIB] + objs + LINK_LIBS)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
