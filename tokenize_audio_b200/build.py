"""In-tree build of the CUDA extension (libmimi_b200.so) with nvcc for sm_100a only."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libmimi_b200.so")
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-shared", "-Xcompiler", "-fPIC",
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
]


def _sources():
    out = [os.path.join(HERE, "..", "include", "mimi_b200.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h", ".inl")):
            out.append(os.path.join(CSRC, f))
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/mimi_b200.cu -> libmimi_b200.so (cross-compiles without a GPU). Returns the path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libmimi_b200.so (and there is no CPU fallback)")
    tmp = f"{LIB_PATH}.{os.getpid()}.tmp"       # pid-unique: two ranks building at once never share a half-written file
    extra = os.environ.get("MIMI_B200_NVCC_EXTRA", "").split()       # e.g. -DMIMI_TCP_DEBUG for tools/gpu_tcp_hang.py
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-o", tmp, os.path.join(CSRC, "mimi_b200.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(r.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
