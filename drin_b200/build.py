"""In-tree nvcc build of libdrin_b200.so for sm_100a (no JIT cache: the built .so travels with the tree)."""
from __future__ import annotations

import glob
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libdrin_b200.so")
STAMP_PATH = os.path.join(PKG_DIR, ".libdrin_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libdrin_b200.so")
    return nvcc


def have_nvcc() -> bool:
    return bool(shutil.which("nvcc")) or os.path.exists("/usr/local/cuda/bin/nvcc")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest() -> str:
    h = hashlib.sha256()
    files = sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(ROOT, "include", "drin_b200.h")]
    for f in files:
        h.update(os.path.basename(f).encode())      # relative names: the tree may live anywhere (gpurun copies it)
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh() -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH)):
        return False
    with open(STAMP_PATH) as fh:
        return fh.read().strip() == _digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if the sources changed.  Safe under torchrun: ranks serialise on a lock file, objects go to a
    per-process directory and the library is moved into place atomically."""
    if not force and is_fresh():
        return LIB_PATH
    import fcntl

    os.makedirs(os.path.join(PKG_DIR, "build"), exist_ok=True)
    with open(os.path.join(PKG_DIR, "build", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_fresh():          # another rank built it while we waited
                return LIB_PATH
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    objs = []
    obj_dir = os.path.join(PKG_DIR, "build", f"obj{os.getpid()}")
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != "--shared"] + ["-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{out}")
        if verbose and out:
            print(out, file=sys.stderr)
    tmp_lib = os.path.join(obj_dir, "libdrin_b200.so")
    link = [_nvcc(), "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp_lib] + objs
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp_lib, LIB_PATH)
    with open(STAMP_PATH + ".tmp", "w") as fh:
        fh.write(_digest())
    os.replace(STAMP_PATH + ".tmp", STAMP_PATH)
    shutil.rmtree(obj_dir, ignore_errors=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
