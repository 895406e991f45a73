"""In-tree build of libpegasus_b200.so with nvcc for sm_100a (no torch extension machinery: the
library is a plain C-ABI shared object, loaded with ctypes)."""
import fcntl
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libpegasus_b200.so")
SOURCES = ["pg_abi.cu", "preprocess.cu", "sort.cu", "binning.cu", "composite.cu", "composite3.cu", "pose.cu", "png.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared"]


def source_hash() -> str:
    """Hash of everything the library is built from (sources, headers, flags): stored beside the .so, so that a stale
    library is noticed whatever happened to the files' mtimes (the GPU box receives a fresh copy of the tree)."""
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    files.append(os.path.join(os.path.dirname(HERE), "include", "pegasus_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    h.update(" ".join(NVCC_FLAGS + SOURCES).encode())
    return h.hexdigest()


def needs_build(out: str = OUT) -> bool:
    if not os.path.exists(out) or not os.path.exists(out + ".srchash"):
        return True
    return open(out + ".srchash").read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False, out: str = OUT, defines=()) -> str:
    """`out` / `defines` build a tuning variant beside the product (selected with PG_LIB_PATH).  Concurrent callers
    (one process per GPU) serialise on a lock file; the library appears atomically (temp file + rename)."""
    if not force and not defines and not needs_build(out):
        return out
    with open(out + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not defines and not needs_build(out):  # another process built it while we waited
            return out
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        tmp = f"{out}.tmp{os.getpid()}"
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [f"-D{d}" for d in defines] + \
              [os.path.join(CSRC, s) for s in SOURCES] + ["-o", tmp]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            if os.path.exists(tmp):
                os.remove(tmp)
            raise RuntimeError("nvcc failed building libpegasus_b200.so")
        os.replace(tmp, out)
        if not defines:
            with open(out + ".srchash", "w") as f:
                f.write(source_hash())
    return out


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
