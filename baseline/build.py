"""In-tree build of baseline/libupstream_style.so (the upstream-style CUDA rasterizer restatement that
bench.py times as `gpu_baseline`).  Plain nvcc flags as upstream's setup.py uses them: FMA contraction
on, no fast-math."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "upstream_style.cu")
OUT = os.path.join(HERE, "libupstream_style.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(SRC), os.path.getmtime(__file__)):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
           "-shared", SRC, "-o", OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libupstream_style.so")
    return OUT


if __name__ == "__main__":
    print(build(force=True))
