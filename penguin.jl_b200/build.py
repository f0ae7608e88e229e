"""In-tree nvcc build of libpenguin_b200.so for sm_100a (called by __graft_entry__.build())."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "penguin_b200.cu")
SO = os.path.join(HERE, "libpenguin_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith((".cu", ".cuh"))] + \
           [os.path.join(HERE, "..", "include", "penguin_b200.h")]


def build(force=False, verbose=False):
    newest = max(os.path.getmtime(p) for p in sources())
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= newest:
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [SRC, "-o", SO, "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libpenguin_b200.so")
    with open(os.path.join(HERE, "csrc", "ptxas_info.log"), "w") as fh:
        fh.write(res.stderr)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
