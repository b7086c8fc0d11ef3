"""Build mocopci_b200/libb200pci.so in-tree with nvcc for sm_100a (no torch headers, seconds).

    python -m mocopci_b200.build [--force]

The library is a plain C-ABI shared object (include/b200pci.h); Python binds it with ctypes.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb200pci.so")
SOURCES = ["common.cu", "knn.cu", "fps.cu", "gather.cu", "emd.cu", "probe.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC"]


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "b200pci.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, force, hdr_mtime):
    s = os.path.join(CSRC, src)
    o = os.path.join(OBJ, src[:-3] + ".o")
    if (not force and os.path.exists(o)
            and os.path.getmtime(o) >= max(os.path.getmtime(s), hdr_mtime)):
        return o, False
    subprocess.check_call([NVCC, *FLAGS, "-c", s, "-o", o])
    return o, True


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdr_mtime = _deps()
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(lambda s: _compile(s, force, hdr_mtime), SOURCES))
    objs = [o for o, _ in res]
    if any(c for _, c in res) or not os.path.exists(LIB):
        subprocess.check_call([NVCC, "-shared", "-o", LIB, *objs, "-lcudart"])
        if verbose:
            print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
