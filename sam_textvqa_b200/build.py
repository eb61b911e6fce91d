"""Builds libsamk.so (all CUDA kernels + the C ABI) for sm_100a with nvcc, in-tree.

    python -m sam_textvqa_b200.build [--force]

The shared object is written next to this file (git-ignored, but shipped to the GPU box by
gpurun).  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("SAMK_LIB") or os.path.join(HERE, "libsamk.so")
OBJ = os.environ.get("SAMK_OBJ_DIR") or os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("SAMK_NVCC_EXTRA", "").split()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_dep():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    paths.append(os.path.join(os.path.dirname(HERE), "include", "samk.h"))
    return max(os.path.getmtime(p) for p in paths)


def _source_hash():
    """sha256 over every file the library is built from (sources, headers, flags): the shipped .so is reused only
    when it was built from exactly these bytes, whatever the file times say."""
    import hashlib
    h = hashlib.sha256(" ".join(FLAGS).encode())
    paths = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(os.path.dirname(HERE), "include", "samk.h")]
    for p in paths:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp = OUT + ".sha256"
    want = _source_hash()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read().strip() == want:
        return OUT
    os.makedirs(OBJ, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) >= _newest_dep()):
            return obj, ""          # (objects are an mtime cache; the library itself is keyed on the content hash)
        r = subprocess.run([NVCC] + FLAGS + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(compile_one, sources()))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    r = subprocess.run([NVCC, "-shared", "-o", OUT] + objs + ["-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(OUT + ".sha256", "w") as f:
        f.write(want + "\n")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
