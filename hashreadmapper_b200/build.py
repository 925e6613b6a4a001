"""Builds libhrm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m hashreadmapper_b200.build [--force] [--verbose]

The library is the product: hand-written CUDA kernels + the C ABI of include/hrm_b200.h.
"""
import os
import subprocess
import sys
import concurrent.futures as cf

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libhrm_b200.so")
SOURCES = ["runtime.cu", "k1_pack.cu", "k2_minhash.cu", "k3_table.cu", "k4_collect.cu", "k4_fused.cu", "k5_shd.cu",
           "k7_verify.cu", "store.cu", "mapper.cu", "sam.cu", "partition.cu", "ingest.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v"]


def _newest_input():
    t = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_input():
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")

    def compile_one(src):
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    objs = []
    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        for src, obj, r in ex.map(compile_one, SOURCES):
            if verbose or r.returncode != 0:
                sys.stderr.write("== %s ==\n%s%s\n" % (src, r.stdout, r.stderr))
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for " + src)
            with open(os.path.join(BUILD, src + ".ptxas.log"), "w") as f:
                f.write(r.stderr)
            objs.append(obj)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-o", LIB] + objs + ["-ldl", "-lz"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
