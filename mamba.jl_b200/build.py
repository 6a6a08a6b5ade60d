"""Builds libmambacuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "mambacuda", "libmambacuda.so")
SOURCES = ["api.cu", "kern_misc.cu", "seeds_fast.cu", "rats_warp.cu", "rats_fast.cu", "pumps_fast.cu", "glm_nuts.cu", "glm_tc.cu", "tpl_line.cu", "tpl_seeds.cu", "tpl_rats.cu", "tpl_pumps.cu", "tpl_glm.cu", "tpl_surgical.cu", "tpl_dyes.cu", "tpl_salm.cu", "tpl_equiv.cu", "tpl_blocker.cu", "tpl_stacks.cu", "tpl_magnesium.cu", "tpl_oxford.cu", "tpl_epil.cu"]
COMPILE_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def newest_mtime(path):
    m = 0.0
    for root, _, files in os.walk(path):
        for f in files:
            if f.endswith((".cu", ".cuh", ".hpp", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(force=False, verbose=False):
    """Compile every translation unit in parallel, then link the shared library."""
    from concurrent.futures import ThreadPoolExecutor
    src_m = max(newest_mtime(CSRC), os.path.getmtime(os.path.join(HERE, "..", "include", "mambacuda.h")))
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= src_m:
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdr_m = src_m_headers()

    def compile_one(s):
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        srcp = os.path.join(CSRC, s)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(srcp), hdr_m):
            return obj
        cmd = [nvcc] + COMPILE_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", srcp, "-o", obj]
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + ["-lcudart"])
    return OUT


def src_m_headers():
    m = os.path.getmtime(os.path.join(HERE, "..", "include", "mambacuda.h"))
    for f in os.listdir(CSRC):
        if f.endswith((".cuh", ".hpp", ".h")):
            m = max(m, os.path.getmtime(os.path.join(CSRC, f)))
    return m


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
