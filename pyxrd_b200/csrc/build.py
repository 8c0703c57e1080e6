"""Build libpyxrd_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["pyxrd_b200.cu"]
DEPENDS = ["pyxrd_b200.cu", "kernels.cuh", "lpf_table.inc", "sincos_table.inc", os.path.join("..", "..", "include", "pyxrd_b200.h")]
OUTPUT = os.path.join(HERE, "libpyxrd_b200.so")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared"]


def find_nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not os.path.exists(OUTPUT):
        return False
    out_t = os.path.getmtime(OUTPUT)
    return all(os.path.getmtime(os.path.join(HERE, d)) <= out_t for d in DEPENDS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUTPUT
    cmd = [find_nvcc()] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUTPUT] + SOURCES
    proc = subprocess.run(cmd, cwd=HERE, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed (%d)" % proc.returncode)
    return OUTPUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
