"""Build oracle/csrc/kernels.c -> oracle/_build/libkrylov_native.so (gcc, OpenMP).  Test / baseline
infrastructure only; called by __graft_entry__.build() and lazily by oracle/native.py."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), "_build")
OUT = os.path.join(OUT_DIR, "libkrylov_native.so")
SRC = os.path.join(HERE, "kernels.c")


def build(force: bool = False) -> str:
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    # not -march=native: the library is built in one container and run on another host; AVX2 + FMA is the
    # common denominator of every server CPU a B200 box ships with
    cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared", "-std=c99", "-Wall", "-Wextra",
           "-o", OUT, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
