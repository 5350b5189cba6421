"""Builds the SIMT-emulator flavour of the kernel sources (g++, no CUDA) for CPU-side logic tests.

TEST INFRASTRUCTURE ONLY: the output lands in tests/emu/_build/ and is loaded explicitly by tests via
cistgcn_b200._cabi.bind(path); the package itself never looks for it."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build", "libcistgcn_emu.so")
SRCS = [os.path.join(ROOT, "cistgcn_b200", "csrc", f) for f in
        ("cistgcn_api.cu", "dstd_block.cuh", "fpn_chain.cuh", "tail.cuh", "simt.h")] + \
       [os.path.join(HERE, "simt_emu.h"), os.path.join(ROOT, "include", "cistgcn_b200.h")]


def build(force: bool = False) -> str:
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not force and os.path.isfile(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in SRCS):
        return OUT
    cmd = ["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-DCISTGCN_EMU", "-Wno-unknown-pragmas",
           "-I", HERE, "-x", "c++", SRCS[0], "-o", OUT]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
