"""Builds the SIMT-emulator flavour of the kernel sources (g++, no CUDA) for CPU-side logic tests.

TEST INFRASTRUCTURE ONLY: the output lands in tests/emu/_build/ and is loaded explicitly by tests via
cistgcn_b200._cabi.bind(path); the package itself never looks for it."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build", "libcistgcn_emu.so")
CSRC = os.path.join(ROOT, "cistgcn_b200", "csrc")
# every translation unit except the tcgen05 ones (inline PTX: GPU only)
UNITS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu") and not f.startswith("fpn_tc_inst"))
SRCS = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))] + \
       [os.path.join(HERE, "simt_emu.h"), os.path.join(ROOT, "include", "cistgcn_b200.h")]


def build(force: bool = False) -> str:
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not force and os.path.isfile(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in SRCS):
        return OUT
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(os.path.dirname(OUT), "obj")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(unit):
        obj = os.path.join(objdir, os.path.splitext(unit)[0] + ".o")
        subprocess.check_call(["g++", "-std=c++20", "-O2", "-fPIC", "-pthread", "-DCISTGCN_EMU", "-Wno-unknown-pragmas",
                               "-I", HERE, "-x", "c++", "-c", os.path.join(CSRC, unit), "-o", obj])
        return obj

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, UNITS))
    subprocess.check_call(["g++", "-shared", "-pthread", "-o", OUT] + objs)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
