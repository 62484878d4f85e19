"""Builds the CPU CTA emulator (tests/emu/libemu_p22.so) with g++."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "libemu_p22.so")
SRC = os.path.join(HERE, "emu_p22.cpp")
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "tfhe-rs_b200", "csrc")


def build() -> str:
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps):
        return OUT
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-O2", "-mavx2", "-mfma", "-std=c++17", "-fPIC", "-shared", "-x", "c++", SRC,
                           "-o", OUT])
    return OUT
