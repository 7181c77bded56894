"""Builds libpmgpu.so for sm_100a with nvcc (cross-compiles without a GPU)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libpmgpu.so")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    srcs.append(os.path.join(_HERE, "..", "include", "pmgpu.h"))
    return any(os.path.getmtime(s) > t for s in srcs if os.path.isfile(s))


def build(force=False, verbose=False):
    if force or needs_build():
        cmd = ["make", "-C", CSRC] + ([] if verbose else ["-s"])
        subprocess.check_call(cmd)
    return LIB
