"""Build the native libraries in-tree (they travel to the GPU box with the snapshot).

    python -m crdr_b200.build          # libcrdr_sm100.so (nvcc, sm_100a) + libcrdr_rans.so (g++)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

CU_SOURCES = ["conv_sm100.cu", "bottleneck_sm100.cu", "wgrad_sm100.cu", "train_sm100.cu", "eltwise.cu", "c_abi.cu"]
CU_DEPS = CU_SOURCES + ["common.cuh", "sm100_device.cuh", os.path.join("..", "..", "include", "crdr_b200.h")]
RANS_SOURCES = ["rans.cpp"]
RANS_DEPS = RANS_SOURCES + [os.path.join("..", "..", "include", "crdr_rans.h")]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in deps if os.path.exists(os.path.join(CSRC, d)))


def build_sm100(force=False, verbose=False):
    out = os.path.join(HERE, "libcrdr_sm100.so")
    if force or _stale(out, CU_DEPS):
        trace = ["-DCRDR_TRACE_EVENTS"] if os.environ.get("CRDR_BUILD_TRACE") else []  # tools/conv_events.py build
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", *trace,
               "-shared", "-Xcompiler", "-fPIC", "-o", out] + [os.path.join(CSRC, s) for s in CU_SOURCES] + ["-lcudart"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    return out


def build_rans(force=False):
    out = os.path.join(HERE, "libcrdr_rans.so")
    if force or _stale(out, RANS_DEPS):
        cmd = ["g++", "-O3", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", out] + \
              [os.path.join(CSRC, s) for s in RANS_SOURCES]
        subprocess.check_call(cmd)
    return out


def build_all(force=False, verbose=False):
    return build_sm100(force, verbose), build_rans(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
