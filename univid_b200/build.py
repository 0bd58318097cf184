"""In-tree build of the CUDA extension: ``python -m univid_b200.build``.

Produces ``univid_b200/libunivid_b200.so`` (the C-ABI library, include/univid_b200.h) and
``univid_b200/csrc/tests/uvb_test`` (stand-alone GPU check) with nvcc for sm_100a only.
nvcc cross-compiles without a GPU, so this also is the CPU-side "does it build" check.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libunivid_b200.so")
TEST_BIN = os.path.join(CSRC, "tests", "uvb_test")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the sm_100a extension cannot be built")
    return nvcc


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources():
    srcs = []
    for root, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h")):
                srcs.append(os.path.join(root, f))
    srcs.append(os.path.join(HERE, "..", "include", "univid_b200.h"))
    return srcs


def build(force=False, verbose=False, with_tests=True):
    """Compile the library (and the stand-alone test binary) if out of date. Returns the .so path."""
    srcs = _sources()
    nvcc = _nvcc()
    if force or not _newer(LIB, srcs):
        cmd = [nvcc] + NVCC_FLAGS + ["-shared", "-o", LIB, os.path.join(CSRC, "c_api.cu")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        subprocess.run(cmd, check=True, cwd=CSRC)
    if with_tests and (force or not _newer(TEST_BIN, srcs + [LIB])):
        cmd = [nvcc] + NVCC_FLAGS + [
            "-o", TEST_BIN, os.path.join(CSRC, "tests", "uvb_test.cu"),
            "-L", HERE, "-lunivid_b200", "-Xlinker", "-rpath=$ORIGIN/../..",
        ]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB


def build_profile_variant(verbose=False):
    """Lab build (not shipped, not loaded by the package): the library and the stand-alone test binary with
    -DUVB_FMHA_PROFILE (barrier-wait cycle counters in the attention kernel) and -DUVB_LAB_VARIANTS (the measured and
    rejected kernel variants: early S release, FMA-pipe exp2) as csrc/tests/prof/."""
    nvcc = _nvcc()
    out = os.path.join(CSRC, "tests", "prof")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "libunivid_b200.so")
    for cmd in ([nvcc] + NVCC_FLAGS + ["-DUVB_FMHA_PROFILE", "-DUVB_LAB_VARIANTS", "-shared", "-o", lib, os.path.join(CSRC, "c_api.cu")],
                [nvcc] + NVCC_FLAGS + ["-DUVB_FMHA_PROFILE", "-o", os.path.join(out, "uvb_test"),
                                       os.path.join(CSRC, "tests", "uvb_test.cu"), "-L", out, "-lunivid_b200",
                                       "-Xlinker", "-rpath=$ORIGIN"]):
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True, cwd=CSRC)
    return out


if __name__ == "__main__":
    if "--profile" in sys.argv:
        print("built", build_profile_variant(verbose=True))
        sys.exit(0)
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
