"""Build the oracle's C helper (oracle/_build/liboracle.so) and, when the
reference tree is mounted, the reference binary (oracle/_ref/audiowmark).

TEST INFRASTRUCTURE ONLY: the product path never imports anything under oracle/.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "_build", "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "audiowmark")
REF_SYNC_DUMP = os.path.join(HERE, "_ref", "sync_dump")      # SyncFinder::search print-out (ref_shims/sync_dump.cc)


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_liboracle(force=False):
    srcs = [os.path.join(HERE, "oracle_c.cc"), os.path.join(HERE, "ref_shims", "fftw_shim.cc")]
    if force or _newer(LIB, srcs + [os.path.join(HERE, "ref_shims", "fftw3.h"), os.path.join(HERE, "ref_shims", "awm_vresampler.hh")]):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        # plain x86-64 baseline, no fast-math: every float op rounds like the reference build
        cmd = ["g++", "-O3", "-std=c++14", "-fopenmp", "-shared", "-fPIC", "-I", HERE, "-o", LIB] + srcs
        subprocess.check_call(cmd)
    return LIB


def build_reference(ref="/root/reference", force=False):
    """Compile the unmodified reference sources (only possible where `ref` exists)."""
    if not os.path.isdir(os.path.join(ref, "src")):
        return REF_BIN if os.path.exists(REF_BIN) else None
    if force or not os.path.exists(REF_BIN) or not os.path.exists(REF_SYNC_DUMP):
        subprocess.check_call(["make", "-s", "-f", os.path.join("oracle", "Makefile.ref"), "-j8", "REF=" + ref], cwd=ROOT)
    return REF_BIN


if __name__ == "__main__":
    print(build_liboracle(force="--force" in sys.argv))
    print(build_reference())
