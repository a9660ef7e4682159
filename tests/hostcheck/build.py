"""Builds tests/hostcheck/libhostcheck.so (g++ host build of csrc/math.cuh, camera_maps.cuh and filter_math.cuh).  TEST-ONLY."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck.cpp")
HDR = os.path.join(HERE, "..", "..", "instantsfm_b200", "csrc", "math.cuh")
HDR2 = os.path.join(HERE, "..", "..", "instantsfm_b200", "csrc", "camera_maps.cuh")
HDR3 = os.path.join(HERE, "..", "..", "instantsfm_b200", "csrc", "filter_math.cuh")
LIB = os.path.join(HERE, "libhostcheck.so")


def load():
    stale = (not os.path.exists(LIB)) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR), os.path.getmtime(HDR2), os.path.getmtime(HDR3))
    if stale:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-x", "c++", "-fPIC", "-shared", "-o", LIB, SRC])
    return ctypes.CDLL(LIB)
