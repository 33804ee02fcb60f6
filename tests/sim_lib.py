"""TEST INFRASTRUCTURE ONLY: puts the sequential stand-in device (tests/sim/libfastq_sim.so: the product's host code over a CPU
implementation of FqDevice) behind fastq_utils_b200's Python binding, so that the engine, the renderer and the multi-rank
orchestration of dist.py can run on a machine without GPUs.  The product itself has no such switch: importing this module is
the only way to get the stand-in, and nothing outside tests/ imports it."""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def use_sim_library():
    import fastq_utils_b200.api as api
    d = os.path.join(ROOT, "tests", "sim")
    subprocess.check_call(["make", "-C", d], stdout=subprocess.DEVNULL)
    api._lib = api.bind(ctypes.CDLL(os.path.join(d, "libfastq_sim.so")))
    return api._lib
