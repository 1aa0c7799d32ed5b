"""Builds tests/mex_stub/libmex_harness.so = the UNMODIFIED gateway source matlab/manisdp_mex.cpp + the functional mex.h
stand-in (mex_stub.cpp), linked against the in-tree libmanisdp_b200.so.  Test infrastructure (tests/test_mex_gateway.py)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "libmex_harness.so")


def build(force=False):
    srcs = [os.path.join(ROOT, "matlab", "manisdp_mex.cpp"), os.path.join(HERE, "mex_stub.cpp")]
    deps = srcs + [os.path.join(HERE, "mex.h"), os.path.join(ROOT, "include", "manisdp_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    libdir = os.path.join(ROOT, "manisdp_matlab_b200")
    cmd = ["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", HERE,
           *srcs, "-o", OUT, "-L", libdir, "-lmanisdp_b200", "-Wl,-rpath," + libdir, "-Wl,-rpath,$ORIGIN/../../manisdp_matlab_b200"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
