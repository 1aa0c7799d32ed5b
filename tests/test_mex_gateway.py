"""CPU: the MATLAB MEX gateway source compiles against include/manisdp_b200.h (syntax + type check with a stub mex.h;
MATLAB itself is absent from the build container) and only calls symbols the header declares."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gateway_compiles_against_header():
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "tests", "mex_stub"), os.path.join(ROOT, "matlab", "manisdp_mex.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_gateway_uses_only_declared_entry_points():
    hdr = open(os.path.join(ROOT, "include", "manisdp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(manisdp_[a-z_A-Z0-9]+)\s*\(", hdr))
    src = open(os.path.join(ROOT, "matlab", "manisdp_mex.cpp")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    used = set(re.findall(r"\b(manisdp_[a-z_A-Z0-9]+)\s*\(", src)) - {"manisdp_mex"}
    assert used <= declared, used - declared


def test_matlab_drivers_keep_reference_signatures():
    sigs = {"ManiSDP_onlyunitdiag.m": "function [X, obj, data] = ManiSDP_onlyunitdiag(C, options)",
            "ManiSDP_unitdiag.m": "function [X, obj, data] = ManiSDP_unitdiag(At, b, c, K, options)",
            "ManiSDP_unittrace.m": "function [X, obj, data] = ManiSDP_unittrace(At, b, c, K, options)",
            "ManiSDP.m": "function [X, obj, data] = ManiSDP(At, b, c, K, options)"}
    for f, s in sigs.items():
        assert open(os.path.join(ROOT, "matlab", f)).readline().strip() == s
