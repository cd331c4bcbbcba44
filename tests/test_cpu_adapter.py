"""The reference-side adapter (adapter/StateHelperB200.cpp) compiles against include/ovp.h and against stand-in declarations of
the reference's StateHelper / State / ov_type / Eigen interfaces (adapter/stubs/), and defines every StateHelper static."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_adapter_syntax_check():
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "adapter", "stubs"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "adapter", "StateHelperB200.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_adapter_defines_every_statehelper_static():
    hdr = open(os.path.join(ROOT, "adapter", "stubs", "state", "StateHelper.h")).read()
    declared = set(re.findall(r"static\s+[\w:<>\s,\*&]+?\b(\w+)\s*\(std::shared_ptr<State> state", hdr))
    src = open(os.path.join(ROOT, "adapter", "StateHelperB200.cpp")).read()
    defined = set(re.findall(r"StateHelper::(\w+)\s*\(std::shared_ptr<State> state", src))
    assert len(declared) == 13, sorted(declared)
    assert declared <= defined, sorted(declared - defined)
