"""The reference-side adapter (adapter/StateHelperB200.cpp) compiles against include/ovp.h and against stand-in declarations of
the reference's StateHelper / State / ov_type / Eigen interfaces (adapter/stubs/), and defines every StateHelper static."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize("src", ["StateHelperB200.cpp", "PlaneFittingB200.cpp", "UpdaterMSCKFB200.cpp", "UpdatersB200.cpp"])
def test_adapter_syntax_check(src):
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "adapter", "stubs"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "adapter", src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_planefitting_adapter_keeps_the_reference_signatures():
    """the stand-in header restates PlaneFitting.h:83-104; the adapter defines both entry points with those parameter lists"""
    src = re.sub(r"\s+", " ", open(os.path.join(ROOT, "adapter", "PlaneFittingB200.cpp")).read())
    assert "bool PlaneFitting::plane_fitting(std::vector<std::shared_ptr<ov_core::Feature>> &feats, Eigen::Vector4d &plane_abcd, int min_inlier_num, double max_plane_solver_condition_number)" in src
    assert ("bool PlaneFitting::optimize_plane(std::vector<std::shared_ptr<ov_core::Feature>> &feats, Eigen::Vector3d &cp_inG, "
            "std::unordered_map<size_t, std::unordered_map<double, ov_core::FeatureInitializer::ClonePose>> &clonesCAM, double sigma_px_norm, "
            "double sigma_c, bool fix_plane, const Eigen::VectorXd &stateI, const Eigen::VectorXd &calib0)") in src


def test_adapter_defines_every_statehelper_static():
    hdr = open(os.path.join(ROOT, "adapter", "stubs", "state", "StateHelper.h")).read()
    declared = set(re.findall(r"static\s+[\w:<>\s,\*&]+?\b(\w+)\s*\(std::shared_ptr<State> state", hdr))
    src = open(os.path.join(ROOT, "adapter", "StateHelperB200.cpp")).read()
    defined = set(re.findall(r"StateHelper::(\w+)\s*\(std::shared_ptr<State> state", src))
    assert len(declared) == 13, sorted(declared)
    assert declared <= defined, sorted(declared - defined)
