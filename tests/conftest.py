import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def chi2_table():
    from ov_plane_b200 import synth
    return synth.chi2_table()


def make_pair(S, chi2, max_state=None, max_meas_rows=None):
    """(gpu ctx, oracle ctx) loaded with the same scenario; returns also the clone handle lists"""
    from ov_plane_b200 import api, synth
    import oracle_backend
    ctx = api.Context(S.options, device=0, max_state=max_state or max(128, S.N + 64), max_meas_rows=max_meas_rows or 60000)
    ctx.set_chi2_table(chi2)
    orc = oracle_backend.OracleContext(S.options)
    orc.set_chi2_table(chi2)
    chg = synth.load_scenario_into(ctx, S)
    cho = synth.load_scenario_into(orc, S)
    return ctx, orc, chg, cho
