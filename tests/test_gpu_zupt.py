"""UpdaterZeroVelocity::try_update (update/UpdaterZeroVelocity.cpp:68-318) through the C ABI against the oracle restatement."""
import numpy as np
import pytest

from conftest import make_pair
from ov_plane_b200 import synth, jpl
from test_gpu_parity import compare_states

pytestmark = pytest.mark.gpu


def _feed(be_list, S, v_imu, t0, t1, rng, moving):
    R = jpl.quat_2_Rot(v_imu[:4])
    g = np.array([0.0, 0.0, 9.81])
    n = int(round((t1 - t0 + 0.03) / 0.0025))
    for k in range(n):
        t = t0 - 0.015 + 0.0025 * k
        wm = v_imu[10:13] + 1.7e-4 / np.sqrt(0.0025) * rng.randn(3) + (np.array([0.2, -0.1, 0.3]) if moving else 0.0)
        am = R @ g + v_imu[13:16] + 2e-3 / np.sqrt(0.0025) * rng.randn(3) + (np.array([0.5, 0.2, -0.4]) if moving else 0.0)
        for be in be_list:
            be.zupt_feed_imu(t, wm, am)


@pytest.mark.parametrize("moving,disparity,expect", [(False, 5.0, True), (True, 5.0, False), (True, 0.2, True)])
def test_zupt_try_update(moving, disparity, expect, chi2_table):
    S = synth.make_scenario("tiny_points", seed=2)
    ctx, orc, chg, cho = make_pair(S, chi2_table)
    rng = np.random.RandomState(7)
    for be in (ctx, orc):
        v, f = be.var_get(be.handle_imu())
        v[7:10] = 1e-3  # nearly at rest
        be.var_set(be.handle_imu(), v, f)
        be.propagator_set_noise(1.6968e-04, 1.9393e-05, 2.0e-3, 3.0e-3, 9.81)
    v_imu, _ = ctx.var_get(ctx.handle_imu())
    t0, t1 = S.timestamp, S.timestamp + 0.1
    _feed([ctx, orc], S, v_imu, t0, t1, rng, moving)
    ag, cg = ctx.zupt_try_update(t1, disparity, 40)
    ao, co = orc.zupt_try_update(t1, disparity, 40)
    print("moving", moving, "disparity", disparity, "-> accepted gpu/oracle", ag, ao, "chi2 %.6f / %.6f" % (cg, co))
    assert ag == ao == expect
    assert abs(cg - co) <= 1e-9 * abs(co)
    assert ctx.get_timestamp() == orc.get_timestamp() == (t1 if expect else t0)
    compare_states(ctx, orc, S, chg, cho, 1e-9)
