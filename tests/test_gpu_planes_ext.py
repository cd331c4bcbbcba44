"""GPU parity of the plane paths beyond in-state planes, of plane merging, and of the sharded (multi-GPU) update halves."""
import numpy as np
import pytest

from conftest import make_pair
from ov_plane_b200 import synth
from test_gpu_parity import relerr, compare_states

pytestmark = pytest.mark.gpu


def plane_gates_agree(g, o):
    """Plane-level chi2 of the reference contains, for every rank-deficient pivot of the stacked H_x, the square of an arbitrary
    (round-off defined) unit projection of the residual; the CUDA path gives those rows zero weight, so its chi2 is
    smaller by a few units and a plane within that band of the threshold may gate differently.  Returns False (borderline)
    in that case after checking that the disagreement has exactly this signature; any other disagreement fails."""
    ok = True
    for i, (a, b) in enumerate(zip(g["plane_status"], o["plane_status"])):
        if a != b:
            dg, do = g["plane_chi2"][i], o["plane_chi2"][i]
            assert a == 1 and b == 0 and 0.0 <= do - dg < 0.25 * do, (i, a, b, dg, do)
            ok = False
    return ok


def _fresh_plane_case(name, seed, chi2_table):
    S = synth.make_scenario(name, seed=seed)
    planes = synth.drop_planes_from_state(S)
    ctx, orc, chg, cho = make_pair(S, chi2_table)
    bg, bo = synth.feature_batch(S, chg), synth.feature_batch(S, cho)
    for b in (bg, bo):
        b["plane_ids"] = np.array([p[0] for p in planes], dtype=np.int64)
        b["plane_cp"] = np.ascontiguousarray([p[1] for p in planes], dtype=np.float64)
    return S, ctx, orc, chg, cho, bg, bo, planes


@pytest.mark.parametrize("name,seed", [("tiny_planes", 0), ("small_planes", 0), ("small_planes", 2), ("small_planes", 3)])
def test_msckf_update_with_planes_not_in_state(name, seed, chi2_table):
    S, ctx, orc, chg, cho, bg, bo, planes = _fresh_plane_case(name, seed, chi2_table)
    g = ctx.msckf_update(bg, 1.0, 1.0)
    o = orc.msckf_update(bo, 1.0, 1.0)
    print(name, seed, "plane status", g["plane_status"], o["plane_status"], "chi2", np.round(g["plane_chi2"], 1), np.round(o["plane_chi2"], 1))
    if not plane_gates_agree(g, o):
        pytest.skip("borderline plane gate (reference chi2 inflated by its round-off rows)")
    assert np.array_equal(g["feat_status"], o["feat_status"])
    e = relerr(ctx.cov(), orc.cov())
    print("cov rel err %.3e" % e)
    assert e < 1e-6


@pytest.mark.parametrize("name,seed", [("tiny_planes", 1), ("tiny_planes", 2), ("small_planes", 0), ("small_planes", 2)])
def test_plane_init(name, seed, chi2_table):
    S, ctx, orc, chg, cho, bg, bo, planes = _fresh_plane_case(name, seed, chi2_table)
    g = ctx.plane_init(bg, 1.0, 1.0)
    o = orc.plane_init(bo, 1.0)
    print(name, seed, "plane init status", g["plane_status"], o["plane_status"])
    if not np.array_equal(g["plane_status"], o["plane_status"]):
        assert all(a >= b for a, b in zip(g["plane_status"], o["plane_status"]))  # only "GPU accepts, reference rejects"
        pytest.skip("borderline plane-initialisation gate (reference chi2 inflated by its round-off rows)")
    assert ctx.cov_rows() == orc.cov_rows()
    for hg, ho in zip(g["new_handles"], o["new_handles"]):
        if hg >= 0:
            assert ctx.var_id(int(hg)) == orc.var_id(int(ho))
            assert np.allclose(ctx.var_get(int(hg))[0], orc.var_get(int(ho))[0], rtol=1e-6, atol=1e-7)  # metres; the 3x3 init system is solved with a different (orthogonal-equivalent) H_L
    e = relerr(ctx.cov(), orc.cov())
    print("cov rel err after plane init %.3e" % e)
    assert e < 1e-6
    # a follow-up MSCKF update now treats them as in-state planes on both sides
    g2 = ctx.msckf_update(bg, 1.0, 1.0)
    o2 = orc.msckf_update(bo, 1.0, 1.0)
    assert np.array_equal(g2["plane_status"], o2["plane_status"]) and np.array_equal(g2["feat_status"], o2["feat_status"])
    assert relerr(ctx.cov(), orc.cov()) < 1e-6


def test_merge_planes_and_marginalize(chi2_table):
    S = synth.make_scenario("tiny_planes", seed=0)
    ctx, orc, chg, cho = make_pair(S, chi2_table)
    # make plane 2 a near-duplicate of plane 1 so that the merge update passes, then merge 2 -> 1
    v1, f1 = ctx.var_get(ctx.plane_handle(1))
    for be in (ctx, orc):
        be.var_set(be.plane_handle(2), v1 + 1e-4, f1 + 1e-4)
    feat2plane = {int(f): 1 for f, p in zip(S.featid, S.planeid) if p}
    ctx.merge_planes_and_marginalize(feat2plane, {1: [2]})
    orc.merge_planes_and_marginalize(feat2plane, {1: [2]})
    assert ctx.cov_rows() == orc.cov_rows() == S.N - 3
    assert ctx.plane_handle(2) == -1 and orc.plane_handle(2) == -1
    assert relerr(ctx.cov(), orc.cov()) < 1e-9
    assert np.allclose(ctx.var_get(ctx.plane_handle(1))[0], orc.var_get(orc.plane_handle(1))[0], atol=1e-12)


def test_sharded_update_halves_equal_single_update(chi2_table):
    """Two shard-compress halves (as two ranks would run them) + the gathered update == one msckf_update of all features."""
    import torch
    S = synth.make_scenario("cfg2_n256_f200", seed=1)
    ctx, orc, chg, cho = make_pair(S, chi2_table)
    ref = synth.make_scenario("cfg2_n256_f200", seed=1)
    from ov_plane_b200 import api
    full = api.Context(ref.options, device=0, max_state=S.N + 64, max_meas_rows=60000)
    full.set_chi2_table(chi2_table)
    chf = synth.load_scenario_into(full, ref)
    gfull = full.msckf_update(synth.feature_batch(ref, chf), 1.0, 1.0)
    n = ctx.msckf_shard_columns(chg)
    G = 2
    blocks = torch.zeros(G * (n + 1) * (n + 1), dtype=torch.float64, device="cuda")
    status = np.zeros(S.F, dtype=np.int32)
    for g in range(G):
        mine = list(range(g, S.F, G))
        r = ctx.msckf_shard_compress(synth.feature_batch(S, chg, mine), chg, blocks.data_ptr() + g * (n + 1) * (n + 1) * 8, 1.0, 1.0)
        status[mine] = r["feat_status"]
    torch.cuda.synchronize()
    ctx.msckf_update_gathered(blocks.data_ptr(), G, chg)
    assert np.array_equal(status, gfull["feat_status"])
    e = relerr(ctx.cov(), full.cov())
    print("sharded vs single cov rel err %.3e" % e)
    assert e < 1e-9
    o = orc.msckf_update(synth.feature_batch(S, cho), 1.0, 1.0)
    assert np.array_equal(status, o["feat_status"])
    assert relerr(ctx.cov(), orc.cov()) < 1e-6
