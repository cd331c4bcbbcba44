"""GPU parity of the plane paths beyond in-state planes, of plane merging, and of the sharded (multi-GPU) update halves."""
import numpy as np
import pytest

from conftest import make_pair
from ov_plane_b200 import synth
from test_gpu_parity import relerr, compare_states, oracle_msckf_update, _check_msckf, _plane_report

pytestmark = pytest.mark.gpu


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
    """Planes that are NOT in the state (projected away, UpdaterMSCKF.cpp:601-604).  Gates are compared on the well-defined
    part of the plane chi2 (test_gpu_parity.oracle_msckf_update); the unmodified reference gate is printed beside it."""
    S, ctx, orc, chg, cho, bg, bo, planes = _fresh_plane_case(name, seed, chi2_table)
    g = ctx.msckf_update(bg, 1.0, 1.0)
    o = oracle_msckf_update(orc, bo, 1.0, 1.0)
    e = _check_msckf(S, ctx, orc, chg, cho, g, o, chi_tol=1e-6)
    _plane_report("%s %d (planes not in state)" % (name, seed), g, o, e)


@pytest.mark.parametrize("name,seed", [("tiny_planes", 1), ("tiny_planes", 2), ("small_planes", 0), ("small_planes", 2)])
def test_plane_init(name, seed, chi2_table):
    import oracle_backend
    S, ctx, orc, chg, cho, bg, bo, planes = _fresh_plane_case(name, seed, chi2_table)
    g = ctx.plane_init(bg, 1.0, 1.0)
    with oracle_backend.GaugeProbe(gate_without=True) as gp:  # StateHelper::initialize gates on the well-defined chi2
        o = orc.plane_init(bo, 1.0)
    print(name, seed, "plane init status", g["plane_status"], o["plane_status"], "round-off-row chi2 share", np.round(gp.junk(), 2))
    assert np.array_equal(g["plane_status"], o["plane_status"])
    assert ctx.cov_rows() == orc.cov_rows()
    for hg, ho in zip(g["new_handles"], o["new_handles"]):
        assert (hg >= 0) == (ho >= 0)
        if hg >= 0:
            assert ctx.var_id(int(hg)) == orc.var_id(int(ho))
            assert np.allclose(ctx.var_get(int(hg))[0], orc.var_get(int(ho))[0], rtol=1e-6, atol=1e-7)  # metres; the 3x3 init system is solved with a different (orthogonal-equivalent) H_L
    e = relerr(ctx.cov(), orc.cov())
    print("cov rel err after plane init %.3e" % e)
    assert e < 1e-6
    # a follow-up MSCKF update now treats them as in-state planes on both sides
    g2 = ctx.msckf_update(bg, 1.0, 1.0)
    o2 = oracle_msckf_update(orc, bo, 1.0, 1.0)
    e2 = _check_msckf(S, ctx, orc, chg, cho, g2, o2, chi_tol=1e-6)
    print("cov rel err after the follow-up update %.3e" % e2)


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
    assert e < 1e-7  # two summation orders of the same Gram matrix
    o = orc.msckf_update(synth.feature_batch(S, cho), 1.0, 1.0)
    assert np.array_equal(status, o["feat_status"])
    assert relerr(ctx.cov(), orc.cov()) < 1e-6


def test_cfg5_sharded_8_ways_equals_single_update_and_oracle(chi2_table):
    """BASELINE config 5 (4000 features sharded 8 ways): the 8 rank-local shard-compress blocks (run here one after another on
    one GPU, exactly what 8 ranks compute) + the gathered update == one single-GPU msckf_update of all 4000 features == the
    oracle, at the north-star tolerance (the oracle's Givens compression of the 141 000 x 470 stack takes ~1.5 min)."""
    import torch
    from ov_plane_b200 import api
    S = synth.make_scenario("cfg5_n512_f4000", seed=0)
    ctx, orc, chg, cho = make_pair(S, chi2_table, max_state=576, max_meas_rows=160000)
    ref = synth.make_scenario("cfg5_n512_f4000", seed=0)
    full = api.Context(ref.options, device=0, max_state=576, max_meas_rows=160000)
    full.set_chi2_table(chi2_table)
    chf = synth.load_scenario_into(full, ref)
    gfull = full.msckf_update(synth.feature_batch(ref, chf), 1.0, 1.0)
    n = ctx.msckf_shard_columns(chg)
    G = 8
    blocks = torch.zeros(G * (n + 1) * (n + 1), dtype=torch.float64, device="cuda")
    status = np.zeros(S.F, dtype=np.int32)
    chi = np.zeros(S.F)
    for g in range(G):
        mine = list(range(g, S.F, G))
        r = ctx.msckf_shard_compress(synth.feature_batch(S, chg, mine), chg, blocks.data_ptr() + g * (n + 1) * (n + 1) * 8, 1.0, 1.0)
        status[mine] = r["feat_status"]
        chi[mine] = r["feat_chi2"]
    torch.cuda.synchronize()
    ctx.msckf_update_gathered(blocks.data_ptr(), G, chg)
    assert np.array_equal(status, gfull["feat_status"])
    e1 = relerr(ctx.cov(), full.cov())
    o = orc.msckf_update(synth.feature_batch(S, cho), 1.0, 1.0)
    assert np.array_equal(status, o["feat_status"]), np.nonzero(status != o["feat_status"])
    m = o["feat_status"] != 2
    assert np.allclose(chi[m], o["feat_chi2"][m], rtol=1e-7, atol=0)
    e2 = compare_states(ctx, orc, S, chg, cho, 1e-6)
    e3 = compare_states(full, orc, S, chf, cho, 1e-6)
    print("cfg5: sharded(8) vs single cov rel err %.3e | sharded vs oracle %.3e | single vs oracle %.3e | accepted %d of %d" % (
        e1, e2, e3, int((status == 1).sum()), S.F))
    assert e1 < 1e-6


def test_library_owned_collective_single_rank(chi2_table):
    """ovp_msckf_update_sharded with a 1-rank NCCL communicator owned by the context (the all-gather runs inside the library):
    equal to ovp_msckf_update of the same features, and to the oracle.  (N = 2, 4, 8 ranks: bench.py `sharded_cfg5`.)"""
    S = synth.make_scenario("cfg2_n256_f200", seed=1)
    ctx, orc, chg, cho = make_pair(S, chi2_table)
    ref = synth.make_scenario("cfg2_n256_f200", seed=1)
    from ov_plane_b200 import api
    full = api.Context(ref.options, device=0, max_state=S.N + 64, max_meas_rows=60000)
    full.set_chi2_table(chi2_table)
    chf = synth.load_scenario_into(full, ref)
    gfull = full.msckf_update(synth.feature_batch(ref, chf), 1.0, 1.0)
    ctx.nccl_init(ctx.nccl_unique_id(), 1, 0)
    r = ctx.msckf_update_sharded(synth.feature_batch(S, chg), chg, 1.0, 1.0)
    assert np.array_equal(r["feat_status"], gfull["feat_status"])
    e = relerr(ctx.cov(), full.cov())
    print("library-owned collective (1 rank) vs single update: cov rel err %.3e" % e)
    assert e < 1e-9
    o = orc.msckf_update(synth.feature_batch(S, cho), 1.0, 1.0)
    assert np.array_equal(r["feat_status"], o["feat_status"])
    compare_states(ctx, orc, S, chg, cho, 1e-6)
    ctx.nccl_finalize()
