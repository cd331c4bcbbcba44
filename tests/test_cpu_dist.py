"""world_size-2 gloo test of the N>1 path (host logic): round-robin sharding, the [R^T ; z^T] block layout, one all-gather, and
the second-level compression + update.  The per-rank compute is done by the CPU oracle here (no GPU in this container); the
result must equal the single-process update of all features."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _columns(o, ch):
    hs = [o.handle_calib(), o.handle_intrinsics()] + list(ch)
    hs = [h for h in hs if o.var_id(h) >= 0]
    hs.sort(key=lambda h: o.var_id(h))
    return hs


def _stack_features(o, S, ch, feats, hs, chi2_table):
    """oracle-side rank-local half: Jacobian -> nullspace -> chi2 gate -> stack in the agreed column order -> compress"""
    col0 = {}
    n = 0
    for h in hs:
        col0[h] = n
        n += o.var_size(h)
    rows_H, rows_r = [], []
    for f in feats:
        a, b = S.meas_offset[f], S.meas_offset[f + 1]
        idx = S.meas_clone_idx[a:b]
        pf = S.p_FinG_original[f]
        Hf, Hx, r, order = o.feature_jacobian_full([ch[i] for i in idx], S.uv[a:b], pf, pf, 0, None, None, 1.0, 0.01)
        Ho, ro = o.nullspace_project_inplace(Hf, Hx, r)
        Pm = o.get_marginal_covariance(order)
        Sm = Ho @ Pm @ Ho.T + np.eye(len(ro))
        chi2 = ro @ np.linalg.solve(Sm, ro)
        if chi2 > chi2_table[len(ro)]:
            continue
        Hbig = np.zeros((len(ro), n))
        c = 0
        for h in order:
            s = o.var_size(h)
            Hbig[:, col0[h]:col0[h] + s] = Ho[:, c:c + s]
            c += s
        rows_H.append(Hbig)
        rows_r.append(ro)
    if not rows_H:
        return np.zeros((n + 1, n + 1))
    H, r = np.vstack(rows_H), np.concatenate(rows_r)
    if H.shape[0] > n:
        H, r = o.measurement_compress_inplace(H, r)
    blk = np.zeros((n + 1, n + 1))  # column-major (n+1)x(n+1): block[j, i] = R[i, j], block[n, i] = z[i]
    blk[:n, :H.shape[0]] = H.T
    blk[n, :H.shape[0]] = r
    return blk


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle_backend as ob
    from ov_plane_b200 import dist as ovd
    from ov_plane_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = synth.make_scenario("tiny_points", seed=3, F=24)
    chi2 = synth.chi2_table()
    o = ob.OracleContext(S.options)
    o.set_chi2_table(chi2)
    ch = synth.load_scenario_into(o, S)
    hs = _columns(o, ch)
    n = sum(o.var_size(h) for h in hs)

    def compress_fn(feats):
        return torch.from_numpy(np.asfortranarray(_stack_features(o, S, ch, feats, hs, chi2)).ravel(order="F").copy())

    def update_fn(allb, G):
        B = allb.numpy().reshape(G, n + 1, n + 1).transpose(0, 2, 1)  # back to (row, col) indexing of each col-major block
        H = np.vstack([B[g][:n, :n].T for g in range(G)])
        r = np.concatenate([B[g][n, :n] for g in range(G)])
        H, r = o.measurement_compress_inplace(H, r)
        o.ekf_update(hs, H, r)

    mine = ovd.sharded_update(compress_fn, update_fn, S.F, rank, world)
    assert mine == list(range(rank, S.F, world))
    q.put((rank, o.cov(), o.var_get(o.handle_imu())[0]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_update_equals_single_process(chi2_table):
    import torch.multiprocessing as mp
    import oracle_backend as ob
    from ov_plane_b200 import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    S = synth.make_scenario("tiny_points", seed=3, F=24)
    o = ob.OracleContext(S.options)
    o.set_chi2_table(chi2_table)
    ch = synth.load_scenario_into(o, S)
    b = synth.feature_batch(S, ch)
    o.msckf_update(b, 1.0, 1.0)
    P_ref = o.cov()
    for rank, P, imu in res:
        assert np.linalg.norm(P - P_ref) / np.linalg.norm(P_ref) < 1e-9, rank
        assert np.allclose(imu, o.var_get(o.handle_imu())[0], atol=1e-10)
    assert np.array_equal(res[0][1], res[1][1])  # replicated update: both ranks hold the identical posterior
