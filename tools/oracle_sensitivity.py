"""Evidence for the plane-gate comparison in tests/ (VERDICT r1 item 1c): the reference's plane chi2 is not reproducible by the
reference algorithm itself.

Builds the CPU oracle (oracle/, the restatement of the reference's Givens path) twice from the same sources - once with the
reference's own flags (baseline x86-64, no FMA) and once with `-mfma -ffp-contract=fast` - and runs UpdaterMSCKF::update on the
same seeded scenarios with both.  Reported per scenario:
  * relative Frobenius difference of the two posteriors (when all gates agree),
  * the plane chi2 of both builds as the reference computes it (incl. the rows whose Jacobian part is round-off),
  * the well-defined part of both (oracle.hpp GaugeProbe: chi2 minus the squared projection of the compressed residual onto the
    left null space of the compressed Jacobian), and the gates.
Usage: python tools/oracle_sensitivity.py [out.md]      (CPU only; ~1 min)"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ov_plane_b200 import synth  # noqa: E402
import oracle_backend as ob  # noqa: E402

REF_FLAGS = "-std=c++17 -O3 -fsee -fomit-frame-pointer -fno-signed-zeros -fno-math-errno -funroll-loops -fPIC".split()


def build_fma():
    out = os.path.join(tempfile.gettempdir(), "liboracle_fma.so")
    subprocess.check_call(["g++"] + REF_FLAGS + ["-mfma", "-ffp-contract=fast", "-shared", "-o", out, os.path.join(ROOT, "oracle", "oracle_capi.cpp")])
    return ob.load_variant(out)


def run(name, seed, variant, chi2):
    S = synth.make_scenario(name, seed=seed)
    o = ob.OracleContext(S.options, variant=variant)
    o.set_chi2_table(chi2)
    ch = synth.load_scenario_into(o, S)
    with ob.GaugeProbe(gate_without=False, variant=variant) as gp:
        r = o.msckf_update(synth.feature_batch(S, ch), 1.0, 1.0)
    m = r["plane_status"] != -1
    return S, o.cov(), r["plane_chi2"][m], r["plane_chi2"][m] - gp.junk(), r["plane_status"][m], [rec[2] for rec in gp.records], [rec[0] for rec in gp.records]


def main():
    chi2 = synth.chi2_table()
    fma = build_fma()
    lines = ["# Sensitivity of the reference's plane chi2 to floating-point contraction (CPU oracle, two builds of the same sources)", "",
             "`python tools/oracle_sensitivity.py` — build A: the reference's flags (no FMA); build B: `-mfma -ffp-contract=fast`.", "",
             "| scenario | N | rows kept / rank | plane chi2 as the reference computes it, A | same, B | max abs diff | well-defined part, A | well-defined part, B | max rel diff | gates A | gates B | posterior rel. diff A vs B |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for name, seed in [("tiny_planes", 0), ("tiny_planes", 3), ("small_planes", 0), ("small_planes", 1), ("small_planes", 2), ("small_planes", 3),
                       ("cfg3_n512_f600_p8", 0)]:
        S, Pa, ca, wa, sa, rank, rows = run(name, seed, None, chi2)
        _, Pb, cb, wb, sb, _, _ = run(name, seed, fma, chi2)
        e = np.linalg.norm(Pa - Pb) / np.linalg.norm(Pa)
        f = lambda v: " ".join("%.2f" % x for x in v)
        lines.append("| %s seed %d | %d | %s | %s | %s | %.2f | %s | %s | %.1e | %s | %s | %.1e%s |" % (
            name, seed, S.N, " ".join("%d/%d" % (a, b) for a, b in zip(rows, rank)), f(ca), f(cb), np.abs(ca - cb).max(), f(wa), f(wb),
            np.abs(wa / wb - 1).max(), "".join(map(str, sa)), "".join(map(str, sb)), e, "" if np.array_equal(sa, sb) else " (a gate flipped)"))
        print(lines[-1], flush=True)
    lines += ["", "Reading: the chi2 the reference gates on moves by several units between two builds of the *same* algorithm, because the",
              "`n - rank` kept rows whose Jacobian part is round-off carry round-off defined projections of the residual; its well-defined",
              "part agrees to ~1e-9 and the posterior to ~1e-11 whenever the gates agree.  A plane whose chi2 lies within that band of the",
              "threshold gates differently in the two builds (small_planes seed 2).  The GPU tests therefore compare gates and chi2 on the",
              "well-defined part (tests/test_gpu_parity.py::oracle_msckf_update) and print the reference's own value beside it."]
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
