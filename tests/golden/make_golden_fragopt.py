"""Writes tests/golden/fragopt_golden.npz: pose.log and output.ctr of the VERBATIM reference FragmentOptimizer
(oracle/_ref/FragmentOptimizer_ref, `make -C oracle fragopt`) in --slac and --rigid mode on the deterministic cases of
tests/helpers.py (FRAGOPT_CASES).  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden_fragopt.py

One OpenMP thread (OMP_THREAD_LIMIT=1), so that the order in which the reference sums its per-pair matrices is fixed."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers          # noqa: E402
import pyoracle         # noqa: E402


def main():
    pyoracle.build()
    out = {}
    for name, mode in helpers.FRAGOPT_CASE_MODES:
        with tempfile.TemporaryDirectory() as d:
            _, args = helpers.fragopt_case(name, d)
            log = pyoracle.run_reference_fragopt(args + ["--" + mode], cwd=d, threads=1)
            if mode != "nonrigid":                             # OptimizeNonrigid writes no pose.log
                out["%s_%s_pose" % (name, mode)] = helpers.read_log_poses(os.path.join(d, "pose.log"))
            out["%s_%s_ctr" % (name, mode)] = np.loadtxt(os.path.join(d, "output.ctr"))
            scores = [float(l.split(" is ")[1].lstrip(": ")) for l in log.splitlines() if "score is" in l]
            out["%s_%s_scores" % (name, mode)] = np.array(scores)
            print(name, mode, out["%s_%s_ctr" % (name, mode)].shape, scores)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fragopt_golden.npz"), **out)


if __name__ == "__main__":
    main()
