"""Reproducibility of the ICP throughput number: N timed passes of the configs[2] share in ONE process; run it from
several fresh processes (tools/repeat_icp.sh) and compare min / median / max across all of them.

    python tools/repeat_icp.py [nfrag=50] [passes=10] [points=200000]

Prints one JSON line: per-pass device milliseconds (CUDA events on the ICP stream around er_icp_align_batch), their
min / median / max and the pose checksum (identical across passes and processes when the results are reproducible).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elasticreconstruction_b200 import capi, synth  # noqa: E402

nfrag = int(sys.argv[1]) if len(sys.argv) > 1 else 50
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 10
points = int(sys.argv[3]) if len(sys.argv) > 3 else 200000
iters = 20
t0 = time.perf_counter()
frags = [synth.make_local_fragment(i, points) for i in range(nfrag)]
t_gen = time.perf_counter() - t0
t0 = time.perf_counter()
clouds = [capi.Cloud(x, n, cell_size=0.03) for (x, n) in frags]
t_up = time.perf_counter() - t0
pairs = [(i, j) for i in range(nfrag) for j in range(i + 1, nfrag)]
T = np.array([synth.pair_init(i, j) for (i, j) in pairs])
tgt, src = [p[0] for p in pairs], [p[1] for p in pairs]
ms, wall, sums = [], [], []
for k in range(passes + 1):                       # pass 0 is the warm-up (first-touch allocations of the arena)
    capi.icp_event_record(0, 0)
    w0 = time.perf_counter()
    To, its, fit = capi.icp_align_batch(clouds, tgt, src, T, 0.03, iters, 1e-6, True)
    w1 = time.perf_counter()
    capi.icp_event_record(0, 1)
    if k:
        ms.append(capi.icp_event_elapsed_ms(0, 0, 1))
        wall.append((w1 - w0) * 1e3)
        sums.append(float(np.abs(To).sum()))
n_it = float(its.sum())
ms = np.array(ms)
print(json.dumps({"pid": os.getpid(), "pairs": len(pairs), "pair_iterations": n_it, "gen_s": round(t_gen, 2), "upload_s": round(t_up, 3),
                  "ms": [round(float(x), 2) for x in ms], "wall_ms": [round(float(x), 2) for x in wall],
                  "min_ms": float(ms.min()), "median_ms": float(np.median(ms)), "max_ms": float(ms.max()),
                  "pair_it_per_s_median": n_it / (float(np.median(ms)) * 1e-3),
                  "max_over_min": float(ms.max() / ms.min()), "checksums_equal": len(set(sums)) == 1, "checksum": sums[0],
                  "env": {k: v for k, v in os.environ.items() if k.startswith("ER_ICP") or k == "OMP_NUM_THREADS"}}), flush=True)
