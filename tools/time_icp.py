"""Quick timing of the ICP path on resident fragments: pair-iterations/s for a few batch shapes.
    python tools/time_icp.py [nfrag] [npairs] [iters]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elasticreconstruction_b200 import capi, synth  # noqa: E402

nfrag = int(sys.argv[1]) if len(sys.argv) > 1 else 12
npairs = int(sys.argv[2]) if len(sys.argv) > 2 else 256
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
frags = [synth.make_local_fragment(i, 200000) for i in range(nfrag)]
clouds = [capi.Cloud(x, n, cell_size=0.03) for (x, n) in frags]
pairs = [(i, j) for i in range(nfrag) for j in range(i + 1, nfrag)]
pairs = (pairs * (npairs // len(pairs) + 1))[:npairs]
T = np.array([synth.pair_init(i, j) for (i, j) in pairs])
tgt, src = [p[0] for p in pairs], [p[1] for p in pairs]
capi.icp_align_batch(clouds, tgt[:16], src[:16], T[:16], 0.03, iters, 1e-6, True)
best = 1e9
for _ in range(3):
    capi.icp_event_record(0, 0)
    To, its, fit = capi.icp_align_batch(clouds, tgt, src, T, 0.03, iters, 1e-6, True)
    capi.icp_event_record(0, 1)
    best = min(best, capi.icp_event_elapsed_ms(0, 0, 1))
print("pairs %d iters %d: %.1f ms -> %.0f pair-iterations/s (%.1f us per pair-iteration)" % (
    len(pairs), int(its.sum()), best, its.sum() / best * 1e3, best * 1e3 / its.sum()), flush=True)
err = max(np.linalg.norm(To[k] - np.linalg.inv(synth.fragment_frame(i)) @ synth.fragment_frame(j)) for k, (i, j) in enumerate(pairs))
print("max pose error vs truth %.2e; checksum %.12f" % (err, float(np.abs(To).sum())), flush=True)
