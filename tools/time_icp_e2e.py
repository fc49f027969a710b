"""Where the end-to-end ICP time goes: fragment upload (host grid build + H2D) with 1/8/16 host threads, then the
first and second align_batch call over all pairs (wall clock).
    python tools/time_icp_e2e.py [nfrag]
"""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elasticreconstruction_b200 import capi, synth  # noqa: E402

nfrag = int(sys.argv[1]) if len(sys.argv) > 1 else 50
frags = [synth.make_local_fragment(i, 200000) for i in range(nfrag)]
pairs = [(i, j) for i in range(nfrag) for j in range(i + 1, nfrag)]
T = np.array([synth.pair_init(i, j) for (i, j) in pairs])
tgt, src = [p[0] for p in pairs], [p[1] for p in pairs]

t0 = time.perf_counter()
c0 = capi.Cloud(frags[0][0], frags[0][1], cell_size=0.03)
print("first upload (context, stream, module load): %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
for nt in (1, 8, 16, 32):
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=nt) as ex:
        clouds = list(ex.map(lambda xn: capi.Cloud(xn[0], xn[1], cell_size=0.03), frags))
    print("upload of %d fragments with %2d threads: %.1f ms" % (nfrag, nt, (time.perf_counter() - t0) * 1e3), flush=True)
for k in range(3):
    t0 = time.perf_counter()
    To, its, fit = capi.icp_align_batch(clouds, tgt, src, T, 0.03, 20, 1e-6, True)
    dt = time.perf_counter() - t0
    print("align_batch call %d: %.1f ms wall for %d pair-iterations -> %.0f /s" % (k, dt * 1e3, its.sum(), its.sum() / dt), flush=True)
