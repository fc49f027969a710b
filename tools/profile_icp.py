"""Small driver for ncu: a few ICP pair-alignments on resident fragments.
    ncu --set full -k regex:icp_iter_kernel -s 2 -c 1 -o gpurun_out/prof_icp python tools/profile_icp.py 8 64
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elasticreconstruction_b200 import capi, synth  # noqa: E402

nfrag = int(sys.argv[1]) if len(sys.argv) > 1 else 8
npairs = int(sys.argv[2]) if len(sys.argv) > 2 else 28
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
frags = [synth.make_local_fragment(i, 200000) for i in range(nfrag)]
clouds = [capi.Cloud(x, n, cell_size=0.03) for (x, n) in frags]
pairs = [(i, j) for i in range(nfrag) for j in range(i + 1, nfrag)]
pairs = (pairs * (npairs // len(pairs) + 1))[:npairs]
T = np.array([synth.pair_init(i, j) for (i, j) in pairs])
To, its, fit = capi.icp_align_batch(clouds, [p[0] for p in pairs], [p[1] for p in pairs], T, 0.03, iters, 1e-6, True)
print("pairs", len(pairs), "iters", its.sum(), "launches", capi.icp_launch_count(0))
