"""Achievable host->device bandwidth of this box for the chunk sizes the TSDF feed uses (page-locked source from the
library's allocator), one and two copy streams: the ceiling of the e2e TSDF number at N = 1 (614 400 B per frame)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elasticreconstruction_b200 import capi  # noqa: E402

torch.cuda.set_device(0)
total = 614_400_000
hb = capi.HostBuffer(total)
src = torch.from_numpy(hb.array(np.uint8))
dst = torch.empty(total, dtype=torch.uint8, device="cuda")
for chunk in (614_400 * 32, 614_400 * 128, total):
    for nstreams in (1, 2):
        streams = [torch.cuda.Stream() for _ in range(nstreams)]
        best = 0.0
        for rep in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            k = 0
            for off in range(0, total, chunk):
                with torch.cuda.stream(streams[k % nstreams]):
                    dst[off:off + chunk].copy_(src[off:off + chunk], non_blocking=True)
                k += 1
            torch.cuda.synchronize()
            best = max(best, total / (time.perf_counter() - t0) / 1e9)
        print("chunk %6.1f MB, %d stream(s): %.1f GB/s  -> %.0f frames/s" % (chunk / 1e6, nstreams, best, best * 1e9 / 614400), flush=True)
