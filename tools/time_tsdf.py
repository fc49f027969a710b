"""Quick A/B timing of the TSDF path on resident frames (no oracle): prints us/frame per variant.
    python tools/time_tsdf.py [nframes]        variants are selected through environment knobs read at create time
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elasticreconstruction_b200 import capi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
poses = synth.quantize_log(synth.loop_trajectory(1000))[300:300 + n]
frames = synth.room_depth(poses)
d = torch.from_numpy(frames.view(np.int16)).cuda()


def run(tag, env, fusion=32, reps=4):
    for k, v in env.items():
        os.environ[k] = v
    vol = capi.TSDFVolume(max_blocks=1024)
    vol.set_fusion(fusion)
    vol.integrate_device(d.data_ptr(), poses, n)        # warm-up (allocates the blocks)
    vol.sync()
    best = 1e9
    for _ in range(reps):
        vol.event_record(0)
        vol.integrate_device(d.data_ptr(), poses, n)
        vol.event_record(1)
        vol.sync()
        best = min(best, vol.event_elapsed_ms(0, 1))
    t, c, i = vol.profile_device(d.data_ptr(), poses, n)
    print("%-28s %7.2f us/frame  (touch %.2f classify %.2f integrate %.2f)" % (tag, 1e3 * best / n, 1e3 * t / n, 1e3 * c / n, 1e3 * i / n), flush=True)
    for k in env:
        os.environ.pop(k, None)
    del vol


def class_stats():
    vol = capi.TSDFVolume(max_blocks=1024)
    vol.integrate_device(d.data_ptr(), poses, n)
    vol.set_stats(True)
    s0 = vol.stats()
    vol.integrate_device(d.data_ptr(), poses, n)
    s1 = vol.stats()
    print({k: (s1[k] - s0[k]) / n for k in s1 if k != "frames"}, flush=True)


if __name__ == "__main__":
    # python tools/time_tsdf.py 256 ER_INT_THREADS=256 ER_INT_THREADS=128 ...   (one run per KEY=VALUE[,KEY=VALUE])
    for spec in sys.argv[2:] or [""]:
        env = dict(kv.split("=") for kv in spec.split(",") if kv)
        run(spec or "default", env)
