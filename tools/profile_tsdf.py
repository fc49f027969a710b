"""Small driver for ncu: integrate N resident frames of the configs[1] sequence (no oracle, no bench logic).

    ncu --set full -k regex:integrate_kernel -s 20 -c 2 -o gpurun_out/prof python tools/profile_tsdf.py 30
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elasticreconstruction_b200 import capi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
fusion = int(sys.argv[2]) if len(sys.argv) > 2 else 32
start = int(sys.argv[3]) if len(sys.argv) > 3 else 300
poses = synth.quantize_log(synth.loop_trajectory(1000))[start:start + n]
frames = synth.room_depth(poses)
d = torch.from_numpy(frames.view(np.int16)).cuda()
vol = capi.TSDFVolume(max_blocks=1024)
vol.set_fusion(fusion)
vol.integrate_device(d.data_ptr(), poses, n)
vol.sync()
print("blocks", vol.num_blocks(), "launches", vol.launch_count())
