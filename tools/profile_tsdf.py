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
if len(sys.argv) > 4 and sys.argv[4] == "ctr":
    # ControlGrid mode (BASELINE configs[3] scene of bench.py) for a capture of reproject_kernel
    import bench
    sc = bench.ctrlgrid_scene(max(n, 100))
    traj = np.array([capi.mat4_mul(sc["pose"][f // sc["interval"]], sc["seg"][f]) for f in range(n)])
    dfr = bench.room_depth_device(traj, torch.device("cuda", 0), lo=0.25, hi=5.75)
    v2 = capi.TSDFVolume(max_blocks=4096)
    v2.set_control_grids(sc["grids"], sc["res"], sc["length"])
    v2.integrate_ctr_device(dfr.data_ptr(), traj, traj[0], sc["seg"][0], sc["seg"][:n], (np.arange(n) // sc["interval"]).astype(np.int32))
    v2.sync()
    print("ctr blocks", v2.num_blocks())
