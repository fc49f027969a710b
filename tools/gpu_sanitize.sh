#!/bin/bash
# compute-sanitizer memcheck + racecheck of the final kernels on small workloads (smoke: 2 TSDF frames + one ICP pair; a
# ControlGrid batch; the async extraction)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/san_drive.py <<'P'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import helpers
from elasticreconstruction_b200 import capi
sc = helpers.scene_room(n=3)
vol = capi.TSDFVolume(max_blocks=256)
vol.integrate_batch(sc["frames"], sc["traj"][:3])
n = vol.extract_count()
buf = capi.HostBuffer((n + 100) * 16)
vol.extract_begin(buf, n + 100)
print("rows", vol.extract_wait(), n)
c = helpers.scene_ctr(num=2, interval=2)
traj = helpers.compose_traj(c["pose_traj"], c["seg_traj"], c["interval"], capi.mat4_mul)
v2 = capi.TSDFVolume(max_blocks=256)
v2.set_control_grids(c["grids"], c["res"], c["length"])
k = len(c["frames"])
v2.integrate_ctr_batch(c["frames"], traj[:k], traj[0], c["seg_traj"][0], c["seg_traj"][:k], [f // c["interval"] for f in range(k)])
print("ctr blocks", v2.num_blocks())
P
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san_drive.py > gpurun_out/r2_san_memcheck_tsdf.log 2>&1; echo "memcheck tsdf rc=$?"; tail -4 gpurun_out/r2_san_memcheck_tsdf.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/san_drive.py > gpurun_out/r2_san_racecheck_tsdf.log 2>&1; echo "racecheck tsdf rc=$?"; tail -4 gpurun_out/r2_san_racecheck_tsdf.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r2_san_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -4 gpurun_out/r2_san_memcheck_smoke.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r2_san_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -4 gpurun_out/r2_san_racecheck_smoke.log
