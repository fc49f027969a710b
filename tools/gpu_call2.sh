#!/bin/bash
# round 2, GPU call 2: parity of the refactored paths, overlap knobs of the TSDF path, reproducibility of the ICP number
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
tail -5 gpurun_out/r2_gputests.log
python tools/time_tsdf.py 512 ER_TSDF_OVERLAP=0 ER_TSDF_OVERLAP=1 ER_TSDF_OVERLAP=1,ER_INT_CTAS=7 ER_TSDF_OVERLAP=1,ER_INT_CTAS=6 ER_TSDF_OVERLAP=0,ER_INT_CTAS=6 > gpurun_out/r2_tsdf_overlap.log 2>&1
cat gpurun_out/r2_tsdf_overlap.log
for k in 1 2 3; do python tools/repeat_icp.py 50 6; done > gpurun_out/r2_icp_repeat.jsonl 2> gpurun_out/r2_icp_repeat.err
ER_ICP_SRC_ORDER=file python tools/repeat_icp.py 50 4 >> gpurun_out/r2_icp_repeat.jsonl 2>> gpurun_out/r2_icp_repeat.err
ER_B200_LIB=elasticreconstruction_b200/variants/libelastic_b200_icp_nostream.so python tools/repeat_icp.py 50 4 >> gpurun_out/r2_icp_repeat.jsonl 2>> gpurun_out/r2_icp_repeat.err
OMP_NUM_THREADS=1 python tools/repeat_icp.py 50 4 >> gpurun_out/r2_icp_repeat.jsonl 2>> gpurun_out/r2_icp_repeat.err
python - <<'P'
import json
for l in open('gpurun_out/r2_icp_repeat.jsonl'):
    d=json.loads(l); print(d['env'], d['ms'], 'median %.0f pair-it/s' % d['pair_it_per_s_median'], 'wall', d['wall_ms'][:3], d['checksums_equal'], d['checksum'])
P
tail -3 gpurun_out/r2_icp_repeat.err
