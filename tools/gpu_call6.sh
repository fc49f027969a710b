#!/bin/bash
# round 2, 2-GPU call: in-process --devices of both executables, the block-partitioned bench mode, union check
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_apps_gpu.py -q -m gpu --timeout 900 -k "devices" > gpurun_out/r2_gputests_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_2gpu.log; tail -6 gpurun_out/r2_gputests_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/r2_bench_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
r=d['roofline']
print('N=2 value %.0f e2e %.0f (%s) scaling %s' % (d['value'], d['e2e']['value'], d['e2e']['feed'][:40], d['scaling']))
print('union', d['union_check'])
print('replica', d['replica'])
print('touch %.2f cls %.2f int %.2f pipelined %.2f' % (r['touch_us_per_frame'], r['classify_us_per_frame'], r['integrate_us_per_frame'], r['pipelined_us_per_frame']))
c=d['ctrlgrid']; print('ctr %.0f frames/s' % c['value'])
i=d['icp']; print('icp', i['value'], i['passes_ms'], 'e2e', i['e2e']['value'])
P
python tools/bench_apps.py --devices 0,1 --out gpurun_out/r2_apps_2gpu.json > gpurun_out/r2_apps_2gpu.log 2>&1; echo "apps rc=$?"; tail -c 1200 gpurun_out/r2_apps_2gpu.log
