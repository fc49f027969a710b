#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
tail -15 gpurun_out/r2_gputests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench_n1.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.0f e2e %.0f rows %d frac %.3f touch %.2f cls %.2f int %.2f pipelined %.2f' % (d['value'], d['e2e']['value'], d['e2e']['result_rows'], r['frac'], r['touch_us_per_frame'], r['classify_us_per_frame'], r['integrate_us_per_frame'], r['pipelined_us_per_frame']))
print('ctr', json.dumps(d['ctrlgrid'])[:600])
i=d['icp']; print('icp', i['value'], i['passes_ms'], i['e2e']['value'], i['cpu_baseline'], i.get('max_pose_diff_vs_oracle'))
print('cpu', d['cpu_baseline'])
P
python bench.py --metric icp --steps 1 --warmup 3 > gpurun_out/r2_bench_icp_n1.json 2>> gpurun_out/r2_bench_n1.err; echo "icp bench rc=$?"
python bench.py --impl reference --metric icp --steps 2 --warmup 1 > gpurun_out/r2_bench_icp_ref.json 2>> gpurun_out/r2_bench_n1.err; echo "icp ref rc=$?"; cat gpurun_out/r2_bench_icp_ref.json | cut -c1-400
