#!/bin/bash
# round 2, GPU call 5: timing of the hierarchical touch kernel, ncu captures of every hot kernel, launch list of the bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_tsdf_gpu.py -q -m gpu --timeout 900 -x > gpurun_out/r2_gputests_touch.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_gputests_touch.log
python tools/time_tsdf.py 512 ER_TOUCH_TMA=0 ER_TOUCH_TMA=1 > gpurun_out/r2_tsdf_touch2.log 2>&1; cat gpurun_out/r2_tsdf_touch2.log
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:integrate_kernel -s 2 -c 1 -o gpurun_out/r2_prof_integrate python tools/profile_tsdf.py 96 > gpurun_out/ncu5.log 2>&1
$NCU -k regex:classify_kernel -s 2 -c 1 -o gpurun_out/r2_prof_classify python tools/profile_tsdf.py 96 >> gpurun_out/ncu5.log 2>&1
$NCU -k regex:touch_kernel -s 2 -c 1 -o gpurun_out/r2_prof_touch python tools/profile_tsdf.py 96 >> gpurun_out/ncu5.log 2>&1
ER_TOUCH_TMA=1 $NCU -k regex:touch_tma_kernel -s 2 -c 1 -o gpurun_out/r2_prof_touch_tma python tools/profile_tsdf.py 96 >> gpurun_out/ncu5.log 2>&1
$NCU -k regex:reproject_kernel -s 2 -c 1 -o gpurun_out/r2_prof_reproject python tools/profile_tsdf.py 96 32 300 ctr >> gpurun_out/ncu5.log 2>&1
$NCU -k regex:icp_iter_kernel -s 3 -c 1 -o gpurun_out/r2_prof_icp python tools/profile_icp.py 8 64 5 >> gpurun_out/ncu5.log 2>&1
ls -la gpurun_out/r2_prof_*.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --icp-fragments 8 --ctr-frames 200 --ctr-steps 1 --cpu-frames 0 > gpurun_out/r2_bench_under_ncu.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/r2_launches.csv
python tools/bench_apps.py --out gpurun_out/r2_apps.json > gpurun_out/r2_apps.log 2>&1; echo "apps rc=$?"; tail -c 1500 gpurun_out/r2_apps.log
python -m pytest tests/test_icp_gpu.py -q -m gpu --timeout 900 -x 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1b.json 2> gpurun_out/r2_bench_n1b.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1b.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_n1b.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.0f e2e %.0f frac %.3f touch %.2f cls %.2f int %.2f pipelined %.2f' % (d['value'], d['e2e']['value'], r['frac'], r['touch_us_per_frame'], r['classify_us_per_frame'], r['integrate_us_per_frame'], r['pipelined_us_per_frame']))
c=d['ctrlgrid']; print('ctr %.0f frames/s reproject %.2f us/frame blocks %d' % (c['value'], c['reproject_us_per_frame'], c['allocated_blocks']))
i=d['icp']; print('icp', i['value'], i['passes_ms'], i['e2e']['value'])
P
