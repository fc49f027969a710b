#!/bin/bash
# round 2, third FragmentOptimizer session (short: ~4 GPU-minutes were left): swizzled staging rows + 96-row batches as the
# default; path tests, A/B against the un-swizzled and the 128-row builds, full-size bench line, smoke, default bench line, ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=$SECONDS
V=elasticreconstruction_b200/variants
timeout 100 python -m pytest tests/test_z_fragopt_gpu.py -x -q --timeout 90 > gpurun_out/r2_fragopt_gputests3.log 2>&1; echo "fragopt pytest rc=$? t=$((SECONDS-S))"; tail -6 gpurun_out/r2_fragopt_gputests3.log
for v in default fo_noswz fo_b128; do
  if [ $v = default ]; then unset ER_B200_LIB; else export ER_B200_LIB=$PWD/$V/libelastic_b200_$v.so; fi
  timeout 60 python tools/profile_fragopt.py 20 100000 1 > gpurun_out/r2_fragopt_ab3_$v.json 2> gpurun_out/r2_fragopt_ab3_$v.err
  echo "variant $v rc=$? t=$((SECONDS-S)) $(python -c "
import json,sys
d=json.load(open('gpurun_out/r2_fragopt_ab3_$v.json')); print('rows', d['rows'], 'kernel ms', d['kernel']['all_ms'], 'TFLOP/s', round(d['roofline']['achieved'],2), 'steps', d['steps_ms'])" 2>&1 | tail -1)"
done
unset ER_B200_LIB
timeout 90 python bench.py --metric fragopt > gpurun_out/r2_bench_fragopt3.json 2> gpurun_out/r2_bench_fragopt3.err; echo "fragopt bench rc=$? t=$((SECONDS-S))"; tail -3 gpurun_out/r2_bench_fragopt3.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench_fragopt3.json')); print('value', d['value'], 'kernel', d['kernel'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', (d.get('cpu_baseline') or {}).get('value'))"
timeout 40 python __graft_entry__.py smoke 2>&1 | tail -3; echo "smoke t=$((SECONDS-S))"
timeout 100 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_final4.json 2> gpurun_out/r2_bench_n1_final4.err; echo "bench rc=$? t=$((SECONDS-S))"; tail -2 gpurun_out/r2_bench_n1_final4.err; python tools/show_bench.py gpurun_out/r2_bench_n1_final4.json 2>&1 | tail -12
timeout 60 ncu --set full --clock-control none --import-source on -k regex:slac_assemble -s 1 -c 1 -f -o gpurun_out/prof_slac_assemble_r2c python tools/profile_fragopt.py 20 100000 1 > gpurun_out/r2_ncu_slac3.log 2>&1; echo "ncu rc=$? t=$((SECONDS-S))"
