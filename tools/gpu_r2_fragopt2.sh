#!/bin/bash
# round 2, second FragmentOptimizer session: two-plane shared-memory layout (new default) + nonrigid mode.
# GPU tests of the path, A/B of the layout / batch variants at mid size, full-size bench line, ncu, nonrigid probe,
# then the whole GPU suite, smoke and the default bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=$SECONDS
V=elasticreconstruction_b200/variants
timeout 300 python -m pytest tests/test_z_fragopt_gpu.py -x -q --timeout 280 > gpurun_out/r2_fragopt_gputests2.log 2>&1; echo "fragopt pytest rc=$? t=$((SECONDS-S))"; tail -12 gpurun_out/r2_fragopt_gputests2.log
for v in default fo_natural fo_b96 fo_b64; do
  if [ $v = default ]; then unset ER_B200_LIB; else export ER_B200_LIB=$PWD/$V/libelastic_b200_$v.so; fi
  if [ $v != default ]; then timeout 120 python -m pytest tests/test_z_fragopt_gpu.py -q -k "kernels_against and not nonrigid" --timeout 100 2>&1 | tail -1; fi
  timeout 120 python tools/profile_fragopt.py 20 100000 1 > gpurun_out/r2_fragopt_ab_$v.json 2> gpurun_out/r2_fragopt_ab_$v.err
  echo "variant $v rc=$? t=$((SECONDS-S)) $(python -c "
import json,sys
d=json.load(open('gpurun_out/r2_fragopt_ab_$v.json')); print('rows', d['rows'], 'kernel ms', d['kernel']['all_ms'], 'TFLOP/s', round(d['roofline']['achieved'],2), 'steps', d['steps_ms'])" 2>&1 | tail -1)"
done
unset ER_B200_LIB
timeout 300 python bench.py --metric fragopt > gpurun_out/r2_bench_fragopt2.json 2> gpurun_out/r2_bench_fragopt2.err; echo "fragopt bench rc=$? t=$((SECONDS-S))"; tail -3 gpurun_out/r2_bench_fragopt2.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench_fragopt2.json')); print('value', d['value'], 'kernel', d['kernel'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', (d.get('cpu_baseline') or {}).get('value'))"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:slac_assemble -s 1 -c 1 -f -o gpurun_out/prof_slac_assemble_r2b python tools/profile_fragopt.py 20 100000 1 > gpurun_out/r2_ncu_slac2.log 2>&1; echo "ncu rc=$? t=$((SECONDS-S))"
timeout 200 python tools/profile_fragopt.py 8 50000 1 nonrigid > gpurun_out/r2_nonrigid_probe.json 2> gpurun_out/r2_nonrigid_probe.err; echo "nonrigid probe rc=$? t=$((SECONDS-S))"; tail -2 gpurun_out/r2_nonrigid_probe.err; cat gpurun_out/r2_nonrigid_probe.json
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -4; echo "smoke t=$((SECONDS-S))"
timeout 400 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2_gputests_final3.log 2>&1; echo "pytest rc=$? t=$((SECONDS-S))"; tail -4 gpurun_out/r2_gputests_final3.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_final3.json 2> gpurun_out/r2_bench_n1_final3.err; echo "bench rc=$? t=$((SECONDS-S))"; tail -2 gpurun_out/r2_bench_n1_final3.err; python tools/show_bench.py gpurun_out/r2_bench_n1_final3.json 2>&1 | tail -12
