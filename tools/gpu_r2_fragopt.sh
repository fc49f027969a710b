#!/bin/bash
# round 2, FragmentOptimizer session: its GPU tests, a small and the full-size bench line, ncu of the assembly kernel,
# then the whole GPU suite, smoke and the default bench line (which now embeds the fragopt leg)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=$SECONDS
timeout 300 python -m pytest tests/test_z_fragopt_gpu.py -x -q --timeout 280 > gpurun_out/r2_fragopt_gputests.log 2>&1; echo "fragopt pytest rc=$? t=$((SECONDS-S))"; tail -15 gpurun_out/r2_fragopt_gputests.log
timeout 120 python tools/profile_fragopt.py 8 50000 2 > gpurun_out/r2_fragopt_small.json 2> gpurun_out/r2_fragopt_small.err; echo "small rc=$? t=$((SECONDS-S))"; tail -3 gpurun_out/r2_fragopt_small.err; cut -c1-1500 gpurun_out/r2_fragopt_small.json
timeout 400 python bench.py --metric fragopt > gpurun_out/r2_bench_fragopt.json 2> gpurun_out/r2_bench_fragopt.err; echo "fragopt bench rc=$? t=$((SECONDS-S))"; tail -3 gpurun_out/r2_bench_fragopt.err; cut -c1-3000 gpurun_out/r2_bench_fragopt.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:slac_assemble -s 1 -c 1 -f -o gpurun_out/prof_slac_assemble_r2 python tools/profile_fragopt.py 20 100000 1 > gpurun_out/r2_ncu_slac.log 2>&1; echo "ncu rc=$? t=$((SECONDS-S))"; tail -2 gpurun_out/r2_ncu_slac.log | cut -c1-300
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_fragopt_launches.csv python tools/profile_fragopt.py 20 100000 1 > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$? t=$((SECONDS-S))"
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -4; echo "smoke t=$((SECONDS-S))"
timeout 400 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2_gputests_final2.log 2>&1; echo "pytest rc=$? t=$((SECONDS-S))"; tail -4 gpurun_out/r2_gputests_final2.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_final2.json 2> gpurun_out/r2_bench_n1_final2.err; echo "bench rc=$? t=$((SECONDS-S))"; tail -2 gpurun_out/r2_bench_n1_final2.err; python tools/show_bench.py gpurun_out/r2_bench_n1_final2.json 2>&1 | tail -30
