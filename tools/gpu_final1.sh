#!/bin/bash
# round 2, final 1-GPU session: full GPU test suite, smoke, the bench line and both reference arms, executable-level timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/r2_gputests_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_gputests_final.log
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_n1_final.err
python tools/show_bench.py gpurun_out/r2_bench_n1_final.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_ref_final.json 2>> gpurun_out/r2_bench_n1_final.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2_bench_ref_final.json
python bench.py --metric icp --steps 1 --warmup 3 > gpurun_out/r2_bench_icp_final.json 2>> gpurun_out/r2_bench_n1_final.err; echo "icp rc=$?"
python bench.py --impl reference --metric icp --steps 3 --warmup 1 > gpurun_out/r2_bench_icp_ref_final.json 2>> gpurun_out/r2_bench_n1_final.err; echo "icp ref rc=$?"; cut -c1-300 gpurun_out/r2_bench_icp_ref_final.json
python tools/bench_apps.py --out gpurun_out/r2_apps_final.json > gpurun_out/r2_apps_final.log 2>&1; echo "apps rc=$?"; tail -c 1800 gpurun_out/r2_apps_final.log
