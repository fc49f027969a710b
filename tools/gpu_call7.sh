#!/bin/bash
# round 2, GPU call 7: A/B of the classify prefetch, ICP cooperative vs per-thread search on cell-sorted sources, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_tsdf_gpu.py -q -m gpu --timeout 900 -x 2>&1 | tail -3
python tools/time_tsdf.py 512 > gpurun_out/r2_tsdf_prefetch.log 2>&1
ER_B200_LIB=elasticreconstruction_b200/variants/libelastic_b200_noprefetch.so python tools/time_tsdf.py 512 >> gpurun_out/r2_tsdf_prefetch.log 2>&1
python tools/time_tsdf.py 512 >> gpurun_out/r2_tsdf_prefetch.log 2>&1
cat gpurun_out/r2_tsdf_prefetch.log
python tools/time_icp.py 12 256 20 > gpurun_out/r2_icp_coop.log 2>&1
ER_ICP_COOP=0 python tools/time_icp.py 12 256 20 >> gpurun_out/r2_icp_coop.log 2>&1
ER_ICP_MINB=6 python tools/time_icp.py 12 256 20 >> gpurun_out/r2_icp_coop.log 2>&1
cat gpurun_out/r2_icp_coop.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"touch|resolve|classify|integrate|extract|reproject|zbuf|icp_iter|grid_|count_inliers|correspond|compact|scan_counts|pair_base|lambda" -c 3000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --icp-fragments 8 --ctr-frames 200 --ctr-steps 1 --cpu-frames 0 > gpurun_out/r2_bench_under_ncu.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/r2_launches.csv
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:classify_kernel -s 2 -c 1 -o gpurun_out/r2_prof_classify_prefetch python tools/profile_tsdf.py 96 > gpurun_out/ncu7.log 2>&1
$NCU -k regex:touch_kernel -s 2 -c 1 -o gpurun_out/r2_prof_touch python tools/profile_tsdf.py 96 >> gpurun_out/ncu7.log 2>&1
