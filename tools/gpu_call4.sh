#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_tsdf_gpu.py tests/test_apps_gpu.py -q -m gpu --timeout 900 > gpurun_out/r2_gputests_tma.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_tma.log
tail -8 gpurun_out/r2_gputests_tma.log
python tools/time_tsdf.py 512 ER_TOUCH_TMA=0 ER_TOUCH_TMA=1 ER_TOUCH_TMA=1,ER_TOUCH_CTAS=2 ER_TOUCH_TMA=1,ER_TOUCH_CTAS=3 ER_TOUCH_TMA=1,ER_TOUCH_CTAS=6 ER_TOUCH_TMA=1,ER_TOUCH_CTAS=8 > gpurun_out/r2_tsdf_tma.log 2>&1
cat gpurun_out/r2_tsdf_tma.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r2_san_memcheck_smoke.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r2_san_memcheck_smoke.log
