#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_tsdf_gpu.py tests/test_apps_gpu.py -q -m gpu --timeout 900 2>&1 | tail -3
python tools/time_tsdf.py 512 ER_TSDF_BUFS=2 ER_TSDF_BUFS=3 ER_TSDF_BUFS=2 ER_TSDF_BUFS=3 > gpurun_out/r2_tsdf_bufs.log 2>&1; cat gpurun_out/r2_tsdf_bufs.log
for Bn in 2 3; do
  ER_TSDF_BUFS=$Bn python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$Bn bench.py --gpus 2 --steps 10 --warmup 3 --no-icp --no-ctr --no-replica > gpurun_out/r2_bench_n2_bufs$Bn.json 2> gpurun_out/r2_bench_n2_bufs$Bn.err; echo "bufs=$Bn rc=$?"
  python tools/show_bench.py gpurun_out/r2_bench_n2_bufs$Bn.json | head -3
done
