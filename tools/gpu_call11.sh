#!/bin/bash
# round 2, 2-GPU call: three-stage pipeline (classify on its own stream) at N = 1 and N = 2, div1000 parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_tsdf_gpu.py -q -m gpu --timeout 900 -x 2>&1 | tail -3
python tools/time_tsdf.py 512 ER_TSDF_CLS_STREAM=0 ER_TSDF_CLS_STREAM=1 ER_TSDF_CLS_STREAM=0 ER_TSDF_CLS_STREAM=1 > gpurun_out/r2_tsdf_cls_stream.log 2>&1; cat gpurun_out/r2_tsdf_cls_stream.log
for S in 0 1; do
  ER_TSDF_CLS_STREAM=$S python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$S bench.py --gpus 2 --steps 10 --warmup 3 --no-icp --no-ctr --no-replica > gpurun_out/r2_bench_n2_cls$S.json 2> gpurun_out/r2_bench_n2_cls$S.err; echo "cls_stream=$S rc=$?"
  python tools/show_bench.py gpurun_out/r2_bench_n2_cls$S.json | head -3
done
