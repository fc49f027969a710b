#!/bin/bash
# round 2, 8-GPU call: the scaling bench at N = 8 and N = 4, --devices tests on 4 devices
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench n$N rc=$?"
  grep -A10 "^\[rank0\]: Traceback" gpurun_out/r2_bench_n$N.err | head -20
  python tools/show_bench.py gpurun_out/r2_bench_n$N.json
done
python -m pytest tests/test_apps_gpu.py -q -m gpu --timeout 900 -k "devices" 2>&1 | tail -3
