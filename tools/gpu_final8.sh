#!/bin/bash
# round 2, final 8-GPU session: the scaling bench at N = 8, 4, 2 with the final code
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n${N}_final.json 2> gpurun_out/r2_bench_n${N}_final.err; echo "bench n$N rc=$?"
  grep -A10 "^\[rank0\]: Traceback" gpurun_out/r2_bench_n${N}_final.err | head -20
  python tools/show_bench.py gpurun_out/r2_bench_n${N}_final.json
done
