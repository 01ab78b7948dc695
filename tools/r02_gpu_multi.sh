#!/bin/bash
# multi-GPU: real-rank parity test + bench lines (weak and strong) at N = $1
mkdir -p gpurun_out
N=$1; tag=${2:-r02m}
if [ "$N" -le 4 ]; then ( timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x 2>&1 | tail -30 ) > gpurun_out/${tag}_pytest_n${N}.txt; tail -5 gpurun_out/${tag}_pytest_n${N}.txt; fi
for mode in weak strong; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 10 --warmup 3 --scaling $mode > gpurun_out/${tag}_bench_${mode}_n${N}.json 2> gpurun_out/${tag}_bench_${mode}_n${N}.err
  head -c 400 gpurun_out/${tag}_bench_${mode}_n${N}.json; echo; tail -2 gpurun_out/${tag}_bench_${mode}_n${N}.err
done
