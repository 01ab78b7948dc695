#!/bin/bash
# 8-GPU box: copy ceiling at 1/2/4/8 ranks, then bench lines (weak + strong) at N = 8, 4, 2 back to back
mkdir -p gpurun_out
tag=${1:-r02j}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29556 tools/h2d_ceiling.py > gpurun_out/${tag}_h2d_ceiling.txt 2> gpurun_out/${tag}_h2d_ceiling.err
cat gpurun_out/${tag}_h2d_ceiling.txt
for N in 8 4 2; do
  for mode in weak strong; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 10 --warmup 3 --scaling $mode > gpurun_out/${tag}_bench_${mode}_n${N}.json 2> gpurun_out/${tag}_bench_${mode}_n${N}.err
    python -c "
import json,sys
d=json.load(open('gpurun_out/${tag}_bench_${mode}_n${N}.json'))
print('$mode N=$N value %.1f G ms %.3f e2e %.1f G parity %s fallbacks %s' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['parity']['match'], d['link_fallbacks']))
" 2>&1 | tail -1
  done
done
( timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q 2>&1 | tail -5 ) > gpurun_out/${tag}_pytest_multirank.txt; tail -3 gpurun_out/${tag}_pytest_multirank.txt
