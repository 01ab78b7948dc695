#!/bin/bash
# 8-GPU box: the strong-scaling lines (one hour in total) at N = 8 and 4
mkdir -p gpurun_out
tag=${1:-r02w}
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong > gpurun_out/${tag}_bench_strong_n${N}.json 2> gpurun_out/${tag}_bench_strong_n${N}.err
  python -c "
import json
d=json.load(open('gpurun_out/${tag}_bench_strong_n${N}.json'))
print('strong N=$N value %.1f G ms %.3f e2e %.1f G parity %s fallbacks %s' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['parity']['match'], d['link_fallbacks']), {k: round(v,3) for k,v in d['stage_ms'].items()}, d['slicer'])
"
done
