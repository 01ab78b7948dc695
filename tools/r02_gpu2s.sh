#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02x}
for seg in 8192 24576; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 10 --warmup 3 --scaling strong --opt segment_len=$seg > gpurun_out/${tag}_bench_strong_n2_seg$seg.json 2> gpurun_out/${tag}_err.txt
python -c "
import json
d=json.load(open('gpurun_out/${tag}_bench_strong_n2_seg$seg.json'))
print('strong N=2 seg $seg value %.1f G ms %.3f parity %s fallbacks %s' % (d['value']/1e9, d['ms_per_step'], d['parity']['match'], d['link_fallbacks']), {k: round(v,3) for k,v in d['stage_ms'].items()}, d['slicer'])
"
done
