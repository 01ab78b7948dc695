#!/bin/bash
# batch test + the extra config lines
mkdir -p gpurun_out
tag=${1:-r02i}
( timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/${tag}_pytest_stages.txt; tail -4 gpurun_out/${tag}_pytest_stages.txt
for c in bpsk_300 qpsk_2400 fsk_9600 afsk_1200; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/${tag}_bench_${c}.json 2> gpurun_out/${tag}_bench_${c}.err
  head -c 250 gpurun_out/${tag}_bench_${c}.json; echo; tail -2 gpurun_out/${tag}_bench_${c}.err
done
