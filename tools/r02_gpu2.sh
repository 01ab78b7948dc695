#!/bin/bash
# GPU tests + one bench line (N=1)
mkdir -p gpurun_out
tag=${1:-r02b}
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/${tag}_pytest.txt
( timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
tail -4 gpurun_out/${tag}_pytest.txt
head -c 600 gpurun_out/${tag}_bench_n1.json; tail -3 gpurun_out/${tag}_bench_n1.err
