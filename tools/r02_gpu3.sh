#!/bin/bash
# selected GPU tests (arguments: tag, then pytest selectors) + one bench line
mkdir -p gpurun_out
tag=$1; shift
( timeout 900 python -m pytest "$@" -m gpu -x -q 2>&1 | tail -50 ) > gpurun_out/${tag}_pytest.txt
( timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
tail -6 gpurun_out/${tag}_pytest.txt
head -c 300 gpurun_out/${tag}_bench_n1.json; echo; tail -3 gpurun_out/${tag}_bench_n1.err
