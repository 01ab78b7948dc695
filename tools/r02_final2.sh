#!/bin/bash
# final single-GPU evidence of round 2, second session: all GPU tests, the bench line (driver flags; it carries the
# full-hour packet-set digest), ncu launch list + full capture (tools/gpu_profile_r02.sh).  (The r02ag run also had
# tools/tile_sweep.py and tools/e2e_trace.py in it.)
mkdir -p gpurun_out
tag=${1:-r02am}
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/${tag}_pytest_gpu.txt; tail -3 gpurun_out/${tag}_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; head -c 200 gpurun_out/${tag}_bench_n1.json; echo
timeout 500 bash tools/gpu_profile_r02.sh ${tag}
