#!/bin/bash
# round 2, first GPU call: all GPU tests, guard calibration on the device, one bench line, the tensor-core FIR experiment
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt 2>&1
( timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 ) > gpurun_out/r02a_pytest.txt
( timeout 120 tools/ubench/fir_umma 0; echo "rc=$?"; timeout 120 tools/ubench/fir_umma 1; echo "rc=$?" ) > gpurun_out/r02a_fir_umma.txt 2>&1
( timeout 300 python tools/guard_bound.py 4 --gpu ) > gpurun_out/r02a_guard_gpu.txt 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err
tail -5 gpurun_out/r02a_pytest.txt
cat gpurun_out/r02a_fir_umma.txt | tail -20
head -c 1500 gpurun_out/r02a_bench_n1.json
