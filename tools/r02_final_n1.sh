#!/bin/bash
# final single-GPU evidence of the round: all GPU tests, the bench line (driver flags), the reference arm, the guard
# calibration on the device, the full-hour packet-set check, the extra config lines
mkdir -p gpurun_out
tag=${1:-r02z}
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/${tag}_pytest_gpu.txt; tail -3 gpurun_out/${tag}_pytest_gpu.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; head -c 200 gpurun_out/${tag}_bench_n1.json; echo
timeout 900 python bench.py --steps 20 --warmup 5 --opt tensor_lpf=0 > gpurun_out/${tag}_bench_n1_ffma_lpf.json 2> gpurun_out/${tag}_bench_n1_ffma_lpf.err; head -c 200 gpurun_out/${tag}_bench_n1_ffma_lpf.json; echo
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err; head -c 200 gpurun_out/${tag}_bench_reference_arm.json; echo
( timeout 300 python tools/guard_bound.py 4 --gpu ) > gpurun_out/${tag}_guard_calibration.txt 2>&1; tail -3 gpurun_out/${tag}_guard_calibration.txt
( timeout 600 python tools/verify_hour.py ) > gpurun_out/${tag}_verify_hour.txt 2>&1; tail -2 gpurun_out/${tag}_verify_hour.txt
for c in bpsk_300 qpsk_2400 fsk_9600 afsk_1200; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/${tag}_bench_${c}.json 2> gpurun_out/${tag}_bench_${c}.err; head -c 160 gpurun_out/${tag}_bench_${c}.json; echo
done
( timeout 300 python tools/tc_clocks.py 600 0 ) > gpurun_out/${tag}_tc_clocks.txt 2>&1; tail -2 gpurun_out/${tag}_tc_clocks.txt
( timeout 300 python tools/slicer_sweep_short.py 450 ) > gpurun_out/${tag}_slicer_sweep_short.txt 2>&1; tail -3 gpurun_out/${tag}_slicer_sweep_short.txt
