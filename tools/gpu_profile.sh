#!/bin/bash
# ncu evidence for profiles/: (1) launch list with per-launch device time, (2) --set full capture of the
# dominant kernels.  Each ncu command runs only after the same command exited 0 without ncu.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --seconds ${1:-3600} --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'afsk_front_kernel|slicer_segments_kernel|gather_write_kernel|guard_fixup_kernel' -s 40 -c 8 -o gpurun_out/prof_r01 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out
