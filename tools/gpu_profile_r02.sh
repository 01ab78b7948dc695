#!/bin/bash
# ncu evidence for profiles/ (round 2): (1) launch list with per-launch device time of one device-resident pass,
# (2) --set full capture of the dominant kernels of the third pass.  Each ncu command runs only after the same command
# exited 0 without ncu.  Arguments: tag, then engine options for quick_run.py (e.g. tensor_lpf=0).
set -u
mkdir -p gpurun_out
tag=${1:-r02}; shift
CMD="python tools/quick_run.py 3600 3 $*"
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'afsk_front_kernel|lpf_tc_kernel|slicer_segments_kernel|gather_write_kernel|guard_fixup_kernel|ax25_gap_kernel' -s 12 -c 6 -o gpurun_out/${tag}_prof -f $CMD > gpurun_out/${tag}_ncu_full.log 2>&1
echo "full capture exit $?"
tail -2 gpurun_out/${tag}_plain.log
ls -la gpurun_out | grep ${tag}
