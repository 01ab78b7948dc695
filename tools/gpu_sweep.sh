#!/bin/bash
# slicer geometry sweep on the 1 h bench workload (device-resident timing only)
mkdir -p gpurun_out
: > gpurun_out/sweep.txt
for seg in 8192 16384 32768; do for warm in 8192 16384 24576 32768; do for chk in 1024 4096; do
  python bench.py --steps 5 --warmup 3 --no-cpu --opt segment_len=$seg --opt warmup_len=$warm --opt checkpoint_len=$chk > gpurun_out/sw.json 2>gpurun_out/sw.err
  python - "$seg" "$warm" "$chk" <<'PY' >> gpurun_out/sweep.txt
import json,sys
try:
    d=json.loads(open('gpurun_out/sw.json').read().strip().splitlines()[-1])
    print(sys.argv[1:], "ms/step %.2f"%d['ms_per_step'], {k:round(v,2) for k,v in d['stage_ms'].items()}, d['slicer'])
except Exception as ex:
    print(sys.argv[1:], "FAILED", ex)
PY
done; done; done
cat gpurun_out/sweep.txt
