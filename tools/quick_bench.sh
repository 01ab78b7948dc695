#!/bin/bash
# quick single-GPU check: headline parity tests + a short bench; prints the stage times
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "superopt or 44k1 or fsk9600_ax25 or noise_only or clipped or ragged" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],3), 'G/s', round(d['value']/1e9,2), {k: round(v,3) for k,v in d['stage_ms'].items()}, 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['ms_per_step'],3))"
