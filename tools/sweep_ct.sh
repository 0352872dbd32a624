#!/bin/bash
# tuning sweep on the full bench workload: first-level bits x count-table slots
for b in 8 9 10; do for s in 8192; do
  ORION_B1=$b ORION_CT_SLOTS=$s timeout 300 python bench.py --steps 2 --warmup 1 --sample-reads 2000 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b1', $b, 'slots', $s, 'ms_per_step', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['phases_ms'].items()})"
done; done
