#!/bin/bash
for c in 4 8 16 32; do for st in 2 3 4; do
echo -n "chunks=$c streams=$st: "
MAGI_E2E_CHUNKS=$c MAGI_E2E_STREAMS=$st python bench.py --no-cpu-baseline --no-hmc --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('e2e %.3e  (%.2f ms/step)' % (d['e2e']['value'], d['e2e']['ms_per_step']))"
done; done
