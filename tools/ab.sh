#!/bin/bash
# A/B the library variants under variants/ with the bench's device-resident evaluation leg.
for v in "$@"; do
  echo -n "$v: "
  MAGI_B200_LIB=$PWD/variants/libmagi_$v.so python bench.py --no-cpu-baseline --no-hmc --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('evals/s %.3e  frac %.3f  e2e %.3e' % (d['value'], d['roofline']['frac'], d['e2e']['value']))"
done
