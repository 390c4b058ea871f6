#!/bin/bash
for v in "$@"; do
  for dense in "" 1; do
  echo -n "$v dense=$dense: "
  MAGI_BENCH_DENSE=$dense MAGI_B200_LIB=$PWD/variants/libmagi_$v.so python bench.py --no-cpu-baseline --no-hmc --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('evals/s %.3e  frac %.3f' % (d['value'], d['roofline']['frac']))"
  done
done
