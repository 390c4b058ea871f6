#!/bin/bash
# A/B the library variants under variants/ with the bench's device-resident evaluation leg (tools/build_variant.sh).
#   tools/ab.sh NAME [NAME ...]      ("intree" = the library in the package)
for v in "$@"; do
  echo -n "$v: "
  if [ "$v" = intree ]; then unset MAGI_B200_LIB; else export MAGI_B200_LIB=$PWD/variants/libmagi_$v.so; fi
  python bench.py --no-cpu-baseline ${AB_HMC:---no-hmc} --no-nuts --no-build --no-other-configs --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); h=d.get('hmc') or {}
print('evals/s %.3e  ms %.3f  frac %.3f  e2e %.3e  parity %.1e  hmc %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['parity_spot']['max_rel_err'], h.get('samples_per_s')))"
done
