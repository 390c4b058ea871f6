#!/bin/bash
# On the GPU box: capture one launch of a kernel with --set full and export the raw and source pages as
# CSV (the .ncu-rep itself can exceed the 64 MiB gpurun_out limit and is dropped).
#   tools/ncu_capture.sh <name> <kernel regex> <skip> -- <command...>
NAME=$1; KERN=$2; SKIP=$3; shift 4
"$@" > gpurun_out/${NAME}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${NAME}_plain.log; exit 1; }
ncu --set full --clock-control none -k regex:$KERN -s $SKIP -c 1 -o /tmp/$NAME "$@" > gpurun_out/${NAME}_ncu.log 2>&1
ncu -i /tmp/$NAME.ncu-rep --page raw --csv > gpurun_out/${NAME}_raw.csv 2>/dev/null
ncu -i /tmp/$NAME.ncu-rep --page source --csv > gpurun_out/${NAME}_source.csv 2>/dev/null
gzip -f gpurun_out/${NAME}_source.csv
ls -la gpurun_out | tail -5
