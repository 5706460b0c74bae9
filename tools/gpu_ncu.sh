#!/bin/bash
# One or more full ncu captures of named kernels on a golden case.
# usage (under gpurun): bash tools/gpu_ncu.sh <tag> <kernel regex> <count> <time_case.py args...>
TAG=$1; KREGEX=$2; COUNT=$3; shift; shift; shift
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -c $COUNT -o /tmp/prof_$TAG \
    python tools/time_case.py "$@" > $OUT/ncu_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > $OUT/ncu_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --print-source sass > $OUT/ncu_${TAG}_source.csv 2>/dev/null
tail -2 $OUT/ncu_$TAG.log
