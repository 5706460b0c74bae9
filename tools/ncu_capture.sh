#!/bin/bash
# usage: tools/ncu_capture.sh <tag> <kernel-regex> <skip> <count> -- <command...>
# Runs <command> plainly first (must exit 0), then once under `ncu --set full`; exports the raw and source pages as CSV into
# gpurun_out/ and removes the (large) .ncu-rep so that gpurun_out stays below the 64 MiB that travel back.
tag=$1; rx=$2; skip=$3; cnt=$4; shift 5
"$@" > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o /tmp/prof_$tag "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page source --csv > gpurun_out/ncu_${tag}_source.csv 2>/dev/null
ls -la /tmp/prof_$tag.ncu-rep gpurun_out/ncu_${tag}_*.csv
