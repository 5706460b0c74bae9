#!/bin/bash
# GPU-box session for the full-FLD path (configs[2]): parity tests of the full-FLD fixtures, timings, one full ncu capture.
TAG=${1:-r2x}
KREGEX=${2:-k_full_lu3}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "c3 or c6 or n1 or n2 or t1 or p1 or g1 or full" > $OUT/gpu_tests_c3_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/gpu_tests_c3_$TAG.log
timeout 300 python tools/time_case.py c3_searth_22b --ncols 4096 --spinup 12 --steps 10 > $OUT/time_c3_$TAG.json 2>&1
timeout 300 python tools/time_case.py c3_searth_22b --ncols 4096 --spinup 12 --steps 10 --ncols 2072 > $OUT/time_c3_2072_$TAG.json 2>&1
timeout 300 python tools/time_case.py c3_searth_22b --ncols 512 --spinup 12 --steps 10 > $OUT/time_c3_512_$TAG.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_c3_$TAG.csv \
    python tools/time_case.py c3_searth_22b --ncols 4096 --spinup 2 --steps 2 > $OUT/ncu_launch_c3_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -c 1 -o /tmp/prof_$TAG \
    python tools/time_case.py c3_searth_22b --ncols 2072 --spinup 5 --steps 1 > $OUT/ncu_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > $OUT/ncu_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv > $OUT/ncu_${TAG}_source.csv 2>/dev/null
tail -3 $OUT/gpu_tests_c3_$TAG.log; cat $OUT/time_c3_$TAG.json
