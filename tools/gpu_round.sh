#!/bin/bash
# One GPU-box session: GPU tests, the bench line, per-case timings, the ncu launch list and one full capture.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [kernel-regex for the full capture] [time_case args for the capture]
TAG=${1:-r2x}
KREGEX=${2:-k_full_lu}
shift; shift
CAPTURE_ARGS=${@:-c3_searth_22b --ncols 2072 --spinup 5 --steps 1}
OUT=gpurun_out
mkdir -p $OUT
lscpu | head -25 > $OUT/lscpu_$TAG.txt; nproc >> $OUT/lscpu_$TAG.txt
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/smi_$TAG.txt
timeout 2400 python -m pytest tests -m gpu -q > $OUT/gpu_tests_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/gpu_tests_$TAG.log
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?" >> $OUT/bench_$TAG.err
timeout 300 python tools/time_case.py c3_searth_22b --ncols 4096 --spinup 12 --steps 10 > $OUT/time_c3_$TAG.json 2>&1
timeout 300 python tools/time_case.py c2m_friction_6spc_mild --ncols 4096 --spinup 50 --steps 50 > $OUT/time_c2m_$TAG.json 2>&1
# launch list of the bench command (per-launch times are cold-cache and serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 32 --warmup 3 --spinup 64 --e2e-steps 2 --no-cpu-baseline --no-secondary > $OUT/ncu_launch_$TAG.log 2>&1
# one full capture of the kernel under work
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -c 1 -o /tmp/prof_$TAG \
    python tools/time_case.py $CAPTURE_ARGS > $OUT/ncu_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > $OUT/ncu_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --print-source sass > $OUT/ncu_${TAG}_source.csv 2>/dev/null
ls -la $OUT | tail -5
