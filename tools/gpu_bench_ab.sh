#!/bin/bash
# A/B of an engine option on the headline bench: bash tools/gpu_bench_ab.sh <tag> <KEY=VALUE> [pytest -k expression]
TAG=$1; OPT=$2; KEXPR=${3:-c4}
OUT=gpurun_out; mkdir -p $OUT
python bench.py --no-cpu-baseline --no-secondary > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err
python bench.py --no-cpu-baseline --no-secondary --opt $OPT > $OUT/bench_${TAG}_alt.json 2>> $OUT/bench_${TAG}.err
python - <<PY
import json
for f in ["$OUT/bench_${TAG}.json", "$OUT/bench_${TAG}_alt.json"]:
    try:
        d = json.load(open(f)); print(f, "value %.4g  ms/step %.4f  e2e %.4g  frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
    except Exception as ex: print(f, "failed", ex)
PY
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "$KEXPR" 2>&1 | tail -3
