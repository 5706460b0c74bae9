#!/bin/bash
# k_friction<6> at 1 / 3 / 4 resident CTAs per SM (launch bounds): time the 6-species friction deck with each build.
OUT=gpurun_out; mkdir -p $OUT
for v in "" _f3 _f4 _f5 _f6; do
  lib=aiolos_b200/libaiolos_b200$v.so
  [ -f $lib ] || continue
  echo "== $lib"
  python tools/time_case.py c2m_friction_6spc_mild --ncols 4096 --spinup 50 --steps 50 --lib $PWD/$lib | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['kernel_ms'], d['ms_per_step'])"
  python tools/time_case.py h0_hd209458b_nophotochem --ncols 4096 --spinup 100 --steps 20 --lib $PWD/$lib | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['kernel_ms'], d['ms_per_step'])"
done
python tools/time_case.py h1_hd209458b_c2ray --ncols 4096 --spinup 300 --steps 20 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('photochem', d['kernel_ms'], d['ms_per_step'])"
