#!/bin/bash
# usage (on the GPU box): tools/tune_c3.sh <variant>...  -- config 3 step-kernel ms (pt_hs_k, hbv_stack) for each build/libshyft_b200_<variant>.so
export SB2_C3_YEARS=${SB2_C3_YEARS:-0.25}
for v in default "$@"; do
  if [ "$v" = default ]; then lib=""; else lib="$PWD/build/libshyft_b200_$v.so"; fi
  SB2_LIB=$lib python tools/bench_configs.py 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$v', d['stack'], 'step_ms', round(d['step_ms'],2), 'interp_ms', round(d['interp_ms'],2))"
done
