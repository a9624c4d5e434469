#!/bin/bash
# usage (on the GPU box): tools/tune.sh <years> <variant>...   -- step-kernel ms per pass for each build/libshyft_b200_<variant>.so
years=$1; shift
for v in default "$@"; do
  if [ "$v" = default ]; then lib=""; else lib="$PWD/build/libshyft_b200_$v.so"; fi
  SB2_LIB=$lib python bench.py --years $years --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); c=d['config']; print('$v', 'step_ms', round(c['step_kernel_ms_per_step'],2), 'interp_ms', round(c['interp_ms_per_step'],2), 'total_ms', round(d['ms_per_step'],2), 'Gcs/s', round(d['value']/1e9,3))"
done
