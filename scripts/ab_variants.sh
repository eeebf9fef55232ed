#!/bin/bash
# A/B of library builds on one box: every sloam_b200/lib/variants/<name>.so is swapped in and
# the default bench run; prints keyframes/s and ms per step.  usage: bash scripts/ab_variants.sh name ...
mkdir -p gpurun_out
cp sloam_b200/lib/libsloam_b200.so /tmp/libsloam_b200.keep
for v in "$@"; do
  cp sloam_b200/lib/variants/$v.so sloam_b200/lib/libsloam_b200.so
  timeout 90 python bench.py --steps 20 --warmup 3 --cpu-sample 4 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), round(d['ms_per_step'],4), d['config'].get('keyframes_ok'), d['config'].get('mean_landmarks'))"
done
cp /tmp/libsloam_b200.keep sloam_b200/lib/libsloam_b200.so
