#!/bin/bash
# A/B of library builds on one box: every sloam_b200/lib/variants/<name>.so is swapped in and
# the default bench run; prints keyframes/s, ms per step (production, 2 lanes) and the serial
# per-kernel times (extra bench.py arguments in $BENCH_ARGS).  usage: bash scripts/ab_variants.sh [--workload W] name ...   ("base" = the built library)
WL=os1-64
if [ "$1" == "--workload" ]; then WL=$2; shift 2; fi
mkdir -p gpurun_out
cp sloam_b200/lib/libsloam_b200.so /tmp/libsloam_b200.keep
for v in "$@"; do
  if [ "$v" != "base" ]; then cp sloam_b200/lib/variants/$v.so sloam_b200/lib/libsloam_b200.so; else cp /tmp/libsloam_b200.keep sloam_b200/lib/libsloam_b200.so; fi
  timeout 120 python bench.py --workload $WL --steps 20 --warmup 3 --cpu-seconds 0.05 $BENCH_ARGS 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read())
print('$v', round(d['value']), 'kf/s', round(d['ms_per_step'],4), 'ms; serial', d['run'].get('serial_ms_per_step'), 'ms;', d['run']['keyframes_ok'], d['run']['mean_landmarks'])
print('   ', ' | '.join('%s %.0f' % (r['kernel'].split('_kernel')[0][:14], r['ms']*1e3) for r in d['kernels']))"
done
cp /tmp/libsloam_b200.keep sloam_b200/lib/libsloam_b200.so
