#!/bin/bash
# Everything profiles/ holds for one round, produced on ONE B200 box (run under gpurun):
#   bash scripts/profile_round.sh <tag>          e.g. r2z
# writes gpurun_out/<tag>_*: bench lines of every workload + the reference arm, the ncu launch
# list of the default bench command, `ncu --set full` captures of the top kernels (OS1-64,
# one lane so that every kernel is a single launch over the whole batch), of the dense-forest
# cylinder kernel and of the 100k association kernel, and the tables made from them.
# Numbers printed by a run under ncu are never bench values.
TAG=${1:-r2z}
O=gpurun_out
mkdir -p $O
for w in os1-64 vlp-16 os1-64-dense os1-128 assoc-100k; do
  python bench.py --workload $w > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err
done
python bench.py --impl reference > $O/${TAG}_bench_os1-64_reference.json 2> /dev/null
python bench.py --workload vlp-16 --impl reference > $O/${TAG}_bench_vlp-16_reference.json 2> /dev/null
# launch list of the default command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $O/${TAG}_launches_os1-64_k1024.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 0.05 \
    > $O/${TAG}_launch_run.log 2>&1
python scripts/launch_summary.py $O/${TAG}_launches_os1-64_k1024.csv > $O/${TAG}_launch_summary.txt 2>&1
# full captures
K='regex:project_split|range_finalize|ground_scatter|ground_cells|ground_fit|cc_rows|cc_label|vertex_kernel|cylinder_kernel|lm_kernel|build_matches'
ncu --set full --clock-control none --import-source on -k "$K" -s 12 -c 12 -o $O/${TAG}_top \
    python bench.py --lanes 1 --steps 1 --warmup 1 --cpu-seconds 0.05 --no-kernel-profile > $O/${TAG}_ncu_top.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:cylinder_kernel|cc_label' -s 2 -c 2 -o $O/${TAG}_dense \
    python bench.py --workload os1-64-dense --lanes 1 --steps 1 --warmup 1 --cpu-seconds 0.05 --no-kernel-profile > $O/${TAG}_ncu_dense.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:assoc_kernel' -s 1 -c 1 -o $O/${TAG}_assoc \
    python bench.py --workload assoc-100k --steps 1 --warmup 1 --cpu-seconds 0.05 > $O/${TAG}_ncu_assoc.log 2>&1
for r in top dense assoc; do
  python scripts/ncu_table.py $O/${TAG}_$r.ncu-rep > $O/${TAG}_ncu_$r.md 2>&1
done
for k in project_split vertex_kernel ground_cells cc_label ground_fit cylinder_kernel; do
  echo "== $k (by stall samples)"; python scripts/ncu_hot.py $O/${TAG}_top.ncu-rep $k 14
done > $O/${TAG}_ncu_hot_lines.txt 2>&1
ls -la $O | grep ${TAG}_
