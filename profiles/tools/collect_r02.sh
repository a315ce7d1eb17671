#!/bin/bash
# Round-2 evidence, run on the GPU box:  bash profiles/tools/collect_r02.sh <tag>
#   1. the plain default bench (the numbers)
#   2. the ncu launch list of the headline step, one batch at a time (shares of the step), after the same command ran plain
#   3. one `ncu --set full` capture of scan_index_kernel and score_entries_kernel (headline), bg_sw_kernel (C2 background leg)
# Nothing printed under ncu is a bench value.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
CMD="python bench.py --workers 1 --steps 2 --warmup 3 --no-cpu-baseline --fasta-targets 0 --dp-problems 0 --config-legs none --min-seconds 0"
$CMD > $out/${tag}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/${tag}_launches.csv $CMD > $out/${tag}_ncu_launches.log 2>&1
for k in scan_index_kernel score_entries_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o $out/${tag}_$k $CMD > $out/${tag}_ncu_$k.log 2>&1
done
BG="python scripts/legs_micro.py background --no-cpu"
$BG > $out/${tag}_bg_plain.json 2> $out/${tag}_bg_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:bg_sw_kernel -s 1 -c 1 -o $out/${tag}_bg_sw_kernel $BG > $out/${tag}_ncu_bg.log 2>&1
ls -la $out | grep "${tag}_"
