#!/bin/bash
# End-of-round-2 evidence, run on the GPU box:  bash profiles/tools/collect_final.sh <tag>
#   1. the plain default bench (the numbers)
#   2. the ncu launch list of the headline step, one batch at a time (shares of the step), after the same command ran plain
#   3. `ncu --set full` captures of scan_index_kernel (roofline.traffic) and edge_lookup_kernel; only the raw pages travel back
# Nothing printed under ncu is a bench value.
tag=${1:-r02final}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
CMD="python bench.py --workers 1 --steps 2 --warmup 3 --no-cpu-baseline --fasta-targets 0 --dp-problems 0 --config-legs none --min-seconds 0"
$CMD > $out/${tag}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/${tag}_launches.csv $CMD > $out/${tag}_ncu_launches.log 2>&1
for k in scan_index_kernel edge_lookup_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o /tmp/${tag}_$k $CMD > $out/${tag}_ncu_$k.log 2>&1
  ncu -i /tmp/${tag}_$k.ncu-rep --page raw --csv > $out/${tag}_${k}_ncu_raw.csv 2>/dev/null
done
ls -la $out | grep "${tag}_"
