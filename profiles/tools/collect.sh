#!/bin/bash
# Round evidence, run on the GPU box:  bash profiles/tools/collect.sh <tag>
#   1. the plain bench (the numbers), 2. the ncu launch list of the same command (shares of the step),
#   3. one `ncu --set full` capture of each dominant kernel.  Nothing printed under ncu is a bench value.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1
for k in scan_index_kernel entry_match_kernel score_items_kernel thermo_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $out/${tag}_$k \
      python bench.py --steps 1 --warmup 2 --no-cpu-baseline > $out/${tag}_ncu_$k.log 2>&1
done
ls $out | grep "^${tag}_"
