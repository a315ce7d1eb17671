#!/bin/bash
# Round evidence, run on the GPU box:  bash profiles/tools/collect.sh <tag>
#   1. the plain bench (the numbers), 2. the ncu launch list of the same command, one batch at a time (shares of the step),
#   3. one `ncu --set full` capture of each dominant kernel.  Nothing printed under ncu is a bench value.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --workers 1 --steps 2 --warmup 1 --no-cpu-baseline --fasta-targets 0 > $out/${tag}_ncu_launches.log 2>&1
for k in scan_index_kernel score_items_kernel scan_edge_fst_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o $out/${tag}_$k \
      python bench.py --workers 1 --steps 1 --warmup 2 --no-cpu-baseline --dp-problems 0 --fasta-targets 0 > $out/${tag}_ncu_$k.log 2>&1
done
ls $out | grep "^${tag}_"
