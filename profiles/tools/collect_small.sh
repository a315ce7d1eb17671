#!/bin/bash
# ncu --set full of the headline step's secondary kernels; only the raw / source pages travel back (the .ncu-rep files are 16 MB each)
tag=${1:-r02}
out=gpurun_out
CMD="python bench.py --workers 1 --steps 2 --warmup 3 --no-cpu-baseline --fasta-targets 0 --dp-problems 0 --config-legs none --min-seconds 0"
$CMD > $out/${tag}_plain.log 2>&1 || exit 1
for k in ${KERNELS:-scan_edge_fst_kernel index_hits_kernel seg_sort_small_kernel seg_scatter_kernel index_query_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o /tmp/${tag}_$k $CMD > $out/${tag}_ncu_$k.log 2>&1
  ncu -i /tmp/${tag}_$k.ncu-rep --page raw --csv > $out/${tag}_${k}_ncu_raw.csv 2>/dev/null
  ncu -i /tmp/${tag}_$k.ncu-rep --page source --csv > $out/${tag}_${k}_ncu_source.csv 2>/dev/null
done
ls -la $out | grep "${tag}_"
