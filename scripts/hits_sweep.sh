#!/bin/bash
# index_hits_kernel / index_query_kernel tuning on the GPU box: resident CTAs per SM of index_hits_kernel (rebuilds the library per point)
out=gpurun_out
BENCH="python bench.py --steps 12 --no-cpu-baseline --dp-problems 0 --fasta-targets 0 --config-legs none --min-seconds 0.3"
for b in ${POINTS:-1 8}; do
  PCRAMP_NVCC_EXTRA="-DIDX_HITS_BLOCKS=$b" python -m pcramp_b200.build --force > $out/hits_build.log 2>&1 || { echo "build failed $b"; continue; }
  $BENCH > $out/hits_$b.json 2> $out/hits_$b.err
  python - <<PY
import json
d=json.loads(open("$out/hits_$b.json").read().strip().splitlines()[-1])
r=d["roofline"]; bd=d["breakdown_ms_per_step"]
print("hits blocks $b: seed stage %.3f ms, db %.3f, score %.3f, step %.4f ms, one at a time %.4f ms, parity %s" % (bd["ms_seed"], bd["ms_db"], bd["ms_score"], d["ms_per_step"], d["pipeline"]["ms_per_step_one_batch_at_a_time"], d.get("parity_at_bench")))
PY
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"index_hits_kernel|index_query_kernel" -c 8 --csv --log-file $out/hits_launch_$b.csv python bench.py --workers 1 --steps 2 --warmup 2 --no-cpu-baseline --dp-problems 0 --fasta-targets 0 --config-legs none --min-seconds 0 > /dev/null 2>&1
  grep -o 'index_[a-z]*_kernel.*' $out/hits_launch_$b.csv | awk -F'"' '{print $1, $(NF-1)}' | tail -4
done
python -m pcramp_b200.build --force > $out/hits_build.log 2>&1
