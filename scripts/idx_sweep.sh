#!/bin/bash
# scan_index_kernel tuning on the GPU box: rows of 32 entries in flight per lane x resident CTAs per SM (rebuilds the library per point)
out=gpurun_out
for cfg in "8 4" "4 4" "4 6" "4 8" "2 8"; do
  set -- $cfg
  PCRAMP_NVCC_EXTRA="-DIDX_ROWS=$1 -DIDX_BLOCKS=$2" python -m pcramp_b200.build --force > $out/idx_build.log 2>&1 || { echo "build failed $cfg"; continue; }
  python bench.py --steps 12 --no-cpu-baseline --dp-problems 0 --fasta-targets 0 --config-legs none --min-seconds 0.3 > $out/idx_$1_$2.json 2> $out/idx_$1_$2.err
  python - <<PY
import json
d=json.loads(open("$out/idx_$1_$2.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("rows $1 blocks $2: kernel alone %.3f ms, beside %.3f ms, step %.3f ms, one at a time %.3f ms" % (r["avg_launch_ms"], r["avg_launch_ms_beside_partial_word_scan"], d["ms_per_step"], d["pipeline"]["ms_per_step_one_batch_at_a_time"]))
PY
done
