#!/bin/bash
# index_hits_kernel tuning on the GPU box: candidates walked side by side per thread (rebuilds the library per point)
out=gpurun_out
BENCH="python bench.py --steps 20 --no-cpu-baseline --dp-problems 0 --fasta-targets 0 --config-legs none --min-seconds 1.0"
for r in ${POINTS:-1 2}; do
  PCRAMP_NVCC_EXTRA="-DIDX_HITS_ROWS=$r" python -m pcramp_b200.build --force > $out/hits_build.log 2>&1 || { echo "build failed $r"; continue; }
  $BENCH > $out/hits_$r.json 2> $out/hits_$r.err
  python - <<PY
import json
d=json.loads(open("$out/hits_$r.json").read().strip().splitlines()[-1])
bd=d["breakdown_ms_per_step"]
print("hits rows $r: seed stage %.3f ms, step %.4f ms, one at a time %.4f ms" % (bd["ms_seed"], d["ms_per_step"], d["pipeline"]["ms_per_step_one_batch_at_a_time"]))
PY
done
python -m pcramp_b200.build --force > $out/hits_build.log 2>&1
