#!/bin/bash
# scan_index_async_kernel tuning on the GPU box: stages of the shared-memory ring (rebuilds the library per point)
out=gpurun_out
for st in 2 3 4; do
  PCRAMP_NVCC_EXTRA="-DIDXA_STAGES=$st" python -m pcramp_b200.build --force > $out/idxa_build.log 2>&1 || { echo "build failed $st"; continue; }
  PCRAMP_TRACE=1 python bench.py --steps 12 --no-cpu-baseline --dp-problems 0 --fasta-targets 0 --config-legs none --min-seconds 0.3 > $out/idxa_$st.json 2> $out/idxa_$st.err
  grep "scan_index_async_kernel" $out/idxa_$st.err | head -1
  python - <<PY
import json
d=json.loads(open("$out/idxa_$st.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("stages $st: kernel alone %.3f ms, beside %.3f ms, step %.3f ms, one at a time %.3f ms" % (r["avg_launch_ms"], r["avg_launch_ms_beside_partial_word_scan"], d["ms_per_step"], d["pipeline"]["ms_per_step_one_batch_at_a_time"]))
PY
done
