#!/bin/bash
# same-box A/B of library options through PCRAMP_OPTIONS: scripts/ab_options.sh "use_edge_table=1" "use_edge_table=0" ...
out=gpurun_out
BENCH="python bench.py --steps 20 --no-cpu-baseline --dp-problems 0 --fasta-targets 0 --config-legs none --min-seconds 1.0"
for rep in 1 2; do
for opt in "$@"; do
  tag=$(echo "$opt" | tr -c 'a-zA-Z0-9_=\n' '_')
  PCRAMP_OPTIONS="$opt" $BENCH > $out/ab_$tag.json 2> $out/ab_$tag.err || { echo "failed: $opt"; tail -3 $out/ab_$tag.err; continue; }
  python - <<PY
import json
d=json.loads(open("$out/ab_$tag.json").read().strip().splitlines()[-1])
print("%-28s step %.4f ms  one at a time %.4f ms  launches/step %s" % ("$opt", d["ms_per_step"], d["pipeline"]["ms_per_step_one_batch_at_a_time"], d["pipeline"]["launches_per_step"]))
PY
done
done
