#!/bin/bash
# K3 / K4 evidence, run on the GPU box:  bash profiles/tools/collect_dp.sh <tag>
#   the plain legs (the numbers), then one `ncu --set full` capture of sw_words_kernel and of thermo_kernel from the same command.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python scripts/legs_micro.py sw dp --no-cpu > $out/${tag}_legs.json 2> $out/${tag}_legs.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:sw_words_kernel -s 1 -c 1 -o $out/${tag}_sw_words_kernel \
    python scripts/legs_micro.py sw dp --no-cpu > $out/${tag}_ncu_sw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:thermo_kernel -s 3 -c 1 -o $out/${tag}_thermo_kernel \
    python scripts/legs_micro.py sw dp --no-cpu > $out/${tag}_ncu_thermo.log 2>&1
ls -la $out | grep "${tag}_"
