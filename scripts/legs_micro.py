"""K3 / K4 alone: the dp_gcups and sw_gcups legs of bench.py on an otherwise empty context (quick kernel iterations on the GPU box).
usage: python scripts/legs_micro.py [sw] [dp] [background] [optimize] [design] [design_c2] [large] [--no-cpu]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from pcramp_b200 import PcrampGpu  # noqa: E402

import bench_legs  # noqa: E402

a = argparse.Namespace(no_cpu_baseline="--no-cpu" in sys.argv, dp_problems=262144, dp_cpu_problems=60000, steps=10, warmup=3, pairs=1000,
                       c4_targets=int(os.environ.get("C4_TARGETS", "64")), c4_length=5000000, hbm_peak=6553.0)
for name, fn in (("background", bench_legs.background_leg), ("optimize", bench_legs.optimize_leg), ("design", bench_legs.design_leg), ("design_c2", bench_legs.design_c2_leg),
                 ("large", bench_legs.large_genome_leg)):
    if name in sys.argv:
        print(json.dumps(fn(a, 0)))
g = PcrampGpu(0)
ext = torch.cuda.ExternalStream(g.stream, device=0)
with torch.cuda.stream(ext):
    if "sw" in sys.argv:
        print(json.dumps(bench.sw_leg(a, g, ext, torch)))
    if "dp" in sys.argv:
        print(json.dumps(bench.dp_leg(a, g, torch, ext, 0, 1, None)))
torch.cuda.synchronize()
del ext
g.close()
