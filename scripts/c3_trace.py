"""the C3 design leg alone (bench_legs.design_c3_leg on the headline's collection), e.g. under PCRAMP_TRACE=1 to see the allocator's share"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcramp_b200 import synth  # noqa: E402
import bench_legs  # noqa: E402

a = argparse.Namespace(no_cpu_baseline=True)
coll = synth.TargetFactory(3, 20000, 30000, n_clades=20, between=0.15, within=0.05).collection()
out = bench_legs.design_c3_leg(a, coll, 0, None)
print(json.dumps({"ms_per_iteration": out["ms_per_iteration"], "ms_first_iteration": out["ms_first_iteration"], "ms_later_iterations": out["ms_later_iterations"], "iterations": [{k: round(v, 1) if isinstance(v, float) else v for k, v in i.items()} for i in out["iterations"]],
                  "index": out["index"]}))
