#!/bin/bash
# stage trace of one whole design run through the command-line host (PCRAMP_TRACE=1)
set -e
cd "$(dirname "$0")/.."
python - <<'P'
import sys, tempfile, os, subprocess
sys.path.insert(0, '.')
from tests import design_cases
case = [c for c in design_cases.cases() if c.name == (sys.argv[1] if len(sys.argv) > 1 else 'c1_seed42_count3')][0]
d = tempfile.mkdtemp()
argv = design_cases.materialise(case, d)
env = dict(os.environ, PCRAMP_TRACE='1')
p = subprocess.run(['pcramp_b200/host/pcramp_b200'] + argv + ['-o', d + '/o.txt', '--timing'], env=env, capture_output=True, text=True)
print(p.stderr)
P
