set -x
python -m pytest tests/test_gpu_thermo.py -m gpu -x -q 2>&1 | tail -3
for mb in 4 5 6 8; do
  PCRAMP_NVCC_EXTRA="-DTHERMO_MIN_BLOCKS=$mb" python -m pcramp_b200.build --force > /dev/null 2>&1
  python bench.py --targets 200 --length 3000 --pairs 100 --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); g=d['dp_gcups']; print('MB=$mb', g['value'], g['ms_per_step'], g['roofline']['avg_launch_ms'])"
done
python -m pcramp_b200.build --force > /dev/null 2>&1
