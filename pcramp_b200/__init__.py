"""pcramp_b200 -- B200 (sm_100a) implementation of PCRamp's primer-pair scoring path.

The product is the C-ABI library in csrc/ (declared in include/pcramp_gpu.h).  This package is the
thin Python host mirror used by the tests and bench.py: `api.PcrampGpu` wraps the C ABI one to one,
`synth` generates the synthetic configurations of BASELINE.json, `words` packs oligos.
"""
from .api import PcrampGpu, GpuError, TARGET, BACKGROUND, MULTIPLEX  # noqa: F401
