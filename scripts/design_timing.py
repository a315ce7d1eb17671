"""Per-stage timing of whole design runs through the command-line host (stderr of --timing) next to the stock program."""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import design_cases  # noqa: E402

HOST = os.path.join(ROOT, "pcramp_b200", "host", "pcramp_b200")
STOCK = os.path.join(ROOT, "oracle", "_ref", "pcramp")

for case in design_cases.cases():
    if len(sys.argv) > 1 and case.name not in sys.argv[1:]:
        continue
    with tempfile.TemporaryDirectory() as d:
        argv = design_cases.materialise(case, d)
        t0 = time.time()
        p = subprocess.run([HOST] + argv + ["-o", os.path.join(d, "o.txt"), "--timing"], capture_output=True, text=True)
        t_host = time.time() - t0
        print("==", case.name, "host %.2f s" % t_host)
        print("\n".join(x for x in p.stderr.splitlines() if "timing" in x or "Design iteration" in x or "Finished" in x))
        if os.path.exists(STOCK) and "--stock" in sys.argv:
            for threads in ("1", "0"):
                a = list(argv)
                a[a.index("--thread") + 1] = threads
                t0 = time.time()
                subprocess.run([STOCK] + a + ["-o", os.path.join(d, "s.txt"), "-v", "silent"], stderr=subprocess.DEVNULL)
                print("   stock --thread %s: %.2f s" % (threads, time.time() - t0))
