"""Golden reports of the STOCK program (oracle/_ref/pcramp: the unmodified reference main.cpp + sources, built by oracle/Makefile)
for the whole-run cases of tests/design_cases.py.  Dev container only:   python tests/golden/make_design_golden.py"""
import os
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests import design_cases  # noqa: E402

STOCK = os.path.join(ROOT, "oracle", "_ref", "pcramp")


def main():
    only = sys.argv[1:]
    for case in design_cases.cases():
        if only and case.name not in only:
            continue
        with tempfile.TemporaryDirectory() as d:
            argv = design_cases.materialise(case, d)
            out = os.path.join(d, "out.txt")
            t0 = time.time()
            subprocess.run([STOCK] + argv + ["-o", out, "-v", "silent"], check=True, stderr=subprocess.DEVNULL)
            lines = design_cases.report_lines(out)
        with open(os.path.join(HERE, "design_%s.txt" % case.name), "w") as fh:
            fh.write("\n".join(lines) + "\n")
        print("%-20s %5.1f s  %d lines, %d assays" % (case.name, time.time() - t0, len(lines), sum(x.startswith("ASSAY") for x in lines)))


if __name__ == "__main__":
    main()
