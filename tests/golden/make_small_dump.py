"""Generate tests/golden/small_dump/: a golden dump (format of shakti_b200/golden.py) of a small seeded
case, written by the CPU oracle.  It is a REGRESSION anchor for both the oracle and the CUDA path (the
reference itself cannot produce it here: no FEniCSx).  The independent pin of the
oracle is element_golden.json (exact sympy integrals).

Run:  python tests/golden/make_small_dump.py
"""
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
for p in (str(ROOT), str(ROOT / "shakti-fenics_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)

from common import make_case, make_oracle  # noqa: E402
from shakti_b200 import golden  # noqa: E402

out = Path(__file__).with_name("small_dump")
if out.exists():
    shutil.rmtree(out)
case = make_case(nx=12, ny=9, seed=21)
golden.write_dump(out, make_oracle(*case), [360.0, 3600.0, 3600.0, 3600.0], producer="oracle (tests/golden/make_small_dump.py)")
print("wrote", out, sum(f.stat().st_size for f in out.iterdir()), "bytes")
