#!/usr/bin/env python3
"""Check the CUDA path (and, with --oracle, the CPU oracle) against a golden dump written from a real FEniCSx
run of the reference by tools/dump_fenicsx_golden.py:

    python tools/check_golden_dump.py <dump_dir> [--oracle]

CSR pattern bit-exact, F and J at 1e-12, fields at 1e-8 (BASELINE.json tolerances); the dump's quadrature
table is validated (degree 7) and handed to the implementation under test."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "shakti-fenics_b200"), str(ROOT / "tests")]

from shakti_b200 import golden  # noqa: E402


def main():
    dump = golden.Dump(sys.argv[1])
    from oracle.quadrature import check_degree          # test infrastructure: this tool is a checker, not product code
    print("quadrature table of the dump:", len(dump.quad[1]), "points, max moment error", check_degree(*dump.quad, degree=7))
    if "--oracle" in sys.argv:
        from common import OracleStepper
        print("oracle :", json.dumps(golden.check(dump, OracleStepper(dump))))
    stepper = golden.ModelStepper(dump)
    print("B200   :", json.dumps(golden.check(dump, stepper, pattern=stepper.pattern() if hasattr(stepper, "pattern") else None)))


if __name__ == "__main__":
    main()
