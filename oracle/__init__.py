"""CPU oracle for the SHAKTI transient hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and there only as the checker or as the timed CPU baseline.

PARITY UNPINNED: the reference (agstub/shakti-fenics) delegates all arithmetic to
DOLFINx/FFCx/Basix/PETSc, none of which is installed here, and it ships no tests or golden
vectors.  This oracle is a restatement of ``source/solvers.py`` +
``source/constitutive.py`` with the documented third-party semantics; it is pinned only by
independent mathematics (exact symbolic integrals, finite-difference Jacobians, patch tests,
manufactured fixtures under ``tests/golden/``).

Files: ``shakti_oracle.py`` (numpy; the checker of the parity tests), ``shakti_oracle_c.c`` + ``cbackend.py``
(the same element formulas compiled with OpenMP, ``make -C oracle``; what bench.py times as the CPU baseline),
``quadrature.py`` (rule tables + validator).
"""
