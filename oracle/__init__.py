"""CPU oracle for the NavierStokes_TRBDF2_DG matrix-free hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import or execute it, and only as the checker
(or as the timed CPU baseline), never as the thing that ships.

PARITY UNPINNED.  The arithmetic of this path lives in deal.II (>= 9.6.0,
``NavierStokes_TRBDF2_DG/CMakeLists.txt:26``), which is neither vendored in the
reference nor installed here, and the reference holds no golden vectors or
tests for ``NavierStokes_TRBDF2_DG``.  The oracle
therefore restates

* the weak forms exactly as written in
  ``NavierStokes_TRBDF2_DG/navier_stokes_TRBDF2_DG.cc`` (cited per function), and
* deal.II's published algorithms for FE_DGQ / FEEvaluation / SolverCG /
  SolverGMRES / PreconditionChebyshev / MGTransferMatrixFree / Multigrid
  (SURVEY.md section 9),

in quadrature-point form with dense 1-D shape matrices and face-centric loops,
i.e. a deliberately *different* evaluation order from the CUDA kernels, and is
pinned by analytic known-answer tests (tests/test_oracle_*.py) instead of by
reference output.
"""
