"""CPU oracle for the EM-planner hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it, and there only as the checker or the timed
CPU baseline - never as a fallback for the CUDA path.

Every function restates (does not copy) the algorithm of the reference
``6Lackiu/EMplanner_Carla`` @ 1f115a5 and cites the reference file:line it
follows.  Parity pinning ("Oracle"):

* S-L DP / densify / corridor / Frenet helpers / S-T DP forward tables are
  pinned against outputs of the *unmodified reference* run in the build
  container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``).
* The QP solver (``oracle/coneqp.py``) restates cvxopt's ``coneqp`` for the
  non-negative orthant.  cvxopt is an un-vendored, un-pinned third-party
  dependency of the reference that is not installable offline, and the
  reference holds no test that pins a QP result: **parity unpinned** at the
  cvxopt boundary.  The QP *matrix construction* is pinned (the reference's
  own ``Quadratic_planning`` is run with the solver intercepted).
"""
