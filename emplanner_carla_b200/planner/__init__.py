"""Drop-in replacement for the reference's `planner` package (6Lackiu/EMplanner_Carla @ 1f115a5).

Put the directory that contains this package (`emplanner_carla_b200/`) in front of the reference
root on sys.path and the driver scripts' `from planner import planning_utils, path_planning`
(test_9.py:26) resolves here: same module names, function names, argument order and defaults,
list-in/list-out types - with the arithmetic running in libemplanner.so on the GPU.
See INTEGRATION.md.
"""
import os as _os
import sys as _sys

try:
    import emplanner_carla_b200 as _pkg  # noqa: F401
except ImportError:  # imported as top-level `planner`: make the parent package importable
    _sys.path.append(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
    import emplanner_carla_b200 as _pkg  # noqa: F401


def _fall_through():
    """Modules of the reference's `planner` package that are outside the hot path (global_planning.py: CARLA
    topology + A*, imported by the driver scripts, test_9.py:28) keep resolving to the reference: every other
    `planner/` directory found on sys.path is appended to this package's search path, so
    `from planner.global_planning import global_path_planner` loads the reference's file while
    `planning_utils`, `path_planning`, `speed_planning`, `speed_planning_test` are served from here."""
    here = _os.path.dirname(_os.path.abspath(__file__))
    for entry in list(_sys.path):
        cand = _os.path.join(entry or _os.getcwd(), "planner")
        if _os.path.isdir(cand) and _os.path.abspath(cand) != here and cand not in __path__:
            __path__.append(cand)


_fall_through()
