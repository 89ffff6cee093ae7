"""Import the UNMODIFIED reference planner modules in the build container.

TEST INFRASTRUCTURE - see oracle/__init__.py.  Only usable where
/root/reference exists (the build container); golden fixtures made with it are
committed under tests/golden/ so nothing on the GPU box needs the reference.

The reference imports `carla` (planning_utils.py:10, annotation at :14) and
`cvxopt` (path_planning.py:10, planning_utils.py:11, speed_planning_test.py:19-20).
Neither is installable offline, so they are replaced by:
  * a stub `carla` module exposing Location / Waypoint;
  * a `cvxopt` package shim whose `matrix` wraps a numpy array with cvxopt's
    column-major flat indexing and whose `solvers.qp` is oracle.coneqp.qp.
"""
import os
import sys
import types
import numpy as np

REFERENCE_ROOT = os.environ.get("EMP_REFERENCE_ROOT", "/root/reference")


class _Matrix:
    """Minimal stand-in for cvxopt.matrix: column-major flat __getitem__, len() = element count."""

    def __init__(self, a, size=None, tc="d"):
        if isinstance(a, _Matrix):
            a = a.a
        self.a = np.array(a, dtype=np.float64)
        if self.a.ndim == 1:
            self.a = self.a.reshape(-1, 1)
        if size is not None:
            self.a = self.a.reshape(size, order="F")

    @property
    def size(self):
        return self.a.shape

    def __len__(self):
        return self.a.size

    def __getitem__(self, k):
        if isinstance(k, tuple):
            return self.a[k]
        flat = self.a.reshape(-1, order="F")
        return flat[k]

    def __setitem__(self, k, v):
        if isinstance(k, tuple):
            self.a[k] = v
        else:
            flat = self.a.reshape(-1, order="F")
            flat[k] = v
            self.a = flat.reshape(self.a.shape, order="F")

    def __array__(self, dtype=None, copy=None):
        return self.a if dtype is None else self.a.astype(dtype)


_installed = False
_mods = None
qp_log = []          # every solvers.qp call: dict(P,q,G,h,A,b,result)
qp_log_enabled = False
diverge_factor = None  # None = cvxopt behaviour; a number = the kernels' documented early exit (oracle/coneqp.py)


def install():
    global _installed
    if _installed:
        return
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    from oracle import coneqp as _coneqp

    carla = types.ModuleType("carla")
    carla.Location = type("Location", (), {})
    carla.Waypoint = type("Waypoint", (), {})
    sys.modules["carla"] = carla

    cvx = types.ModuleType("cvxopt")
    sol = types.ModuleType("cvxopt.solvers")
    sol.options = {}

    def qp(P, q, G=None, h=None, A=None, b=None, **kw):
        arr = lambda m: None if m is None else np.asarray(m.a if isinstance(m, _Matrix) else m, dtype=np.float64)
        res = _coneqp.qp(arr(P), arr(q), arr(G), arr(h), arr(A), arr(b), diverge_factor=diverge_factor)
        if qp_log_enabled:
            qp_log.append(dict(P=arr(P), q=arr(q), G=arr(G), h=arr(h), A=arr(A), b=arr(b), result=res))
        out = dict(res)
        out["x"] = _Matrix(res["x"])
        return out

    sol.qp = qp
    cvx.matrix = _Matrix
    cvx.solvers = sol
    sys.modules["cvxopt"] = cvx
    sys.modules["cvxopt.solvers"] = sol
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def modules():
    """Return (path_planning, planning_utils, speed_planning_test) of the reference."""
    global _mods
    if _mods is not None:
        return _mods
    install()
    import warnings
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    # the repo ships its own drop-in `planner` package; make sure the reference's wins here
    def _is_ref(mod):
        f = getattr(mod, "__file__", None)
        if f:
            return f.startswith(REFERENCE_ROOT)
        return any(str(p).startswith(REFERENCE_ROOT) for p in getattr(mod, "__path__", []))
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if (k == "planner" or k.startswith("planner.")) and not _is_ref(sys.modules[k])}
    if REFERENCE_ROOT in sys.path:
        sys.path.remove(REFERENCE_ROOT)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        from planner import path_planning, planning_utils, speed_planning_test
        mods = (path_planning, planning_utils, speed_planning_test)
    finally:
        # leave sys.modules as found for the repo's own drop-in package
        for k in [k for k in sys.modules if k == "planner" or k.startswith("planner.")]:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.path.remove(REFERENCE_ROOT)
    _mods = mods
    return mods


def speed_planning_module():
    """The reference's planner/speed_planning.py (the non-running S-T stub), imported like modules()."""
    install()
    modules()
    import importlib
    def _is_ref(mod):
        f = getattr(mod, "__file__", None)
        if f:
            return f.startswith(REFERENCE_ROOT)
        return any(str(p).startswith(REFERENCE_ROOT) for p in getattr(mod, "__path__", []))
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if (k == "planner" or k.startswith("planner.")) and not _is_ref(sys.modules[k])}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        mod = importlib.import_module("planner.speed_planning")
    finally:
        for k in [k for k in sys.modules if k == "planner" or k.startswith("planner.")]:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.path.remove(REFERENCE_ROOT)
    return mod
