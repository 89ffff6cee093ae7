"""Oracle: cvxopt `solvers.qp` -> `coneprog.coneqp` restated for the 'l' cone only.

TEST INFRASTRUCTURE - see oracle/__init__.py.   **Parity unpinned.**

cvxopt is a third-party dependency of the reference (call sites
path_planning.py:211-214, planning_utils.py:353, speed_planning_test.py:502);
it is not vendored, not version-pinned (no requirements file) and not
installable offline, and the reference has no test that pins a QP answer.
This file restates the published algorithm of cvxopt 1.2/1.3
`coneprog.coneqp` (Mehrotra predictor-corrector with Nesterov-Todd scaling)
specialised to the non-negative orthant, with the default `chol2` KKT solver
(`misc.kkt_chol2`: Cholesky of S = P + G' W^-2 G, then of K = A S^-1 A'),
default options maxiters=100, abstol=1e-7, reltol=1e-6, feastol=1e-7,
refinement=0, STEP=0.99, EXPON=3.

    minimize 1/2 x'Px + q'x   s.t.  Gx + s = h, s >= 0,  Ax = b

Deviation (documented in DESIGN.md): a factorisation failure
(ArithmeticError in cvxopt) ends the solve with status 'unknown' exactly as in
cvxopt; additionally non-finite iterates end it with 'unknown' (cvxopt would
raise or keep iterating on NaNs - either way the reference never inspects the
status, path_planning.py:211-219).
"""
import numpy as np

MAXITERS, ABSTOL, RELTOL, FEASTOL = 100, 1e-7, 1e-6, 1e-7
STEP, EXPON = 0.99, 3


class _Chol2:
    """misc.kkt_chol2 for dims = {'l': m}: factor once per scaling, solve many."""

    def __init__(self, P, G, A):
        self.P, self.G, self.A = P, G, A
        self.p = A.shape[0]

    def factor(self, d):
        # Gs = W^-1 G ; S = P + Gs'Gs ; K = A S^-1 A'
        self.d = d
        self.Gs = self.G / d[:, None]
        S = self.P + self.Gs.T @ self.Gs
        self.L = np.linalg.cholesky(S)                   # raises LinAlgError <-> ArithmeticError
        if self.p:
            self.Asct = np.linalg.solve(self.L, self.A.T)
            self.LK = np.linalg.cholesky(self.Asct.T @ self.Asct)
        return self

    def solve(self, bx, by, bz):
        """[P A' G'; A 0 0; G 0 -W'W][ux; uy; W^-1 uz] = [bx; by; bz]; returns ux, uy, uz."""
        z = bz / self.d
        x = np.linalg.solve(self.L, bx + self.Gs.T @ z)
        if self.p:
            y = self.Asct.T @ x - by
            y = np.linalg.solve(self.LK.T, np.linalg.solve(self.LK, y))
            x = x - self.Asct @ y
        else:
            y = np.zeros(0)
        x = np.linalg.solve(self.L.T, x)
        z = self.Gs @ x - z
        return x, y, z


def qp(P, q, G, h, A=None, b=None, maxiters=MAXITERS, abstol=ABSTOL, reltol=RELTOL,
       feastol=FEASTOL, trace=None, kkt=None, diverge_factor=None):
    """Returns dict(x, y, s, z, status, iterations, gap, pcost, dcost, pres, dres).

    diverge_factor (None/0 = cvxopt behaviour): documented early exit of the GPU kernels
    (EmpQpParams.diverge_factor, csrc/sl_qp.cu): a duality gap above diverge_factor x the
    smallest gap of the earlier iterations ends the solve with status 'unknown' and the PREVIOUS
    iterate x (the last one reached before the gap grew; y, s, z are left at the current one).  coneqp has no infeasibility certificate: on an infeasible problem cvxopt goes on for
    a few more iterations until its Cholesky breaks down (status 'unknown' as well), through
    iterates that are no longer a numerically well defined function of the input (two exact-
    arithmetic-equivalent factorisations differ by O(1 m) there, profiles/parity_r2.json); the
    gap of a converging solve decreases monotonically (worst ratio seen: tests/test_qp_parity)."""
    P = np.asarray(P, dtype=np.float64)
    n = P.shape[0]
    q = np.asarray(q, dtype=np.float64).reshape(-1)
    G = np.asarray(G, dtype=np.float64)
    h = np.asarray(h, dtype=np.float64).reshape(-1)
    m = G.shape[0]
    if A is None:
        A, b = np.zeros((0, n)), np.zeros(0)
    A = np.asarray(A, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64).reshape(-1)

    resx0 = max(1.0, np.sqrt(q @ q))
    resy0 = max(1.0, np.sqrt(b @ b))
    resz0 = max(1.0, np.sqrt(h @ h))

    kkt = kkt or _Chol2(P, G, A)
    # initial point: W = I
    try:
        f = kkt.factor(np.ones(m))
    except np.linalg.LinAlgError:
        raise ValueError("Rank(A) < p or Rank([P; A; G]) < n")
    x, y, z = f.solve(-q, b.copy(), h.copy())
    s = -z
    nrms = np.sqrt(s @ s)
    ts = -s.min()
    if ts >= -1e-8 * max(nrms, 1.0):
        s = s + (1.0 + ts)
    nrmz = np.sqrt(z @ z)
    tz = -z.min()
    if tz >= -1e-8 * max(nrmz, 1.0):
        z = z + (1.0 + tz)
    gap = s @ z
    gap_min = gap

    def result(status, it):
        return dict(x=x, y=y, s=s, z=z, status=status, iterations=it, gap=gap,
                    pcost=pcost, dcost=dcost, pres=pres, dres=dres)

    d = lmbda = None
    for it in range(maxiters + 1):
        rx = P @ x + q
        f0 = 0.5 * (x @ rx + x @ q)
        rx = rx + A.T @ y + G.T @ z
        resx = np.sqrt(rx @ rx)
        ry = A @ x - b
        resy = np.sqrt(ry @ ry)
        rz = s + G @ x - h
        resz = np.sqrt(rz @ rz)
        pcost = f0
        dcost = f0 + y @ ry + z @ rz - gap
        if pcost < 0.0:
            relgap = gap / -pcost
        elif dcost > 0.0:
            relgap = gap / dcost
        else:
            relgap = None
        pres = max(resy / resy0, resz / resz0)
        dres = resx / resx0
        if trace is not None:
            trace.append(dict(it=it, pcost=pcost, dcost=dcost, gap=gap, pres=pres, dres=dres, x=x.copy()))
        if pres <= feastol and dres <= feastol and (gap <= abstol or (relgap is not None and relgap <= reltol)):
            return result("optimal", it)
        if it == maxiters:
            return result("unknown", it)
        if not (np.isfinite(gap) and np.isfinite(resx) and np.isfinite(resz)):
            return result("unknown", it)
        if diverge_factor and it > 0 and gap > diverge_factor * gap_min:
            x = x_prev                                  # the step that blew the gap up is discarded
            return result("unknown", it)
        gap_min = min(gap_min, gap)
        x_prev = x

        if it == 0:
            d = np.sqrt(s / z)                      # misc.compute_scaling
            lmbda = np.sqrt(s * z)
        lmbdasq = lmbda * lmbda
        try:
            f3 = kkt.factor(d)
        except np.linalg.LinAlgError:
            return result("unknown", it)

        mu = gap / m
        sigma, eta = 0.0, 0.0
        ws3 = None
        for i in (0, 1):
            ds = np.zeros(m)
            if i == 1:
                ds = ds - ws3
            ds = ds - lmbdasq
            ds = ds + sigma * mu
            dx = (-1.0 + eta) * rx
            dy = (-1.0 + eta) * ry
            dz = (-1.0 + eta) * rz
            # f4_no_ir
            ds = ds / lmbda
            dz = dz - d * ds
            dx, dy, dz = f3.solve(dx, dy, dz)
            ds = ds - dz
            dsdz = ds @ dz
            if i == 0:
                ws3 = ds * dz
            ds = ds / lmbda                          # misc.scale2
            dz = dz / lmbda
            ts = -ds.min()
            tz = -dz.min()
            t = max(0.0, ts, tz)
            if t == 0:
                step = 1.0
            elif i == 0:
                step = min(1.0, 1.0 / t)
            else:
                step = min(1.0, STEP / t)
            if i == 0:
                sigma = min(1.0, max(0.0, 1.0 - step + dsdz / gap * step ** 2)) ** EXPON
                eta = 0.0
        x = x + step * dx
        y = y + step * dy
        ds = (1.0 + step * ds) * lmbda               # scale2(..., inverse='I')
        dz = (1.0 + step * dz) * lmbda
        with np.errstate(invalid="ignore"):
            sq_s = np.sqrt(ds)                       # misc.update_scaling ('l' block)
            sq_z = np.sqrt(dz)
        d = d * sq_s / sq_z
        lmbda = sq_s * sq_z
        s = lmbda * d
        z = lmbda / d
        gap = lmbda @ lmbda
    raise AssertionError("unreachable")
