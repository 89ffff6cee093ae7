"""Executable specification (numpy float64) of the arithmetic of csrc/sl_qp.cu.

Test helper only - NOT a fallback and not importable from the product package.
It mirrors the kernel step by step (per-station "lane" data, unscaled Mehrotra
iteration, closed-form 2x2 Cholesky of the (l,dl) block, jerk split, dual
block-tridiagonal Cholesky) so the algorithm can be checked against
oracle/coneqp.py on the CPU, where no GPU is available.
"""
import numpy as np


def G_mul(x, d1, d2):
    a = x[:, 0] + d1 * x[:, 1]
    b = x[:, 0] - d2 * x[:, 1]
    return np.stack([a, a, b, b, -a, -a, -b, -b, x[:, 0], x[:, 1], x[:, 2], -x[:, 0], -x[:, 1], -x[:, 2]], axis=1)


def Gt_mul(v, d1, d2):
    pa = (v[:, 0] + v[:, 1]) - (v[:, 4] + v[:, 5])
    pb = (v[:, 2] + v[:, 3]) - (v[:, 6] + v[:, 7])
    return np.stack([pa + pb + v[:, 8] - v[:, 11], d1 * pa - d2 * pb + v[:, 9] - v[:, 12], v[:, 10] - v[:, 13]], axis=1)


def solve_qp(lmin, lmax, st, ds=2.0, w_l=1000, w_ddl=3000, w_dddl=150, w_c=250, w_el=40, w_edl=40, w_eddl=40,
             d1=3, d2=3, host_w=3, maxiters=100, abstol=1e-7, reltol=1e-6, feastol=1e-7, trace=None, diverge_factor=None):
    """Returns (x [n,3], status, iterations)."""
    lmin = np.asarray(lmin, dtype=np.float64)
    lmax = np.asarray(lmax, dtype=np.float64)
    n = len(lmin)
    m = 14 * n
    fi = int(np.ceil(d1 / ds))
    bi = int(np.ceil(d2 / ds))
    idx = np.arange(n)
    U = lmax[np.minimum(idx + fi, n - 1)]
    Lo = lmin[np.maximum(idx - bi, 0)]
    hw = host_w / 2
    ub = np.full((n, 3), 1e5)
    nlb = np.full((n, 3), 1e5)
    ub[0] = st
    nlb[0] = -np.array(st)
    ub[n - 1] = 0
    nlb[n - 1] = 0
    h = np.stack([U - hw, U + hw, U - hw, U + hw, -Lo + hw, -Lo - hw, -Lo + hw, -Lo - hw], axis=1)
    h = np.concatenate([h, ub, nlb], axis=1)
    Pd = np.zeros((n, 3))
    Pd[:, 0] = 2 * (w_l + w_c)
    Pd[:, 2] = 2 * w_ddl + 2 * w_dddl * 2
    Pd[0, 2] -= 2 * w_dddl
    Pd[n - 1, 2] -= 2 * w_dddl
    Pd[n - 1] += 2 * np.array([w_el, w_edl, w_eddl])
    cc = -2 * w_dddl
    Pnod = Pd.copy()
    Pnod[:, 2] = 2 * w_ddl
    Pnod[n - 1, 2] += 2 * w_eddl
    q = np.zeros((n, 3))
    q[:, 0] = w_c * (-2 * ((lmin + lmax) / 2))
    a_, b_, e_ = ds, ds * ds / 3, ds / 2
    G0 = ds * ds / 6
    E = np.array([[1, a_, b_], [0, 1, e_], [0, 0, 1.0]])
    Fm = np.array([[-1, 0, G0], [0, -1, e_], [0, 0, -1.0]])
    jinv = 1.0 / (2 * w_dddl)

    def Pmul(x):
        r = Pd * x
        r[:-1, 2] += cc * x[1:, 2]
        r[1:, 2] += cc * x[:-1, 2]
        return r

    def Amul(x):
        r0 = x[:-1, 0] + a_ * x[:-1, 1] + b_ * x[:-1, 2] - x[1:, 0] + G0 * x[1:, 2]
        r1 = x[:-1, 1] + e_ * x[:-1, 2] - x[1:, 1] + e_ * x[1:, 2]
        return np.stack([r0, r1], axis=1)

    def Atmul(y):
        r = np.zeros((n, 3))
        r[:-1, 0] += y[:, 0]
        r[:-1, 1] += a_ * y[:, 0] + y[:, 1]
        r[:-1, 2] += b_ * y[:, 0] + e_ * y[:, 1]
        r[1:, 0] -= y[:, 0]
        r[1:, 1] -= y[:, 1]
        r[1:, 2] += G0 * y[:, 0] + e_ * y[:, 1]
        return r

    def factor(w):
        wa = w[:, 0] + w[:, 1] + w[:, 4] + w[:, 5]
        wb = w[:, 2] + w[:, 3] + w[:, 6] + w[:, 7]
        s01 = d1 * wa - d2 * wb
        s22 = Pnod[:, 2] + w[:, 10] + w[:, 13]
        D0 = Pnod[:, 0] + w[:, 8] + w[:, 11]
        D1 = Pnod[:, 1] + w[:, 9] + w[:, 12]
        r00sq = D0 + wa + wb
        r00 = np.sqrt(r00sq)
        r01 = s01 / r00
        r11sq = D1 + (D0 * (d1 * d1 * wa + d2 * d2 * wb) + wa * wb * (d1 + d2) ** 2) / r00sq
        r11 = np.sqrt(r11sq)
        Q = np.zeros((n, 3, 3))
        Q[:, 0, 0] = 1 / r00
        Q[:, 0, 1] = -r01 / (r00 * r11)
        Q[:, 1, 1] = 1 / r11
        Q[:, 2, 2] = 1 / np.sqrt(s22)
        Um = np.einsum('ab,nbc->nac', E, Q)
        Vm = np.einsum('ab,nbc->nac', Fm, Q)
        EXE = np.einsum('nab,ncb->nac', Um, Um)
        FXF = np.einsum('nab,ncb->nac', Vm, Vm)
        FXE = np.einsum('nab,ncb->nac', Vm, Um)
        Kd = EXE[:-1] + FXF[1:]
        Kd[:, 2, 2] += jinv
        B = FXE[1:-1]
        L = np.zeros((n - 1, 3, 3))
        M = np.zeros((n - 1, 3, 3))
        ok = True
        for i in range(n - 1):
            D = Kd[i].copy()
            if i > 0:
                M[i] = np.linalg.solve(L[i - 1], B[i - 1])
                D = D - M[i].T @ M[i]
            try:
                L[i] = np.linalg.cholesky(D)
            except np.linalg.LinAlgError:
                ok = False
                break
        return dict(Q=Q, L=L, M=M, ok=ok)

    def kkt_solve(F, g, by):
        Q, L, M = F['Q'], F['L'], F['M']
        Xg = np.einsum('nab,nb->na', Q, np.einsum('nba,nb->na', Q, g))
        r = np.einsum('ab,nb->na', E, Xg[:-1]) + np.einsum('ab,nb->na', Fm, Xg[1:])
        r[:, 0] -= by[:, 0]
        r[:, 1] -= by[:, 1]
        t = np.zeros((n - 1, 3))
        for i in range(n - 1):
            rr = r[i] if i == 0 else r[i] - M[i].T @ t[i - 1]
            t[i] = np.linalg.solve(L[i], rr)
        yy = np.zeros((n - 1, 3))
        for i in range(n - 2, -1, -1):
            rr = t[i] if i == n - 2 else t[i] - M[i + 1] @ yy[i + 1]
            yy[i] = np.linalg.solve(L[i].T, rr)
        Aty = np.zeros((n, 3))
        Aty[:-1] += np.einsum('ba,nb->na', E, yy)
        Aty[1:] += np.einsum('ba,nb->na', Fm, yy)
        x = np.einsum('nab,nb->na', Q, np.einsum('nba,nb->na', Q, g - Aty))
        return x, yy[:, :2].copy()

    F = factor(np.ones((n, 14)))
    x, y = kkt_solve(F, -q + Gt_mul(h, d1, d2), np.zeros((n - 1, 2)))
    z = G_mul(x, d1, d2) - h
    s = -z
    nrms = np.sqrt((s * s).sum())
    ts = -s.min()
    if ts >= -1e-8 * max(nrms, 1.0):
        s = s + (1.0 + ts)
    tz = -z.min()
    if tz >= -1e-8 * max(nrms, 1.0):
        z = z + (1.0 + tz)
    gap = (s * z).sum()
    gap_min = gap
    resx0 = max(1.0, np.sqrt((q * q).sum()))
    resy0 = 1.0
    resz0 = max(1.0, np.sqrt((h * h).sum()))
    for it in range(maxiters + 1):
        Pxq = Pmul(x) + q
        f0 = 0.5 * ((x * Pxq).sum() + (x * q).sum())
        rx = Pxq + Atmul(y) + Gt_mul(z, d1, d2)
        resx = np.sqrt((rx * rx).sum())
        ry = Amul(x)
        resy = np.sqrt((ry * ry).sum())
        rz = s + G_mul(x, d1, d2) - h
        resz = np.sqrt((rz * rz).sum())
        pcost = f0
        dcost = f0 + (y * ry).sum() + (z * rz).sum() - gap
        relgap = gap / -pcost if pcost < 0 else (gap / dcost if dcost > 0 else None)
        pres = max(resy / resy0, resz / resz0)
        dres = resx / resx0
        if trace is not None:
            trace.append((it, gap, pres, dres))
        if pres <= feastol and dres <= feastol and (gap <= abstol or (relgap is not None and relgap <= reltol)):
            return x, 'optimal', it
        if it == maxiters:
            return x, 'unknown', it
        if not np.isfinite(gap + resx + resz):
            return x, 'unknown', it
        if diverge_factor and it > 0 and gap > diverge_factor * gap_min:
            return x_prev, 'unknown', it
        gap_min = min(gap_min, gap)
        x_prev = x
        isv = 1.0 / s
        w = z * isv
        F = factor(w)
        if not F['ok']:
            return x, 'unknown', it
        mu = gap / m
        sigma = 0.0
        for i in (0, 1):
            bs = -s * z + sigma * mu
            if i == 1:
                bs = bs - ws3
            tq = (bs + z * rz) * isv
            g = -rx - Gt_mul(tq, d1, d2)
            dx, dy = kkt_solve(F, g, -ry)
            Gdx = G_mul(dx, d1, d2)
            dz = tq + w * Gdx
            dsv = -rz - Gdx
            dsdz = (dsv * dz).sum()
            if i == 0:
                ws3 = dsv * dz
            ts = (-dsv * isv).max()
            tz = (-dz / z).max()
            t = max(0.0, ts, tz)
            if t == 0:
                step = 1.0
            else:
                step = min(1.0, 1.0 / t) if i == 0 else min(1.0, 0.99 / t)
            if i == 0:
                sigma = min(1.0, max(0.0, 1.0 - step + dsdz / gap * step ** 2)) ** 3
        x = x + step * dx
        y = y + step * dy
        s = s + step * dsv
        z = z + step * dz
        gap = (s * z).sum()
