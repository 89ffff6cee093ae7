"""Oracle: S-L lattice DP + densify (numpy float64, batched over frames).

TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates reference planner/path_planning.py:
  DP_algorithm        :276-375   -> dp_batch()
  cal_start_cost      :435-514   -> _start_cost_*()
  cal_neighbor_cost   :517-585   -> _column_edges_*()
  cal_obs_cost        :588-609   -> _obs_cost()
  enrich_DP_s_l       :378-432   -> enrich()
and planner/planning_utils.py:
  cal_quintic_coefficient :671-703 -> _quintic_inv() (mode "inv") or the
                                      closed form (mode "closed")

Two coefficient modes:
  "inv"    evaluates the quintic exactly like the reference: 6x6
           ``np.linalg.inv(A) @ B`` in the absolute-s power basis.  Used only
           to pin this restatement bit-for-bit against the reference.
  "closed" uses the analytic quintic in the local variable u = s - s_start.
           This is the arithmetic the CUDA kernel implements; it differs from
           "inv" only by the rounding noise of the ill-conditioned inverse
           (cond(A) up to 1e16+), quantified in tests/test_oracle_golden.py.
All arithmetic is IEEE float64 with no fused multiply-add (numpy never
contracts), in a fixed association order that csrc/sl_dp.cu mirrors.
"""
import numpy as np

DANGER2 = 16.0   # danger_dis**2, path_planning.py:588,602
SAFE2 = 36.0     # safe_dis**2,   path_planning.py:588,606
NEAR_COST = 5000.0   # path_planning.py:608
LEFT_BIAS = 10000.0  # path_planning.py:317-318,341-342
NSAMP = 10           # path_planning.py:486,559


def node_l(R, sample_l):
    """Lateral offset of lattice row r: ((row+1)/2 - 1 - r)*sample_l (path_planning.py:326,478)."""
    c_row = (R + 1) / 2 - 1
    return (c_row - np.arange(R, dtype=np.float64)) * sample_l


def _poly_tables():
    t = np.arange(NSAMP, dtype=np.float64) / 10.0
    omt = 1.0 - t
    f0 = (t * t * t) * (10.0 + t * (-15.0 + 6.0 * t))
    f1 = 30.0 * ((t * t) * (omt * omt))
    f2 = 60.0 * (t * (omt * (1.0 - 2.0 * t)))
    a1 = 0.0
    a2 = 0.0
    for m in range(NSAMP):
        a1 = a1 + f1[m] * f1[m]
        a2 = a2 + f2[m] * f2[m]
    return f0, f1, f2, a1, a2


def _obs_cost(w_col, d2_list):
    """cal_obs_cost (path_planning.py:588-609): sequential over the 10 samples,
    `<=16` adds w_col and stops, (16,36) adds 5000/d2.  d2_list: list of arrays."""
    acc = np.zeros_like(d2_list[0])
    alive = np.ones(d2_list[0].shape, dtype=bool)
    for d2 in d2_list:
        hit = alive & (d2 <= DANGER2)
        near = alive & (d2 > DANGER2) & (d2 < SAFE2)
        with np.errstate(divide="ignore", invalid="ignore"):
            acc = acc + np.where(hit, w_col, np.where(near, NEAR_COST / d2, 0.0))
        alive = alive & ~hit
    return acc


def _collision(w_col, s_m, l_m, obs_s, obs_l, nobs):
    """Sum over obstacles, in list order, of cal_obs_cost (path_planning.py:576-583).
    s_m: [B,10]; l_m: [B,...,10] broadcastable; obs_*: [B,Kmax]; nobs: [B]."""
    extra = l_m.ndim - 2
    coll = np.zeros(l_m.shape[:-1])
    Kmax = obs_s.shape[1]
    for o in range(Kmax):
        act = o < nobs
        if not act.any():
            break
        os_ = np.where(act, obs_s[:, o], 1e30).reshape((-1,) + (1,) * extra)
        ol_ = np.where(act, obs_l[:, o], 1e30).reshape((-1,) + (1,) * extra)
        d2 = []
        for m in range(NSAMP):
            dlon = os_ - s_m[:, m].reshape((-1,) + (1,) * extra)
            dlat = ol_ - l_m[..., m]
            d2.append(dlon * dlon + dlat * dlat)
        coll = coll + _obs_cost(w_col, d2)
    return coll


def _quintic_inv(l0, dl0, ddl0, l1, dl1, ddl1, s0, s1):
    """cal_quintic_coefficient as written (planning_utils.py:692-703)."""
    A = np.array(
        [[1, s0, pow(s0, 2), pow(s0, 3), pow(s0, 4), pow(s0, 5)],
         [0, 1, 2 * s0, 3 * pow(s0, 2), 4 * pow(s0, 3), 5 * pow(s0, 4)],
         [0, 0, 2, 6 * s0, 12 * pow(s0, 2), 20 * pow(s0, 3)],
         [1, s1, pow(s1, 2), pow(s1, 3), pow(s1, 4), pow(s1, 5)],
         [0, 1, 2 * s1, 3 * pow(s1, 2), 4 * pow(s1, 3), 5 * pow(s1, 4)],
         [0, 0, 2, 6 * s1, 12 * pow(s1, 2), 20 * pow(s1, 3)]])
    B = np.array([l0, dl0, ddl0, l1, dl1, ddl1]).reshape((6, 1))
    return list((np.linalg.inv(A) @ B).squeeze())


def _edge_cost_inv(obs_s, obs_l, s0, l0, dl0, ddl0, s1, l1, T, w_col, w_s, w_ref):
    """One edge exactly as the reference evaluates it (path_planning.py:484-514 /
    :556-585): power basis in absolute s, `dddl` with the `s*2` slip (:498,:571)."""
    c = _quintic_inv(l0, dl0, ddl0, l1, 0, 0, s0, s1)
    s = np.zeros((NSAMP, 1))
    for m in range(NSAMP):
        s[m][0] = s0 + m * T / 10
    l = c[0] + c[1] * s + c[2] * (s ** 2) + c[3] * (s ** 3) + c[4] * (s ** 4) + c[5] * (s ** 5)
    dl = c[1] + 2 * c[2] * s + 3 * c[3] * (s ** 2) + 4 * c[4] * (s ** 3) + 5 * c[5] * (s ** 4)
    ddl = 2 * c[2] + 6 * c[3] * s + 12 * c[4] * (s ** 2) + 20 * c[5] * (s ** 3)
    dddl = 6 * c[3] + 24 * c[4] * s + 60 * c[5] * (s * 2)
    smooth = w_s[0] * (dl.T @ dl) + w_s[1] * (ddl.T @ ddl) + w_s[2] * (dddl.T @ dddl)
    ref = w_ref * (l.T @ l)
    coll = 0
    for o in range(len(obs_s)):
        d2 = (obs_s[o] - s) ** 2 + (obs_l[o] - l) ** 2
        one = 0
        for v in d2.squeeze():
            if v <= DANGER2:
                one += w_col
                break
            elif DANGER2 < v < SAFE2:
                one += NEAR_COST / v
        coll += one
    return float((smooth + coll + ref).squeeze())


def quintic_closed(l0, dl0, ddl0, l1, dl1, ddl1, s0, s1):
    """cal_quintic_coefficient (planning_utils.py:671-703) in closed form: local coefficients in u = s - s0,
    expanded to the absolute-s power basis the reference returns."""
    T = s1 - s0
    iT = 1.0 / T
    iT3 = (iT * iT) * iT
    iT4 = iT3 * iT
    iT5 = iT4 * iT
    D = l1 - l0
    T2 = T * T
    a = [l0, dl0, ddl0 / 2.0,
         ((20.0 * D - (8.0 * dl1 + 12.0 * dl0) * T) - (3.0 * ddl0 - ddl1) * T2) * (0.5 * iT3),
         ((-30.0 * D + (14.0 * dl1 + 16.0 * dl0) * T) + (3.0 * ddl0 - 2.0 * ddl1) * T2) * (0.5 * iT4),
         ((12.0 * D - (6.0 * (dl1 + dl0)) * T) - (ddl0 - ddl1) * T2) * (0.5 * iT5)]
    m = -s0
    m2, m3, m4, m5 = m * m, m * m * m, (m * m * m) * m, ((m * m * m) * m) * m
    return [a[0] + a[1] * m + a[2] * m2 + a[3] * m3 + a[4] * m4 + a[5] * m5,
            a[1] + 2.0 * a[2] * m + 3.0 * a[3] * m2 + 4.0 * a[4] * m3 + 5.0 * a[5] * m4,
            a[2] + 3.0 * a[3] * m + 6.0 * a[4] * m2 + 10.0 * a[5] * m3,
            a[3] + 4.0 * a[4] * m + 10.0 * a[5] * m2,
            a[4] + 5.0 * a[5] * m,
            a[5]]


def edge_cost_closed(obs_s, obs_l, s0, l0, dl0, ddl0, s1, l1, sample_s, w_col, w_s, w_ref, danger2=DANGER2, safe2=SAFE2):
    """One edge in closed form (what csrc/misc.cu::edge_cost_kernel and column 0 of K1 compute):
    cal_start_cost / cal_neighbor_cost (path_planning.py:435-585) with the quintic evaluated in u = s - s0.
    danger2 / safe2 = 9 / 25 gives planner/speed_planning.py cal_start_cost (:96-169) on an (s, t) quintic."""
    T = s1 - s0
    iT = 1.0 / T
    iT3 = (iT * iT) * iT
    iT4 = iT3 * iT
    iT5 = iT4 * iT
    D = l1 - l0
    T2 = T * T
    a1, a2 = dl0, ddl0 / 2.0
    a3 = ((20.0 * D - (8.0 * 0.0 + 12.0 * dl0) * T) - (3.0 * ddl0 - 0.0) * T2) * (0.5 * iT3)
    a4 = ((-30.0 * D + (14.0 * 0.0 + 16.0 * dl0) * T) + (3.0 * ddl0 - 2.0 * 0.0) * T2) * (0.5 * iT4)
    a5 = ((12.0 * D - (6.0 * (0.0 + dl0)) * T) - (ddl0 - 0.0) * T2) * (0.5 * iT5)
    c4 = a4 - (5.0 * a5) * s0
    c3 = (a3 - (4.0 * a4) * s0) + (10.0 * a5) * (s0 * s0)
    alpha, beta = 6.0 * c3, 24.0 * c4 + 120.0 * a5
    sdl = sddl = sj = sl2 = 0.0
    sm, lm = [], []
    for m in range(NSAMP):
        u = (m * sample_s) / 10.0
        s = s0 + u
        l = l0 + u * (a1 + u * (a2 + u * (a3 + u * (a4 + u * a5))))
        dl = a1 + u * (2.0 * a2 + u * (3.0 * a3 + u * (4.0 * a4 + u * (5.0 * a5))))
        ddl = 2.0 * a2 + u * (6.0 * a3 + u * (12.0 * a4 + u * (20.0 * a5)))
        jm = alpha + beta * s
        sdl, sddl, sj, sl2 = sdl + dl * dl, sddl + ddl * ddl, sj + jm * jm, sl2 + l * l
        sm.append(s)
        lm.append(l)
    smooth = (w_s[0] * sdl + w_s[1] * sddl) + w_s[2] * sj
    coll = 0.0
    for o in range(len(obs_s)):
        one = 0.0
        for m in range(NSAMP):
            dlon, dlat = obs_s[o] - sm[m], obs_l[o] - lm[m]
            d2 = dlon * dlon + dlat * dlat
            if d2 <= danger2:
                one = one + w_col
                break
            elif d2 < safe2:
                one = one + NEAR_COST / d2
        coll = coll + one
    return (smooth + coll) + w_ref * sl2


def dp_batch(obs_s, obs_l, nobs, s0, l0, dl0, ddl0, sampling_res=2, w_col=1e12,
             w_smooth=(300, 1000, 5000), w_ref=20, row=12, col=6, sample_s=15,
             sample_l=1.5, mode="closed", return_tables=False):
    """Batched DP_algorithm (path_planning.py:276-375), without the densify tail.

    obs_s, obs_l: [B,Kmax] (entries >= nobs[b] ignored); nobs: [B] int;
    s0,l0,dl0,ddl0: [B].  Returns rows [B,col] int32 (row index per column;
    -1 for K=0 frames, which take the straight branch :362-363), node_s [B,col],
    node_l [B,col], and final-column min cost [B].
    """
    obs_s = np.atleast_2d(np.asarray(obs_s, dtype=np.float64))
    obs_l = np.atleast_2d(np.asarray(obs_l, dtype=np.float64))
    nobs = np.atleast_1d(np.asarray(nobs, dtype=np.int64))
    s0, l0, dl0, ddl0 = [np.atleast_1d(np.asarray(v, dtype=np.float64)) for v in (s0, l0, dl0, ddl0)]
    B = s0.shape[0]
    R, C, T = int(row), int(col), float(sample_s)
    nl = node_l(R, sample_l)
    half = R >> 1
    bias = np.where(np.arange(R) < half, LEFT_BIAS, 0.0)
    w0, w1, w2 = [float(v) for v in w_smooth]
    um = (np.arange(NSAMP, dtype=np.float64) * T) / 10.0           # i*sample_s/10 (:493,:566)

    cost = np.full((B, R, C), np.inf)
    parent = np.ones((B, R, C), dtype=np.int32)                     # :304

    if mode == "inv":
        for b in range(B):
            K = int(nobs[b])
            if K == 0:
                continue
            os_, ol_ = list(obs_s[b, :K]), list(obs_l[b, :K])
            for i in range(R):
                cost[b, i, 0] = _edge_cost_inv(os_, ol_, s0[b], l0[b], dl0[b], ddl0[b],
                                               s0[b] + T, nl[i], T, w_col, (w0, w1, w2), w_ref)
                if i < half:
                    cost[b, i, 0] += LEFT_BIAS
            for j in range(1, C):
                for i in range(R):
                    for k in range(R):
                        e = _edge_cost_inv(os_, ol_, s0[b] + j * T, nl[k], 0, 0,
                                           s0[b] + (j + 1) * T, nl[i], T, w_col, (w0, w1, w2), w_ref)
                        tmp = cost[b, k, j - 1] + e
                        if i < half:
                            tmp += LEFT_BIAS
                        if tmp < cost[b, i, j]:
                            cost[b, i, j] = tmp
                            parent[b, i, j] = k
    else:
        f0, f1, f2, A1, A2 = _poly_tables()
        iT = 1.0 / T
        iT2 = iT * iT
        iT3 = iT2 * iT
        iT4 = iT3 * iT
        iT5 = iT4 * iT
        # ---- column 0: general quintic from the plan start (:468-514)
        T2 = T * T
        D0 = nl[None, :] - l0[:, None]                               # [B,R]
        a1 = dl0[:, None]
        a2 = ddl0[:, None] / 2.0
        vT = dl0[:, None] * T
        aT2 = ddl0[:, None] * T2
        a3 = ((20.0 * D0 - 12.0 * vT) - 3.0 * aT2) * (0.5 * iT3)
        a4 = ((-30.0 * D0 + 16.0 * vT) + 3.0 * aT2) * (0.5 * iT4)
        a5 = ((12.0 * D0 - 6.0 * vT) - aT2) * (0.5 * iT5)
        sb = s0[:, None]
        c4 = a4 - 5.0 * a5 * sb
        c3 = (a3 - 4.0 * a4 * sb) + 10.0 * a5 * (sb * sb)
        alpha = 6.0 * c3
        beta = 24.0 * c4 + 120.0 * a5
        s_m0 = s0[:, None] + um[None, :]                             # [B,10]
        sdl = np.zeros((B, R))
        sddl = np.zeros((B, R))
        sj = np.zeros((B, R))
        sl2 = np.zeros((B, R))
        l_m = np.empty((B, R, NSAMP))
        for m in range(NSAMP):
            u = um[m]
            lm = l0[:, None] + u * (a1 + u * (a2 + u * (a3 + u * (a4 + u * a5))))
            dlm = a1 + u * (2.0 * a2 + u * (3.0 * a3 + u * (4.0 * a4 + u * (5.0 * a5))))
            ddlm = 2.0 * a2 + u * (6.0 * a3 + u * (12.0 * a4 + u * (20.0 * a5)))
            jm = alpha + beta * s_m0[:, m][:, None]
            sdl = sdl + dlm * dlm
            sddl = sddl + ddlm * ddlm
            sj = sj + jm * jm
            sl2 = sl2 + lm * lm
            l_m[:, :, m] = lm
        smooth = (w0 * sdl + w1 * sddl) + w2 * sj
        coll = _collision(w_col, s_m0, l_m, obs_s, obs_l, nobs)
        cost[:, :, 0] = ((smooth + coll) + w_ref * sl2) + bias[None, :]
        # ---- columns 1..C-1 (:320-346).  Edge tables indexed [i (target), k (source)].
        Dik = nl[:, None] - nl[None, :]                              # l_i - l_k
        l_tab = nl[None, :, None] + Dik[:, :, None] * f0[None, None, :]   # [R,R,10]
        refc = np.zeros((R, R))
        for m in range(NSAMP):
            refc = refc + l_tab[:, :, m] * l_tab[:, :, m]
        refc = w_ref * refc
        D2 = Dik * Dik
        base_s = (w0 * A1) * iT2 + (w1 * A2) * iT4
        b4 = -15.0 * iT4
        b3 = 10.0 * iT3
        ch5 = 6.0 * iT5
        for j in range(1, C):
            sj0 = s0 + j * T                                         # pre_node_s (:330)
            ch4 = b4 - 5.0 * ch5 * sj0
            ch3 = (b3 - 4.0 * b4 * sj0) + 10.0 * ch5 * (sj0 * sj0)
            al = 6.0 * ch3
            be = 24.0 * ch4 + 120.0 * ch5
            s_m = sj0[:, None] + um[None, :]
            J = np.zeros(B)
            for m in range(NSAMP):
                v = al + be * s_m[:, m]
                J = J + v * v
            Sj = base_s + w2 * J                                     # [B]
            smooth = D2[None, :, :] * Sj[:, None, None]
            coll = _collision(w_col, s_m, np.broadcast_to(l_tab, (B, R, R, NSAMP)), obs_s, obs_l, nobs)
            edge = (smooth + coll) + refc[None, :, :]
            cand = (cost[:, None, :, j - 1] + edge) + bias[None, :, None]   # [B,i,k]
            kbest = np.argmin(cand, axis=2)                          # first min == strict '<' ascending k (:344)
            best = np.take_along_axis(cand, kbest[:, :, None], axis=2)[:, :, 0]
            cost[:, :, j] = best
            parent[:, :, j] = kbest

    rows = np.full((B, C), -1, dtype=np.int32)
    last = np.argmin(cost[:, :, -1], axis=1)                         # :349
    rows[:, C - 1] = last
    cur = last
    for j in range(C - 1, 0, -1):                                    # :355-359
        cur = parent[np.arange(B), cur, j]
        rows[:, j - 1] = cur
    has_obs = nobs > 0
    c_row = (R + 1) / 2 - 1
    rowsf = np.where(has_obs[:, None], rows.astype(np.float64), c_row)    # :363
    rows = np.where(has_obs[:, None], rows, -1).astype(np.int32)
    node_s = s0[:, None] + (np.arange(C, dtype=np.float64)[None, :] + 1.0) * T   # :369
    node_lv = (c_row - rowsf) * sample_l                             # :370
    final_cost = np.where(has_obs, cost[np.arange(B), last, -1], 0.0)
    if return_tables:
        return rows, node_s, node_lv, final_cost, cost, parent
    return rows, node_s, node_lv, final_cost


def max_enriched(col, sample_s, sampling_res):
    """Upper bound on the densified length (int(end-start) can round to floor(T) or floor(T)+-1)."""
    per = int(np.floor(float(sample_s) / float(sampling_res))) + 1
    return int(col) * per + 1


def enrich(node_s, node_lv, s0, l0, dl0, ddl0, resolution=2, mode="closed"):
    """enrich_DP_s_l (path_planning.py:378-432) for ONE frame.
    node_s/node_lv: length-C arrays.  Returns (s_list, l_list) as numpy arrays."""
    S = [float(s0)] + [float(v) for v in node_s]
    L = [float(l0)] + [float(v) for v in node_lv]
    out_s, out_l = [], []
    f_res = resolution
    for i in range(1, len(S)):
        ss, es = S[i - 1], S[i]
        sl_, el = L[i - 1], L[i]
        sdl, sddl = (float(dl0), float(ddl0)) if i == 1 else (0.0, 0.0)
        cnt = int(es - ss)                                           # :405,:423
        ks = np.arange(0, cnt, f_res)
        s = ss + ks
        if mode == "inv":
            c = _quintic_inv(sl_, sdl, sddl, el, 0, 0, ss, es)
            l = c[0] + c[1] * s + c[2] * (s ** 2) + c[3] * (s ** 3) + c[4] * (s ** 4) + c[5] * (s ** 5)
        else:
            T = es - ss
            iT = 1.0 / T
            iT3 = (iT * iT) * iT
            iT4 = iT3 * iT
            iT5 = iT4 * iT
            D = el - sl_
            vT = sdl * T
            aT2 = sddl * (T * T)
            a2 = sddl / 2.0
            a3 = ((20.0 * D - 12.0 * vT) - 3.0 * aT2) * (0.5 * iT3)
            a4 = ((-30.0 * D + 16.0 * vT) + 3.0 * aT2) * (0.5 * iT4)
            a5 = ((12.0 * D - 6.0 * vT) - aT2) * (0.5 * iT5)
            u = ks.astype(np.float64)
            l = sl_ + u * (sdl + u * (a2 + u * (a3 + u * (a4 + u * a5))))
        out_s += list(s)
        out_l += list(l)
    out_s.append(S[-1])
    out_l.append(L[-1])
    return np.array(out_s, dtype=np.float64), np.array(out_l, dtype=np.float64)


def dp_algorithm(obs_s_list, obs_l_list, plan_start_s, plan_start_l, plan_start_dl, plan_start_ddl,
                 sampling_res=2, w_collision_cost=1e12, w_smooth_cost=(300, 1000, 5000),
                 w_reference_cost=20, row=12, col=6, sample_s=15, sample_l=1.5, mode="closed"):
    """Single-frame DP_algorithm with the reference's signature (path_planning.py:276-279)."""
    K = len(obs_s_list)
    os_ = np.full((1, max(K, 1)), np.nan)
    ol_ = np.full((1, max(K, 1)), np.nan)
    os_[0, :K] = obs_s_list
    ol_[0, :K] = obs_l_list
    rows, ns, nlv, _ = dp_batch(os_, ol_, [K], [plan_start_s], [plan_start_l], [plan_start_dl],
                                [plan_start_ddl], sampling_res, w_collision_cost, w_smooth_cost,
                                w_reference_cost, row, col, sample_s, sample_l, mode=mode)
    s, l = enrich(ns[0], nlv[0], plan_start_s, plan_start_l, plan_start_dl, plan_start_ddl,
                  resolution=sampling_res, mode=mode)
    return s, l, rows[0]
