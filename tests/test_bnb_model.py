"""CPU model of the exact branch and bound used by the two DP kernels (csrc/sl_dp.cu, csrc/st_dp.cu).

Claim: if every candidate's lower bound satisfies lb <= exact, then evaluating exactly only (a) per round the
not-yet-evaluated candidate(s) with the smallest lb that is still <= ub and (b) finally every candidate left with
lb <= ub, and leaving lb in place for the rest, gives the same first-minimum (strict '<' over ascending k) as the
full sweep - including ties between exact values and ties lb == ub.  The model follows the kernels' control flow;
the adversarial inputs are small integers, so ties are everywhere."""
import numpy as np


def full_sweep(exact):
    """Reference: first minimum over ascending k (path_planning.py:344, speed_planning_test.py:145)."""
    best, kb = np.inf, 0
    for k, v in enumerate(exact):
        if v < best:
            best, kb = v, k
    return best, kb


def bnb(lb, exact, rounds, m, always_exact=()):
    """The kernels' procedure for one target row.  always_exact: candidates evaluated unconditionally (the S-T
    kernel's source row 0).  Returns (best, kb, number of exact evaluations)."""
    cand = np.array(lb, dtype=float)
    done = np.zeros(len(lb), bool)
    n_eval = 0
    for k in always_exact:
        cand[k], done[k] = exact[k], True
        n_eval += 1
    ub = np.inf
    for r in range(rounds):
        last = r == rounds - 1
        if r > 0:
            ub = min([cand[k] for k in range(len(lb)) if done[k]], default=np.inf)
        picked = []
        if not last:
            for _ in range(m):
                best, kb = np.inf, -1
                for k in range(len(lb)):
                    if not done[k] and k not in picked and cand[k] < best:
                        best, kb = cand[k], k
                if kb < 0 or not best <= ub:
                    break
                picked.append(kb)
        else:
            picked = [k for k in range(len(lb)) if not done[k] and cand[k] <= ub]
        for k in picked:
            cand[k], done[k] = exact[k], True
            n_eval += 1
    best, kb = full_sweep(cand)
    return best, kb, n_eval


def test_branch_and_bound_equals_full_sweep_with_ties():
    rng = np.random.default_rng(3)
    saved = total = 0
    for trial in range(4000):
        n = int(rng.integers(1, 41))
        lb = rng.integers(0, 12, n).astype(float)
        extra = rng.integers(0, 6, n).astype(float) * (rng.random(n) < 0.5)      # exact - lb >= 0, often 0
        exact = lb + extra
        want = full_sweep(exact)
        # S-L kernel: one bounding round with the single best candidate, then the survivors
        got = bnb(lb, exact, rounds=2, m=1)
        assert got[:2] == want, (trial, lb, exact)
        # S-T kernel: source row 0 always exact, PRUNE_M = 1, PRUNE_ROUNDS = 3
        got3 = bnb(lb, exact, rounds=3, m=1, always_exact=(0,))
        assert got3[:2] == want, (trial, lb, exact)
        # other settings of the tuning macros
        assert bnb(lb, exact, rounds=2, m=4, always_exact=(0,))[:2] == want
        assert bnb(lb, exact, rounds=4, m=2)[:2] == want
        saved += n - got3[2]
        total += n
    assert saved > 0.3 * total                                 # and it does prune


def test_lower_bound_is_monotone_in_floating_point():
    """The bound used by the kernels: fl(fl(x + c_acc) + c_ref) + cp is non-decreasing in x >= 0 (rounding is
    monotone), so dropping the non-negative obstacle term can never overshoot - checked on adversarial magnitudes
    (1e12 collision weights against O(1) costs)."""
    rng = np.random.default_rng(4)
    for _ in range(20000):
        c_acc, c_ref, cp = rng.uniform(0, 10.0 ** rng.integers(-3, 13), 3)
        x = rng.uniform(0, 10.0 ** rng.integers(-16, 13))
        lb = ((0.0 + c_acc) + c_ref) + cp
        ex = ((x + c_acc) + c_ref) + cp
        assert lb <= ex
        assert (0.0 + c_acc) == c_acc


def test_settled_column_shortcut():
    """S-L kernel, columns whose work-list vote is empty: the row's minimum is its first-evaluated candidate k0 (the first
    minimum of the lower bounds), exact - no final scan.  With ties everywhere: whenever no other candidate has
    lb <= exact[k0], (exact[k0], k0) is what the full sweep returns."""
    rng = np.random.default_rng(5)
    settled = 0
    for trial in range(20000):
        n = int(rng.integers(1, 17))
        lb = rng.integers(0, 12, n).astype(float)
        exact = lb + rng.integers(0, 6, n).astype(float) * (rng.random(n) < 0.5)
        _, k0 = full_sweep(lb)
        ub = exact[k0]
        if any(lb[k] <= ub for k in range(n) if k != k0):
            continue                                            # the kernel builds a work list and re-scans the row
        settled += 1
        assert full_sweep(exact) == (ub, k0), (trial, lb, exact)
    assert settled > 2000
