"""CPU model of the S-L DP kernel's (candidate, sample) pair evaluation (csrc/sl_dp.cu, exact_pairs) against the oracle's
cal_obs_cost restatement (oracle/sl_dp.py:_obs_cost, path_planning.py:588-609), bit for bit.

The kernel classifies every (candidate, sample) pair on its own lane and parks one term per pair: the near-miss cost
5000/d2, +0.0 for a sample that is out of reach, -w_collision for a hit (the sign bit is the hit flag); the candidate's
lane then adds |term| in sample order while it is alive and dies at the first set sign bit.  This file pins the
argument that the sum has the bits of the reference's loop with its `break` - including w_collision = 0 (the flag is
then the sign of -0.0), NaN obstacles (neither near nor hit) and samples beyond the 6 m longitudinal reach."""
import numpy as np

from oracle import sl_dp

DANGER2, SAFE2, NEAR = sl_dp.DANGER2, sl_dp.SAFE2, sl_dp.NEAR_COST


def _pair_terms(dl2, dlat, w_col):
    """Phase 1 of exact_pairs: one term per pair, no knowledge of the other samples."""
    d2 = dl2 + dlat * dlat
    reach = dl2 < SAFE2
    hit = reach & (d2 <= DANGER2)
    near = reach & ~hit & (d2 < SAFE2)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(hit, -w_col, np.where(near, NEAR / d2, 0.0))


def _ordered_merge(terms):
    """Phase 2: terms [..., 10] -> cost [...]; alive until the first set sign bit (inclusive)."""
    one = np.zeros(terms.shape[:-1])
    alive = np.ones(terms.shape[:-1], dtype=bool)
    for m in range(terms.shape[-1]):
        v = terms[..., m]
        one = np.where(alive, one + np.abs(v), one)
        alive = alive & ~np.signbit(v)
    return one


def _cases(rng, n):
    dlon = rng.uniform(-8.0, 8.0, (n, 10))
    dlon[: n // 8] = np.sort(rng.uniform(-7.0, 7.0, (n // 8, 10)), axis=1)          # an obstacle passing by
    dlat = rng.uniform(-7.0, 7.0, (n, 10))
    dlat[n // 8: n // 4] *= 0.2                                                    # head-on: hits early
    dlon[-4:] = np.nan                                                             # NaN obstacle slot
    # exact boundaries of the two range tests
    dlon[-8, :] = 0.0
    dlat[-8, :] = [4.0, 6.0, np.nextafter(4.0, 5), np.nextafter(6.0, 5), 4.0, 0.0, 5.0, 5.0, 5.0, 5.0]
    return dlon * dlon, dlat


def test_pair_merge_has_the_bits_of_the_sequential_loop():
    rng = np.random.default_rng(20261017)
    dl2, dlat = _cases(rng, 4096)
    for w_col in (1e12, 1.0, 0.0):
        ref = sl_dp._obs_cost(w_col, [dl2[:, m] + dlat[:, m] * dlat[:, m] for m in range(10)])
        got = _ordered_merge(_pair_terms(dl2, dlat, w_col))
        assert np.array_equal(ref.view(np.int64), got.view(np.int64)), w_col
    # the cases cover what they claim to
    d2 = dl2 + dlat * dlat
    first_hit = np.argmax(d2 <= DANGER2, axis=1)
    assert ((d2 <= DANGER2).any(1) & (first_hit > 0) & (first_hit < 9)).sum() > 100       # a break in mid-loop
    assert ((d2 > DANGER2) & (d2 < SAFE2)).sum() > 1000 and np.isnan(d2).any()


def test_zero_weight_hit_still_stops_the_sum():
    """w_collision = 0: the hit adds 0 but the reference still breaks; the flag must survive as the sign of -0.0."""
    dl2 = np.zeros((1, 10))
    dlat = np.array([[5.0, 5.0, 3.0, 5.0, 5.0, 5.0, 5.0, 5.0, 5.0, 5.0]])           # near, near, hit, near ...
    got = _ordered_merge(_pair_terms(dl2, dlat, 0.0))
    assert got[0] == NEAR / 25.0 + NEAR / 25.0
    assert got[0] == sl_dp._obs_cost(0.0, [dl2[:, m] + dlat[:, m] ** 2 for m in range(10)])[0]


def test_divide_by_ten_and_by_rows_by_multiply_shift():
    """The kernel's index arithmetic: pair -> (candidate, sample) and edge -> (target, source) without integer division."""
    p = np.arange(1024)
    assert np.array_equal((p * 205) >> 11, p // 10)
    for R in range(1, 33):
        magic = 65536 // R + 1
        e = np.arange(R * R)
        assert np.array_equal((e * magic) >> 16, e // R), R
