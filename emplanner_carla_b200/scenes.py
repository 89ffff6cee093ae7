"""Synthetic Frenet scenes for parity tests and the benchmark.

Deterministic in (config_id, frame index): every field has its own PCG64
stream, so the first n frames of a larger batch are exactly the n-frame batch
(the CPU baseline times a prefix of the GPU batch).
"""
from dataclasses import dataclass
import numpy as np

KMAX = 8
BASE_SEED = 20261017


@dataclass(frozen=True)
class Lattice:
    """S-L lattice parameters = keyword defaults of DP_algorithm (reference path_planning.py:276-279)."""
    name: str
    row: int
    col: int
    sample_s: float
    sample_l: float = 1.5
    sampling_res: float = 2

    def kwargs(self):
        return dict(row=self.row, col=self.col, sample_s=self.sample_s, sample_l=self.sample_l,
                    sampling_res=self.sampling_res)

    @property
    def horizon(self):
        return self.col * self.sample_s

    @property
    def edges(self):
        return self.row + (self.col - 1) * self.row * self.row


LATTICE_A = Lattice("A_12x6", 12, 6, 15)        # what every reference caller uses
LATTICE_B = Lattice("B_60x9", 9, 60, 1.5)       # BASELINE.json configs[0]
LATTICE_C = Lattice("C_60x11", 11, 60, 1.5)     # BASELINE.json configs[1..2]
LATTICES = {"A": LATTICE_A, "B": LATTICE_B, "C": LATTICE_C}


def _rng(config_id, field):
    return np.random.default_rng([BASE_SEED + int(config_id), int(field)])


def sl_scenes(batch, lattice, config_id, k_fixed=None):
    """Returns dict of float64/int32 arrays: obs_s[B,8], obs_l[B,8] (NaN padded), nobs[B],
    s0, l0, dl0, ddl0 [B]."""
    B = int(batch)
    s0 = _rng(config_id, 0).uniform(0.0, 3.0, B)
    l0 = _rng(config_id, 1).uniform(-0.5, 0.5, B)
    dl0 = _rng(config_id, 2).uniform(-0.05, 0.05, B)
    ddl0 = _rng(config_id, 3).uniform(-0.01, 0.01, B)
    if k_fixed is None:
        nobs = _rng(config_id, 4).integers(0, KMAX + 1, B).astype(np.int32)
    else:
        nobs = np.full(B, int(k_fixed), dtype=np.int32)
    jitter = _rng(config_id, 5).uniform(0.0, 4.0, (B, KMAX))
    lat = _rng(config_id, 6).uniform(-3.0, 3.0, (B, KMAX))
    k = np.arange(KMAX, dtype=np.float64)[None, :]
    spacing = (lattice.horizon - 25.0) / np.maximum(nobs, 1)[:, None]
    obs_s = 10.0 + k * spacing + jitter
    active = k < nobs[:, None]
    obs_s = np.where(active, obs_s, np.nan)
    obs_l = np.where(active, lat, np.nan)
    return dict(obs_s=obs_s, obs_l=obs_l, nobs=nobs, s0=s0, l0=l0, dl0=dl0, ddl0=ddl0)


def st_scenes(batch, config_id, k_fixed=KMAX):
    """S-T DP scenes (BASELINE.json configs[3]): obstacle entry/exit (s,t) pairs, NaN = absent."""
    B = int(batch)
    s_in = _rng(config_id, 10).uniform(10.0, 40.0, (B, KMAX))
    s_out = s_in + _rng(config_id, 11).uniform(5.0, 25.0, (B, KMAX))
    t_in = _rng(config_id, 12).uniform(0.5, 4.0, (B, KMAX))
    t_out = t_in + _rng(config_id, 13).uniform(1.0, 3.0, (B, KMAX))
    v0 = _rng(config_id, 14).uniform(3.0, 12.0, B)
    if k_fixed is None:
        nobs = _rng(config_id, 15).integers(0, KMAX + 1, B).astype(np.int32)
    else:
        nobs = np.full(B, int(k_fixed), dtype=np.int32)
    active = np.arange(KMAX)[None, :] < nobs[:, None]
    nanify = lambda a: np.where(active, a, np.nan)
    return dict(s_in=nanify(s_in), s_out=nanify(s_out), t_in=nanify(t_in), t_out=nanify(t_out),
                v0=v0, nobs=nobs)


def cycle_scenes(batch, lattice, config_id=4, st_k_fixed=None):
    """Full EM cycle scenes (BASELINE.json configs[4]): config-3 S-L inputs + config-4 S-T inputs per frame,
    plus the plan-start acceleration and the plan time stamp."""
    sl = sl_scenes(batch, lattice, config_id)
    st = st_scenes(batch, config_id, k_fixed=st_k_fixed)
    a0 = _rng(config_id, 20).uniform(-1.0, 1.0, int(batch))
    now = 0.1 * np.arange(int(batch), dtype=np.float64)
    return dict(sl=sl, st=st, a0=a0, now=now)


def reference_line(n_nodes=400, spacing=2.0, amp=8.0, wavelength=60.0):
    """Synthetic global route y = amp*sin(x/wavelength) as [(x, y), ...]."""
    xs = np.arange(n_nodes, dtype=np.float64) * spacing
    return np.stack([xs, amp * np.sin(xs / wavelength)], axis=1)
