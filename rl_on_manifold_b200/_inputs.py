"""Seeded synthetic (q, dq, alpha) draws of the shapes BASELINE.json names — the pure part.

No import of the package or of libatacom_b200.so: `bench.py --impl reference` loads this file by path so that the
reference arm's process never maps the product's native code.  `synthetic.py` re-exports everything here and adds
the device-side pieces.  Inputs are drawn on the CPU with a seeded torch.Generator, so the NumPy oracle, the
reference arm and the kernels all see bit-identical fp32 values.
"""
import math

import torch

# joint angles that put the striker tip at (0.65, 0, 0.1505) in the robot frame, pointing down
# (the pose env_single.py:39-44 solves for by CLIK), found once with the oracle's FK
IIWA_HOME = (0.0, 0.2596, 0.0, -1.2247, 0.0, 1.4384, 0.0)
IIWA_Q_MAX = (2.9670597283903604, 2.0943951023931953, 2.9670597283903604, 2.0943951023931953,
              2.9670597283903604, 2.0943951023931953, 3.0543261909900763)           # urdf/iiwa_1.urdf:74..297
IIWA_VEL_MAX = (1.4835298641951802, 1.4835298641951802, 1.7453292519943295, 1.3089969389957472,
                2.2689280275926285, 2.356194490192345, 2.356194490192345)           # urdf/iiwa_1.urdf:74..297
PLANAR_Q_MAX = (2.9670597283903604, 2.0943951023931953, 2.0943951023931953)
PLANAR_VEL_MAX = (1.4835298641951802, 1.4835298641951802, 1.7453292519943295)


def _u(gen, shape, lo, hi):
    return torch.rand(shape, generator=gen, dtype=torch.float32) * (hi - lo) + lo


def state_batch(family, B, seed, n_ctrl_joints=6, params=None):
    """(q, dq, alpha) on the CPU, fp32, for one env family.  `params` (an AtacomParams) overrides the joint and
    velocity limits the draws are scaled by; the defaults are the families' own constants (fp32-rounded, as the
    C struct stores them)."""
    gen = torch.Generator().manual_seed(int(seed))
    f32 = lambda v: torch.tensor(list(v), dtype=torch.float64).float()
    if family == "circle":
        th = _u(gen, (B,), -math.pi / 6, 7 * math.pi / 6)
        rad = 1.0 + _u(gen, (B,), -1e-3, 1e-3)
        q = torch.stack([torch.cos(th), torch.sin(th)], 1) * rad[:, None]
        v = _u(gen, (B,), -1.0, 1.0)
        dq = torch.stack([-torch.sin(th), torch.cos(th)], 1) * v[:, None] + _u(gen, (B, 2), -0.01, 0.01)
        alpha = _u(gen, (B, 1), -10.0, 10.0)
    elif family == "planar":
        qmax = torch.tensor(list(params.env[5:8])) if params is not None else torch.tensor(PLANAR_Q_MAX)
        vmax = torch.tensor(list(params.vel_max[:3])) if params is not None else f32(PLANAR_VEL_MAX)
        q = _u(gen, (B, 3), -0.8, 0.8) * qmax
        dq = _u(gen, (B, 3), -0.5, 0.5) * vmax
        alpha = _u(gen, (B, 3), -10.0, 10.0)
    elif family == "iiwa":
        n = n_ctrl_joints
        vmax = torch.tensor(list(params.vel_max[:n])) if params is not None else f32(IIWA_VEL_MAX[:n])
        qmax = torch.tensor(IIWA_Q_MAX[:n])
        q = torch.tensor(IIWA_HOME[:n]) + _u(gen, (B, n), -0.35, 0.35)
        q = torch.minimum(torch.maximum(q, -0.95 * qmax), 0.95 * qmax)
        dq = _u(gen, (B, n), -0.5, 0.5) * vmax
        alpha = _u(gen, (B, n - 1), -10.0, 10.0)
    else:
        raise ValueError(family)
    return q.float().contiguous(), dq.float().contiguous(), alpha.contiguous()


def slack_mix(s, seed, interior=0.85, s_floor=0.3, s_active=0.02):
    """Apply the interior / boundary mix to slacks `s` [B, G] (any device); deterministic in seed: 85 % of the
    environments keep every slack >= 0.3, 15 % get one slack drawn from U(0, 0.02) (an active constraint)."""
    B, G = s.shape
    gen = torch.Generator().manual_seed(int(seed) + 7919)
    boundary = torch.rand(B, generator=gen) >= interior
    which = torch.randint(0, G, (B,), generator=gen)
    small = torch.rand(B, generator=gen, dtype=torch.float32) * s_active
    boundary, which, small = boundary.to(s.device), which.to(s.device), small.to(s.device)
    out = torch.clamp(s, min=s_floor)
    rows = torch.nonzero(boundary, as_tuple=True)[0]
    out[rows, which[rows]] = small[rows]
    return out.contiguous()


def point_reach_batch(B, seed, n_objects=4):
    """(q, dq, p, dp, action) on the CPU for env C."""
    gen = torch.Generator().manual_seed(int(seed))
    q = _u(gen, (B, 2), 1.0, 9.0)
    dq = _u(gen, (B, 2), -1.0, 1.0)
    p = _u(gen, (B, 2 * n_objects), 2.0, 8.0)
    dp = _u(gen, (B, 2 * n_objects), -1.0, 1.0)
    action = _u(gen, (B, 2), -1.0, 1.0)
    return q, dq, p, dp, action
