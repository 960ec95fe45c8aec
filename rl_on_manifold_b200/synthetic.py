"""Seeded synthetic (q, dq, s, alpha) batches of the shapes BASELINE.json names.

The draws themselves live in `_inputs.py` (pure torch, importable without the native library).  The slack
variables start from the reference's own reset rule (atacom.py:145-149), evaluated by the slack-init kernel,
and are then pushed into an interior / boundary mix: 85 % of the environments keep every slack >= 0.3,
15 % get one slack drawn from U(0, 0.02) (an active constraint).
"""
from . import _lib, projection
from ._inputs import (IIWA_HOME, IIWA_Q_MAX, IIWA_VEL_MAX, PLANAR_Q_MAX, PLANAR_VEL_MAX, point_reach_batch,  # noqa: F401
                      slack_mix, state_batch)


def device_batch(family, B, seed, device, n_ctrl_joints=6, params=None):
    """Full (q, dq, s, alpha) batch resident on `device`, slacks from the slack-init kernel + mix."""
    if params is None:
        params = _lib.default_params(family, n_ctrl_joints) if family == "iiwa" else _lib.default_params(family)
    q, dq, alpha = state_batch(family, B, seed, n_ctrl_joints, params)
    q, dq, alpha = q.to(device), dq.to(device), alpha.to(device)
    s = projection.slack_init(family, q, dq, params, n_ctrl_joints=n_ctrl_joints)
    s = slack_mix(s, seed)
    return q, dq, s, alpha


def point_reach_device_batch(B, seed, device, n_objects=4, params=None):
    """(q, dq, p, dp, s, action) on `device` for env C, slacks from PointReachAtacom.reset's rule
    (collision_avoidance_atacom.py:25) through the slack-init kernel."""
    if params is None:
        params = _lib.default_params("point_reach")
    q, dq, p, dp, action = (t.to(device) for t in point_reach_batch(B, seed, n_objects))
    s = projection.point_reach_slack_init(q, p, params)
    return q, dq, p, dp, s, action
