"""Generate tests/golden/reference_golden.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference's own classes are imported under the shims of oracle/ref_loader.py and
driven through their public methods; inputs and outputs are stored as float64.
The GPU box has no /root/reference: tests there read the committed .npz.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore", category=SyntaxWarning)

from oracle import ref_loader  # noqa: E402

R = ref_loader.load()
out = {}

# ---------------------------------------------------------------- rref vs the reference's rref
rng = np.random.default_rng(20211108)
mats, res05, resdef = [], [], []
for _ in range(40):
    m, n = rng.integers(1, 7), rng.integers(1, 10)
    V = rng.normal(size=(m, n))
    if rng.uniform() < 0.5:
        V[:, rng.integers(n)] *= 0.01            # make the tolerance branch fire sometimes
    pad = np.full((6, 9), np.nan)
    pad[:m, :n] = V
    mats.append(pad)
    a = np.full((6, 9), np.nan); a[:m, :n] = R.rref(V, row_vectors=True, tol=0.05)
    b = np.full((6, 9), np.nan); b[:m, :n] = R.rref(V, row_vectors=True)
    res05.append(a); resdef.append(b)
out["rref_in"], out["rref_tol005"], out["rref_default"] = map(np.array, (mats, res05, resdef))

# ---------------------------------------------------------------- Circle A trajectory (CircleEnvAtacom)
def circle_traj(cls, act_dim, steps, seed, fixed=None):
    env = cls()
    st0 = env.reset().copy()
    rng = np.random.default_rng(seed)
    acts, states, ss, rews, aa, ab, ae = [], [st0], [env.s.copy()], [], [], [], []
    for i in range(steps):
        a = np.array(fixed[i]) if fixed is not None and i < len(fixed) else rng.uniform(-1.3, 1.3, act_dim)
        st, r, _, _ = env.step(a)
        acts.append(a); states.append(st.copy()); ss.append(env.s.copy()); rews.append(r)
        aa.append(env._act_a.copy()); ab.append(env._act_b.copy()); ae.append(env._act_err.copy())
    logs = np.array(env.get_constraints_logs())          # (c_avg, c_max, c_dq_max) over the trajectory
    return dict(actions=np.array(acts), states=np.array(states), s=np.array(ss), rewards=np.array(rews),
                act_a=np.array(aa), act_b=np.array(ab), act_err=np.array(ae), logs=logs)

for key, val in circle_traj(R.CircleEnvAtacom, 1, 500, 1, fixed=[[0.7], [-0.2], [1.5], [0.3]]).items():
    out["circleA_" + key] = val
for key, val in circle_traj(R.CircleEnvErrorCorrection, 2, 300, 2, fixed=[[0.7, 0.1], [-0.2, 0.5]]).items():
    out["circleE_" + key] = val

# single projections at arbitrary states, through the reference's step_action_function
env = R.CircleEnvAtacom()
env.reset()
rng = np.random.default_rng(5)
Q, DQ, S, AL, U, S2 = [], [], [], [], [], []
for i in range(64):
    th = rng.uniform(-np.pi / 6, 7 * np.pi / 6)
    q = np.array([np.cos(th), np.sin(th)]) * (1 + rng.uniform(-1e-3, 1e-3))
    dq = np.array([-np.sin(th), np.cos(th)]) * rng.uniform(-1, 1) + rng.uniform(-0.01, 0.01, 2)
    env.q, env.dq = q.copy(), dq.copy()
    env._compute_slack_variables()
    if i % 4 == 0:
        env.s = np.array([0.0]) if i % 8 == 0 else rng.uniform(0, 0.02, 1)
    s0 = env.s.copy()
    alpha = rng.uniform(-10, 10, 1)
    u = env.step_action_function(np.concatenate([q, dq]), alpha)
    Q.append(q); DQ.append(dq); S.append(s0); AL.append(alpha); U.append(u); S2.append(env.s.copy())
out["circleP_q"], out["circleP_dq"], out["circleP_s"], out["circleP_alpha"], out["circleP_u"], out["circleP_s_new"] = \
    map(np.array, (Q, DQ, S, AL, U, S2))

# ---------------------------------------------------------------- Collision C trajectory (PointReachAtacom)
np.random.seed(1)
env = R.PointReachAtacom(n_objects=4, random_walk=True)
st = env.reset().copy()
rng = np.random.default_rng(3)
pre, acts, s_hist, u_hist, post, obj_draws, rews = [], [], [env.s.copy()], [], [], [], []
_uniform = np.random.uniform
for i in range(300):
    a = np.array([0.5, -0.5]) if i == 0 else (np.array([1.0, 1.0]) if i == 1 else rng.uniform(-1, 1, 2))
    pre.append(env._state.copy())
    draws = []

    def _recording_uniform(*args, **kwargs):             # the obstacles' random-walk draws of this step
        v = _uniform(*args, **kwargs)
        draws.append(np.array(v, dtype=np.float64))
        return v

    np.random.uniform = _recording_uniform
    try:
        _, r, _, _ = env.step(a)
    finally:
        np.random.uniform = _uniform
    acts.append(a); s_hist.append(env.s.copy()); u_hist.append(env._action.copy()); post.append(env._state.copy())
    obj_draws.append(np.concatenate(draws)); rews.append(r)
out["collC_pre"], out["collC_actions"], out["collC_s"], out["collC_u"], out["collC_post"] = \
    map(np.array, (pre, acts, s_hist, u_hist, post))
out["collC_obj_draws"], out["collC_rewards"] = np.array(obj_draws), np.array(rews)   # draws: U(-1,1), [T, 2 * n_objects]
out["collC_logs"] = np.array(env.get_constraints_logs())

# ---------------------------------------------------------------- generic wrapper, synthetic ConstraintsSets
class _Base(R.Environment):
    def __init__(self, dim):
        box = R.Box(-np.ones(2 * dim), np.ones(2 * dim))
        super().__init__(R.MDPInfo(box, R.Box(-np.ones(dim), np.ones(dim)), 0.99, 100))
        self._dim = dim

    def reset(self, state=None):
        return np.zeros(2 * self._dim)

    def _create_observation(self, s):
        return s


class _Wrap(R.AtacomEnvWrapper):
    def _get_q(self, st): return st[:self.dims['q']]
    def _get_dq(self, st): return st[self.dims['q']:]
    def acc_to_ctrl_action(self, ddq): return ddq


def generic_cases(n, F, G, count, seed, s_small_frac):
    rng = np.random.default_rng(seed)
    rec = {k: [] for k in ("c", "J", "b", "dq", "s", "alpha", "ddq", "s_new", "act_a", "act_b", "act_err")}
    K_f, K_g = rng.uniform(0.1, 1.0, F), rng.uniform(0.3, 2.0, G)
    K_c = rng.uniform(20, 200)
    vel_max, acc_max = rng.uniform(1, 2, n), np.ones(n) * 10.0
    for _ in range(count):
        c, J, b = rng.normal(size=F + G) * 0.2, rng.normal(size=(F + G, n)), rng.normal(size=F + G)
        cur = dict(c=c, J=J, b=b)
        f = g = None
        if F:
            f = R.ConstraintsSet(n)
            f.add_constraint(R.ViabilityConstraint(n, F, lambda q: cur["c"][:F], lambda q: cur["J"][:F],
                                                   lambda q, dq: cur["b"][:F], K_f))
        if G:
            g = R.ConstraintsSet(n)
            g.add_constraint(R.ViabilityConstraint(n, G, lambda q: cur["c"][F:], lambda q: cur["J"][F:],
                                                   lambda q, dq: cur["b"][F:], K_g))
        w = _Wrap(_Base(n), n, vel_max, acc_max, f=f, g=g, Kc=K_c, Kq=2 * acc_max / vel_max, time_step=0.01)
        w.q, w.dq = rng.normal(size=n), rng.uniform(-1, 1, n) * vel_max
        if G:
            w.s = rng.uniform(0.3, 2.0, G)
            if rng.uniform() < s_small_frac:
                w.s[rng.integers(G)] = rng.uniform(0, 0.02)
        else:
            w._compute_slack_variables()         # G = 0: the reference leaves s = None (atacom.py:146)
        s0 = w.s.copy() if G else np.zeros(0)
        alpha = rng.uniform(-10, 10, n - F)
        ddq = w.step_action_function(np.zeros(2 * n), alpha)
        for k_, v in (("c", c), ("J", J), ("b", b), ("dq", w.dq), ("s", s0), ("alpha", alpha), ("ddq", ddq),
                      ("s_new", w.s.copy() if G else np.zeros(0)), ("act_a", w._act_a), ("act_b", w._act_b),
                      ("act_err", w._act_err)):
            rec[k_].append(np.array(v, dtype=np.float64))
    meta = np.concatenate([[n, F, G, K_c], K_f, K_g, vel_max, acc_max])
    return {k: np.array(v) for k, v in rec.items()}, meta


for seed, (tag, (n, F, G)) in enumerate({"g423": (4, 2, 3), "g313": (3, 1, 3), "g306": (3, 0, 6),
                                        "g6111": (6, 1, 11), "g310": (3, 1, 0)}.items()):
    rec, meta = generic_cases(n, F, G, 48, 100 + seed, 0.25)
    out[tag + "_meta"] = meta
    for k, v in rec.items():
        out[tag + "_" + k] = v

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")
