"""Oracle: the ATACOM projection step, one environment at a time, float64 NumPy.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, with fresh code,

  atacom/atacom.py:106-108      step(): action clip and alpha_max scaling
  atacom/atacom.py:117-121      acc_truncation
  atacom/atacom.py:123-139      step_action_function
  atacom/atacom.py:145-149      _compute_slack_variables
  atacom/atacom.py:151-165      _construct_Jc_psi
  atacom/atacom.py:167-196      _compute_error_correction / _compute_c
  atacom/constraints.py:33-43   ViabilityConstraint.fun / K_J / b
  atacom/error_correction_wrapper.py:102-134   variant "E"
  atacom/environments/collision_avoidance/collision_avoidance_atacom.py:18-48,72-127  variant "C"

An env family plugs in through `evaluate(q, dq) -> ConstraintEval` (the values
the reference obtains from its `fun / J / b` callbacks).
"""
from dataclasses import dataclass, field

import numpy as np

from . import nullspace as ns


@dataclass
class ConstraintEval:
    """Raw constraint callbacks' outputs at (q, dq): c(q), J(q), b(q,dq)."""
    c_f: np.ndarray
    J_f: np.ndarray
    b_f: np.ndarray
    c_g: np.ndarray
    J_g: np.ndarray
    b_g: np.ndarray


@dataclass
class Spec:
    """Constructor arguments of AtacomEnvWrapper (atacom.py:10-71) resolved to arrays."""
    n: int
    F: int
    G: int
    K_f: np.ndarray
    K_g: np.ndarray
    K_c: np.ndarray
    K_q: np.ndarray
    vel_max: np.ndarray
    acc_max: np.ndarray
    dt: float
    tol: float = 0.05          # atacom.py:128
    C: int = field(init=False)
    N: int = field(init=False)
    k: int = field(init=False)

    def __post_init__(self):
        self.C = self.F + self.G
        self.N = self.n + self.G
        self.k = self.n - self.F
        as_vec = lambda v, m: np.ones(m) * v if np.isscalar(v) else np.asarray(v, dtype=np.float64)
        self.K_f = as_vec(self.K_f, self.F)
        self.K_g = as_vec(self.K_g, self.G)
        self.K_c = as_vec(self.K_c, self.C)
        self.K_q = as_vec(self.K_q, self.n)
        self.vel_max = as_vec(self.vel_max, self.n)
        self.acc_max = as_vec(self.acc_max, self.n)

    @property
    def alpha_max(self):
        return np.ones(self.k) * self.acc_max.max()      # atacom.py:71


def viability_terms(spec, ev, dq):
    """constraints.py:33-43: K_J, fun(origin_constr=False), b  for f and g."""
    Jdq_f = ev.J_f @ dq if spec.F else np.zeros(0)
    Jdq_g = ev.J_g @ dq if spec.G else np.zeros(0)
    A_f = spec.K_f[:, None] * ev.J_f
    A_g = spec.K_g[:, None] * ev.J_g
    ct_f = ev.c_f + spec.K_f * Jdq_f
    ct_g = ev.c_g + spec.K_g * Jdq_g
    psi = np.concatenate([Jdq_f + spec.K_f * ev.b_f, Jdq_g + spec.K_g * ev.b_g])
    return A_f, A_g, ct_f, ct_g, psi


def stack_Jc(spec, A_f, A_g, s):
    """atacom.py:151-165."""
    Jc = np.zeros((spec.C, spec.N))
    Jc[:spec.F, :spec.n] = A_f
    Jc[spec.F:, :spec.n] = A_g
    Jc[spec.F:, spec.n:] = np.diag(s)
    return Jc


def acc_limits(spec, dq):
    """atacom.py:117-120."""
    up = np.maximum(np.minimum(spec.acc_max, -spec.K_q * (dq - spec.vel_max)), -spec.acc_max)
    lo = np.minimum(np.maximum(-spec.acc_max, -spec.K_q * (dq + spec.vel_max)), spec.acc_max)
    return lo, up


def clip_acc(spec, dq, ddq):
    lo, up = acc_limits(spec, dq)
    return np.minimum(np.maximum(ddq, lo), up)     # np.clip semantics: upper wins if crossed


def slack_init(spec, ev, dq):
    """atacom.py:145-149: s = sqrt(max(-2 (g + K_g J_g dq), 0))."""
    ct_g = ev.c_g + spec.K_g * (ev.J_g @ dq)
    return np.sqrt(np.maximum(-2.0 * ct_g, 0.0))


def scale_action(spec, action, variant="atacom"):
    """atacom.py:106-108 / error_correction_wrapper.py:102-104."""
    a = np.clip(action, -1.0, 1.0)
    return a * (spec.alpha_max if variant == "atacom" else spec.acc_max)


def null_coordinates(Jc, k, tol, basis="svd", pinv_Q=None, trace=None):
    """Nc (N x k): RREF-canonicalised null-space basis, atacom.py:127-128."""
    if pinv_Q is None:
        pinv, Q, _ = ns.svd_pinv_null(Jc)
    else:
        pinv, Q = pinv_Q
    if basis == "svd":
        V = np.zeros((k, Jc.shape[1]))
        kk = min(k, Q.shape[1])
        V[:kk] = Q[:, :kk].T                      # Nc[:, :k] silently truncates (atacom.py:128)
    elif basis == "lapack":
        if Q.shape[1] != k:                       # rank-deficient Jc: LAPACK's extra null vectors come out of its
            V = np.zeros((k, Jc.shape[1]))        # SVD iteration and are not restated; flagged by the callers
            V[:min(k, Q.shape[1])] = Q[:, :min(k, Q.shape[1])].T
        else:
            V = ns.lapack_null_basis(Jc).T
    elif basis == "canonical":
        V = ns.canonical_null_basis(Jc, k, 0.0 if tol is None else tol, pinv=pinv)
    else:
        raise ValueError(basis)
    return ns.tol_rref(V, tol=tol, trace=trace).T


def atacom_step(spec, ev, dq, s, alpha, basis="svd", variant="atacom"):
    """One call of step_action_function (atacom.py:123-139; variant 'ec':
    error_correction_wrapper.py:117-134).  `alpha` is the already scaled
    tangent action.  Returns a dict; `s_new` is the integrated slack."""
    A_f, A_g, ct_f, ct_g, psi = viability_terms(spec, ev, dq)
    Jc = stack_Jc(spec, A_f, A_g, s)
    pinv, Q, rank = ns.svd_pinv_null(Jc)
    c = np.concatenate([ct_f, ct_g + 0.5 * s ** 2])          # atacom.py:183-196
    act_err = -pinv @ (spec.K_c * c)                          # atacom.py:181
    trace = {}
    if variant == "atacom":
        Nc = null_coordinates(Jc, spec.k, spec.tol, basis, (pinv, Q), trace)
        act_a = -pinv @ psi                                   # atacom.py:130
        act_b = Nc @ alpha                                    # atacom.py:131
    elif variant == "ec":
        Nc = None
        act_a = np.zeros(spec.N)
        act_b = np.concatenate([alpha, np.zeros(spec.G)])
    else:
        raise ValueError(variant)
    w = act_a + act_b + act_err
    s_new = s + w[spec.n:] * spec.dt                          # atacom.py:135
    ddq = clip_acc(spec, dq, w[:spec.n])                      # atacom.py:137
    return dict(ddq=ddq, s_new=s_new, w=w, act_a=act_a, act_b=act_b, act_err=act_err,
                Jc=Jc, Nc=Nc, rank=rank, trace=trace, c=c, psi=psi)


# --------------------------------------------------------------------------- variant C

POINT_REACH_RADIUS2 = 0.6 ** 2      # collision_avoidance_atacom.py:75


def point_reach_terms(q, dq, p, dp, K):
    """collision_avoidance_atacom.py:72-127 for obstacles p (G x 2), dp (G x 2).

    c_o = 0.36 - |q-p_i|^2, J_q = -2 (q-p_i), J_p = +2 (q-p_i),
    dc = J_p dp_i + J_q dq, and the as-written (positions, not velocities)
    bp + bq = (p.Hpp + q.Hqp).p + (q.Hqq + p.Hqp).q with Hqq=Hpp=-2I, Hqp=2I.
    """
    d = q[None, :] - p
    c_o = POINT_REACH_RADIUS2 - (d ** 2).sum(1)
    J_q = -2.0 * d
    J_p = 2.0 * d
    dc = (J_p * dp).sum(1) + J_q @ dq
    bp = ((-2.0 * p + 2.0 * q[None, :]) * p).sum(1)
    bq = ((-2.0 * q[None, :] + 2.0 * p) * q[None, :]).sum(1)
    psi = dc + K * (bp + bq)
    return c_o, J_q, dc, psi


def point_reach_slack_init(q, p):
    """collision_avoidance_atacom.py:25."""
    c_o = POINT_REACH_RADIUS2 - ((q[None, :] - p) ** 2).sum(1)
    return np.sqrt(np.maximum(-2.0 * c_o, 0.0))


def point_reach_step(q, dq, p, dp, s, action, dt=0.01, K=0.5, Kc=100.0, basis="svd", tol=None):
    """PointReachAtacom.step up to the call into the base env
    (collision_avoidance_atacom.py:29-47).  Returns w (2+G), s_new and the
    control the base env applies, clip(w[:2], -1, 1) * 10
    (collision_avoidance_base.py:44-45)."""
    G = p.shape[0]
    K = np.ones(G) * K
    c_o, J_q, dc, psi = point_reach_terms(q, dq, p, dp, K)
    Jc = np.zeros((G, 2 + G))
    Jc[:, :2] = J_q
    Jc[:, 2:] = np.diag(s)
    pinv, Q, rank = ns.svd_pinv_null(Jc)
    c = c_o + 0.5 * s ** 2 + K * dc
    trace = {}
    if basis == "svd":
        if Q.shape[1] != 2:
            raise ValueError("rank-deficient Jc: the reference fails here with a shape mismatch")
        Nc = ns.tol_rref(Q.T, tol=tol, trace=trace).T
    else:
        Nc = null_coordinates(Jc, 2, tol, basis, (pinv, Q), trace)
    w = -pinv @ (psi + Kc * c) + Nc @ action
    s_new = s + w[2:] * dt
    u = np.clip(w[:2], -1.0, 1.0) * 10.0
    return dict(w=w, s_new=s_new, u=u, Jc=Jc, Nc=Nc, trace=trace, c_o=c_o, rank=rank)
