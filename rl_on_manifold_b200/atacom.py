"""AtacomEnvWrapper for a batch of B environments — same constructor, methods and hook protocol as
the reference's atacom/atacom.py, with the per-step NumPy algebra replaced by one CUDA kernel launch.

Reference surface kept (atacom/atacom.py):
    AtacomEnvWrapper(base_env, dim_q, vel_max, acc_max, f, g, Kc, Kq, time_step)      :10
    _get_q / _get_dq / acc_to_ctrl_action (abstract)                                   :81-88
    seed / reset / render / stop / step / info / set_logger / get_constraints_logs     :90-115,141-143,198-216
    step_action_function(sim_state, alpha), installed on base_env as a hook            :79,123-139
    acc_truncation, _compute_slack_variables, dims, K_c, K_q, alpha_max, s, q, dq      :25-71,117-121,145-149

State lives in float32 CUDA tensors shaped [B, dim].  A base environment is any object with
`info`, `reset(state)`, `step(action)`, `_create_observation(sim_state)`, `seed`, `render`, `stop`
and a `step_action_function` attribute it calls back — the protocol of the reference's base envs
(circle_base.py:53-57, env_base.py:161-165).  NumPy inputs of shape [dim] are accepted for B = 1
and answered in NumPy, so a single-environment MushroomRL `Core` loop drives it unchanged.
"""
import numpy as np
import torch

from . import _lib, projection
from .constraints import ConstraintsSet
from .mdp import Box


def _vec(v, m):
    return np.ones(m) * v if np.isscalar(v) else np.asarray(v, dtype=np.float64)


class AtacomEnvWrapper:
    """Environment wrapper of ATACOM (batched)."""

    variant = _lib.VARIANT_ATACOM

    def __init__(self, base_env, dim_q, vel_max, acc_max, f=None, g=None, Kc=1., Kq=10., time_step=0.01,
                 family=None, n_ctrl_joints=None, params=None):
        self.env = base_env
        self.dims = {'q': dim_q, 'f': 0, 'g': 0}
        # the reference accepts a bare ViabilityConstraint as well as a ConstraintsSet (atacom.py:18-19)
        self.f = ConstraintsSet.wrap(f, dim_q)
        self.g = ConstraintsSet.wrap(g, dim_q)
        f, g = self.f, self.g
        self.time_step = time_step
        self._logger = None

        if self.f is not None:
            assert self.dims['q'] == self.f.dim_q, "Input dimension is different in f"
            self.dims['f'] = self.f.dim_out
        if self.g is not None:
            assert self.dims['q'] == self.g.dim_q, "Input dimension is different in g"
            self.dims['g'] = self.g.dim_out
        self.dims['null'] = self.dims['q'] - self.dims['f']
        self.dims['c'] = self.dims['f'] + self.dims['g']

        self.K_c = _vec(Kc, self.dims['c'])
        self.vel_max = _vec(vel_max, dim_q)
        assert np.shape(self.vel_max)[0] == dim_q
        self.acc_max = _vec(acc_max, dim_q)
        assert np.shape(self.acc_max)[0] == dim_q
        self.K_q = _vec(Kq, dim_q)
        assert np.shape(self.K_q)[0] == dim_q
        self.alpha_max = np.ones(self.dims['null']) * self.acc_max.max()

        self._mdp_info = self.env.info.copy()
        self._mdp_info.action_space = Box(low=-np.ones(self._action_dim()), high=np.ones(self._action_dim()))

        # which kernel: a built-in device functor, or the generic one fed by batched callbacks
        self._family = family or getattr(f, "family", None) or getattr(g, "family", None)
        self._n_ctrl_joints = n_ctrl_joints or dim_q
        n, F, G = dim_q, self.dims['f'], self.dims['g']
        if self._family is None:
            if not projection.generic_supported(n, F, G):
                raise ValueError("no kernel compiled for a generic ConstraintsSet of shape (n,F,G)=(%d,%d,%d); "
                                 "see atacom_generic_supported()" % (n, F, G))
            for cs in (f, g):
                if cs is not None and not cs.has_callbacks:
                    raise ValueError("generic ConstraintsSet needs batched fun / J / b callbacks")
        self.params = params.copy() if params is not None else self._build_params()

        self.device = getattr(base_env, "device", torch.device("cuda", torch.cuda.current_device())
                              if torch.cuda.is_available() else None)
        if self.device is None or torch.device(self.device).type != "cuda":
            raise _lib.AtacomError("AtacomEnvWrapper needs a CUDA device: there is no CPU fallback")

        self.state = self._as_batch(self.env.reset())
        self.B = self.state.shape[0]
        self.q = torch.zeros(self.B, n, device=self.device)
        self.dq = torch.zeros(self.B, n, device=self.device)
        self.s = torch.zeros(self.B, G, device=self.device) if G > 0 else None
        self._ddq = torch.zeros(self.B, n, device=self.device)
        self._status = torch.zeros(self.B, dtype=torch.uint8, device=self.device)
        self._w_dbg = None
        self.debug = False
        self._act_a = None          # with debug=True: -Jc^+ psi - Jc^+ K_c c  (the reference's _act_a + _act_err)
        self._act_b = None          # with debug=True: Nc alpha
        self._act_err = None

        self.constr_logs = list()
        self.env.step_action_function = self.step_action_function

    # ------------------------------------------------------------------ parameters
    def _action_dim(self):
        return self.dims['null']

    def _build_params(self):
        if self._family is not None:
            p = (_lib.default_params(self._family, self._n_ctrl_joints) if self._family == "iiwa"
                 else _lib.default_params(self._family))
        else:
            p = _lib.AtacomParams()
            p.rref_tol, p.clip_acc = 0.05, 1
        F, G, n = self.dims['f'], self.dims['g'], self.dims['q']
        if self.f is not None:
            p.K_f[:F] = [float(x) for x in self.f.K]
        if self.g is not None:
            p.K_g[:G] = [float(x) for x in self.g.K]
        p.K_c[:F + G] = [float(x) for x in self.K_c]
        p.K_q[:n] = [float(x) for x in self.K_q]
        p.vel_max[:n] = [float(x) for x in self.vel_max]
        p.acc_max[:n] = [float(x) for x in self.acc_max]
        p.dt = float(self.time_step)
        p.variant = self.variant
        return p

    # ------------------------------------------------------------------ abstract hooks (atacom.py:81-88)
    def _get_q(self, state):
        raise NotImplementedError

    def _get_dq(self, state):
        raise NotImplementedError

    def acc_to_ctrl_action(self, ddq):
        raise NotImplementedError

    # ------------------------------------------------------------------ MDP facade
    def seed(self, seed):
        self.env.seed(seed)

    def reset(self, state=None, mask=None):
        """atacom.py:93-98.  `mask` (bool [B], batched extension): only those environments start a new episode —
        the base env re-initialises their rows and `_compute_slack_variables` their slacks (the kernels' mask)."""
        if mask is None:
            self.state = self._as_batch(self.env.reset(state))
        else:
            mask = torch.as_tensor(mask, device=self.device).bool()
            self.state = self._as_batch(self.env.reset(state, mask=mask))
        self.q = self._get_q(self.state).contiguous()
        self.dq = self._get_dq(self.state).contiguous()
        self._compute_slack_variables(None if mask is None else mask.to(torch.uint8).contiguous())
        return self._ret(self.state)

    def render(self):
        self.env.render()

    def stop(self):
        self.env.stop()

    def step(self, action):
        self._numpy_io = isinstance(action, np.ndarray)
        action = self._as_batch(action)
        low = torch.as_tensor(self.info.action_space.low, dtype=torch.float32, device=self.device)
        high = torch.as_tensor(self.info.action_space.high, dtype=torch.float32, device=self.device)
        alpha = torch.minimum(torch.maximum(action, low), high)
        alpha = alpha * torch.as_tensor(self._action_scale(), dtype=torch.float32, device=self.device)

        self.state, reward, absorb, info = self.env.step(alpha)
        self.state = self._as_batch(self.state)
        self.q = self._get_q(self.state).contiguous()
        self.dq = self._get_dq(self.state).contiguous()
        if not hasattr(self.env, "get_constraints_logs"):
            self._update_constraint_stats(self.q, self.dq)
        return self._ret(self.state.clone()), self._ret(reward), self._ret(absorb), info

    def _action_scale(self):
        return self.alpha_max

    def acc_truncation(self, dq, ddq):
        """atacom.py:117-121 on tensors (the kernels apply the same rule in their epilogue)."""
        t = lambda a: torch.as_tensor(a, dtype=dq.dtype, device=dq.device)
        acc_max, K_q, vel_max = t(self.acc_max), t(self.K_q), t(self.vel_max)
        acc_u = torch.maximum(torch.minimum(acc_max, -K_q * (dq - vel_max)), -acc_max)
        acc_l = torch.minimum(torch.maximum(-acc_max, -K_q * (dq + vel_max)), acc_max)
        return torch.minimum(torch.maximum(ddq, acc_l), acc_u)

    # ------------------------------------------------------------------ the hot path (atacom.py:123-139)
    def step_action_function(self, sim_state, alpha):
        self.state = self._as_batch(self.env._create_observation(sim_state))
        self._refresh_before_projection()
        alpha = self._as_batch(alpha).contiguous()
        n, G = self.dims['q'], self.dims['g']
        if self.debug and self._w_dbg is None:
            self._w_dbg = torch.zeros(self.B, 2 * (n + G), device=self.device)
        dbg = self._w_dbg if self.debug else None
        q, dq = self.q.contiguous(), self.dq.contiguous()
        if self._family is not None:
            projection.step(self._family, q, dq, self.s, alpha, self.params, n_ctrl_joints=self._n_ctrl_joints,
                            ddq=self._ddq, s_out=self.s, status=self._status, w_dbg=dbg)
        else:
            c, J, b = self._raw_constraints(q, dq)
            projection.generic_step(n, self.dims['f'], G, c, J, b, dq, self.s, alpha, self.params,
                                    ddq=self._ddq, s_out=self.s, status=self._status, w_dbg=dbg)
        if self.debug:
            N = n + G
            self._act_a = self._w_dbg[:, :N]
            self._act_b = self._w_dbg[:, N:]
            self._act_err = torch.zeros_like(self._act_a)
        return self.acc_to_ctrl_action(self._ddq)

    def _refresh_before_projection(self):
        """The reference does NOT refresh q, dq here (atacom.py:123-126): they are the values set by the last
        step()/reset().  For envs whose q is a live view of the simulator state (circle) subclasses override."""

    def _raw_constraints(self, q, dq):
        parts = [cs.raw(q, dq) for cs in (self.f, self.g) if cs is not None]
        c = torch.cat([p[0] for p in parts], 1).contiguous()
        J = torch.cat([p[1] for p in parts], 1).contiguous()
        b = torch.cat([p[2] for p in parts], 1).contiguous()
        return c, J, b

    @property
    def info(self):
        return self._mdp_info

    @property
    def status(self):
        """uint8 [B] ATACOM_ST_* bits of the last projection."""
        return self._status

    # ------------------------------------------------------------------ slack variables (atacom.py:145-149)
    def _compute_slack_variables(self, mask=None):
        if self.dims['g'] == 0:
            self.s = None
            return
        q, dq = self.q.contiguous(), self.dq.contiguous()
        if self.s is None or self.s.shape[0] != q.shape[0]:
            self.s = torch.zeros(q.shape[0], self.dims['g'], device=self.device)
        if self._family is not None:
            projection.slack_init(self._family, q, dq, self.params, n_ctrl_joints=self._n_ctrl_joints, s=self.s,
                                  mask=mask)
        else:
            s_new = torch.sqrt(torch.clamp(-2 * self.g.fun(q, dq, origin_constr=False), min=0))
            if mask is None:
                self.s.copy_(s_new)
            else:
                self.s[mask.bool()] = s_new[mask.bool()]

    # ------------------------------------------------------------------ constraint statistics (atacom.py:198-216)
    def set_logger(self, logger):
        self._logger = logger

    def _origin_constraints(self, q, dq):
        """c(q) with origin_constr=True: [B, F+G] (atacom.py:183-196)."""
        parts = []
        if self.f is not None:
            parts.append(self.f.fun(q, dq, origin_constr=True))
        if self.g is not None:
            parts.append(self.g.fun(q, dq, origin_constr=True))
        return torch.cat(parts, 1)

    def _update_constraint_stats(self, q, dq):
        if self._family is not None:      # built-in functor: one kernel, per-env (c_i, c_dq_i)
            self.constr_logs.append(projection.constraint_stats(self._family, q.contiguous(), dq.contiguous(),
                                                                self.params, n_ctrl_joints=self._n_ctrl_joints))
            return
        c_i = self._origin_constraints(q, dq)
        F = self.dims['f']
        c_i = torch.cat([c_i[:, :F].abs(), c_i[:, F:]], 1)
        vel_max = torch.as_tensor(self.vel_max, dtype=dq.dtype, device=dq.device)
        c_dq_i = dq.abs() - vel_max
        self.constr_logs.append(torch.stack([c_i.max(1).values, c_dq_i.max(1).values], 1))

    def get_constraints_logs(self):
        """(c_avg, c_max, c_dq_max) of the epoch (atacom.py:207-216).  Under env-index sharding (`self.shard` set
        to a sharding.EnvShard) the three numbers cover the GLOBAL batch: one all-reduce per epoch — SUM on the
        sum and the count, MAX on the two maxima (sharding.reduce_constraint_logs)."""
        if not hasattr(self.env, "get_constraints_logs"):
            logs = torch.stack(self.constr_logs, 0)                       # [T, B, 2]
            total, count = float(logs[..., 0].double().sum()), logs[..., 0].numel()
            c_max = float(logs[..., 0].max())
            c_dq_max = float(logs[..., 1].max())
            self.constr_logs.clear()
            if self.shard is not None:
                from .sharding import reduce_constraint_logs
                total, count, c_max, c_dq_max = reduce_constraint_logs(total, count, c_max, c_dq_max, self.shard.group)
            return total / max(count, 1), c_max, c_dq_max
        return self.env.get_constraints_logs()

    shard = None            # sharding.EnvShard when this wrapper holds one rank's slice of a global batch

    # ------------------------------------------------------------------ helpers
    _numpy_io = False

    def _as_batch(self, x):
        if isinstance(x, np.ndarray):
            x = torch.as_tensor(x, dtype=torch.float32, device=self.device if hasattr(self, "device") else None)
        if isinstance(x, torch.Tensor):
            if x.dim() == 1:
                x = x[None, :]
            if x.dtype != torch.float32:
                x = x.float()
            if hasattr(self, "device") and x.device != torch.device(self.device):
                x = x.to(self.device)
        return x

    def _ret(self, x):
        """Answer in NumPy [dim] when driven with NumPy actions at B = 1 (MushroomRL Core compatibility)."""
        if self._numpy_io and isinstance(x, torch.Tensor) and self.B == 1:
            x = x.detach().cpu().numpy()
            return x[0] if x.ndim > 0 and x.shape[0] == 1 else x
        return x
