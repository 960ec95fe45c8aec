"""Circular-motion environments for B parallel points (reference:
atacom/environments/circular_motion/circle_base.py, circle_atacom.py, circle_error_correction.py).

The base dynamics are a 2-D double integrator (circle_base.py:53-67); state [B, 4] = (x, y, dx, dy) in a
float32 CUDA tensor.  Viewer / plotting code of the reference is out of scope."""
import math

import numpy as np
import torch

from ..atacom import AtacomEnvWrapper
from ..constraints import ConstraintsSet, ViabilityConstraint
from ..error_correction_wrapper import ErrorCorrectionEnvWrapper
from ..mdp import Box, MDPInfo


class CircularMotion:
    """circle_base.py:11-115, batched."""

    def __init__(self, time_step=0.01, horizon=500, gamma=0.99, random_init=False, n_envs=1, device=None):
        self.time_step = time_step
        self.random_init = random_init
        self.n_envs = n_envs
        self.device = torch.device(device if device is not None else "cuda")
        inf_array = np.ones(4) * np.inf
        self._mdp_info = MDPInfo(Box(low=-inf_array, high=inf_array), Box(low=-np.ones(2), high=np.ones(2)),
                                 gamma, horizon)
        self.step_action_function = None
        self.action_scale = torch.tensor([10., 10.], device=self.device)
        self._gen = torch.Generator(device="cpu")
        self.constr_logs = list()
        self._stats = None           # device accumulators of the fused roll-out path (projection.new_stats)
        self._state = None

    @property
    def info(self):
        return self._mdp_info

    def seed(self, seed):
        self._gen.manual_seed(int(seed))

    def reset(self, state=None, mask=None):
        """mask (bool [B], optional): re-initialise only those environments (per-env episode termination)."""
        if mask is not None and self._state is not None:
            old = self._state
            new = CircularMotion.reset(self, state).clone()
            m = torch.as_tensor(mask, device=self.device).bool()
            old[m] = new[m]
            self._state = old
            return self._state
        B = self.n_envs
        if state is None:
            if self.random_init:                         # circle_base.py:39-45
                u = lambda lo, hi: torch.rand(B, generator=self._gen) * (hi - lo) + lo
                y = u(-0.5, 1.0)
                x = torch.sqrt(1 - y ** 2) * torch.sign(u(-1.0, 1.0))
                dx = u(-1.0, 1.0)
                dy = -x * dx / y
                v = torch.stack([dx, dy], 1)
                v = v / v.norm(dim=1, keepdim=True) * u(0.0, 1.0)[:, None]
                self._state = torch.cat([x[:, None], y[:, None], v], 1).float().to(self.device)
            else:
                self._state = torch.tensor([-1., 0., 0., 0.], device=self.device).repeat(B, 1)
        else:
            state = torch.as_tensor(state, dtype=torch.float32, device=self.device)
            state = state[None, :] if state.dim() == 1 else state
            ok = ((state[:, 0] ** 2 + state[:, 1] ** 2 - 1).abs() < 1e-6) & \
                 ((state[:, 0] * state[:, 2] - state[:, 1] * state[:, 3]).abs() < 1e-6)   # circle_base.py:49
            if not bool(ok.all()):
                raise ValueError("Can not reset to the state: ", state)
            self._state = state.clone()
        return self._state

    def step(self, action):
        self.check_constraint()
        if self.step_action_function is not None:
            action = self.step_action_function(self._state, action)
        self._action = torch.clamp(action, -1.0, 1.0) * self.action_scale
        dt = self.time_step
        self._state[:, :2] += self._state[:, 2:4] * dt + self._action * dt ** 2 / 2
        self._state[:, 2:4] += self._action * dt
        goal = torch.tensor([1., 0.], device=self.device)
        reward = torch.exp(-(goal - self._state[:, :2]).norm(dim=1))
        absorbing = torch.zeros(self._state.shape[0], dtype=torch.bool, device=self.device)
        return self._state, reward, absorbing, dict()

    def render(self):
        pass

    def stop(self):
        pass

    def _create_sim_state(self):
        return self._state

    def _create_observation(self, state):
        return state

    def check_constraint(self):
        q, dq = self._state[:, :2], self._state[:, 2:4]
        c = self.get_c(q, dq)
        c = torch.cat([c[:, :1].abs(), c[:, 1:]], 1)
        self.constr_logs.append(c)

    @staticmethod
    def get_c(q, dq):
        return torch.cat([(q[:, :1] ** 2 + q[:, 1:2] ** 2 - 1), (-q[:, 1:2] - 0.5), dq.abs() - 1], 1)

    def get_constraints_logs(self):
        """(c_avg, c_max, c_dq_max) since the last call (circle_base.py:109-115): per-step logs of step() merged
        with the device accumulators of the fused roll-out path."""
        total, n, c_max, c_dq_max = 0.0, 0, -math.inf, -math.inf
        if self.constr_logs:
            logs = torch.stack(self.constr_logs, 0)                    # [T, B, 4]
            per_step = logs[..., :2].max(-1).values
            total, n = float(per_step.sum()), per_step.numel()
            c_max, c_dq_max = float(logs[..., :2].max()), float(logs[..., 2:].max())
            self.constr_logs.clear()
        if self._stats is not None:
            v = self._stats.cpu().tolist()
            total, n, c_max, c_dq_max = total + v[0], n + int(v[3]), max(c_max, v[1]), max(c_dq_max, v[2])
            self._stats = None
        if self.shard is not None:                                     # one all-reduce per epoch
            from ..sharding import reduce_constraint_logs
            total, n, c_max, c_dq_max = reduce_constraint_logs(total, n, c_max, c_dq_max, self.shard.group)
        return total / max(n, 1), c_max, c_dq_max

    shard = None            # sharding.EnvShard when this env holds one rank's slice of a global batch


def _circle_sets():
    # circle_atacom.py:9-16; the callbacks stay available for constraint statistics / generic use
    circle = ViabilityConstraint(2, 1, fun=lambda q: q[:, :1] ** 2 + q[:, 1:2] ** 2 - 1,
                                 J=lambda q: 2 * q[:, None, :],
                                 b=lambda q, dq: 2 * (dq ** 2).sum(1, keepdim=True), K=0.1)
    height = ViabilityConstraint(2, 1, fun=lambda q: -q[:, 1:2] - 0.5,
                                 J=lambda q: torch.tensor([[[0., -1.]]], device=q.device).expand(q.shape[0], 1, 2),
                                 b=lambda q, dq: torch.zeros(q.shape[0], 1, device=q.device), K=2)
    f = ConstraintsSet(2, family="circle")
    f.add_constraint(circle)
    g = ConstraintsSet(2, family="circle")
    g.add_constraint(height)
    return f, g


class _CircleMixin:
    def rollout(self, actions):
        """T agent steps in ONE kernel launch (atacom_circle_rollout): `actions` [T, B, action_dim] raw agent
        actions (CUDA float32).  Equivalent to T calls of step(); returns rewards [T, B].  State, slack and the
        constraint log advance exactly as step() would advance them."""
        from .. import projection
        actions = torch.as_tensor(actions, dtype=torch.float32, device=self.device).contiguous()
        base = self.env
        if base._stats is None:
            base._stats = projection.new_stats(self.device)
        state = base._state.contiguous()
        rewards = projection.circle_rollout(state, self.s, actions, self.params, stats=base._stats,
                                            status=self._status)
        if state.data_ptr() != base._state.data_ptr():
            base._state.copy_(state)
        self.state = base._state
        self.q = self._get_q(self.state).contiguous()
        self.dq = self._get_dq(self.state).contiguous()
        return rewards

    def _get_q(self, state):
        return state[:, :2]

    def _get_dq(self, state):
        return state[:, 2:4]

    def acc_to_ctrl_action(self, ddq):
        return ddq / torch.as_tensor(self.acc_max, dtype=ddq.dtype, device=ddq.device)   # circle_atacom.py:26-27

    def _refresh_before_projection(self):
        # in the reference q = state[:2] is a VIEW of the base env's state array, so it is live at hook time
        self.q = self._get_q(self.state).contiguous()
        self.dq = self._get_dq(self.state).contiguous()


class CircleEnvAtacom(_CircleMixin, AtacomEnvWrapper):
    """circle_atacom.py:7-18."""

    def __init__(self, horizon=500, gamma=0.99, random_init=False, Kc=100, time_step=0.01, n_envs=1, device=None):
        base_env = CircularMotion(random_init=random_init, horizon=horizon, gamma=gamma, n_envs=n_envs, device=device)
        f, g = _circle_sets()
        super().__init__(base_env=base_env, dim_q=2, f=f, g=g, Kc=Kc, acc_max=10, vel_max=1, Kq=20,
                         time_step=time_step, family="circle")


class CircleEnvErrorCorrection(_CircleMixin, ErrorCorrectionEnvWrapper):
    """circle_error_correction.py:7-21."""

    def __init__(self, horizon=500, gamma=0.99, random_init=False, Kc=100., time_step=0.01, n_envs=1, device=None):
        base_env = CircularMotion(random_init=random_init, horizon=horizon, gamma=gamma, n_envs=n_envs, device=device)
        f, g = _circle_sets()
        super().__init__(base_env=base_env, dim_q=2, f=f, g=g, Kc=Kc, acc_max=10, vel_max=1, Kq=20,
                         time_step=time_step, family="circle")
