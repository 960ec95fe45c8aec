"""Collision-avoidance point environment for B parallel agents (reference:
atacom/environments/collision_avoidance/collision_avoidance_base.py and collision_avoidance_atacom.py).

State [B, 4 (1 + n_objects)]: agent (x, y, dx, dy) then per obstacle (x, y, dx, dy).  PointReachAtacom
does not use AtacomEnvWrapper in the reference either: it inlines the projection in step()
(collision_avoidance_atacom.py:29-48); here that step is one launch of the point-reach kernel."""
import math

import numpy as np
import torch

from .. import _lib, projection
from ..mdp import Box, MDPInfo


class PointGoalReach:
    """collision_avoidance_base.py:6-76, batched."""

    def __init__(self, time_step=0.01, horizon=500, gamma=0.99, n_objects=2, random_walk=False, n_envs=1,
                 device=None):
        self.time_step = time_step
        self.n_objects = n_objects
        self.random_walk = random_walk
        self.n_envs = n_envs
        self.device = torch.device(device if device is not None else "cuda")
        self.state_dim = 4 * (1 + n_objects)
        self._mdp_info = MDPInfo(Box(low=-np.ones(self.state_dim) * 10, high=np.ones(self.state_dim) * 10),
                                 Box(low=-np.ones(2), high=np.ones(2)), gamma, horizon)
        self.step_action_function = None
        self.action_scale = torch.tensor([10., 10.], device=self.device)
        self._obj_radius = 2
        self._time = 0.
        self._gen = torch.Generator(device="cpu")
        self._state = None

    @property
    def info(self):
        return self._mdp_info

    def seed(self, seed):
        self._gen.manual_seed(int(seed))

    def reset(self, state=None, mask=None):
        """mask (bool [B], optional): re-initialise only those environments (per-env episode termination); the
        obstacles' clock `_time` is shared by the batch and restarts only on a full reset."""
        if mask is not None and self._state is not None:
            old, old_c, t = self._state, self._obj_circle_center, self._time
            new = PointGoalReach.reset(self, state)
            m = torch.as_tensor(mask, device=self.device).bool()
            old[m] = new[m]
            old_c[m] = self._obj_circle_center[m]
            self._state, self._obj_circle_center, self._time = old, old_c, t
            return self._state
        self._time = 0.
        B, G = self.n_envs, self.n_objects
        st = torch.zeros(B, self.state_dim)
        st[:, 0:2] = 1.0
        obj = st[:, 4:].view(B, G, 4)
        obj[:, :, :2] = torch.rand(B, G, 2, generator=self._gen) * 6 + 2      # uniform(2, 8), base.py:36
        self._state = st.to(self.device)
        centre = self._state[:, 4:].view(B, G, 4)[:, :, :2].clone()
        centre[:, :, 0] -= self._obj_radius
        self._obj_circle_center = centre
        if state is not None:
            self._state = torch.as_tensor(state, dtype=torch.float32, device=self.device).reshape(B, -1).clone()
        return self._state

    def step(self, action):
        if self.step_action_function is not None:
            action = self.step_action_function(self._state, action)
        self._action = torch.clamp(action, -1.0, 1.0) * self.action_scale
        dt = self.time_step
        st = self._state
        st[:, :2] += st[:, 2:4] * dt                                        # base.py:47-48
        st[:, 2:4] += self._action * dt
        flip = (st[:, 0:2] <= 0) | (st[:, 0:2] >= 10)
        st[:, 2:4] = torch.where(flip, -st[:, 2:4], st[:, 2:4])
        obj = st[:, 4:].view(st.shape[0], self.n_objects, 4)
        if self.random_walk:                                                # base.py:57-66
            obj[:, :, :2] += obj[:, :, 2:] * dt
            obj[:, :, :2] = torch.clamp(obj[:, :, :2], 2, 10)
            act = (torch.rand(obj.shape[0], self.n_objects, 2, generator=self._gen) * 2 - 1).to(self.device) * 10
            flip = (obj[:, :, :2] <= 2) | (obj[:, :, :2] >= 10)
            obj[:, :, 2:] = torch.where(flip, -obj[:, :, 2:], obj[:, :, 2:])
            obj[:, :, 2:] += act * dt
            obj[:, :, 2:] = torch.clamp(obj[:, :, 2:], -1, 1)
        else:                                                               # base.py:67-72
            ang = self._time * 2 * math.pi
            obj[:, :, 0] = self._obj_circle_center[:, :, 0] + self._obj_radius * math.cos(ang)
            obj[:, :, 1] = self._obj_circle_center[:, :, 1] + self._obj_radius * math.sin(ang)
            obj[:, :, 2] = -2 * self._obj_radius * math.pi * math.sin(ang)
            obj[:, :, 3] = 2 * self._obj_radius * math.pi * math.cos(ang)
        self._time += dt
        goal = torch.tensor([9., 9.], device=self.device)
        reward = -(goal - st[:, :2]).norm(dim=1) / (8 * math.sqrt(2))
        absorbing = torch.zeros(st.shape[0], dtype=torch.bool, device=self.device)
        return st, reward, absorbing, dict()

    def render(self):
        pass

    def stop(self):
        pass

    def _create_sim_state(self):
        return self._state

    def _create_observation(self, state):
        return state


class PointReachAtacom(PointGoalReach):
    """collision_avoidance_atacom.py:8-48."""

    def __init__(self, time_step=0.01, horizon=1000, gamma=0.99, n_objects=4, random_walk=False, n_envs=1,
                 device=None):
        super().__init__(time_step=time_step, horizon=horizon, gamma=gamma, n_objects=n_objects,
                         random_walk=random_walk, n_envs=n_envs, device=device)
        self.params = _lib.default_params("point_reach")
        self.params.dt = float(time_step)
        self.s = torch.zeros(n_envs, n_objects, device=self.device)
        self._w = torch.zeros(n_envs, 2, device=self.device)
        self.status = torch.zeros(n_envs, dtype=torch.uint8, device=self.device)
        self.constr_logs = list()
        self._stats = None

    def _split(self, state):
        B = state.shape[0]
        obj = state[:, 4:].view(B, self.n_objects, 4)
        return (state[:, :2].contiguous(), state[:, 2:4].contiguous(),
                obj[:, :, :2].reshape(B, -1).contiguous(), obj[:, :, 2:].reshape(B, -1).contiguous())

    def reset(self, state=None, mask=None):
        super().reset(state, mask)
        self.q, self.dq, self.p, self.dp = self._split(self._state)
        m = None if mask is None else torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        projection.point_reach_slack_init(self.q, self.p, self.params, s=self.s, mask=m)      # :25
        return self._state

    def step(self, action):
        numpy_io = isinstance(action, np.ndarray)
        action = torch.as_tensor(action, dtype=torch.float32, device=self.device)
        action = (action[None, :] if action.dim() == 1 else action).contiguous()
        self.q, self.dq, self.p, self.dp = self._split(self._state)
        d = self.q[:, None, :] - self.p.view(-1, self.n_objects, 2)
        c_origin = self.params.env[0] - (d ** 2).sum(-1)                               # get_c, :72-76
        self.constr_logs.append(torch.stack([c_origin.max(1).values, torch.zeros_like(c_origin[:, 0])], 1))
        projection.point_reach_step(self.q, self.dq, self.p, self.dp, self.s, action, self.params, w=self._w,
                                    s_out=self.s, status=self.status)
        out = super().step(self._w)
        if numpy_io and self.n_envs == 1:
            return tuple(o[0].cpu().numpy() if isinstance(o, torch.Tensor) else o for o in out)
        return out

    def rollout(self, actions, obstacle_draws=None):
        """T steps in ONE kernel launch (atacom_point_reach_rollout): `actions` [T, B, 2]; with random_walk the
        obstacles' U(-1, 1) draws [T, B, 2 n_objects] are taken from `obstacle_draws` or drawn here from the env's
        generator.  Equivalent to T calls of step(); returns rewards [T, B]."""
        actions = torch.as_tensor(actions, dtype=torch.float32, device=self.device).contiguous()
        T, B, G = actions.shape[0], self.n_envs, self.n_objects
        centers = None
        if self.random_walk:
            if obstacle_draws is None:
                obstacle_draws = torch.rand(T, B, 2 * G, generator=self._gen) * 2 - 1
            obstacle_draws = torch.as_tensor(obstacle_draws, dtype=torch.float32, device=self.device).contiguous()
        else:
            obstacle_draws = None
            centers = self._obj_circle_center.reshape(B, 2 * G).contiguous()
        if self._stats is None:
            self._stats = projection.new_stats(self.device)
        rewards = projection.point_reach_rollout(self._state, self.s, actions, self.params,
                                                 obstacle_draws=obstacle_draws, obstacle_centers=centers,
                                                 time0=self._time, stats=self._stats, status=self.status)
        self._time += T * self.time_step
        self.q, self.dq, self.p, self.dp = self._split(self._state)
        return rewards

    def get_constraints_logs(self):
        total, n, c_max, c_dq_max = 0.0, 0, -math.inf, 0.0
        if self.constr_logs:
            logs = torch.stack(self.constr_logs, 0)
            total, n, c_max = float(logs[..., 0].sum()), logs[..., 0].numel(), float(logs[..., 0].max())
            self.constr_logs.clear()
        if self._stats is not None:
            v = self._stats.cpu().tolist()
            total, n, c_max = total + v[0], n + int(v[3]), max(c_max, v[1])
            self._stats = None
        if self.shard is not None:                                     # one all-reduce per epoch
            from ..sharding import reduce_constraint_logs
            total, n, c_max, c_dq_max = reduce_constraint_logs(total, n, c_max, c_dq_max, self.shard.group)
        return total / max(n, 1), c_max, c_dq_max

    shard = None            # sharding.EnvShard when this env holds one rank's slice of a global batch
