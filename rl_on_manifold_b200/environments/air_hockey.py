"""Air-hockey ATACOM wrappers (reference: planar_air_hockey/atacom_air_hockey.py,
iiwa_air_hockey/iiwa_hit_atacom.py) for B parallel robots.

The reference's base environments are PyBullet rigid-body simulations (MushroomRL `PyBullet`, un-vendored;
CPU, one robot) — out of scope here.  What is kept is everything on the projection path: constructor
gains, constraint sets (evaluated by the fused planar / iiwa kernels), observation layout
[puck(3), puck_vel(3), q(n), dq(n)] (env_single.py:119) and the hook protocol.  `JointSpaceEnv` is a
kinematic stand-in base env (joint-space double integrator, n_intermediate_steps hook calls per agent
step like MushroomRL's PyBullet.step) so the wrappers can be driven end to end; with a real simulator
pass your own base_env.  `acc_to_ctrl_action` returns the joint acceleration: the reference's inverse
dynamics call (iiwa_hit_atacom.py:58-63) belongs to the simulator."""
import numpy as np
import torch

from ..atacom import AtacomEnvWrapper
from ..constraints import ConstraintsSet, ViabilityConstraint
from ..mdp import Box, MDPInfo
from .. import _lib


class JointSpaceEnv:
    """Kinematic base env: obs [B, 6 + 2n] = [puck(3), puck_vel(3), q(n), dq(n)]; every agent step runs
    `n_intermediate_steps` sub-steps, each calling the hook and integrating ddq with `timestep`."""

    def __init__(self, n_joints, q_init, gamma=0.99, horizon=120, timestep=1 / 240., n_intermediate_steps=4,
                 n_envs=1, device=None):
        self.n, self.n_envs = n_joints, n_envs
        self.timestep, self.n_intermediate_steps = timestep, n_intermediate_steps
        self.device = torch.device(device if device is not None else "cuda")
        dim = 6 + 2 * n_joints
        self._mdp_info = MDPInfo(Box(-np.ones(dim) * np.inf, np.ones(dim) * np.inf),
                                 Box(-np.ones(n_joints), np.ones(n_joints)), gamma, horizon)
        self.step_action_function = None
        self._q_init = torch.as_tensor(q_init, dtype=torch.float32, device=self.device)
        self._state = None

    @property
    def info(self):
        return self._mdp_info

    def seed(self, seed):
        torch.manual_seed(int(seed))

    def reset(self, state=None, mask=None):
        """mask (bool [B], optional): re-initialise only those environments (per-env episode termination)."""
        if state is not None:
            new = torch.as_tensor(state, dtype=torch.float32, device=self.device).reshape(self.n_envs, -1).clone()
        else:
            new = torch.zeros(self.n_envs, 6 + 2 * self.n, device=self.device)
            new[:, 6:6 + self.n] = self._q_init
        if mask is None or self._state is None:
            self._state = new
        else:
            m = torch.as_tensor(mask, device=self.device).bool()
            self._state[m] = new[m]
        return self._state

    def step(self, action):
        n, dt = self.n, self.timestep
        for _ in range(self.n_intermediate_steps):
            ddq = self.step_action_function(self._state, action) if self.step_action_function is not None else action
            self._state[:, 6:6 + n] += self._state[:, 6 + n:] * dt + 0.5 * ddq * dt * dt
            self._state[:, 6 + n:] += ddq * dt
        reward = torch.zeros(self.n_envs, device=self.device)
        absorbing = torch.zeros(self.n_envs, dtype=torch.bool, device=self.device)
        return self._state, reward, absorbing, dict()

    def _create_observation(self, state):
        return state

    def render(self):
        pass

    def stop(self):
        pass


IIWA_HOME = (0.0, 0.2596, 0.0, -1.2247, 0.0, 1.4384, 0.0)


class AirHockeyIiwaAtacom(AtacomEnvWrapper):
    """iiwa_hit_atacom.py:10-63.  n_ctrl_joints = 6 reproduces isolated_joint_7=True (env_single.py:17-18)."""

    def __init__(self, task='H', gamma=0.99, horizon=120, timestep=1 / 240., n_intermediate_steps=4, Kc=240.,
                 n_ctrl_joints=6, base_env=None, n_envs=1, device=None, bias_mode=_lib.BIAS_OMEGA_X_V):
        if task != 'H':
            raise NotImplementedError                    # iiwa_hit_atacom.py:20-21
        n = n_ctrl_joints
        p = _lib.default_params("iiwa", n)
        if base_env is None:
            base_env = JointSpaceEnv(n, IIWA_HOME[:n], gamma, horizon, timestep, n_intermediate_steps, n_envs, device)
        f = ConstraintsSet(n, family="iiwa")
        f.add_constraint(ViabilityConstraint(n, 1, K=0.1))                               # :25-28
        g = ConstraintsSet(n, family="iiwa")
        g.add_constraint(ViabilityConstraint(n, 5, K=0.5))                               # :30-31
        g.add_constraint(ViabilityConstraint(n, n, K=1))                                 # :32-33
        acc_max = np.ones(n) * 10
        vel_max = np.array(list(p.vel_max[:n]), dtype=np.float64)
        self.n_ctrl_joints = n
        super().__init__(base_env, n, f=f, g=g, Kc=Kc, vel_max=vel_max, acc_max=acc_max, Kq=4 * acc_max / vel_max,
                         time_step=timestep, family="iiwa", n_ctrl_joints=n)
        self.params.bias_mode = bias_mode

    def _get_q(self, state):
        n = self.n_ctrl_joints
        return state[:, -2 * n:-n]                                                       # :52-53

    def _get_dq(self, state):
        return state[:, -self.n_ctrl_joints:]                                            # :55-56

    def acc_to_ctrl_action(self, ddq):
        return ddq



class AirHockeyPlanarAtacom(AtacomEnvWrapper):
    """atacom_air_hockey.py:11-76 (planar 3R arm; URDF constants are parameters)."""

    def __init__(self, task='H', gamma=0.99, horizon=120, timestep=1 / 240., n_intermediate_steps=4, Kc=240.,
                 base_env=None, n_envs=1, device=None, bias_mode=_lib.BIAS_OMEGA_X_V):
        p = _lib.default_params("planar")
        if base_env is None:
            base_env = JointSpaceEnv(3, (-1.0, 1.6, 0.5), gamma, horizon, timestep, n_intermediate_steps, n_envs, device)
        g = ConstraintsSet(3, family="planar")
        g.add_constraint(ViabilityConstraint(3, 3, K=0.5))                               # :30-31
        g.add_constraint(ViabilityConstraint(3, 3, K=1.0))                               # :32-33
        acc_max = np.ones(3) * 10
        vel_max = np.array(list(p.vel_max[:3]), dtype=np.float64)
        super().__init__(base_env, 3, f=None, g=g, Kc=Kc, vel_max=vel_max, acc_max=acc_max,
                         Kq=2 * acc_max / vel_max, time_step=timestep, family="planar")
        self.params.bias_mode = bias_mode

    def _get_q(self, state):
        return state[:, 6:9]                                                             # :66-67

    def _get_dq(self, state):
        return state[:, 9:12]                                                            # :69-70

    def acc_to_ctrl_action(self, ddq):
        return ddq
