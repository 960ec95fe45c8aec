"""Batched roll-out driver: what `mushroom_rl.core.Core` does for the reference's experiment scripts
(examples/circle_exp.py:15-82), for B environments stepping in lock-step on the GPU.

MushroomRL's Core drives ONE environment through `mdp.reset / mdp.step` and hands `(state, action, reward,
next_state, absorbing, last)` tuples to the agent.  `BatchedCore` does the same with [B, ...] tensors: the agent
is any object with `draw_action(state [B, ds]) -> action [B, da]` (CUDA tensors) and, for `learn`, `fit(dataset)`.
The dataset is a dict of tensors with a leading time axis — `parse_dataset` of the reference
(`mushroom_rl.utils.dataset`) without the Python list in between.  `compute_J` is the reference's discounted
return (examples/circle_exp.py:74-75) per episode and environment.
"""
import torch


class BatchedCore:
    def __init__(self, agent, mdp):
        self.agent, self.mdp = agent, mdp
        self._state = None
        self._t = 0                                   # steps into the current episode (all envs in lock-step)

    def reset(self):
        self._state = self.mdp.reset()
        if not isinstance(self._state, torch.Tensor):
            self._state = torch.as_tensor(self._state)
        self._state = self._state.clone()
        self._t = 0

    def _collect(self, n_steps):
        """n_steps environment steps in every environment; episodes end at the horizon (mdp.info.horizon) or,
        per environment, at an absorbing state (the whole batch is reset when the horizon is reached; an
        absorbing environment keeps stepping — ATACOM environments are never absorbing, atacom.py:106-115)."""
        if self._state is None:
            self.reset()
        horizon = self.mdp.info.horizon
        keys = ("state", "action", "reward", "next_state", "absorbing", "last")
        data = {k: [] for k in keys}
        for _ in range(n_steps):
            action = self.agent.draw_action(self._state)
            next_state, reward, absorbing, _ = self.mdp.step(action)
            self._t += 1
            last = torch.full_like(absorbing, self._t >= horizon) | absorbing
            for k, v in zip(keys, (self._state, action, reward, next_state, absorbing, last)):
                data[k].append(v.clone() if isinstance(v, torch.Tensor) else torch.as_tensor(v))
            if self._t >= horizon:
                self.reset()
            else:
                self._state = next_state.clone()
        return {k: torch.stack(v, 0) for k, v in data.items()}

    def evaluate(self, n_steps=None, n_episodes=None):
        """Core.evaluate: fresh episodes, no fitting.  n_episodes counts episodes PER environment."""
        self.reset()
        if n_steps is None:
            n_steps = n_episodes * self.mdp.info.horizon
        return self._collect(n_steps)

    def learn(self, n_steps, n_steps_per_fit):
        """Core.learn: collect n_steps per environment, calling agent.fit every n_steps_per_fit steps."""
        done = 0
        while done < n_steps:
            chunk = min(n_steps_per_fit, n_steps - done)
            self.agent.fit(self._collect(chunk))
            done += chunk


def compute_J(dataset, gamma=1.0):
    """Discounted return of every episode of every environment: a [n_episodes, B] tensor
    (mushroom_rl.utils.dataset.compute_J, batched; gamma = 1 gives the undiscounted return R)."""
    reward, last = dataset["reward"], dataset["last"]
    T = reward.shape[0]
    out, acc, disc = [], torch.zeros_like(reward[0], dtype=torch.float64), 1.0
    for t in range(T):
        acc = acc + disc * reward[t].double()
        disc *= gamma
        if bool(last[t].all()) or t == T - 1:
            if bool(last[t].all()):
                out.append(acc)
                acc, disc = torch.zeros_like(acc), 1.0
    if not out:
        out.append(acc)
    return torch.stack(out, 0)


def compute_metrics(core, n_episodes, gamma=None):
    """(J, R, c_avg, c_max, c_dq_max) as examples/circle_exp.py:71-82 computes them (means over episodes and
    environments; the agent's entropy E is the agent's business)."""
    dataset = core.evaluate(n_episodes=n_episodes)
    gamma = core.mdp.info.gamma if gamma is None else gamma
    J = float(compute_J(dataset, gamma).mean())
    R = float(compute_J(dataset, 1.0).mean())
    c_avg, c_max, c_dq_max = core.mdp.get_constraints_logs()
    return J, R, c_avg, c_max, c_dq_max


class UniformAgent:
    """Random agent over the MDP's action box (the reference's smoke tests use one, circle_terminated.py:43-68)."""

    def __init__(self, mdp, seed=0, device=None):
        self.low = torch.as_tensor(mdp.info.action_space.low, dtype=torch.float32)
        self.high = torch.as_tensor(mdp.info.action_space.high, dtype=torch.float32)
        self._gen = torch.Generator().manual_seed(seed)
        self.device = device
        self.n_fits = 0

    def draw_action(self, state):
        B = state.shape[0]
        a = torch.rand(B, self.low.numel(), generator=self._gen) * (self.high - self.low) + self.low
        return a.to(state.device if self.device is None else self.device)

    def fit(self, dataset):
        self.n_fits += 1
