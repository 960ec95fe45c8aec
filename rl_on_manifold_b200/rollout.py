"""Batched roll-out driver: what `mushroom_rl.core.Core` does for the reference's experiment scripts
(examples/circle_exp.py:15-82, planar_air_hockey_exp.py:15-70), for B environments stepping in lock-step on the GPU.

MushroomRL's Core drives ONE environment through `mdp.reset / mdp.step` and hands `(state, action, reward,
next_state, absorbing, last)` tuples to the agent.  `BatchedCore` does the same with [B, ...] tensors: the agent
is any object with `draw_action(state [B, ds]) -> action [B, da]` (CUDA tensors) and, for `learn`, `fit(dataset)`.
The dataset is a dict of tensors with a leading time axis — `parse_dataset` of the reference
(`mushroom_rl.utils.dataset`) without the Python list in between.  Episodes end PER ENVIRONMENT, at the horizon or
at an absorbing state: the environments that ended are reset through `mdp.reset(mask=...)` (masked slack
re-initialisation in the kernels) while the others keep stepping.  `compute_J` is the reference's discounted
return (examples/circle_exp.py:74-75) per episode and environment; `MinMaxPreprocessor` is the observation
normaliser of the air-hockey experiments (planar_air_hockey_exp.py:30-32).
"""
import torch


class BatchedCore:
    def __init__(self, agent, mdp, preprocessors=None):
        self.agent, self.mdp = agent, mdp
        self.preprocessors = list(preprocessors) if preprocessors is not None else []
        self._state = None
        self._t = None                                # steps into the current episode, per environment [B]

    def _preprocess(self, state):
        for p in self.preprocessors:                  # Core._preprocess of MushroomRL
            state = p(state)
        return state

    def reset(self):
        state = self.mdp.reset()
        if not isinstance(state, torch.Tensor):
            state = torch.as_tensor(state)
        self._state = self._preprocess(state.clone())
        self._t = torch.zeros(state.shape[0], dtype=torch.long, device=state.device)

    def _collect(self, n_steps):
        """n_steps environment steps in every environment.  Per environment an episode ends at the horizon
        (mdp.info.horizon) or at an absorbing state; those environments alone are reset (mdp.reset(mask=...))."""
        if self._state is None:
            self.reset()
        horizon = self.mdp.info.horizon
        keys = ("state", "action", "reward", "next_state", "absorbing", "last")
        data = {k: [] for k in keys}
        for _ in range(n_steps):
            action = self.agent.draw_action(self._state)
            next_state, reward, absorbing, _ = self.mdp.step(action)
            next_state = self._preprocess(torch.as_tensor(next_state).clone())
            absorbing = torch.as_tensor(absorbing, device=next_state.device).bool()
            self._t += 1
            last = (self._t >= horizon) | absorbing
            for k, v in zip(keys, (self._state, action, reward, next_state, absorbing, last)):
                data[k].append(v.clone() if isinstance(v, torch.Tensor) else torch.as_tensor(v))
            if bool(last.all()):
                self.reset()
            elif bool(last.any()):
                fresh = self._preprocess(torch.as_tensor(self.mdp.reset(mask=last)).clone())
                self._state = torch.where(last[:, None], fresh, next_state)
                self._t = torch.where(last, torch.zeros_like(self._t), self._t)
            else:
                self._state = next_state
        return {k: torch.stack(v, 0) for k, v in data.items()}

    def evaluate(self, n_steps=None, n_episodes=None):
        """Core.evaluate: fresh episodes, no fitting.  n_episodes counts horizon-long episodes PER environment
        (environments that absorb early start more episodes within the same number of steps)."""
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


def episode_returns(dataset, gamma=1.0):
    """Discounted return of EVERY episode of every environment, as mushroom_rl.utils.dataset.compute_J builds its
    list: an episode closes where `last` is set — per environment — and the unfinished trailing episode of each
    environment is appended too.  Returns (J [n], env [n]): the returns in order of closing and the environment
    each belongs to."""
    reward, last = dataset["reward"], dataset["last"].bool()
    T, B = reward.shape[0], reward.shape[1]
    acc = torch.zeros(B, dtype=torch.float64, device=reward.device)
    disc = torch.ones(B, dtype=torch.float64, device=reward.device)
    Js, envs = [], []
    idx = torch.arange(B, device=reward.device)
    for t in range(T):
        acc = acc + disc * reward[t].double()
        disc = disc * gamma
        close = last[t] if t < T - 1 else torch.ones_like(last[t])
        if bool(close.any()):
            Js.append(acc[close])
            envs.append(idx[close])
            acc = torch.where(close, torch.zeros_like(acc), acc)
            disc = torch.where(close, torch.ones_like(disc), disc)
    return torch.cat(Js), torch.cat(envs)


def compute_J(dataset, gamma=1.0):
    """mushroom_rl.utils.dataset.compute_J, batched (gamma = 1 gives the undiscounted return R).  When every
    environment closes its episodes at the same steps (no absorbing states: every ATACOM environment of the
    reference) the result is a [n_episodes, B] tensor; otherwise the flat list of `episode_returns`."""
    J, env = episode_returns(dataset, gamma)
    B = dataset["reward"].shape[1]
    if J.numel() % B == 0 and bool((env.view(-1, B) == torch.arange(B, device=env.device)).all()):
        return J.view(-1, B)
    return J


def compute_metrics(core, n_episodes, gamma=None):
    """(J, R, c_avg, c_max, c_dq_max) as examples/circle_exp.py:71-82 computes them (means over episodes and
    environments; the agent's entropy E is the agent's business)."""
    dataset = core.evaluate(n_episodes=n_episodes)
    gamma = core.mdp.info.gamma if gamma is None else gamma
    J = float(compute_J(dataset, gamma).mean())
    R = float(compute_J(dataset, 1.0).mean())
    c_avg, c_max, c_dq_max = core.mdp.get_constraints_logs()
    return J, R, c_avg, c_max, c_dq_max


class MinMaxPreprocessor:
    """Observation normaliser of the air-hockey experiments (planar_air_hockey_exp.py:30-32 builds
    `MinMaxPreprocessor(mdp_info=mdp.info)` from MushroomRL and hands it to Core): a bounded observation
    dimension is mapped to [-1, 1] through its bounds, (obs - (high + low) / 2) / ((high - low) / 2); an
    unbounded one falls back to running standardisation, (obs - mean) / std clipped to +-clip_obs, with mean / std
    updated from every batch seen.  Batched: one call normalises [B, dim] and folds all B rows into the running
    statistics (parallel Welford merge).  `get_state / set_state` is what the reference's `prepro.save / load`
    (planar_air_hockey_exp.py:45-46,63-64) persists."""

    def __init__(self, mdp_info, clip_obs=10.0, alpha=1e-32):
        low = torch.as_tensor(mdp_info.observation_space.low, dtype=torch.float64)
        high = torch.as_tensor(mdp_info.observation_space.high, dtype=torch.float64)
        self.bounded = ~(torch.isinf(low) | torch.isinf(high))
        self.obs_mean = torch.where(self.bounded, (high + low) / 2, torch.zeros_like(low))
        self.obs_delta = torch.where(self.bounded, (high - low) / 2, torch.ones_like(low))
        self.clip_obs, self.alpha = clip_obs, alpha
        self.count = 0.0
        self.mean = torch.zeros_like(low)
        self.m2 = torch.zeros_like(low)

    def _update(self, obs64):
        n = obs64.shape[0]
        bm = obs64.mean(0)
        bm2 = ((obs64 - bm) ** 2).sum(0)
        tot = self.count + n
        delta = bm - self.mean
        self.mean = self.mean + delta * (n / tot)
        self.m2 = self.m2 + bm2 + delta ** 2 * (self.count * n / tot)
        self.count = tot

    @property
    def std(self):
        return torch.sqrt(torch.clamp(self.m2 / max(self.count, 1.0), min=self.alpha))

    def __call__(self, obs):
        single = obs.dim() == 1
        x = (obs[None, :] if single else obs).double()
        for name in ("bounded", "obs_mean", "obs_delta", "mean", "m2"):
            setattr(self, name, getattr(self, name).to(x.device))
        out = (x - self.obs_mean) / self.obs_delta
        if not bool(self.bounded.all()):
            self._update(x)
            run = torch.clamp((x - self.mean) / self.std, -self.clip_obs, self.clip_obs)
            out = torch.where(self.bounded, out, run)
        out = out.to(obs.dtype)
        return out[0] if single else out

    def get_state(self):
        return dict(count=self.count, mean=self.mean.cpu().clone(), m2=self.m2.cpu().clone())

    def set_state(self, state):
        self.count, self.mean, self.m2 = state["count"], state["mean"].clone(), state["m2"].clone()


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
