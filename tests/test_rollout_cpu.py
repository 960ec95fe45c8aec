"""Host logic of the batched roll-out driver, on CPU tensors: per-environment episode
termination, compute_J the way MushroomRL builds it, the MinMaxPreprocessor equivalent, and the cross-rank
reduction of the constraint log on a world_size-2 gloo group."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rl_on_manifold_b200.mdp import Box, MDPInfo
from rl_on_manifold_b200.rollout import BatchedCore, MinMaxPreprocessor, UniformAgent, compute_J, episode_returns
from rl_on_manifold_b200.sharding import reduce_constraint_logs, reduce_stats


def _mushroom_compute_J(reward, last, gamma):
    """mushroom_rl.utils.dataset.compute_J restated for one environment's list of (reward, last)."""
    js, j, steps = [], 0.0, 0
    for i in range(len(reward)):
        j += gamma ** steps * reward[i]
        steps += 1
        if last[i] or i == len(reward) - 1:
            js.append(j)
            j, steps = 0.0, 0
    return js


def test_compute_J_per_environment_boundaries_and_trailing_episode():
    rng = np.random.default_rng(0)
    T, B = 40, 6
    reward = rng.normal(size=(T, B))
    last = rng.uniform(size=(T, B)) < 0.15
    data = dict(reward=torch.from_numpy(reward), last=torch.from_numpy(last))
    for gamma in (1.0, 0.9):
        J, env = episode_returns(data, gamma)
        for b in range(B):
            want = _mushroom_compute_J(reward[:, b], last[:, b], gamma)
            got = J[env == b].tolist()
            np.testing.assert_allclose(got, want, rtol=1e-12)
    # lock-step episodes (every ATACOM env of the reference): a [n_episodes, B] tensor, trailing partial episode kept
    last2 = np.zeros((T, B), bool)
    last2[9], last2[19], last2[29] = True, True, True
    J = compute_J(dict(reward=torch.from_numpy(reward), last=torch.from_numpy(last2)), 0.99)
    assert J.shape == (4, B)
    np.testing.assert_allclose(J[3, 2], sum(0.99 ** t * reward[30 + t, 2] for t in range(10)), rtol=1e-12)
    assert compute_J(data, 1.0).dim() == 1            # ragged boundaries: the flat list


class _ToyMDP:
    """Counter environment: state = [steps in episode, env id]; env b absorbs after b + 2 steps."""

    def __init__(self, B, horizon):
        self.B = B
        self._info = MDPInfo(Box(-np.ones(2) * np.inf, np.ones(2) * np.inf), Box(-np.ones(1), np.ones(1)), 0.9, horizon)
        self.resets = []
        self.state = None

    @property
    def info(self):
        return self._info

    def reset(self, state=None, mask=None):
        fresh = torch.stack([torch.zeros(self.B), torch.arange(self.B, dtype=torch.float32)], 1)
        if mask is None or self.state is None:
            self.state = fresh
            self.resets.append(None)
        else:
            self.state = torch.where(mask[:, None], fresh, self.state)
            self.resets.append(mask.clone())
        return self.state

    def step(self, action):
        self.state = self.state + torch.tensor([1.0, 0.0])
        absorbing = self.state[:, 0] >= self.state[:, 1] + 2
        return self.state, torch.ones(self.B), absorbing, {}


def test_batched_core_resets_only_the_environments_that_ended():
    B, H = 4, 5
    mdp = _ToyMDP(B, H)
    core = BatchedCore(UniformAgent(mdp, device="cpu"), mdp)
    data = core.evaluate(n_steps=12)
    steps = data["state"][:, :, 0]                    # steps into the episode BEFORE each transition
    for b in range(B):
        period = min(b + 2, H)                        # env b absorbs after b + 2 steps; the horizon caps it at 5
        assert steps[:, b].tolist() == [float(t % period) for t in range(12)]
        assert data["last"][:, b].tolist() == [(t % period) == period - 1 for t in range(12)]
        assert data["absorbing"][:, b].tolist() == [(t % period) == period - 1 and b + 2 <= H for t in range(12)]
    # the dataset keeps the true next state; only the driver's current state is replaced by the reset state
    assert (data["next_state"][:, :, 0] == steps + 1).all()
    masks = [m for m in mdp.resets if m is not None]
    assert masks and all(m.dtype == torch.bool and not bool(m.all()) for m in masks)
    J = compute_J(data, 1.0)
    assert J.dim() == 1 and float(J.sum()) == 12 * B  # reward 1 per step, every step in exactly one episode


def test_minmax_preprocessor_bounded_and_running_dimensions():
    low = np.array([-2.0, -np.inf, 0.0])
    high = np.array([2.0, np.inf, 10.0])
    info = MDPInfo(Box(low, high), Box(-np.ones(1), np.ones(1)), 0.99, 10)
    pre = MinMaxPreprocessor(info)
    rng = np.random.default_rng(1)
    seen = []
    for _ in range(5):
        obs = torch.from_numpy(np.stack([rng.uniform(-2, 2, 64), rng.normal(3.0, 5.0, 64), rng.uniform(0, 10, 64)], 1))
        out = pre(obs)
        seen.append(obs[:, 1].numpy())
        np.testing.assert_allclose(out[:, 0].numpy(), obs[:, 0].numpy() / 2.0, rtol=1e-12)
        np.testing.assert_allclose(out[:, 2].numpy(), (obs[:, 2].numpy() - 5.0) / 5.0, rtol=1e-12)
        allv = np.concatenate(seen)
        want = np.clip((obs[:, 1].numpy() - allv.mean()) / allv.std(), -10, 10)
        np.testing.assert_allclose(out[:, 1].numpy(), want, rtol=1e-9)
    st = pre.get_state()
    pre2 = MinMaxPreprocessor(info)
    pre2.set_state(st)
    assert float(pre2.count) == 5 * 64 and torch.allclose(pre2.std, pre.std)
    one = pre(torch.tensor([1.0, 3.0, 5.0], dtype=torch.float64))       # a single observation (B = 1 plumbing)
    assert one.shape == (3,) and abs(float(one[0]) - 0.5) < 1e-12 and abs(float(one[2])) < 1e-12
    # as a Core preprocessor: the agent and the dataset see normalised states
    mdp = _ToyMDP(3, 4)
    core = BatchedCore(UniformAgent(mdp, device="cpu"), mdp, preprocessors=[lambda s: s * 0.5])
    d = core.evaluate(n_steps=3)
    assert torch.equal(d["state"][1, :, 0], torch.full((3,), 0.5))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _stats_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank r logged 10 * (r + 1) samples with sum 3 * (r + 1), max c = 0.1 * (r + 1), max c_dq = -r
    total, count, c_max, c_dq = reduce_constraint_logs(3.0 * (rank + 1), 10 * (rank + 1), 0.1 * (rank + 1), -float(rank))
    stats = torch.tensor([3.0 * (rank + 1), 0.1 * (rank + 1), -float(rank), 10.0 * (rank + 1)], dtype=torch.float64)
    reduce_stats(stats)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array([total, count, c_max, c_dq] + stats.tolist()))
    dist.destroy_process_group()


def test_constraint_log_all_reduce_two_rank_gloo(tmp_path):
    """get_constraints_logs under env-index sharding (atacom.py:207-216 over the global batch): SUM on the sum and
    the count, MAX on the maxima; every rank ends up with the same global triple."""
    world = 2
    mp.start_processes(_stats_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, start_method="fork")
    for r in range(world):
        v = np.load(tmp_path / ("r%d.npy" % r))
        np.testing.assert_allclose(v[:4], [9.0, 30, 0.2, 0.0])
        np.testing.assert_allclose(v[4:], [9.0, 0.2, 0.0, 30.0])          # { sum, max c, max c_dq, count }
    # without a process group it is the identity
    assert reduce_constraint_logs(1.0, 2, 3.0, 4.0) == (1.0, 2, 3.0, 4.0)
