"""SURVEY.md §8f rows on the GPU: fused roll-outs of the two environments whose base dynamics are plain
arithmetic (circle, point reach) against the trajectories recorded from the UNMODIFIED reference
(tests/golden/), and the constraint-statistics kernels against the oracle."""
import math

import numpy as np
import pytest
import torch

from rl_on_manifold_b200 import _lib, projection, synthetic
from tests import helpers

pytestmark = pytest.mark.gpu


def _f(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)


def _windows(x, W):
    """[T, ...] -> [W, T - W + 1, ...]: window i holds x[i], ..., x[i + W - 1] along axis 0."""
    T = x.shape[0]
    return np.stack([x[w:T - W + 1 + w] for w in range(W)], 0)


@pytest.mark.parametrize("variant,tag,na", [(_lib.VARIANT_ATACOM, "circleA", 1),
                                            (_lib.VARIANT_ERROR_CORRECTION, "circleE", 2)])
def test_circle_rollout_reference_trajectory(cuda_device, golden, variant, tag, na):
    """Fused circle roll-out against the episode recorded from the reference.  The closed loop (K_c dt = 1,
    clipped random actions, slacks crossing zero) amplifies a 1e-16 perturbation to 1e-4 within 250 steps even
    between two float64 implementations, so parity is checked (i) teacher-forced: every recorded state advanced by
    ONE fused step must land on the next recorded state, reward and slack, and the constraint log accumulated over
    those launches must equal the reference's epoch log; (ii) over windows of 8 fused steps in one launch."""
    dev = cuda_device
    p = _lib.default_params("circle")
    p.variant = variant
    acts, states, s_ref, rew, logs = (golden[tag + "_" + k] for k in ("actions", "states", "s", "rewards", "logs"))
    T = acts.shape[0]
    state = _f(states[:-1], dev)
    s = _f(s_ref[:-1], dev)
    stats = projection.new_stats(dev)
    status = torch.zeros(T, dtype=torch.uint8, device=dev)
    rewards = projection.circle_rollout(state, s, _f(acts, dev)[None].contiguous(), p, stats=stats, status=status)
    torch.cuda.synchronize()
    e_state = np.abs(state.cpu().numpy() - states[1:]).max()
    e_s = np.abs(s.cpu().numpy() - s_ref[1:]).max()
    e_r = np.abs(rewards.cpu().numpy()[0] - rew).max()
    assert e_state < 2e-6 and e_s < 2e-5 and e_r < 1e-6, (e_state, e_s, e_r)      # fp32 storage of state / action
    c_avg, c_max, c_dq_max = projection.read_stats(stats)
    assert abs(c_avg - logs[0]) < 1e-6 and abs(c_max - logs[1]) < 1e-6 and abs(c_dq_max - logs[2]) < 1e-6
    assert ((status.cpu().numpy() & (_lib.ST_NONFINITE | _lib.ST_DENSE_PATH)) == 0).all()
    # (ii) windows of W fused steps
    W = 8
    nw = T - W + 1
    state = _f(states[:nw], dev)
    s = _f(s_ref[:nw], dev)
    rewards = projection.circle_rollout(state, s, _f(_windows(acts, W), dev), p)
    err = np.abs(state.cpu().numpy() - states[W:W + nw]).max(1)
    assert np.median(err) < 1e-6 and np.quantile(err, 0.99) < 1e-3, (np.median(err), err.max())
    assert np.median(np.abs(rewards.cpu().numpy()[-1] - rew[W - 1:])) < 1e-6
    print("\n[%s] one fused step from each of %d recorded states: |state - ref| %.1e, |s - ref| %.1e, |reward - ref| "
          "%.1e; log (%.6f, %.6f, %.6f); %d-step windows: median %.1e, max %.1e"
          % (tag, T, e_state, e_s, e_r, c_avg, c_max, c_dq_max, W, np.median(err), err.max()))


def test_circle_rollout_equals_stepwise_kernels(cuda_device):
    """Fused roll-out == the step kernel + torch base dynamics, environment by environment (config 2:
    4096 circle environments)."""
    dev = cuda_device
    B, T = 4096, 40
    p = _lib.default_params("circle")
    q, dq, s, _ = synthetic.device_batch("circle", B, 31, dev, params=p)
    gen = torch.Generator().manual_seed(5)
    actions = (torch.rand(T, B, 1, generator=gen) * 2.6 - 1.3).to(dev)
    state = torch.cat([q, dq], 1).contiguous()
    st_f, s_f = state.clone(), s.clone()
    rew_f = projection.circle_rollout(st_f, s_f, actions, p)
    st, ss = state.double(), s.clone()
    rew = []
    for t in range(T):
        alpha = actions[t].clamp(-1, 1) * 10.0
        qf, dqf = st[:, :2].float().contiguous(), st[:, 2:].float().contiguous()
        ddq, ss = projection.step("circle", qf, dqf, ss, alpha, p)
        u = (ddq.double() / 10.0).clamp(-1, 1) * 10.0
        st[:, :2] += st[:, 2:] * 0.01 + u * 0.01 ** 2 / 2
        st[:, 2:] += u * 0.01
        rew.append(torch.exp(-torch.sqrt((1 - st[:, 0]) ** 2 + st[:, 1] ** 2)))
    # the step-wise path rounds q, dq, s to fp32 every step and the closed loop amplifies that where a
    # constraint is active (see test_circle_rollout_reference_trajectory): most environments agree to fp32
    # rounding, a few per cent drift
    err = (st_f.double() - st).abs().max(1).values
    assert err.median() < 1e-6 and err.quantile(0.9) < 1e-4, (err.median(), err.quantile(0.9))
    assert (rew_f.double() - torch.stack(rew)).abs().median() < 1e-6


def test_point_reach_rollout_reference_trajectory(cuda_device, golden):
    """Same two checks for PointReachAtacom with the obstacles' random-walk draws recorded from the reference."""
    dev = cuda_device
    p = _lib.default_params("point_reach")
    pre, post, acts, s_ref, draws, rew, logs = (golden["collC_" + k] for k in
                                                ("pre", "post", "actions", "s", "obj_draws", "rewards", "logs"))
    T = acts.shape[0]
    assert np.array_equal(pre[1:], post[:-1])
    state = _f(pre, dev)
    s = _f(s_ref[:-1], dev)
    stats = projection.new_stats(dev)
    rewards = projection.point_reach_rollout(state, s, _f(acts, dev)[None].contiguous(), p,
                                             obstacle_draws=_f(draws, dev)[None].contiguous(), stats=stats)
    torch.cuda.synchronize()
    e_state = np.abs(state.cpu().numpy() - post).max()
    e_s = np.abs(s.cpu().numpy() - s_ref[1:]).max()
    e_r = np.abs(rewards.cpu().numpy()[0] - rew).max()
    assert e_state < 5e-6 and e_s < 2e-5 and e_r < 1e-6, (e_state, e_s, e_r)      # coordinates up to 10 in fp32
    c_avg, c_max, _ = projection.read_stats(stats)
    assert abs(c_avg - logs[0]) < 1e-5 and abs(c_max - logs[1]) < 1e-5
    W = 8
    nw = T - W + 1
    state = _f(pre[:nw], dev)
    s = _f(s_ref[:nw], dev)
    projection.point_reach_rollout(state, s, _f(_windows(acts, W), dev), p,
                                   obstacle_draws=_f(_windows(draws, W), dev))
    err = np.abs(state.cpu().numpy() - post[W - 1:]).max(1)
    assert np.median(err) < 1e-5 and np.quantile(err, 0.99) < 1e-3, (np.median(err), err.max())
    print("\n[collC] one fused step from each of %d recorded states: |state - ref| %.1e, |s - ref| %.1e; "
          "%d-step windows: median %.1e, max %.1e" % (T, e_state, e_s, W, np.median(err), err.max()))


def test_point_reach_rollout_full_batch(cuda_device):
    """Config 5: 65 536 point-reach environments, 4 moving obstacles; fused roll-out against the step kernel +
    torch dynamics on a sample, finiteness and wall / speed-limit invariants on all."""
    dev = cuda_device
    B, T, G = 65536, 25, 4
    p = _lib.default_params("point_reach")
    q, dq, P, DP, _ = (t.to(dev) for t in synthetic.point_reach_batch(B, 9, G))
    state = torch.cat([q, dq, torch.stack([P.view(B, G, 2), DP.view(B, G, 2)], 2).reshape(B, 4 * G)], 1).contiguous()
    s = projection.point_reach_slack_init(q, P, p)
    gen = torch.Generator().manual_seed(1)
    actions = (torch.rand(T, B, 2, generator=gen) * 2 - 1).to(dev)
    draws = (torch.rand(T, B, 2 * G, generator=gen) * 2 - 1).to(dev)
    st_f, s_f = state.clone(), s.clone()
    stats = projection.new_stats(dev)
    rewards = projection.point_reach_rollout(st_f, s_f, actions, p, obstacle_draws=draws, stats=stats)
    assert torch.isfinite(st_f).all() and torch.isfinite(s_f).all() and torch.isfinite(rewards).all()
    obj = st_f[:, 4:].view(B, G, 4)
    assert (obj[:, :, 2:].abs() <= 1.0).all() and (obj[:, :, :2] >= 2.0).all() and (obj[:, :, :2] <= 10.0).all()
    assert projection.read_stats(stats)[1] <= 0.36 + 1e-6          # c = 0.36 - |q - p|^2 <= 0.36
    # sample: the first steps through the single-step (fp32) kernel and torch arithmetic
    T2 = 4
    st_f, s_f = state.clone(), s.clone()
    projection.point_reach_rollout(st_f, s_f, actions[:T2].contiguous(), p, obstacle_draws=draws[:T2].contiguous())
    idx = torch.arange(0, B, 257, device=dev)
    st, ss = state[idx].double(), s[idx].clone()
    for t in range(T2):
        o = st[:, 4:].view(-1, G, 4)
        w, ss = projection.point_reach_step(st[:, :2].float().contiguous(), st[:, 2:4].float().contiguous(),
                                            o[:, :, :2].reshape(-1, 2 * G).float().contiguous(),
                                            o[:, :, 2:].reshape(-1, 2 * G).float().contiguous(), ss,
                                            actions[t, idx].contiguous(), p)
        u = w.double().clamp(-1, 1) * 10
        st[:, :2] += st[:, 2:4] * 0.01
        st[:, 2:4] += u * 0.01
        flip = (st[:, :2] <= 0) | (st[:, :2] >= 10)
        st[:, 2:4] = torch.where(flip, -st[:, 2:4], st[:, 2:4])
        o[:, :, :2] += o[:, :, 2:] * 0.01
        o[:, :, :2] = o[:, :, :2].clamp(2, 10)
        flip = (o[:, :, :2] <= 2) | (o[:, :, :2] >= 10)
        o[:, :, 2:] = torch.where(flip, -o[:, :, 2:], o[:, :, 2:])
        o[:, :, 2:] += draws[t, idx].double().view(-1, G, 2) * 10 * 0.01
        o[:, :, 2:] = o[:, :, 2:].clamp(-1, 1)
    err = (st_f[idx].double() - st).abs().max(1).values
    assert err.median() < 2e-5 and (err < 1e-3).float().mean() > 0.95, (err.median(), (err < 1e-3).float().mean())


@pytest.mark.parametrize("family,B", [("circle", 4096), ("planar", 16384), ("iiwa6", 65536), ("iiwa7", 4096)])
def test_constraint_stats_vs_oracle(cuda_device, family, B):
    dev = cuda_device
    fam, nj = ("iiwa", int(family[-1])) if family.startswith("iiwa") else (family, 6)
    p = _lib.default_params("iiwa", nj) if fam == "iiwa" else _lib.default_params(fam)
    q, dq, _ = synthetic.state_batch(fam, B, 17, nj, p)
    stats = projection.new_stats(dev)
    per_env = projection.constraint_stats(fam, q.to(dev), dq.to(dev), p, n_ctrl_joints=nj, stats=stats)
    pe = per_env.cpu().numpy()
    spec = helpers.oracle_spec(family)
    idx = np.arange(0, B, max(1, B // 256))
    for i in idx:
        ev = helpers.oracle_eval(family, q[i].double().numpy(), dq[i].double().numpy())
        c = np.concatenate([np.abs(ev.c_f), ev.c_g])                   # atacom.py:202-203
        assert abs(pe[i, 0] - c.max()) < 2e-6
        assert abs(pe[i, 1] - (np.abs(dq[i].double().numpy()) - spec.vel_max).max()) < 2e-6
    c_avg, c_max, c_dq_max = projection.read_stats(stats)
    assert abs(c_avg - pe[:, 0].astype(np.float64).mean()) < 1e-6
    assert abs(c_max - pe[:, 0].max()) < 1e-6 and abs(c_dq_max - pe[:, 1].max()) < 1e-6
    # accumulating a second batch keeps the running reduction
    projection.constraint_stats(fam, q.to(dev), dq.to(dev), p, n_ctrl_joints=nj, per_env=per_env, stats=stats)
    assert stats[3].item() == 2 * B and abs(projection.read_stats(stats)[0] - c_avg) < 1e-6
