"""The host-side mirror of the reference's interface (AtacomEnvWrapper and the env classes) on the GPU:
the reference's own trajectories (golden vectors) replayed through the wrappers, batched roll-outs,
and a user-defined ConstraintsSet through the generic kernel."""
import numpy as np
import pytest
import torch

from oracle import atacom_oracle as ao
from oracle import envs as oenv
from rl_on_manifold_b200 import _lib
from rl_on_manifold_b200.atacom import AtacomEnvWrapper
from rl_on_manifold_b200.constraints import ConstraintsSet, ViabilityConstraint
from rl_on_manifold_b200.environments import (AirHockeyIiwaAtacom, AirHockeyPlanarAtacom, CircleEnvAtacom,
                                              CircleEnvErrorCorrection, PointReachAtacom)
from rl_on_manifold_b200.mdp import Box, MDPInfo
from tests import helpers

pytestmark = pytest.mark.gpu


def test_circle_env_a_reference_trajectory_b1_numpy_io(cuda_device, golden):
    """config 1 of BASELINE.json: CircleEnvAtacom, one env, NumPy in / NumPy out like a MushroomRL Core loop."""
    env = CircleEnvAtacom(n_envs=1, device=cuda_device)
    assert env.dims == {'q': 2, 'f': 1, 'g': 1, 'null': 1, 'c': 2}
    np.testing.assert_allclose(env.K_c, [100, 100]); np.testing.assert_allclose(env.K_q, [20, 20])
    np.testing.assert_allclose(env.alpha_max, [10])
    assert env.info.action_space.low.shape == (1,)
    st = env.reset()
    np.testing.assert_allclose(st.cpu().numpy()[0], [-1, 0, 0, 0])
    np.testing.assert_allclose(env.s.cpu().numpy(), [[1.0]])
    states, s_ref, rew = golden["circleA_states"], golden["circleA_s"], golden["circleA_rewards"]
    for i, a in enumerate(golden["circleA_actions"][:40]):
        state, r, absorbing, info = env.step(np.asarray(a))
        assert isinstance(state, np.ndarray) and state.shape == (4,)
        # closed loop in fp32 against the float64 trajectory: rounding is amplified by the K_c = 100 loop
        np.testing.assert_allclose(state, states[i + 1], atol=2e-4)
        np.testing.assert_allclose(env.s.cpu().numpy()[0], s_ref[i + 1], atol=2e-3)
        assert abs(float(r) - rew[i]) < 1e-4
    c_avg, c_max, c_dq_max = env.get_constraints_logs()
    assert c_max < 1e-2


def test_circle_env_e_reference_trajectory(cuda_device, golden):
    env = CircleEnvErrorCorrection(n_envs=1, device=cuda_device)
    assert env.info.action_space.low.shape == (2,)
    env.reset()
    for i, a in enumerate(golden["circleE_actions"][:20]):
        state, r, _, _ = env.step(np.asarray(a))
        np.testing.assert_allclose(state, golden["circleE_states"][i + 1], atol=2e-4)
        np.testing.assert_allclose(env.s.cpu().numpy()[0], golden["circleE_s"][i + 1], atol=2e-3)


def test_circle_env_batched_rollout_stays_on_manifold(cuda_device):
    """config 2: 4096 circle envs, random agent; ATACOM keeps every env on the circle and above y = -0.5."""
    B = 4096
    env = CircleEnvAtacom(n_envs=B, random_init=True, device=cuda_device)
    env.seed(3)
    env.reset()
    gen = torch.Generator(device="cpu").manual_seed(0)
    for t in range(200):
        a = (torch.rand(B, 1, generator=gen) * 2.4 - 1.2).to(cuda_device)
        state, r, absorbing, _ = env.step(a)
    assert torch.isfinite(state).all() and state.shape == (B, 4)
    radius_err = (state[:, 0] ** 2 + state[:, 1] ** 2 - 1).abs()
    assert radius_err.max() < 2e-2 and (state[:, 1] > -0.52).all()
    c_avg, c_max, c_dq_max = env.get_constraints_logs()
    assert c_max < 5e-2


def test_point_reach_env_reference_trajectory(cuda_device, golden):
    env = PointReachAtacom(n_objects=4, random_walk=True, n_envs=1, device=cuda_device)
    pre, acts, s_ref, post = (golden["collC_" + k] for k in ("pre", "actions", "s", "post"))
    env.reset(state=pre[0])
    np.testing.assert_allclose(env.s.cpu().numpy()[0], s_ref[0], rtol=1e-6)
    for i in range(30):
        env._state = torch.as_tensor(pre[i], dtype=torch.float32, device=cuda_device)[None].clone()
        env.s.copy_(torch.as_tensor(s_ref[i], dtype=torch.float32, device=cuda_device)[None])
        env.random_walk = False                      # obstacle motion is random in the reference; agent part only
        state, r, _, _ = env.step(acts[i].astype(np.float32))
        np.testing.assert_allclose(env.s.cpu().numpy()[0], s_ref[i + 1], rtol=2e-5, atol=1e-5)
        np.testing.assert_allclose(state[:4], post[i][:4], rtol=1e-5, atol=1e-5)


def test_point_reach_batched_avoids_obstacles(cuda_device):
    B = 2048
    env = PointReachAtacom(n_objects=4, random_walk=False, n_envs=B, device=cuda_device)
    env.seed(1)
    env.reset()
    gen = torch.Generator(device="cpu").manual_seed(0)
    for t in range(100):
        state, r, _, _ = env.step((torch.rand(B, 2, generator=gen) * 2 - 1).to(cuda_device))
    assert torch.isfinite(state).all()
    c_avg, c_max, _ = env.get_constraints_logs()
    assert c_max < 0.36                                # never reaches an obstacle centre; violations stay bounded


@pytest.mark.parametrize("n", [6, 7])
def test_iiwa_wrapper_rollout(cuda_device, n):
    """config 4 shape through the wrapper: 4 projections per agent step with q, dq frozen."""
    B = 1024
    env = AirHockeyIiwaAtacom(n_ctrl_joints=n, n_envs=B, device=cuda_device)
    assert env.dims == {'q': n, 'f': 1, 'g': 5 + n, 'null': n - 1, 'c': 6 + n}
    env.reset()
    s0 = env.s.clone()
    assert (s0 > 0).all()                              # home pose is strictly feasible
    gen = torch.Generator(device="cpu").manual_seed(0)
    for t in range(30):
        a = (torch.rand(B, n - 1, generator=gen) * 2 - 1).to(cuda_device)
        state, r, _, _ = env.step(a)
    assert torch.isfinite(state).all()
    q = env._get_q(state).double().cpu().numpy()
    dq = env._get_dq(state).double().cpu().numpy()
    worst_f, worst_g = 0.0, -1e9
    for i in range(0, B, 64):
        ev = oenv.iiwa_eval(q[i], dq[i])
        worst_f = max(worst_f, abs(ev.c_f[0]))
        worst_g = max(worst_g, ev.c_g.max())
    assert worst_f < 5e-3                              # striker tip stays on the table plane
    assert worst_g < 1e-2                              # inequality constraints hold
    vel_max = torch.as_tensor(env.vel_max, device=cuda_device)
    assert (env._get_dq(state).abs() <= vel_max * 1.05).all()


def test_planar_wrapper_rollout(cuda_device):
    B = 1024
    env = AirHockeyPlanarAtacom(n_envs=B, device=cuda_device)
    env.reset()
    gen = torch.Generator(device="cpu").manual_seed(0)
    for t in range(30):
        state, r, _, _ = env.step((torch.rand(B, 3, generator=gen) * 2 - 1).to(cuda_device))
    assert torch.isfinite(state).all()
    q, dq = env._get_q(state).double().cpu().numpy(), env._get_dq(state).double().cpu().numpy()
    assert max(oenv.planar_eval(q[i], dq[i]).c_g.max() for i in range(0, B, 64)) < 1e-2


class _VecEnv:
    """Minimal batched base env with the reference's hook protocol."""

    def __init__(self, dim, B, device):
        self.device, self.B, self.dim = device, B, dim
        self.info = MDPInfo(Box(-np.ones(2 * dim), np.ones(2 * dim)), Box(-np.ones(dim), np.ones(dim)), 0.99, 100)
        self.step_action_function = None
        self._state = torch.zeros(B, 2 * dim, device=device)

    def reset(self, state=None):
        if state is not None:
            self._state = state.clone()
        return self._state

    def step(self, action):
        u = self.step_action_function(self._state, action)
        self.last_u = u
        return self._state, torch.zeros(self.B, device=self.device), torch.zeros(self.B, dtype=torch.bool), {}

    def _create_observation(self, s):
        return s

    def seed(self, seed):
        pass


class _GenericWrap(AtacomEnvWrapper):
    def _get_q(self, st):
        return st[:, :self.dims['q']]

    def _get_dq(self, st):
        return st[:, self.dims['q']:]

    def acc_to_ctrl_action(self, ddq):
        return ddq


def test_generic_constraints_set_through_wrapper(cuda_device):
    """A user-defined ConstraintsSet with batched tensor callbacks (n=3, F=1, G=3): sphere equality plus three
    box inequalities, through AtacomEnvWrapper -> generic kernel, against the oracle."""
    n, B = 3, 512
    f = ConstraintsSet(n)
    f.add_constraint(ViabilityConstraint(n, 1, fun=lambda q: (q ** 2).sum(1, keepdim=True) - 1.0,
                                         J=lambda q: 2 * q[:, None, :],
                                         b=lambda q, dq: 2 * (dq ** 2).sum(1, keepdim=True), K=0.2))
    g = ConstraintsSet(n)
    g.add_constraint(ViabilityConstraint(n, 3, fun=lambda q: q - 0.9,
                                         J=lambda q: torch.eye(3, device=q.device).expand(q.shape[0], 3, 3),
                                         b=lambda q, dq: torch.zeros_like(q), K=0.5))
    base = _VecEnv(n, B, cuda_device)
    env = _GenericWrap(base, n, vel_max=1.5, acc_max=10., f=f, g=g, Kc=50., Kq=13., time_step=0.01)
    gen = torch.Generator(device="cpu").manual_seed(1)
    q = torch.nn.functional.normalize(torch.randn(B, n, generator=gen), dim=1) * 0.8
    dq = torch.randn(B, n, generator=gen) * 0.3
    env.reset(state=torch.cat([q, dq], 1).to(cuda_device))
    s0 = env.s.clone().cpu().numpy().astype(np.float64)
    action = torch.rand(B, 2, generator=gen) * 2 - 1
    env.step(action.to(cuda_device))
    ddq = base.last_u.cpu().numpy()
    spec = ao.Spec(n=3, F=1, G=3, K_f=0.2, K_g=0.5, K_c=50., K_q=13., vel_max=1.5, acc_max=10., dt=0.01)
    qn, dqn = q.double().numpy(), dq.double().numpy()
    worst = 0.0
    for i in range(B):
        ev = ao.ConstraintEval(c_f=np.array([(qn[i] ** 2).sum() - 1]), J_f=2 * qn[i][None], b_f=np.array([2 * (dqn[i] ** 2).sum()]),
                               c_g=qn[i] - 0.9, J_g=np.eye(3), b_g=np.zeros(3))
        np.testing.assert_allclose(s0[i], ao.slack_init(spec, ev, dqn[i]), atol=1e-6)
        o = ao.atacom_step(spec, ev, dqn[i], s0[i], ao.scale_action(spec, action[i].double().numpy()), basis="svd")
        if any(p > 1e-9 and abs(p - 0.05) < 5e-4 for (_, _, p) in o["trace"]["dropped"] + o["trace"]["pivots"]):
            continue
        worst = max(worst, np.abs(ddq[i] - o["ddq"]).max() / max(1.0, np.abs(o["w"]).max()))
    assert worst < 5e-5


def test_env_rollout_equals_repeated_step(cuda_device, golden):
    """env.rollout(actions) (one fused launch) against T calls of env.step(): states, rewards, constraint log."""
    # circle, replaying the first steps of the reference episode in 64 identical environments
    acts = torch.tensor(golden["circleA_actions"][:30], dtype=torch.float32, device=cuda_device)
    B = 64
    a = CircleEnvAtacom(n_envs=B, device=cuda_device); a.reset()
    b = CircleEnvAtacom(n_envs=B, device=cuda_device); b.reset()
    rew_a = a.rollout(acts[:, None, :].repeat(1, B, 1))
    rew_b = torch.stack([b.step(acts[t][None, :].repeat(B, 1))[1] for t in range(acts.shape[0])])
    np.testing.assert_allclose(a.state.cpu().numpy()[0], golden["circleA_states"][30], atol=2e-6)
    np.testing.assert_allclose(a.state.cpu().numpy(), b.state.cpu().numpy(), atol=2e-4)
    np.testing.assert_allclose(a.s.cpu().numpy(), b.s.cpu().numpy(), atol=2e-3)
    np.testing.assert_allclose(rew_a.cpu().numpy(), rew_b.cpu().numpy(), atol=1e-4)
    la, lb = a.get_constraints_logs(), b.get_constraints_logs()
    np.testing.assert_allclose(la, lb, atol=1e-4)
    # a later step() keeps working on the advanced state, and the merged log covers both paths
    a.rollout(acts[:5, None, :].repeat(1, B, 1)); a.step(acts[5][None, :].repeat(B, 1))
    assert np.isfinite(a.get_constraints_logs()).all()
    # point reach with the recorded obstacle draws
    draws = torch.tensor(golden["collC_obj_draws"][:30], dtype=torch.float32, device=cuda_device)
    pacts = torch.tensor(golden["collC_actions"][:30], dtype=torch.float32, device=cuda_device)
    env = PointReachAtacom(n_objects=4, random_walk=True, n_envs=1, device=cuda_device)
    env.reset(golden["collC_pre"][0])
    rew = env.rollout(pacts[:, None, :], draws[:, None, :])
    np.testing.assert_allclose(env._state.cpu().numpy()[0], golden["collC_post"][29], atol=2e-5)
    np.testing.assert_allclose(env.s.cpu().numpy()[0], golden["collC_s"][30], atol=2e-5)
    np.testing.assert_allclose(rew.cpu().numpy()[:, 0], golden["collC_rewards"][:30], atol=1e-6)


def test_batched_core_circle_experiment_loop(cuda_device):
    """The experiment loop of examples/circle_exp.py (evaluate -> metrics -> learn) with a random agent over
    1024 circle environments: dataset layout, episode boundaries, J / R, the constraint log."""
    from rl_on_manifold_b200.rollout import BatchedCore, UniformAgent, compute_J, compute_metrics
    B, H = 1024, 50
    mdp = CircleEnvAtacom(horizon=H, gamma=0.99, n_envs=B, random_init=True, device=cuda_device)
    mdp.seed(0)
    agent = UniformAgent(mdp, seed=1)
    core = BatchedCore(agent, mdp)
    data = core.evaluate(n_episodes=2)
    assert data["state"].shape == (2 * H, B, 4) and data["action"].shape == (2 * H, B, 1)
    assert data["last"][H - 1].all() and data["last"][2 * H - 1].all() and int(data["last"].sum()) == 2 * B
    assert not data["absorbing"].any()
    # next_state of step t is the state of step t + 1 inside an episode
    assert torch.equal(data["next_state"][:H - 1], data["state"][1:H])
    J, R = compute_J(data, 0.99), compute_J(data, 1.0)
    assert J.shape == (2, B) and (R >= J).all() and torch.isfinite(J).all()
    ref = sum(0.99 ** t * data["reward"][t, 0].double() for t in range(H))
    assert abs(float(J[0, 0]) - float(ref)) < 1e-9
    mdp.get_constraints_logs()
    Jm, Rm, c_avg, c_max, c_dq_max = compute_metrics(core, n_episodes=1)
    assert c_max < 2e-2 and c_dq_max < 0.5 and 0 < Jm < Rm          # ATACOM keeps the agent on the manifold
    core.learn(n_steps=30, n_steps_per_fit=10)
    assert agent.n_fits == 3


@pytest.mark.parametrize("cls,n", [(AirHockeyIiwaAtacom, 6), (AirHockeyPlanarAtacom, 3)])
def test_air_hockey_constraint_log_is_finite_and_matches_oracle(cuda_device, cls, n):
    """get_constraints_logs() of the air-hockey wrappers (atacom.py:201-216 with origin_constr=True): finite, and
    equal to max(|f|, g) of the oracle's constraint functions along the roll-out."""
    B = 64
    env = cls(n_envs=B, device=cuda_device)
    env.reset()
    gen = torch.Generator().manual_seed(0)
    want_c, want_dq = [], []
    fam = "iiwa6" if n == 6 else "planar"
    for _ in range(5):
        a = (torch.rand(B, env.info.action_space.low.shape[0], generator=gen) * 2 - 1).to(cuda_device)
        env.step(a)
        qn, dqn = env.q.double().cpu().numpy(), env.dq.double().cpu().numpy()
        for i in range(B):
            ev = helpers.oracle_eval(fam, qn[i], dqn[i])
            want_c.append(max(np.abs(ev.c_f).max() if ev.c_f.size else -np.inf, ev.c_g.max()))
            want_dq.append((np.abs(dqn[i]) - env.vel_max).max())
    c_avg, c_max, c_dq_max = env.get_constraints_logs()
    assert np.isfinite([c_avg, c_max, c_dq_max]).all()
    assert abs(c_avg - np.mean(want_c)) < 1e-5 and abs(c_max - np.max(want_c)) < 1e-5
    assert abs(c_dq_max - np.max(want_dq)) < 1e-5


def test_masked_reset_reinitialises_only_the_selected_environments(cuda_device):
    """Per-environment episode termination: reset(mask=...) re-initialises the state and — through
    the slack-init kernels' mask — the slacks of the selected environments only."""
    B = 96
    mask = torch.arange(B, device=cuda_device) % 3 == 0
    for env in (CircleEnvAtacom(n_envs=B, random_init=True, device=cuda_device),
                AirHockeyIiwaAtacom(n_envs=B, device=cuda_device),
                PointReachAtacom(n_objects=4, random_walk=True, n_envs=B, device=cuda_device)):
        env.seed(1)
        env.reset()
        gen = torch.Generator().manual_seed(2)
        adim = env.info.action_space.low.shape[0]
        for _ in range(3):
            env.step((torch.rand(B, adim, generator=gen) * 2 - 1).to(cuda_device))
        base = env.env if hasattr(env, "env") else env
        st0, s0 = base._state.clone(), env.s.clone()
        env.reset(mask=mask)
        st1, s1 = base._state, env.s
        assert torch.equal(st1[~mask], st0[~mask]) and torch.equal(s1[~mask], s0[~mask])
        assert not torch.equal(st1[mask], st0[mask])
        # the re-initialised slacks are what a full reset computes for those states
        fresh = type(env).__new__(type(env))
        if isinstance(env, PointReachAtacom):
            want = __import__("rl_on_manifold_b200").projection.point_reach_slack_init(*[env._split(st1)[i] for i in (0, 2)],
                                                                                       env.params)
        else:
            from rl_on_manifold_b200 import projection
            want = projection.slack_init(env._family, env._get_q(st1).contiguous(), env._get_dq(st1).contiguous(),
                                         env.params, n_ctrl_joints=env._n_ctrl_joints)
        assert torch.equal(s1[mask], want[mask])
        del fresh


def test_bare_viability_constraint_and_missing_callbacks(cuda_device):
    """The reference accepts a bare ViabilityConstraint for f / g (atacom.py:18-19); a constraint without callbacks
    outside a built-in family raises a clear error instead of a TypeError deep inside."""
    g = ViabilityConstraint(3, 3, fun=lambda q: q - 0.9, J=lambda q: torch.eye(3, device=q.device).expand(q.shape[0], 3, 3),
                            b=lambda q, dq: torch.zeros_like(q), K=0.5)
    f = ViabilityConstraint(3, 1, fun=lambda q: (q ** 2).sum(1, keepdim=True) - 1, J=lambda q: 2 * q[:, None, :],
                            b=lambda q, dq: 2 * (dq ** 2).sum(1, keepdim=True), K=0.2)

    class _Base:
        def __init__(self, B, dev):
            self.device, self.B = dev, B
            self._info = MDPInfo(Box(-np.ones(6) * np.inf, np.ones(6) * np.inf), Box(-np.ones(3), np.ones(3)), 0.99, 10)
            self.step_action_function = None

        info = property(lambda self: self._info)

        def reset(self, state=None):
            self.state = torch.zeros(self.B, 6, device=self.device)
            self.state[:, 0] = 1.0
            return self.state

        def _create_observation(self, s):
            return s

    class _W(AtacomEnvWrapper):
        def _get_q(self, st): return st[:, :3]
        def _get_dq(self, st): return st[:, 3:]
        def acc_to_ctrl_action(self, ddq): return ddq

    w = _W(_Base(8, cuda_device), 3, vel_max=1.5, acc_max=10., f=f, g=g, Kc=50., Kq=13.)
    assert isinstance(w.f, ConstraintsSet) and w.dims == {'q': 3, 'f': 1, 'g': 3, 'null': 2, 'c': 4}
    w.reset()
    out = w.step_action_function(w.state, torch.zeros(8, 2, device=cuda_device))
    assert out.shape == (8, 3) and torch.isfinite(out).all()
    with pytest.raises(ValueError, match="callbacks"):
        ViabilityConstraint(3, 1, K=1.0).fun(torch.zeros(2, 3), torch.zeros(2, 3))
