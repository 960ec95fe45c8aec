"""The oracle against the reference: golden vectors taken from the reference's own classes
(tests/golden/make_golden.py), the live reference when /root/reference is present, and the
properties that tie the kernels' canonical null basis to the reference's SVD basis."""
import numpy as np
import pytest

from oracle import atacom_oracle as ao
from oracle import envs as oenv
from oracle import nullspace as ns
from oracle import ref_loader
from tests import helpers


def test_rref_matches_reference_golden(golden):
    for V, r05, rdef in zip(golden["rref_in"], golden["rref_tol005"], golden["rref_default"]):
        m = int((~np.isnan(V[:, 0])).sum())
        n = int((~np.isnan(V[0])).sum())
        V = V[:m, :n]
        np.testing.assert_allclose(ns.tol_rref(V, tol=0.05), r05[:m, :n], rtol=0, atol=1e-12)
        np.testing.assert_allclose(ns.tol_rref(V, tol=None), rdef[:m, :n], rtol=0, atol=1e-12)


def test_circle_a_survey_known_answers(golden):
    # SURVEY.md §8c table, recorded from CircleEnvAtacom
    st, s = golden["circleA_states"], golden["circleA_s"]
    np.testing.assert_allclose(st[0], [-1, 0, 0, 0])
    np.testing.assert_allclose(s[0], [1.0])
    np.testing.assert_allclose(st[1], [-1, 3.5e-4, 0, 0.07], atol=1e-15)
    np.testing.assert_allclose(s[1], [1.14])
    np.testing.assert_allclose(golden["circleA_act_b"][0], [0, 7, 14], atol=1e-12)
    np.testing.assert_allclose(st[2], [-0.99999961141, 9.6651064670e-4, 7.7718245269e-5, 0.05330212934], atol=1e-10)
    np.testing.assert_allclose(s[2], [1.1030300515], atol=1e-9)
    np.testing.assert_allclose(golden["circleA_rewards"][0], 0.1353352791, atol=1e-9)


def _circle_rollout(golden, key, variant):
    """One-step parity along the reference trajectory (each step restarts from the recorded state: the
    closed loop amplifies 1e-16 rounding differences through small rref pivots)."""
    spec = oenv.circle_spec()
    acts, states, s_ref = golden[key + "_actions"], golden[key + "_states"], golden[key + "_s"]
    s0 = ao.slack_init(spec, oenv.circle_eval(states[0][:2], states[0][2:]), states[0][2:])
    np.testing.assert_allclose(s0, s_ref[0], atol=1e-14)
    worst = 0.0
    for i, a in enumerate(acts):
        state, s = states[i], s_ref[i]
        alpha = ao.scale_action(spec, a, variant)
        o = ao.atacom_step(spec, oenv.circle_eval(state[:2], state[2:]), state[2:], s, alpha, variant=variant)
        for name in ("act_a", "act_b", "act_err"):
            ref = golden[key + "_" + name][i]
            worst = max(worst, np.abs(o[name] - ref).max() / max(1.0, np.abs(ref).max()))
        nxt, rew = oenv.circle_base_step(state, o["ddq"] / spec.acc_max)
        worst = max(worst, np.abs(nxt - states[i + 1]).max(), np.abs(o["s_new"] - s_ref[i + 1]).max(),
                    abs(rew - golden[key + "_rewards"][i]))
    return worst


def test_circle_a_trajectory_golden(golden):
    assert _circle_rollout(golden, "circleA", "atacom") < 1e-10


def test_circle_e_trajectory_golden(golden):
    assert _circle_rollout(golden, "circleE", "ec") < 1e-10


def test_circle_single_projections_golden(golden):
    spec = oenv.circle_spec()
    for q, dq, s, al, u, s2 in zip(*(golden["circleP_" + k] for k in ("q", "dq", "s", "alpha", "u", "s_new"))):
        o = ao.atacom_step(spec, oenv.circle_eval(q, dq), dq, s, al)
        np.testing.assert_allclose(o["ddq"] / spec.acc_max, u, atol=1e-10)
        np.testing.assert_allclose(o["s_new"], s2, atol=1e-10)


def test_collision_c_trajectory_golden(golden):
    pre, acts, s_ref, u_ref = (golden["collC_" + k] for k in ("pre", "actions", "s", "u"))
    # reset slack and first two steps
    np.testing.assert_allclose(s_ref[0], [8.9697322363, 4.1376184651, 3.3440775409, 5.2095362995], atol=1e-9)
    np.testing.assert_allclose(s_ref[1], [8.9920445308, 4.1384486547, 3.3369686473, 5.212973761], atol=1e-9)
    st = pre[0]
    s = ao.point_reach_slack_init(st[:2], st[4:].reshape(4, 4)[:, :2])
    np.testing.assert_allclose(s, s_ref[0], atol=1e-12)
    for i in range(len(acts)):
        st = pre[i]
        ob = st[4:].reshape(4, 4)
        o = ao.point_reach_step(st[:2], st[2:4], ob[:, :2], ob[:, 2:], s, acts[i])
        s = o["s_new"]
        np.testing.assert_allclose(s, s_ref[i + 1], atol=1e-9)
        np.testing.assert_allclose(o["u"], u_ref[i], atol=1e-9)


GENERIC = ["g423", "g313", "g306", "g6111", "g310"]


def generic_spec(meta):
    n, F, G = (int(x) for x in meta[:3])
    K_c = meta[3]
    rest = meta[4:]
    K_f, K_g, vel_max, acc_max = rest[:F], rest[F:F + G], rest[F + G:F + G + n], rest[F + G + n:]
    return ao.Spec(n=n, F=F, G=G, K_f=K_f, K_g=K_g, K_c=K_c, K_q=2 * acc_max / vel_max, vel_max=vel_max,
                   acc_max=acc_max, dt=0.01)


@pytest.mark.parametrize("tag", GENERIC)
def test_generic_wrapper_golden(golden, tag):
    """AtacomEnvWrapper.step_action_function driven with synthetic ConstraintsSets of several shapes."""
    spec = generic_spec(golden[tag + "_meta"])
    F = spec.F
    n_fired = 0
    for i in range(golden[tag + "_c"].shape[0]):
        g = {k: golden["%s_%s" % (tag, k)][i] for k in ("c", "J", "b", "dq", "s", "alpha", "ddq", "s_new",
                                                          "act_a", "act_b", "act_err")}
        ev = ao.ConstraintEval(c_f=g["c"][:F], J_f=g["J"][:F], b_f=g["b"][:F], c_g=g["c"][F:], J_g=g["J"][F:],
                               b_g=g["b"][F:])
        o = ao.atacom_step(spec, ev, g["dq"], g["s"], g["alpha"], basis="svd")
        fired = any(p > 1e-9 for (_, _, p) in o["trace"]["dropped"])
        n_fired += fired
        np.testing.assert_allclose(o["act_a"], g["act_a"], atol=1e-9)
        np.testing.assert_allclose(o["act_err"], g["act_err"], atol=1e-8)
        if not fired:   # where the tolerance branch fires the result depends on LAPACK's null basis
            np.testing.assert_allclose(o["act_b"], g["act_b"], atol=1e-8)
            np.testing.assert_allclose(o["ddq"], g["ddq"], atol=1e-8)
            np.testing.assert_allclose(o["s_new"], g["s_new"], atol=1e-9)
            oc = ao.atacom_step(spec, ev, g["dq"], g["s"], g["alpha"], basis="canonical")
            if not any(p > 1e-9 for (_, _, p) in oc["trace"]["dropped"]):
                np.testing.assert_allclose(oc["act_b"], g["act_b"], atol=1e-8)
    assert n_fired < golden[tag + "_c"].shape[0]


def test_canonical_basis_is_a_valid_reference_input():
    """The kernels' basis is orthonormal, spans null(Jc), and the reference's rref applied to it makes
    the same pivot / drop decisions it was built with."""
    rng = np.random.default_rng(0)
    spec = oenv.iiwa_spec(6)
    for t in range(60):
        q = np.array([0.0, 0.26, 0.0, -1.22, 0.0, 1.44]) + rng.uniform(-0.35, 0.35, 6)
        dq = rng.uniform(-0.5, 0.5, 6)
        ev = oenv.iiwa_eval(q, dq)
        s = np.maximum(ao.slack_init(spec, ev, dq), 0.3)
        if t % 2:
            s[rng.integers(11)] = rng.uniform(0, 0.02)
        A_f, A_g, *_ = ao.viability_terms(spec, ev, dq)
        Jc = ao.stack_Jc(spec, A_f, A_g, s)
        tr_c, tr_l = {}, {}
        V = ns.canonical_null_basis(Jc, spec.k, 0.05, trace=tr_c)
        np.testing.assert_allclose(V @ V.T, np.eye(spec.k), atol=1e-10)
        assert np.abs(Jc @ V.T).max() < 1e-10
        ns.tol_rref(V, tol=0.05, trace=tr_l)
        assert [c for (_, c, _) in tr_c["pivots"]] == [c for (_, c, _) in tr_l["pivots"]]


@pytest.mark.parametrize("n", [6, 7])
def test_iiwa_kinematics_finite_differences(n):
    rng = np.random.default_rng(n)
    for _ in range(5):
        q = rng.uniform(-0.8, 0.8, n) * oenv.IIWA_Q_MAX[:n]
        dq = rng.uniform(-0.5, 0.5, n) * oenv.IIWA_VEL_MAX[:n]
        ev = oenv.iiwa_eval(q, dq, bias="jdot_qdot")      # the full dJ/dt dq is what finite differences see
        cfun = lambda x: np.concatenate([oenv.iiwa_eval(x, dq).c_f, oenv.iiwa_eval(x, dq).c_g])
        h = 1e-6
        Jfd = np.stack([(cfun(q + h * e) - cfun(q - h * e)) / (2 * h) for e in np.eye(n)], 1)
        assert np.abs(np.vstack([ev.J_f, ev.J_g]) - Jfd).max() < 1e-8
        h = 1e-4
        bfd = (cfun(q + h * dq) - 2 * cfun(q) + cfun(q - h * dq)) / h ** 2
        assert np.abs(np.concatenate([ev.b_f, ev.b_g]) - bfd).max() < 1e-5
        # default mode (what the reference executes: classical acceleration with data.a = 0): omega x v of each
        # frame, checked against finite-difference velocities of the oracle's own FK
        ev2 = oenv.iiwa_eval(q, dq)
        qf, dqf = np.zeros(7), np.zeros(7)
        qf[:n], dqf[:n] = q, dq
        fr = oenv.chain_fk(oenv.IIWA_ORIGINS, qf)

        def frame_point(x, joint, off):
            f = oenv.chain_fk(oenv.IIWA_ORIGINS, x)
            return f[joint - 1][1] + f[joint - 1][0] @ np.asarray(off, float)

        want = []
        for joint, off in ((7, oenv.IIWA_TIP), (4, (0, 0, 0)), (7, (0, 0, 0))):
            v = (frame_point(qf + 1e-6 * dqf, joint, off) - frame_point(qf - 1e-6 * dqf, joint, off)) / 2e-6
            w = sum(fr[i][0][:, 2] * dqf[i] for i in range(joint))
            want.append(np.cross(w, v))
        b = np.array([want[0][2], -want[0][0], -want[0][1], want[0][1], -want[1][2], -want[2][2]])
        assert np.abs(np.concatenate([ev2.b_f, ev2.b_g[:5]]) - b).max() < 1e-7


def test_iiwa_fk_facts():
    # SURVEY.md §8c [probe]: q = 0 -> tip (0,0,1.846), link_4 z = 0.78, link_7 z = 1.261; tip independent of q7
    fr = oenv.chain_fk(oenv.IIWA_ORIGINS, np.zeros(7))
    np.testing.assert_allclose(fr[6][1] + fr[6][0] @ oenv.IIWA_TIP, [0, 0, 1.846], atol=1e-12)
    np.testing.assert_allclose(fr[3][1][2], 0.78, atol=1e-12)
    np.testing.assert_allclose(fr[6][1][2], 1.261, atol=1e-12)
    q = np.array([0.3, -0.5, 0.2, 1.0, -0.4, 0.7, 0.0])
    tips = []
    for q7 in (0.0, 1.3):
        q[6] = q7
        fr = oenv.chain_fk(oenv.IIWA_ORIGINS, q)
        tips.append(fr[6][1] + fr[6][0] @ oenv.IIWA_TIP)
    np.testing.assert_allclose(tips[0], tips[1], atol=1e-12)


def test_planar_kinematics_finite_differences():
    rng = np.random.default_rng(2)
    q, dq = rng.uniform(-0.8, 0.8, 3), rng.uniform(-1, 1, 3)
    ev = oenv.planar_eval(q, dq, bias="jdot_qdot")
    cfun = lambda x: oenv.planar_eval(x, dq).c_g
    h = 1e-6
    Jfd = np.stack([(cfun(q + h * e) - cfun(q - h * e)) / (2 * h) for e in np.eye(3)], 1)
    assert np.abs(ev.J_g - Jfd).max() < 1e-8
    h = 1e-4
    assert np.abs(ev.b_g - (cfun(q + h * dq) - 2 * cfun(q) + cfun(q - h * dq)) / h ** 2).max() < 1e-5


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present (GPU box)")
def test_oracle_against_live_reference():
    R = ref_loader.load()
    rng = np.random.default_rng(11)
    # rref on random matrices, both tolerances (the reference's own rref_test idea, null_space_coordinate.py:172-179)
    for _ in range(50):
        V = rng.normal(size=(rng.integers(1, 8), rng.integers(1, 10)))
        np.testing.assert_allclose(ns.tol_rref(V, tol=0.05), R.rref(V, tol=0.05), atol=1e-12)
        np.testing.assert_allclose(ns.tol_rref(V), R.rref(V), atol=1e-12)
        np.testing.assert_allclose(ns.tol_rref(V), R.rref_sympy(V), atol=1e-8)
        Jc = rng.normal(size=(3, 7))
        B, Q = R.pinv_null(Jc)
        pinv, Qo, _ = ns.svd_pinv_null(Jc)
        np.testing.assert_allclose(pinv, B, atol=1e-13)
        np.testing.assert_allclose(Qo, Q, atol=1e-13)
    # a fresh Circle-A trajectory
    env = R.CircleEnvAtacom()
    state = env.reset().copy()
    spec = oenv.circle_spec()
    s = ao.slack_init(spec, oenv.circle_eval(state[:2], state[2:]), state[2:])
    for _ in range(200):
        a = rng.uniform(-1.2, 1.2, 1)
        ref_state, _, _, _ = env.step(a)
        o = ao.atacom_step(spec, oenv.circle_eval(state[:2], state[2:]), state[2:], s, ao.scale_action(spec, a))
        s = o["s_new"]
        state, _ = oenv.circle_base_step(state, o["ddq"] / spec.acc_max)
        np.testing.assert_allclose(state, ref_state, atol=1e-12)
        np.testing.assert_allclose(s, env.s, atol=1e-12)


@pytest.mark.parametrize("shape", [(2, 3), (6, 9), (12, 17), (13, 19), (4, 6), (5, 7), (3, 5), (2, 4), (1, 3), (5, 8)])
def test_lapack_null_basis_is_scipy_svd_null_basis(shape):
    """The reference's null basis (SciPy svd -> LAPACK gesdd, null_space_coordinate.py:9-18) is NOT arbitrary: for
    a full-row-rank matrix it is the trailing columns of the product of the right Householder reflectors of the
    bidiagonalisation (or, for N >= 11 M / 6, of the LQ factorisation).  The restatement reproduces SciPy's
    vectors, signs included."""
    m, n = shape
    rng = np.random.default_rng(m * 100 + n)
    for t in range(60):
        A = rng.normal(size=(m, n))
        if t % 3 == 0 and m > 1:                      # the block structure of Jc: [[A_f, 0], [A_g, diag(s)]]
            nq = n - (m - 1)
            A[0, nq:] = 0.0
            A[1:, nq:] = np.diag(np.abs(rng.normal(size=m - 1)) * (0.01 if t % 2 else 1.0))
        _, Q, rank = ns.svd_pinv_null(A)
        assert rank == m
        np.testing.assert_allclose(ns.lapack_null_basis(A), Q, atol=1e-11)


@pytest.mark.parametrize("family", ["planar", "iiwa6", "iiwa7"])
def test_lapack_basis_reproduces_the_reference_on_both_strata(family):
    """tol_rref of the restated LAPACK basis == tol_rref of SciPy's SVD basis (what atacom.py:127-128 computes),
    on the synthetic benchmark batch including every environment where the tolerance branch fires."""
    B = 400
    q, dq, s, alpha = helpers.synthetic_cpu(family, B, seed=77)
    svd = helpers.oracle_batch(family, q, dq, s, alpha, basis="svd")
    lap = helpers.oracle_batch(family, q, dq, s, alpha, basis="lapack")
    ok = ~svd["rank_def"] & (svd["margin"] > 1e-6)
    assert svd["fired"][ok].sum() > 20
    assert helpers.rel_err(lap["w_null"], svd["w_null"])[ok].max() < 1e-9
    assert helpers.rel_err(lap["ddq"], svd["ddq"], svd["w"])[ok].max() < 1e-9
    assert (lap["fired"] == svd["fired"])[ok].all()


def test_canonical_mode_stratum_two_acceptance():
    """SURVEY.md §7.3-1 for the opt-in CANONICAL basis mode (the default mode needs no such test: it reproduces the
    reference's own LAPACK basis).  Where the reference's rref fires its tolerance branch, the canonical member of the
    output family is compared with the reference's (SciPy SVD basis): same pivot columns in the large majority,
    constraint residual |Jc Nc| comparable, and the distance to the reference within the reference's own spread
    under 16 random rotations of its null basis.  Stratum sizes are printed."""
    family, B = "iiwa6", 500
    q, dq, s, alpha = helpers.synthetic_cpu(family, B, 4242)
    spec = helpers.oracle_spec(family)
    rng = np.random.default_rng(0)
    rows = []
    n_ok = 0
    for i in range(B):
        ev = helpers.oracle_eval(family, q[i].astype(float), dq[i].astype(float))
        A_f, A_g, _, _, _ = ao.viability_terms(spec, ev, dq[i].astype(float))
        Jc = ao.stack_Jc(spec, A_f, A_g, s[i].astype(float))
        pinv, Q, rank = ns.svd_pinv_null(Jc)
        if rank < spec.C:
            continue
        n_ok += 1
        tr_s, tr_c = {}, {}
        Nc_s = ao.null_coordinates(Jc, spec.k, spec.tol, "svd", (pinv, Q), tr_s)
        Nc_c = ao.null_coordinates(Jc, spec.k, spec.tol, "canonical", (pinv, Q), tr_c)
        fired = [any(p > 1e-9 for (_, _, p) in t["dropped"]) for t in (tr_s, tr_c)]
        if not any(fired):
            np.testing.assert_allclose(Nc_c, Nc_s, atol=1e-9 * max(1.0, np.abs(Nc_s).max()))     # stratum I: basis-free
            continue
        a = alpha[i].astype(float)
        spread = 0.0
        for _ in range(16):
            M = np.linalg.qr(rng.normal(size=(spec.k, spec.k)))[0]
            spread = max(spread, np.abs(ns.tol_rref((Q[:, :spec.k] @ M).T, tol=spec.tol).T @ a - Nc_s @ a).max())
        rows.append(dict(same=[c for (_, c, _) in tr_s["pivots"]] == [c for (_, c, _) in tr_c["pivots"]],
                         res_s=np.abs(Jc @ Nc_s).max(), res_c=np.abs(Jc @ Nc_c).max(),
                         d=np.abs(Nc_c @ a - Nc_s @ a).max(), spread=spread))
    same = np.array([r["same"] for r in rows])
    d, spread = np.array([r["d"] for r in rows]), np.array([r["spread"] for r in rows])
    res_c, res_s = np.array([r["res_c"] for r in rows]), np.array([r["res_s"] for r in rows])
    print("\n[%s] %d environments, stratum II %d (%.1f %%): same pivot columns %.1f %%, |ours - ref| <= spread %.1f %%, "
          "median |Jc Nc| ours %.3g ref %.3g" % (family, n_ok, len(rows), 100.0 * len(rows) / n_ok, 100 * same.mean(),
                                                 100 * (d <= spread + 1e-9).mean(), np.median(res_c), np.median(res_s)))
    assert 0.1 < len(rows) / n_ok < 0.35
    assert same.mean() > 0.85
    assert (d <= spread + 1e-9).mean() > 0.93
    assert np.median(res_c) < 1.25 * np.median(res_s)
