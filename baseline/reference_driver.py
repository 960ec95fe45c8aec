"""Drive the UNMODIFIED reference's hot path on host cores: `AtacomEnvWrapper.step_action_function`
(atacom/atacom.py:123-139) with everything it reaches in the reference's own code — `_construct_Jc_psi`,
`ConstraintsSet / ViabilityConstraint.fun / K_J / b` (constraints.py), `pinv_null` (SciPy SVD) and `rref`
(utils/null_space_coordinate.py), `_compute_error_correction`, `acc_truncation` — one environment per call, the
way the reference runs it.  Used by `bench.py --impl reference` and by bench.py's `cpu_baseline` leg.

The package is imported from baseline/_ref (staged by baseline/stage_reference.py; falls back to /root/reference)
under the import shims of oracle/ref_loader.py (mushroom_rl and matplotlib are not installed).  The wrapper is
built with the ConstraintsSets of AirHockeyIiwaAtacom (iiwa_hit_atacom.py:23-40: f = 1 row K=0.1; g = 5 rows
K=0.5 + n rows K=1; Kc=240, Kq=4 acc/vel, dt=1/240).  Its leaf callbacks fun / J / b call pinocchio in the
reference; pinocchio is not available, so they return values computed BEFORE the timed loop by the oracle's
NumPy kinematics (oracle/envs.py) — the timed work is the reference's wrapper + constraint stacking + SVD + rref
only, which favours the CPU side (the real callbacks add ~30 pinocchio calls per step).

This process never imports rl_on_manifold_b200 (the input generator is loaded by file path), so the product's
native library is not mapped into the reference arm.
"""
import importlib.util
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref")

_W = {}     # per-process state of a pool worker


def reference_root():
    if os.path.isdir(os.path.join(STAGED, "atacom")):
        return STAGED
    return os.environ.get("ATACOM_REFERENCE_ROOT", "/root/reference")


def reference_present():
    return os.path.isdir(os.path.join(reference_root(), "atacom"))


def load_inputs_module():
    """rl_on_manifold_b200/_inputs.py by path: pure torch, does not touch the package or its .so."""
    spec = importlib.util.spec_from_file_location("_atacom_inputs",
                                                  os.path.join(ROOT, "rl_on_manifold_b200", "_inputs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _load_reference():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    os.environ["ATACOM_REFERENCE_ROOT"] = reference_root()
    from oracle import ref_loader
    ref_loader.REFERENCE_ROOT = reference_root()
    return ref_loader.load()


def build_iiwa_wrapper(R, n=6):
    """The reference's AtacomEnvWrapper with AirHockeyIiwaAtacom's constructor arguments
    (iiwa_hit_atacom.py:23-40) on a stub base env; the callbacks read `cur` (set per environment)."""
    from oracle import envs as oenv
    cur = {}

    class _Base(R.Environment):
        def __init__(self):
            box = R.Box(-np.ones(6 + 2 * n) * np.inf, np.ones(6 + 2 * n) * np.inf)
            super().__init__(R.MDPInfo(box, R.Box(-np.ones(n), np.ones(n)), 0.99, 120))

        def reset(self, state=None):
            return np.zeros(6 + 2 * n)

        def _create_observation(self, s):
            return s

    class _Wrap(R.AtacomEnvWrapper):
        def _get_q(self, st):
            return st[-2 * n:-n]

        def _get_dq(self, st):
            return st[-n:]

        def acc_to_ctrl_action(self, ddq):       # the reference's inverse dynamics belongs to the simulator
            return ddq

    f = R.ConstraintsSet(n)
    f.add_constraint(R.ViabilityConstraint(dim_q=n, dim_out=1, fun=lambda q: cur["ev"].c_f, J=lambda q: cur["ev"].J_f,
                                           b=lambda q, dq: cur["ev"].b_f, K=0.1))
    g = R.ConstraintsSet(n)
    g.add_constraint(R.ViabilityConstraint(dim_q=n, dim_out=5, fun=lambda q: cur["ev"].c_g[:5],
                                           J=lambda q: cur["ev"].J_g[:5], b=lambda q, dq: cur["ev"].b_g[:5], K=0.5))
    g.add_constraint(R.ViabilityConstraint(dim_q=n, dim_out=n, fun=lambda q: cur["ev"].c_g[5:],
                                           J=lambda q: cur["ev"].J_g[5:], b=lambda q, dq: cur["ev"].b_g[5:], K=1))
    acc_max = np.ones(n) * 10
    vel_max = oenv.IIWA_VEL_MAX[:n].copy()
    w = _Wrap(_Base(), n, f=f, g=g, Kc=240., vel_max=vel_max, acc_max=acc_max, Kq=4 * acc_max / vel_max,
              time_step=1 / 240.)
    return w, cur


def prepare_shard(args):
    """Pool task: build this worker's slice [lo, hi) of the seeded IiwaAirHockey-7H batch — inputs from
    _inputs.state_batch, constraint values from the oracle's FK, slacks from the reference's OWN reset rule
    (_compute_slack_variables, atacom.py:145-149) + the interior / boundary mix — and the wrapper."""
    lo, hi, B, seed, n = args
    try:
        from threadpoolctl import threadpool_limits
        _W["limit"] = threadpool_limits(1)
    except Exception:                                         # pragma: no cover
        pass
    import torch
    torch.set_num_threads(1)
    from oracle import envs as oenv
    R = _load_reference()
    inputs = load_inputs_module()
    q, dq, alpha = (t.double().numpy() for t in inputs.state_batch("iiwa", B, seed, n))
    w, cur = build_iiwa_wrapper(R, n)
    evs = [oenv.iiwa_eval(q[i], dq[i]) for i in range(lo, hi)]
    G = 5 + n
    s_full = np.zeros((B, G), np.float32)
    for j, i in enumerate(range(lo, hi)):
        cur["ev"] = evs[j]
        w.q, w.dq = q[i], dq[i]
        w._compute_slack_variables()
        s_full[i] = w.s
    s = inputs.slack_mix(torch.from_numpy(s_full), seed).double().numpy()[lo:hi]
    _W.update(w=w, cur=cur, evs=evs, q=q[lo:hi].copy(), dq=dq[lo:hi].copy(), s=s, alpha=alpha[lo:hi].copy(), n=n)
    return hi - lo


def run_shard(passes=1, keep=False):
    """Pool task: `passes` sweeps over the worker's slice, one step_action_function call per environment.
    Returns (seconds, checksum[, ddq, s_new])."""
    w, cur, n = _W["w"], _W["cur"], _W["n"]
    q, dq, s, alpha, evs = _W["q"], _W["dq"], _W["s"], _W["alpha"], _W["evs"]
    m = q.shape[0]
    sim_state = np.zeros(6 + 2 * n)
    out_ddq = np.zeros((m, n)) if keep else None
    out_s = np.zeros((m, 5 + n)) if keep else None
    t0 = time.perf_counter()
    acc = 0.0
    for _ in range(passes):
        for i in range(m):
            cur["ev"] = evs[i]
            w.q, w.dq = q[i], dq[i]
            w.s = s[i].copy()                      # the wrapper integrates its slack in place (atacom.py:135)
            ddq = w.step_action_function(sim_state, alpha[i])
            acc += ddq[0]
            if keep:
                out_ddq[i], out_s[i] = ddq, w.s
    dt = time.perf_counter() - t0
    return (dt, acc, out_ddq, out_s) if keep else (dt, acc)


class ReferencePool:
    """`cores` worker processes, each holding a slice of a B-environment batch and its own reference wrapper."""

    def __init__(self, B, seed=1234, n=6, cores=None):
        import multiprocessing as mp
        self.cores = cores or os.cpu_count() or 1
        self.B = B
        # one single-worker pool per slice, so a task always runs on the worker that holds that slice
        ctx = mp.get_context("fork")
        self.pools = [ctx.Pool(1) for _ in range(self.cores)]
        bounds = np.linspace(0, B, self.cores + 1).astype(int)
        jobs = [p.apply_async(prepare_shard, ((int(bounds[i]), int(bounds[i + 1]), B, seed, n),))
                for i, p in enumerate(self.pools)]
        assert sum(j.get() for j in jobs) == B

    def step(self, passes=1):
        """One pass (or `passes`) of the whole batch over all workers; returns wall seconds seen by the caller."""
        t0 = time.perf_counter()
        jobs = [p.apply_async(run_shard, (passes,)) for p in self.pools]
        res = [j.get() for j in jobs]
        return time.perf_counter() - t0, max(r[0] for r in res)

    def outputs(self):
        jobs = [p.apply_async(run_shard, (1, True)) for p in self.pools]
        res = [j.get() for j in jobs]
        return np.concatenate([r[2] for r in res]), np.concatenate([r[3] for r in res])

    def close(self):
        for p in self.pools:
            p.terminate()
            p.join()
