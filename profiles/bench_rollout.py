#!/usr/bin/env python
"""Throughput of the fused roll-out kernels at BASELINE.json's configs 2 and 5, next to the
step-kernel + torch base-dynamics path they replace.  Prints one line per configuration."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_on_manifold_b200 import _lib, projection, synthetic  # noqa: E402
from rl_on_manifold_b200.environments import CircleEnvAtacom, PointReachAtacom  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def main():
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(0)
    # config 2: CircleMotion env A, 4096 environments
    B, T = 4096, 500
    env = CircleEnvAtacom(n_envs=B, random_init=True, device=dev)
    env.seed(1); env.reset()
    acts = (torch.rand(T, B, 1, generator=gen) * 2.6 - 1.3).to(dev)
    t_f = timed(lambda: env.rollout(acts))
    env.reset()
    t0 = time.perf_counter()
    for t in range(50):
        env.step(acts[t])
    torch.cuda.synchronize()
    t_s = (time.perf_counter() - t0) / 50 * T
    print("circle A, B=%d, T=%d: fused roll-out %.3f ms (%.2f G env-steps/s); step() loop %.1f ms (%.1f M env-steps/s)"
          % (B, T, t_f * 1e3, B * T / t_f / 1e9, t_s * 1e3, B * T / t_s / 1e6))
    # config 5: collision avoidance, 65536 environments, 4 obstacles
    B, T, G = 65536, 100, 4
    env = PointReachAtacom(n_objects=G, random_walk=True, n_envs=B, device=dev)
    env.seed(2); env.reset()
    acts = (torch.rand(T, B, 2, generator=gen) * 2 - 1).to(dev)
    draws = (torch.rand(T, B, 2 * G, generator=gen) * 2 - 1).to(dev)
    t_f = timed(lambda: env.rollout(acts, draws))
    env.reset()
    t0 = time.perf_counter()
    for t in range(20):
        env.step(acts[t])
    torch.cuda.synchronize()
    t_s = (time.perf_counter() - t0) / 20 * T
    print("collision C, B=%d, T=%d: fused roll-out %.3f ms (%.2f G env-steps/s); step() loop %.1f ms (%.1f M env-steps/s)"
          % (B, T, t_f * 1e3, B * T / t_f / 1e9, t_s * 1e3, B * T / t_s / 1e6))
    # constraint statistics kernel at the iiwa benchmark batch
    p = _lib.default_params("iiwa", 6)
    q, dq, s, alpha = synthetic.device_batch("iiwa", 65536, 3, dev, 6, p)
    stats = projection.new_stats(dev)
    per_env = torch.empty(65536, 2, device=dev)
    t_c = timed(lambda: [projection.constraint_stats("iiwa", q, dq, p, per_env=per_env, stats=stats) for _ in range(50)])
    print("iiwa constraint statistics, B=65536: %.2f us per launch" % (t_c / 50 * 1e6))
    # fused simulator sub-steps: K = 4 hook calls per agent step.  Feasible states only (every slack of the
    # reference's reset rule > 0.05, i.e. the environments are ON the constraint manifold as in operation): the
    # benchmark's interior / boundary mix overrides the reset slacks, which the first sub-step then corrects with
    # K_c dt = 1 and pushes several constraints active at once — the regime of the out-of-line general routine.
    qs, dqs, als = (t.to(dev) for t in synthetic.state_batch("iiwa", 1 << 20, 7, 6, p))
    ss = projection.slack_init("iiwa", qs, dqs, p)
    keep = torch.nonzero((ss > 0.05).all(1), as_tuple=True)[0][:65536]
    q, dq, alpha, s = (t[keep].contiguous() for t in (qs, dqs, als, ss))
    B = q.shape[0]
    K = 4
    ddq = torch.empty(K, B, 6, device=dev)
    s_out = torch.empty_like(s)
    st = torch.zeros(B, dtype=torch.uint8, device=dev)
    ws = torch.empty(_lib.lib.atacom_iiwa_substeps_workspace_doubles(6) * B, device=dev, dtype=torch.float64)
    projection.iiwa_substeps(q, dq, s, alpha, p, K, ddq=ddq, s_out=s_out, status=st, workspace=ws)
    print("feasible batch: B=%d, slack pivots in %d environments after %d sub-steps" % (B, int(((st & 4) != 0).sum()), K))
    for label, kw in (("parked kinematics", dict(workspace=ws)), ("recomputed kinematics", dict(use_workspace=False))):
        t_k = timed(lambda: [projection.iiwa_substeps(q, dq, s, alpha, p, K, ddq=ddq, s_out=s_out, **kw)
                             for _ in range(50)])
        print("iiwa fused sub-steps (%s), B=%d, K=4: %.2f us per launch = %.2f us per projection (%.2f G env-steps/s)"
              % (label, B, t_k / 50 * 1e6, t_k / 50 / K * 1e6, B * K * 50 / t_k / 1e9))
    ddq1 = torch.empty(B, 6, device=dev)
    t_1 = timed(lambda: [projection.step("iiwa", q, dq, s, alpha, p, ddq=ddq1, s_out=s_out) for _ in range(200)])
    print("iiwa single step on the same batch (eager launches, warm inputs): %.2f us per projection" % (t_1 / 200 * 1e6))


if __name__ == "__main__":
    main()
