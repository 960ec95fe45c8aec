#!/usr/bin/env python
"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python profiles/sanitizer_driver.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_on_manifold_b200 import _lib, projection, synthetic  # noqa: E402

dev = torch.device("cuda:0")
for fam, nj, B in (("iiwa", 6, 1000), ("iiwa", 7, 333), ("planar", 6, 500), ("circle", 6, 300)):
    p = _lib.default_params(fam, nj) if fam == "iiwa" else _lib.default_params(fam)
    q, dq, s, alpha = synthetic.device_batch(fam, B, 5, dev, nj, p)
    st = torch.zeros(B, dtype=torch.uint8, device=dev)
    ddq, so = projection.step(fam, q, dq, s, alpha, p, n_ctrl_joints=nj, status=st)
    projection.constraint_stats(fam, q, dq, p, n_ctrl_joints=nj, stats=projection.new_stats(dev))
    if fam == "iiwa":                       # bulk-copy instantiation on mapped host memory, and staged copies
        h = [t.cpu().pin_memory() for t in (q, dq, s, alpha)]
        ddq_h, s_h = torch.empty(B, nj).pin_memory(), torch.empty(B, 5 + nj).pin_memory()
        for mode in ("zero_copy", "staged"):
            ctx = projection.HostContext(B, chunks=2, mode=mode)
            ctx.iiwa_step(nj, *h, ddq_h, s_h, p)
            ctx.close()
            assert torch.equal(ddq_h, ddq.cpu()) and torch.equal(s_h, so.cpu())
p = _lib.default_params("circle")
q, dq, s, _ = synthetic.device_batch("circle", 200, 1, dev, params=p)
state = torch.cat([q, dq], 1).contiguous()
projection.circle_rollout(state, s, torch.rand(5, 200, 1, device=dev) * 2 - 1, p, stats=projection.new_stats(dev))
p = _lib.default_params("point_reach")
qq, dd, P, DP, act = (t.to(dev) for t in synthetic.point_reach_batch(200, 2, 4))
s = projection.point_reach_slack_init(qq, P, p)
projection.point_reach_step(qq, dd, P, DP, s, act, p)
state = torch.cat([qq, dd, torch.stack([P.view(200, 4, 2), DP.view(200, 4, 2)], 2).reshape(200, 16)], 1).contiguous()
projection.point_reach_rollout(state, s, torch.rand(5, 200, 2, device=dev) * 2 - 1, p,
                               obstacle_draws=torch.rand(5, 200, 8, device=dev) * 2 - 1)
torch.cuda.synchronize()
print("sanitizer driver done")
