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
os.environ.setdefault("ATACOM_ZC_WINDOW", "5")      # ordered admission of the zero-copy launch: warps really wait
for fam, nj, B in (("iiwa", 6, 1000), ("iiwa", 7, 333), ("planar", 6, 500), ("circle", 6, 300)):
    p = _lib.default_params(fam, nj) if fam == "iiwa" else _lib.default_params(fam)
    q, dq, s, alpha = synthetic.device_batch(fam, B, 5, dev, nj, p)
    st = torch.zeros(B, dtype=torch.uint8, device=dev)
    ddq, so = projection.step(fam, q, dq, s, alpha, p, n_ctrl_joints=nj, status=st)
    projection.constraint_stats(fam, q, dq, p, n_ctrl_joints=nj, stats=projection.new_stats(dev))
    if fam != "circle":
        # two / three / four constraints exactly active: closed forms, the warp-cooperative general routine (few
        # askers per warp) and the serial one (every lane asks)
        G = s.shape[1]
        for rows, cols in ((slice(0, B, 7), (0, G - 1)), (slice(1, B, 50), (0, 1, G - 1)), (slice(64, 128), (1, 2, G - 2)),
                           (slice(2, B, 90), (0, 1, 2, G - 1))):
            if len(cols) > G - 2:
                continue
            s2 = s.clone()
            for c in cols:
                s2[rows, c] = 0.0
            ddq2, so2 = projection.step(fam, q, dq, s2, alpha, p, n_ctrl_joints=nj, status=st)
            assert torch.isfinite(ddq2).all() and torch.isfinite(so2).all()
        if fam == "iiwa":
            projection.iiwa_substeps(q, dq, s2, alpha, p, 3, n_ctrl_joints=nj)      # fused sub-steps, same rare paths
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
