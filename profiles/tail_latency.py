"""Launch time of the iiwa step kernel when M environments of the batch have three constraints exactly active
(three slack pivots: the general null-space routine of atacom_dual.cuh) — the tail latency of the rare path.

Run on a GPU box:  python profiles/tail_latency.py > gpurun_out/tail_latency.log
Each batch is also checked against the float64 oracle on its three-active environments.
"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from rl_on_manifold_b200 import _lib, projection, synthetic
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import helpers

dev = torch.device("cuda:0")
B = 65536
p = _lib.default_params("iiwa", 6)
q, dq, s0, alpha = synthetic.device_batch("iiwa", B, 1234, dev, 6, p)
g = torch.Generator().manual_seed(5)
for M in (0, 1, 64, 2048, 65536):
    s = s0.clone()
    rows = torch.randperm(B, generator=g)[:M].to(dev)
    for r in range(3):
        cols = ((torch.arange(M, device=dev) * 3 + r * 4) % 11)          # three distinct slack columns per env
        s[rows, cols] = 0.0
    st = torch.zeros(B, dtype=torch.uint8, device=dev)
    ddq = torch.empty_like(q); so = torch.empty_like(s)
    f = lambda: projection.step("iiwa", q, dq, s, alpha, p, ddq=ddq, s_out=so, status=st)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    stc = st.cpu().numpy()
    msg = "envs with 3 active constraints: %5d -> %.1f us per launch; slack-pivot envs %d, rank-deficient %d, deferred flag %d" % (
        M, us, int(((stc & _lib.ST_SLACK_PIVOT) != 0).sum()), int(((stc & _lib.ST_RANK_DEFICIENT) != 0).sum()),
        int(((stc & _lib.ST_DENSE_PATH) != 0).sum()))
    if M:
        idx = rows[:min(M, 256)].cpu().numpy()
        ref = helpers.oracle_batch("iiwa6", *(t[idx].cpu().numpy().astype(np.float64) for t in (q, dq, s, alpha)),
                                   basis="canonical")
        ok = ~ref["rank_def"] & (ref["margin"] > 1e-3)
        e = helpers.rel_err(so[idx].cpu().numpy().astype(np.float64), ref["s_new"])
        msg += "; max rel err of s' vs oracle on %d of them: %.2e" % (int(ok.sum()), float(e[ok].max()) if ok.any() else 0.0)
    print(msg, flush=True)
