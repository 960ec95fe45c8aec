import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from rl_on_manifold_b200 import _lib, projection, synthetic
dev = torch.device("cuda:0"); B = 65536
p = _lib.default_params("iiwa", 6)
q, dq, s, alpha = synthetic.device_batch("iiwa", B, 1234, dev, 6, p)
host = [t.cpu().pin_memory() for t in (q, dq, s, alpha)]
ddq_h = torch.empty(B, 6).pin_memory(); s_h = torch.empty(B, 11).pin_memory()
ctx = projection.HostContext(B, chunks=1, mode="zero_copy")
f = lambda: ctx.iiwa_step(6, *host, ddq_h, s_h, p)
for _ in range(10): f()
best = 1e9
for rep in range(5):
    t0 = time.perf_counter()
    for _ in range(100): f()
    best = min(best, (time.perf_counter() - t0) / 100)
ref = projection.step("iiwa", q, dq, s, alpha, p)
print("%s: %.1f us/call equal=%s" % (os.path.basename(_lib.LIB_PATH), best * 1e6, torch.equal(ddq_h, ref[0].cpu()) and torch.equal(s_h, ref[1].cpu())))
