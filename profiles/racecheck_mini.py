"""Smallest run of the two-launch step and of the fused zero-copy launch for compute-sanitizer --tool racecheck."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_on_manifold_b200 import _lib, projection, synthetic
dev = torch.device("cuda:0"); B = 512
p = _lib.default_params("iiwa", 6)
q, dq, s, alpha = synthetic.device_batch("iiwa", B, 5, dev, 6, p)
st = torch.zeros(B, dtype=torch.uint8, device=dev)
ddq, so = projection.step("iiwa", q, dq, s, alpha, p, status=st)
torch.cuda.synchronize()
h = [t.cpu().pin_memory() for t in (q, dq, s, alpha)]
ddq_h, s_h = torch.empty(B, 6).pin_memory(), torch.empty(B, 11).pin_memory()
ctx = projection.HostContext(B, chunks=1, mode="zero_copy")
ctx.iiwa_step(6, *h, ddq_h, s_h, p)
print("deferred:", int(((st & 32) != 0).sum()), "equal:", torch.equal(ddq_h, ddq.cpu()) and torch.equal(s_h, so.cpu()))
