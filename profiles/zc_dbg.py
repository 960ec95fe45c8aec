import os, sys, torch
sys.path.insert(0, os.getcwd())
from rl_on_manifold_b200 import _lib, projection, synthetic
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
p = _lib.default_params("iiwa", 6)
q, dq, s, alpha = synthetic.device_batch("iiwa", B, 1234, dev, 6, p)
host = [t.cpu().pin_memory() for t in (q, dq, s, alpha)]
ddq_h = torch.empty(B, 6).pin_memory(); s_h = torch.empty(B, 11).pin_memory()
for mode in ("zero_copy", "staged", "zero_copy", "staged"):
    ctx = projection.HostContext(B, chunks=1, mode=mode)
    try:
        for _ in range(3): ctx.iiwa_step(6, *host, ddq_h, s_h, p)
        torch.cuda.synchronize()
        print(mode, "ok")
    except Exception as e:
        print(mode, "FAILED", e)
        import ctypes
        break
    ctx.close()
ref = projection.step("iiwa", q, dq, s, alpha, p)
print("equal:", torch.equal(ddq_h, ref[0].cpu()), torch.equal(s_h, ref[1].cpu()))
