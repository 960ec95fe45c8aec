"""Duration of the step kernel and the fix-up kernel against the batch size (CUDA events, warm): separates the
latency of one environment's dependent chain from throughput.  python profiles/fix_scaling.py"""
import sys
import torch
sys.path.insert(0, ".")
from rl_on_manifold_b200 import _lib, projection, synthetic

dev = torch.device("cuda:0")
p = _lib.default_params("iiwa", 6)
pc = p.copy(); pc.basis_mode = _lib.BASIS_CANONICAL
for B in (1024, 4096, 16384, 32768, 65536, 131072):
    q, dq, s, alpha = synthetic.device_batch("iiwa", B, 7, dev, 6, p)
    out = {}
    for name, prm in (("lapack", p), ("canonical", pc)):
        for _ in range(5):
            projection.step("iiwa", q, dq, s, alpha, prm)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            projection.step("iiwa", q, dq, s, alpha, prm)
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / 50 * 1e3
    print("B %6d: lapack mode %.1f us/step, canonical %.1f us/step, difference (fix-up) %.1f us" % (B, out["lapack"], out["canonical"], out["lapack"] - out["canonical"]))
