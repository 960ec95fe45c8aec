"""Zero-copy host path: sweep of the admission window (ATACOM_ZC_WINDOW, warps whose bulk loads may be in flight)
and of the block size of the launch (ATACOM_ZC_TPB, 0 = the device path's choice).

Run on a GPU box:  python profiles/zc_window_sweep.py [B] > gpurun_out/zc_sweep.log
Every setting is checked bit for bit against the device-buffer path (status included), then timed by wall clock
around the blocking C-ABI call (what the caller of atacom_iiwa_step_host sees).
"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from rl_on_manifold_b200 import _lib, projection, synthetic

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for n in (6, 7):
    p = _lib.default_params("iiwa", n)
    q, dq, s, alpha = synthetic.device_batch("iiwa", B, 1234, dev, n, p)
    ref = projection.step("iiwa", q, dq, s, alpha, p, n_ctrl_joints=n)
    host = [t.cpu().pin_memory() for t in (q, dq, s, alpha)]
    ddq_h = torch.empty(B, n).pin_memory()
    s_h = torch.empty(B, 5 + n).pin_memory()
    combos = [(w, 0) for w in (0, 32, 64, 96, 128, 192, 256, 384, 512, 1024)]
    combos += [(w, t) for t in (224, 128, 64) for w in (64, 128, 256)]      # smaller blocks: shorter tail
    for window, tpb in combos:
        os.environ["ATACOM_ZC_WINDOW"] = str(window)
        os.environ["ATACOM_ZC_TPB"] = str(tpb)
        ctx = projection.HostContext(B, chunks=1, mode="zero_copy")
        ddq_h.zero_(); s_h.zero_()
        f = lambda: ctx.iiwa_step(n, *host, ddq_h, s_h, p)
        for _ in range(5):
            f()
        ok = torch.equal(ddq_h, ref[0].cpu()) and torch.equal(s_h, ref[1].cpu())
        best = 1e9
        tot = 0.0
        for rep in range(5):
            t0 = time.perf_counter()
            for _ in range(40):
                f()
            dt = (time.perf_counter() - t0) / 40
            best = min(best, dt)
            tot += dt
        ok2 = torch.equal(ddq_h, ref[0].cpu()) and torch.equal(s_h, ref[1].cpu())
        print("n=%d B=%d window %4d block %3d: mean %.1f us, best %.1f us -> %.1f M env-steps/s  bit-identical to device path: %s"
              % (n, B, window, tpb, tot / 5 * 1e6, best * 1e6, B / best / 1e6, ok and ok2), flush=True)
        ctx.close()
