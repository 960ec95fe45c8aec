"""End-to-end diagnosis on one box: PCIe copy bandwidth, then the host-buffer entry point in every data path, with the
reference-exact null basis (two launches per step) and the canonical one (a single launch) side by side.
    python profiles/e2e_diag.py > gpurun_out/e2e_diag.txt"""
import os, sys, time, torch, subprocess
sys.path.insert(0, os.getcwd())
from rl_on_manifold_b200 import _lib, projection, synthetic
print(subprocess.run("nvidia-smi topo -m | head -8; lscpu | grep -i -E 'numa|^CPU\\(s\\)|Model name'; cat /proc/self/status | grep -i cpus_allowed_list", shell=True, capture_output=True, text=True).stdout)
dev = torch.device("cuda:0")
B = 65536
p = _lib.default_params("iiwa", 6)
q, dq, s, alpha = synthetic.device_batch("iiwa", B, 1234, dev, 6, p)
host = [t.cpu().pin_memory() for t in (q, dq, s, alpha)]
ddq_h = torch.empty(B, 6).pin_memory(); s_h = torch.empty(B, 11).pin_memory()
big_h = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); big_d = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
def bw(fn, nbytes, it=20):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / it
    return nbytes / dt / 1e9, dt * 1e6
print("H2D 64MB: %.1f GB/s (%.0f us)" % bw(lambda: big_d.copy_(big_h, non_blocking=True), 64 << 20))
print("D2H 64MB: %.1f GB/s (%.0f us)" % bw(lambda: big_h.copy_(big_d, non_blocking=True), 64 << 20))
sm_h = big_h[:1 << 20]; sm_d = big_d[:1 << 20]
print("H2D 1MB: %.1f GB/s (%.0f us)" % bw(lambda: sm_d.copy_(sm_h, non_blocking=True), 1 << 20, 100))
print("D2H 1MB: %.1f GB/s (%.0f us)" % bw(lambda: sm_h.copy_(sm_d, non_blocking=True), 1 << 20, 100))
for basis, name in ((_lib.BASIS_LAPACK, "lapack (two launches)"), (_lib.BASIS_CANONICAL, "canonical (one launch)")):
    p.basis_mode = basis
    for mode, chunks in (("zero_copy", 1), ("staged", 1), ("staged", 2), ("staged", 4), ("hybrid", 1), ("hybrid", 2)):
        ctx = projection.HostContext(B, chunks=chunks, mode=mode)
        f = lambda: ctx.iiwa_step(6, *host, ddq_h, s_h, p)
        for _ in range(5): f()
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter()
            for _ in range(50): f()
            best = min(best, (time.perf_counter() - t0) / 50)
        print("%-24s %-9s chunks %d: %.0f us/step -> %.1f M env-steps/s" % (name, mode, chunks, best * 1e6, B / best / 1e6))
        ctx.close()
p.basis_mode = _lib.BASIS_LAPACK
ctx = projection.HostContext(B, chunks=1, mode="zero_copy")
ctx.iiwa_step(6, *host, ddq_h, s_h, p)
ref = projection.step("iiwa", q, dq, s, alpha, p)
print("zero-copy == device path:", torch.equal(ddq_h, ref[0].cpu()), torch.equal(s_h, ref[1].cpu()))
