"""Where the host-side microseconds of the zero-copy host-buffer call go (atacom_iiwa_step_host, 65 536 environments):
wall time of the call from Python at the full and at a tiny batch, and of cudaPointerGetAttributes.

Run on a GPU box:  python profiles/host_overhead.py
"""
import ctypes, os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from rl_on_manifold_b200 import _lib, projection, synthetic

dev = torch.device("cuda:0")
B = 65536
p = _lib.default_params("iiwa", 6)
q, dq, s, alpha = synthetic.device_batch("iiwa", B, 1234, dev, 6, p)
host = [t.cpu().pin_memory() for t in (q, dq, s, alpha)]
ddq_h = torch.empty(B, 6).pin_memory(); s_h = torch.empty(B, 11).pin_memory()
ctx = projection.HostContext(B, chunks=1, mode="zero_copy")
f = lambda: ctx.iiwa_step(6, *host, ddq_h, s_h, p)
for _ in range(10): f()
def wall(fn, it):
    t0 = time.perf_counter()
    for _ in range(it): fn()
    return (time.perf_counter() - t0) / it * 1e6
print("full call, B=65536: %.1f us" % min(wall(f, 50) for _ in range(5)))
small = [t[:32] for t in host]
h = lambda: ctx.iiwa_step(6, *small, ddq_h[:32], s_h[:32], p)
for _ in range(10): h()
print("32 environments (marshalling + pointer checks + launch + sync): %.2f us" % min(wall(h, 500) for _ in range(3)))
rt = ctypes.CDLL("libcudart.so.12")
class Attr(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int), ("device", ctypes.c_int), ("devicePointer", ctypes.c_void_p), ("hostPointer", ctypes.c_void_p)]
a = Attr()
ptr = ctypes.c_void_p(host[0].data_ptr())
k = lambda: rt.cudaPointerGetAttributes(ctypes.byref(a), ptr)
print("cudaPointerGetAttributes through ctypes: %.2f us (ctypes no-op call: %.2f us)" % (
    min(wall(k, 5000) for _ in range(3)), min(wall(lambda: rt.cudaGetLastError(), 5000) for _ in range(3))))
ctx.close()
