#!/usr/bin/env python
"""Minimal driver for Nsight Compute: a few launches of the IiwaAirHockey-7H step kernel at the benchmark
batch size on cold inputs, nothing else in the process (no subprocesses, no clock sampler).

    ncu --set full --clock-control none --import-source on -k regex:atacom_step_kernel -s 6 -c 2 \
        -o gpurun_out/prof python profiles/profile_step.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_on_manifold_b200 import _lib, projection, synthetic  # noqa: E402


def time_kernel(B=65536, steps=200, family="iiwa", nj=6, ring_n=23):
    """Device time per launch over `steps` back-to-back launches on a ring of distinct batches (cold inputs)."""
    dev = torch.device("cuda:0")
    p = _lib.default_params(family, nj) if family == "iiwa" else _lib.default_params(family)
    base = synthetic.device_batch(family, B, 1234, dev, nj, p)
    ring = [tuple(t.clone() for t in base) for _ in range(ring_n)]
    outs = [(torch.empty_like(base[0]), torch.empty_like(base[2])) for _ in range(ring_n)]
    status = torch.zeros(B, dtype=torch.uint8, device=dev)
    projection.step(family, *base, p, n_ctrl_joints=nj, status=status)
    torch.cuda.synchronize()
    st = status.cpu()
    for i in range(20):
        q, dq, s, alpha = ring[i % ring_n]
        projection.step(family, q, dq, s, alpha, p, n_ctrl_joints=nj, ddq=outs[i % ring_n][0], s_out=outs[i % ring_n][1])
    torch.cuda.synchronize()
    def run(i):
        q, dq, s, alpha = ring[i % ring_n]
        projection.step(family, q, dq, s, alpha, p, n_ctrl_joints=nj, ddq=outs[i % ring_n][0], s_out=outs[i % ring_n][1])

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(3):
            run(i)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(steps):
            run(i)
    graph.replay()
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps * 1e3)
    print("%s B=%d: %.2f us/launch (CUDA graph of 200, best of 5); status: dropped %.3f slack-pivot %.3f fallback %.4f rank-def %.5f"
          % (family, B, best, ((st & 2) != 0).float().mean(), ((st & 4) != 0).float().mean(),
             ((st & 16) != 0).float().mean(), ((st & 1) != 0).float().mean()))


def main(B=65536, launches=10, family="iiwa", nj=6):
    dev = torch.device("cuda:0")
    p = _lib.default_params(family, nj) if family == "iiwa" else _lib.default_params(family)
    ring = [synthetic.device_batch(family, B, 1234 + i, dev, nj, p) for i in range(4)]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    for i in range(launches):
        flush.fill_(i & 0xFF)                       # evict L2 between launches
        q, dq, s, alpha = ring[i % len(ring)]
        projection.step(family, q, dq, s, alpha, p, n_ctrl_joints=nj)
    torch.cuda.synchronize()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "time":
        print("lib:", _lib.LIB_PATH)
        if len(sys.argv) > 2 and sys.argv[2] == "sweep":       # latency- or throughput-bound?  time vs batch size
            for B in (4736, 9472, 18944, 37888, 65536, 131072, 262144):
                time_kernel(B=B)
            sys.exit(0)
        time_kernel()
        if len(sys.argv) > 2 and sys.argv[2] == "all":
            time_kernel(family="planar", B=16384)
            time_kernel(family="circle", B=4096)
    else:
        main()
