#!/usr/bin/env python
"""Benchmark of the batched ATACOM projection step (BASELINE.json metric: env-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one projection (AtacomEnvWrapper.step_action_function, atacom/atacom.py:123-139) of a
batch of B = 65 536 IiwaAirHockey-7H environments (n=6, F=1, G=11) per GPU.  Weak scaling: every rank
owns its own 65 536 environments; at N > 1 each step ends with one NCCL all-gather of the projected
accelerations ddq.  Rank 0 prints one JSON line.

--impl reference times the reference's own per-environment NumPy path (the float64 oracle port of it,
oracle/atacom_oracle.py — the reference itself is Python and cannot travel to the GPU box) on all host
cores, on bounded samples of the same workload.
"""
import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "atacom_projection_env_steps_per_sec"
UNIT = "env-steps/s"
BATCH_PER_GPU = 65536
N_JOINTS = 6
DIMS = dict(n=6, F=1, G=11, k=5)
# algorithmic bytes per env-step, fp32, inputs read once + outputs written once:
# q6 + dq6 + s11 + alpha5 = 28 floats in, ddq6 + s11 = 17 floats out
BYTES_IN, BYTES_OUT = 28 * 4, 17 * 4
BYTES_PER_ENV_STEP = BYTES_IN + BYTES_OUT
L2_BYTES = 126 * 1024 * 1024


def workload_name(n_gpus):
    return "IiwaAirHockey-7H projection (n=6,F=1,G=11), batch %d per GPU x %d GPU" % (BATCH_PER_GPU, n_gpus)


# ----------------------------------------------------------------------------- CPU baseline (oracle port)

def _cpu_worker(args):
    """Per-environment float64 NumPy path, the way the reference runs it (SVD + rref per call)."""
    q, dq, s, alpha, reps = args
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(1)
    except Exception:                                         # pragma: no cover
        limiter = None
    from oracle import atacom_oracle as ao
    from oracle import envs as oenv
    spec = oenv.iiwa_spec(N_JOINTS)
    # The reference evaluates its constraint callbacks through pinocchio (C++), which is not available;
    # the oracle's NumPy kinematics would be ~10x slower than that, so the callbacks' values are
    # computed beforehand and only the wrapper + linear algebra is timed (this favours the CPU side).
    evs = [oenv.iiwa_eval(q[i], dq[i]) for i in range(q.shape[0])]
    t0 = time.perf_counter()
    acc = 0.0
    for _ in range(reps):
        for i in range(q.shape[0]):
            out = ao.atacom_step(spec, evs[i], dq[i], s[i], alpha[i], basis="svd")
            acc += out["ddq"][0]
    dt = time.perf_counter() - t0
    del limiter
    return dt, acc


def _cpu_sample(n_envs, seed):
    import numpy as np
    from oracle import atacom_oracle as ao
    from oracle import envs as oenv
    from rl_on_manifold_b200 import synthetic
    import torch
    q, dq, alpha = synthetic.state_batch("iiwa", n_envs, seed, N_JOINTS)
    q, dq, alpha = (t.double().numpy() for t in (q, dq, alpha))
    spec = oenv.iiwa_spec(N_JOINTS)
    s = np.stack([ao.slack_init(spec, oenv.iiwa_eval(q[i], dq[i]), dq[i]) for i in range(n_envs)])
    s = synthetic.slack_mix(torch.from_numpy(s.astype(np.float32)), seed).double().numpy()
    return q, dq, s, alpha


def cpu_env_steps_per_sec(pool, cores, sample, reps=1):
    """Run one bounded sample (`reps` passes over it) split over `cores` processes; returns
    (env-steps/s, seconds of the slowest worker's timed loop)."""
    import numpy as np
    q, dq, s, alpha = sample
    idx = np.array_split(np.arange(q.shape[0]), cores)
    res = pool.map(_cpu_worker, [(q[i], dq[i], s[i], alpha[i], reps) for i in idx])
    wall = max(r[0] for r in res)
    return reps * q.shape[0] / wall, wall


def run_cpu_baseline(envs_per_core=256, reps=24):
    cores = os.cpu_count() or 1
    sample = _cpu_sample(envs_per_core * cores, 1234)
    with mp.get_context("fork").Pool(cores) as pool:
        cpu_env_steps_per_sec(pool, cores, tuple(a[:8 * cores] for a in sample))     # warm the workers
        value, wall = cpu_env_steps_per_sec(pool, cores, sample, reps)
    return dict(value=value, unit=UNIT, cores=cores, kind="port",
                sample="%d passes over the first %d envs of the seed-1234 IiwaAirHockey-7H batch, float64 oracle "
                       "port of atacom.py:123-139 (SciPy SVD + rref per env; constraint callbacks precomputed, "
                       "pinocchio absent), %d processes x 1 thread, %.1f s timed"
                       % (reps, envs_per_core * cores, cores, wall))


def reference_arm(args):
    """bench.py --impl reference: the reference's CPU path (oracle port) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = 128 * cores                              # distinct envs; 4 passes per step
    sample = _cpu_sample(per_step, 1234)
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(max(args.warmup, 1)):
            cpu_env_steps_per_sec(pool, cores, tuple(a[:16 * cores] for a in sample))
        walls = []
        for _ in range(args.steps):
            walls.append(cpu_env_steps_per_sec(pool, cores, sample, 4)[1])
    total = sum(walls)
    per_step *= 4
    value = per_step * args.steps / total
    desc = ("%d envs per step (bounded sample of the %d-env batch), float64 oracle port of atacom.py:123-139 "
            "(SciPy SVD + rref per env; constraint callbacks precomputed, pinocchio absent), "
            "%d processes x 1 thread" % (per_step, BATCH_PER_GPU, cores))
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * total / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic", impl="reference",
                config=dict(workload=workload_name(args.gpus), sample=desc),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=desc),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    emit(line)


# ----------------------------------------------------------------------------- clocks

class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self, interval_ms=100):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", str(interval_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def pause(self):
        """Stop polling but keep what was sampled (start() may be called again)."""
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
            self.proc = None

    def stop(self):
        self.pause()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------- ours

def ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run" % (args.gpus, world))

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_cpu_baseline()               # before CUDA is initialised (fork-safe)

    import torch
    import torch.distributed as dist
    from rl_on_manifold_b200 import _lib, projection, synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    n, G, k = DIMS["n"], DIMS["G"], DIMS["k"]
    params = _lib.default_params("iiwa", N_JOINTS)
    # a ring of distinct batches whose footprint exceeds L2 twice over: every step's inputs are cold
    ring = max(2, -(-2 * L2_BYTES // (B * BYTES_PER_ENV_STEP)))
    sets = []
    for r in range(ring):
        q, dq, s, alpha = synthetic.device_batch("iiwa", B, 1234 + rank + 1000 * r, dev, N_JOINTS, params)
        sets.append(dict(q=q, dq=dq, s=s, alpha=alpha, ddq=torch.empty_like(q), s_out=torch.empty_like(s)))
    gathered = torch.empty(world * B, n, device=dev) if world > 1 else None
    fused, fused_note = None, "not attempted"
    if world > 1 and args.gather == "fused":
        try:
            from rl_on_manifold_b200.sharding import SymmetricGather
            fused = SymmetricGather(B, n)
            fused_note = "fused peer-store epilogue over NVLink symmetric memory + one device barrier per step"
        except Exception as exc:
            fused_note = "fused gather unavailable (%s: %s); NCCL all-gather used" % (type(exc).__name__, exc)

    def step(i):
        d = sets[i % ring]
        projection.step("iiwa", d["q"], d["dq"], d["s"], d["alpha"], params, n_ctrl_joints=N_JOINTS,
                        ddq=d["ddq"], s_out=d["s_out"])
        if world > 1:
            dist.all_gather_into_tensor(gathered, d["ddq"])

    def step_fused(i):
        d = sets[i % ring]
        fused.step(d["q"], d["dq"], d["s"], d["alpha"], params, n_ctrl_joints=N_JOINTS, s_out=d["s_out"])

    def kernel_only(i):
        d = sets[i % ring]
        projection.step("iiwa", d["q"], d["dq"], d["s"], d["alpha"], params, n_ctrl_joints=N_JOINTS,
                        ddq=d["ddq"], s_out=d["s_out"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    launches0 = _lib.launch_count()
    eager_ms = timed(step, args.steps, args.warmup)
    launches = _lib.launch_count() - launches0 - args.warmup
    total_ms, launch_mode = eager_ms, "eager launches from Python (ctypes)"
    gather_mode = "single GPU" if world == 1 else "one NCCL all-gather of ddq per step"
    nccl_ms = eager_ms if world > 1 else None
    if fused is not None:
        launches1 = _lib.launch_count()
        fused_ms = timed(step_fused, args.steps, args.warmup)
        if fused_ms < total_ms:
            total_ms, gather_mode = fused_ms, fused_note
            launches = _lib.launch_count() - launches1 - args.warmup
    kernel_ms = eager_ms if world == 1 else timed(kernel_only, args.steps, args.warmup)

    def timed_graph(fn, steps):
        """The same K steps captured once in a CUDA graph and replayed: removes the per-launch host cost."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(3):
                fn(i)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(steps):
                fn(i)
        graph.replay()                                   # warm-up replay
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    if not args.no_graph:
        try:
            graph_ms = timed_graph(kernel_only, args.steps)
            if world > 1:
                t = torch.tensor([graph_ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                graph_ms = float(t.item())
            if graph_ms < kernel_ms:
                kernel_ms = graph_ms
            if world == 1 and graph_ms < total_ms:
                total_ms = graph_ms
                launch_mode = "CUDA graph of %d launches, one replay" % args.steps
                if os.environ.get("ATACOM_PDL", "1") != "0":
                    launch_mode += ", programmatic dependent launch"
        except Exception as exc:                         # pragma: no cover
            launch_mode += " (graph capture failed: %s)" % type(exc).__name__
        if world > 1 and fused is not None:
            # the fused step (kernel with peer-store epilogue + symmetric-memory barrier) replayed as a graph
            try:
                fg_ms = timed_graph(step_fused, args.steps)
                t = torch.tensor([fg_ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                fg_ms = float(t.item())
                if rank == 0:
                    print("[bench] fused step: eager %.2f us, CUDA graph %.2f us; NCCL path eager %.2f us"
                          % (1e3 * fused_ms / args.steps, 1e3 * fg_ms / args.steps, 1e3 * nccl_ms / args.steps),
                          file=sys.stderr)
                if fg_ms < total_ms:
                    total_ms, gather_mode = fg_ms, fused_note
                    launch_mode = "CUDA graph of %d fused steps, one replay" % args.steps
            except Exception as exc:                     # pragma: no cover
                launch_mode += " (fused-step graph capture failed: %s)" % type(exc).__name__

    # End to end through the host-buffer entry point: pinned host arrays in, pinned host arrays out, every
    # byte of the step's inputs and outputs crosses PCIe inside the timed region.  Two data paths of the same
    # call are timed: zero-copy (the kernel reads and writes the caller's page-locked buffers directly; what
    # HostContext picks for pinned buffers) and staged (device buffers + copy engines, replayed as a CUDA graph).
    # nvidia-smi polling every 100 ms slows driver calls and PCIe traffic measurably (2x on some boxes), so the
    # sampler runs at 1 s here.
    if rank == 0:
        sampler.pause()
        sampler.start(1000)
    host = {k_: sets[0][k_].cpu().pin_memory() for k_ in ("q", "dq", "s", "alpha")}
    ddq_h = torch.empty(B, n).pin_memory()
    s_h = torch.empty(B, G).pin_memory()

    def time_e2e(mode, chunks):
        ctx = projection.HostContext(B, chunks=chunks, mode=mode)

        def e2e_step():
            ctx.iiwa_step(N_JOINTS, host["q"], host["dq"], host["s"], host["alpha"], ddq_h, s_h, params)
            return float(ddq_h[0, 0])                                # read the result on the host

        for _ in range(max(3, args.warmup)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        ctx.close()
        return ms

    e2e_staged_ms = time_e2e("staged", args.chunks)
    e2e_ms = time_e2e("auto", args.chunks)
    e2e_api = "atacom_iiwa_step_host, pinned host buffers, zero-copy (kernel loads/stores cross PCIe via cp.async.bulk)"
    if e2e_staged_ms < e2e_ms:
        e2e_ms, e2e_api = e2e_staged_ms, "atacom_iiwa_step_host, pinned host buffers, staged copies (CUDA graph)"

    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        k_ms = kernel_ms / args.steps
        achieved = BYTES_PER_ENV_STEP * B / (k_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get("bytes_per_launch")
        except Exception:
            pass
        line = dict(
            metric=METRIC, value=world * B * args.steps / (total_ms * 1e-3), unit=UNIT, n_gpus=world,
            steps=args.steps, warmup=args.warmup, ms_per_step=total_ms / args.steps, higher_is_better=True,
            scaling="weak", vs_baseline=None, dtype="f64 (kinematics and projection; fp32 I/O and velocity products)",
            data="synthetic",
            config=dict(workload=workload_name(world), batch_per_gpu=B, global_batch=world * B,
                        parallelism="env-shard x%d; %s" % (world, gather_mode) if world > 1 else "single GPU",
                        gather=fused_note if world > 1 else None,
                        nccl_all_gather_ms_per_step=(nccl_ms / args.steps) if nccl_ms else None,
                        cold_inputs="ring of %d distinct batches (%.0f MB > 2x L2) rotated every step"
                                    % (ring, ring * B * BYTES_PER_ENV_STEP / 1e6),
                        bytes_per_env_step=BYTES_PER_ENV_STEP, launch=launch_mode,
                        eager_ms_per_step=eager_ms / args.steps),
            e2e=dict(value=world * B * args.steps / (e2e_ms * 1e-3), unit=UNIT, h2d_bytes_per_step=B * BYTES_IN,
                     d2h_bytes_per_step=B * BYTES_OUT, api=e2e_api,
                     staged_value=world * B * args.steps / (e2e_staged_ms * 1e-3), staged_chunks=args.chunks),
            gpu_launches=int(launches),
            roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                          traffic=traffic, peak_source=peak_src, kernel="atacom_step_kernel<IiwaEnv<6>>",
                          kernel_us=k_ms * 1e3, frac_of_8TBs_nominal=achieved / 8000.0,
                          note="nominal bound; the kernel is FP64-issue and latency bound (~4 k warp-instructions "
                               "per 32 environments against 5.8 KB of traffic; ncu: FP64 pipe 32 % busy, 0.38 "
                               "instructions per cycle per scheduler, DRAM 4 %): DESIGN.md section 6"),
            clocks=clocks,
        )
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def _reserve_stdout():
    """Libraries (NCCL's version banner, torchrun warnings) write to fd 1; the contract is ONE JSON line on
    stdout.  Point fd 1 at stderr for the rest of the process and keep the real stdout for the result line."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


RESULT_OUT = None


def emit(line):
    out = RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global RESULT_OUT
    RESULT_OUT = _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--chunks", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
