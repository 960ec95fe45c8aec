#!/usr/bin/env python
"""Benchmark of the batched ATACOM projection step (BASELINE.json metric: env-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload iiwa|circle|planar|point_reach] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one projection (AtacomEnvWrapper.step_action_function, atacom/atacom.py:123-139) of one batch of
environments.  Default workload (the one the metric is quoted on, BASELINE.json configs[3]): B = 65 536
IiwaAirHockey-7H environments (n=6, F=1, G=11) per GPU (weak scaling) or in total (--scaling strong: 65 536 / N per
GPU); at N > 1 each step ends with the gather of the projected accelerations ddq, fused into the
kernel epilogue over NVLink symmetric memory.  The other workloads are BASELINE.json's configs[1], [2], [4].
Rank 0 prints ONE JSON line.

Timing: the K steps are captured in CUDA graphs over a ring of distinct input batches larger than 2x L2 (every
step reads cold data); R >= 5 replays of exactly K steps are timed with CUDA events, each bracketed by a barrier and
a device synchronisation, max over ranks per replay; `value` is computed from the MEDIAN replay (the minimum is
reported beside it).

--impl reference times the reference's own CPU implementation on all host cores: for the default workload the
UNMODIFIED reference package (baseline/_ref, staged by baseline/stage_reference.py) driven through
AtacomEnvWrapper.step_action_function (baseline/reference_driver.py); the oracle port otherwise.
"""
import argparse
import json
import math
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
L2_BYTES = 126 * 1024 * 1024

# algorithmic bytes per env-step, fp32, inputs read once + outputs written once
WORKLOADS = {
    "iiwa": dict(name="IiwaAirHockey-7H projection (n=6,F=1,G=11)", batch=65536, floats_in=28, floats_out=17,
                 kernel="atacom_step_kernel<IiwaEnv<6>>", out_dim=6),
    "circle": dict(name="CircularMotion env A projection (n=2,F=1,G=1)", batch=4096, floats_in=6, floats_out=3,
                   kernel="atacom_step_kernel<CircleEnv>", out_dim=2),
    "planar": dict(name="PlanarAirHockey env H projection (n=3,F=0,G=6)", batch=16384, floats_in=15, floats_out=9,
                   kernel="atacom_step_kernel<PlanarEnv>", out_dim=3),
    "point_reach": dict(name="CollisionAvoidance env C projection (n=2,G=4 obstacles)", batch=65536, floats_in=26,
                        floats_out=6, kernel="point_reach_step_kernel<4>", out_dim=2),
}
N_JOINTS = 6


def workload_label(wl, B, n_gpus, scaling):
    return "%s, batch %d per GPU x %d GPU (%s scaling)" % (WORKLOADS[wl]["name"], B, n_gpus, scaling)


def local_batch(args):
    B = args.batch or WORKLOADS[args.workload]["batch"]
    if args.scaling == "strong":
        if B % args.gpus:
            raise SystemExit("--scaling strong: the global batch %d is not divisible by %d GPUs" % (B, args.gpus))
        B //= args.gpus
    return B


# ----------------------------------------------------------------------------- CPU arms

def _port_worker(task):
    """Oracle port (float64 NumPy restatement, oracle/), one environment per call; callbacks' values precomputed."""
    wl, arrays, reps = task
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(1)
    except Exception:                                         # pragma: no cover
        limiter = None
    from oracle import atacom_oracle as ao
    from oracle import envs as oenv
    m = arrays[0].shape[0]
    acc = 0.0
    if wl == "point_reach":
        q, dq, p, dp, s, act = arrays
        t0 = time.perf_counter()
        for _ in range(reps):
            for i in range(m):
                acc += ao.point_reach_step(q[i], dq[i], p[i].reshape(-1, 2), dp[i].reshape(-1, 2), s[i], act[i])["w"][0]
        return time.perf_counter() - t0, acc
    q, dq, s, alpha = arrays
    spec = dict(iiwa=lambda: oenv.iiwa_spec(N_JOINTS), circle=oenv.circle_spec, planar=oenv.planar_spec)[wl]()
    ev_fn = dict(iiwa=oenv.iiwa_eval, circle=oenv.circle_eval, planar=oenv.planar_eval)[wl]
    evs = [ev_fn(q[i], dq[i]) for i in range(m)]
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(m):
            acc += ao.atacom_step(spec, evs[i], dq[i], s[i], alpha[i], basis="svd")["ddq"][0]
    dt = time.perf_counter() - t0
    del limiter
    return dt, acc


def _port_sample(wl, n_envs, seed):
    """Seeded inputs of a workload as float64 NumPy arrays (slacks from the oracle's reset rule + the mix).
    Loads the pure input generator by path: no product code in a CPU arm."""
    import numpy as np
    import torch
    from baseline import reference_driver as rd
    from oracle import atacom_oracle as ao
    from oracle import envs as oenv
    inputs = rd.load_inputs_module()
    if wl == "point_reach":
        q, dq, p, dp, act = (t.double().numpy() for t in inputs.point_reach_batch(n_envs, seed, 4))
        s = np.stack([ao.point_reach_slack_init(q[i], p[i].reshape(-1, 2)) for i in range(n_envs)])
        return q, dq, p, dp, s.astype(np.float32).astype(np.float64), act
    q, dq, alpha = (t.double().numpy() for t in inputs.state_batch(wl, n_envs, seed, N_JOINTS))
    spec = dict(iiwa=lambda: oenv.iiwa_spec(N_JOINTS), circle=oenv.circle_spec, planar=oenv.planar_spec)[wl]()
    ev_fn = dict(iiwa=oenv.iiwa_eval, circle=oenv.circle_eval, planar=oenv.planar_eval)[wl]
    s = np.stack([ao.slack_init(spec, ev_fn(q[i], dq[i]), dq[i]) for i in range(n_envs)])
    s = inputs.slack_mix(torch.from_numpy(s.astype(np.float32)), seed).double().numpy()
    return q, dq, s, alpha


PORT_DESC = ("float64 oracle port of atacom.py:123-139 (SciPy SVD + rref per env; constraint callbacks' values "
             "precomputed), %d processes x 1 thread")


def port_env_steps_per_sec(pool, cores, wl, sample, reps=1):
    import numpy as np
    idx = np.array_split(np.arange(sample[0].shape[0]), cores)
    t0 = time.perf_counter()
    pool.map(_port_worker, [(wl, tuple(a[i] for a in sample), reps) for i in idx])
    wall = time.perf_counter() - t0
    return reps * sample[0].shape[0] / wall, wall


def run_cpu_baseline(wl, budget_s=12.0):
    """The CPU path beside the GPU line, on a bounded sample (about `budget_s` seconds of all host cores).  Default
    workload: the reference itself (kind "reference") with the oracle port's number beside it; else the port."""
    from baseline import reference_driver as rd
    cores = os.cpu_count() or 1
    out = None
    if wl == "iiwa" and rd.reference_present():
        n_envs = 256 * cores
        pool = rd.ReferencePool(n_envs, seed=1234, n=N_JOINTS, cores=cores)
        try:
            pool.step()                                                   # warm the workers
            wall1, _ = pool.step()
            passes = max(1, int(math.ceil(budget_s / max(wall1, 1e-3))))
            wall, _ = pool.step(passes)
        finally:
            pool.close()
        out = dict(value=passes * n_envs / wall, unit=UNIT, cores=cores, kind="reference",
                   sample="%d passes over the first %d envs of the seed-1234 %s batch through the unmodified "
                          "reference's AtacomEnvWrapper.step_action_function (baseline/_ref; leaf callbacks return "
                          "values precomputed by the oracle's NumPy FK, pinocchio absent), %d processes x 1 thread, "
                          "%.1f s timed" % (passes, n_envs, WORKLOADS[wl]["name"], cores, wall))
        budget_s = 5.0
    n_envs = 128 * cores
    sample = _port_sample(wl, n_envs, 1234)
    with mp.get_context("fork").Pool(cores) as pool:
        port_env_steps_per_sec(pool, cores, wl, tuple(a[:8 * cores] for a in sample))     # warm the workers
        _, wall1 = port_env_steps_per_sec(pool, cores, wl, sample)
        reps = max(1, int(math.ceil(budget_s / max(wall1, 1e-3))))
        value, wall = port_env_steps_per_sec(pool, cores, wl, sample, reps)
    port = dict(value=value, unit=UNIT, cores=cores, kind="port",
                sample="%d passes over the first %d envs of the seed-1234 %s batch, " % (reps, n_envs, WORKLOADS[wl]["name"])
                       + PORT_DESC % cores + ", %.1f s timed" % wall)
    if out is None:
        return port
    out["port_value"], out["port_sample"] = port["value"], port["sample"]
    return out


def reference_arm(args):
    """bench.py --impl reference: the reference's CPU path on all host cores, same config as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from baseline import reference_driver as rd
    wl = args.workload
    cores = os.cpu_count() or 1
    B = local_batch(args)
    # a step = the full batch when the whole run stays within a few minutes, else a bounded sample of it
    per_step = B
    budget = 40
    if args.steps + args.warmup > budget:
        per_step = max(64 * cores, B * budget // (args.steps + args.warmup))
    if os.environ.get("ATACOM_REF_ENVS_PER_STEP"):
        per_step = int(os.environ["ATACOM_REF_ENVS_PER_STEP"])
    per_step = min(per_step, B)
    walls = []
    if wl == "iiwa" and rd.reference_present():
        kind = "reference"
        desc = ("%d envs per step%s through the unmodified reference's AtacomEnvWrapper.step_action_function "
                "(atacom.py:123-139, staged copy under baseline/_ref; leaf callbacks return values precomputed by the "
                "oracle's NumPy FK — pinocchio absent — so only the reference's wrapper, constraint stacking, SVD and "
                "rref are timed), float64, %d processes x 1 thread"
                % (per_step, "" if per_step == B else " (bounded sample of the %d-env batch)" % B, cores))
        pool = rd.ReferencePool(per_step, seed=1234, n=N_JOINTS, cores=cores)
        try:
            for _ in range(args.warmup):
                pool.step()
            for _ in range(args.steps):
                walls.append(pool.step()[0])
        finally:
            pool.close()
    else:
        kind = "port"
        desc = "%d envs per step%s, " % (per_step, "" if per_step == B else " (bounded sample of the %d-env batch)" % B) \
               + PORT_DESC % cores + ("; the reference tree is not staged on this box" if wl == "iiwa" else "")
        sample = _port_sample(wl, per_step, 1234)
        with mp.get_context("fork").Pool(cores) as pool:
            for _ in range(args.warmup):
                port_env_steps_per_sec(pool, cores, wl, sample)
            for _ in range(args.steps):
                walls.append(port_env_steps_per_sec(pool, cores, wl, sample)[1])
    total = sum(walls)
    value = per_step * args.steps / total
    cfg = dict(workload=workload_label(wl, B, args.gpus, args.scaling), batch_per_gpu=B, global_batch=B * args.gpus,
               envs_per_step=per_step, sample=desc)
    emit(dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
              ms_per_step=1e3 * statistics.median(walls), higher_is_better=True, scaling=args.scaling, vs_baseline=None,
              dtype="f64", data="synthetic", impl="reference", config=cfg,
              cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind=kind, sample=desc),
              e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0))


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
    wl = args.workload
    W = WORKLOADS[wl]
    B = local_batch(args)
    bytes_in, bytes_out = 4 * W["floats_in"], 4 * W["floats_out"]
    bytes_env = bytes_in + bytes_out
    K = args.steps

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_cpu_baseline(wl)             # before CUDA is initialised (fork-safe)

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

    # ---- inputs: a ring of distinct batches whose footprint exceeds L2 twice over; step (g, i) of graph g uses
    # slot (g K + i) mod ring, so a slot is only re-read after > 2x L2 of other data went by: every step is cold
    ring = max(2, -(-2 * L2_BYTES // (B * bytes_env)))
    n_graphs = max(1, -(-ring // K))
    n_out = W["out_dim"]
    fam = wl
    if wl == "point_reach":
        params = _lib.default_params("point_reach")
        q, dq, p, dp, s, act = synthetic.point_reach_device_batch(ring * B, 1234 + rank, dev, 4, params)
        big = dict(q=q, dq=dq, p=p, dp=dp, s=s, alpha=act)
    else:
        params = _lib.default_params("iiwa", N_JOINTS) if wl == "iiwa" else _lib.default_params(wl)
        q, dq, s, alpha = synthetic.device_batch(fam, ring * B, 1234 + rank, dev, N_JOINTS, params)
        big = dict(q=q, dq=dq, s=s, alpha=alpha)
    big["out"] = torch.empty(ring * B, n_out, device=dev)
    big["s_out"] = torch.empty_like(big["s"])
    sets = [{k_: v[r * B:(r + 1) * B] for k_, v in big.items()} for r in range(ring)]
    gathered = torch.empty(world * B, n_out, device=dev) if world > 1 else None

    def kernel_only(i, prm=None):
        d = sets[i % ring]
        if wl == "point_reach":
            projection.point_reach_step(d["q"], d["dq"], d["p"], d["dp"], d["s"], d["alpha"], prm or params, w=d["out"],
                                        s_out=d["s_out"])
        else:
            projection.step(fam, d["q"], d["dq"], d["s"], d["alpha"], prm or params, n_ctrl_joints=N_JOINTS,
                            ddq=d["out"], s_out=d["s_out"])

    # the canonical-basis mode (one launch per step, equal to the reference only where rref's tolerance branch stays
    # silent): timed beside the default mode, reported in config, never substituted for it
    params_canonical = params.copy()
    params_canonical.basis_mode = _lib.BASIS_CANONICAL

    def kernel_canonical(i):
        kernel_only(i, params_canonical)

    def step_nccl(i):
        kernel_only(i)
        if world > 1:
            dist.all_gather_into_tensor(gathered, sets[i % ring]["out"])

    fused, fused_note, fused_other = None, None, None
    if world > 1 and wl == "iiwa" and args.gather != "nccl":
        try:
            from rl_on_manifold_b200.sharding import SymmetricGather
            root = 0 if args.gather_to == "root" else None
            fused = SymmetricGather(B, n_out, deferred=(args.gather == "fused-deferred"), root=root)
            fused_note = fused.describe()
            # for comparison: the all-gather (every rank receives every row) when the default gathers to the learner,
            # else the other barrier schedule
            fused_other = (SymmetricGather(B, n_out, deferred=(args.gather == "fused-deferred")) if root is not None else
                           SymmetricGather(B, n_out, deferred=(args.gather != "fused-deferred")))
        except Exception as exc:
            fused_note = "fused gather unavailable (%s: %s); NCCL all-gather used" % (type(exc).__name__, exc)

    def step_fused(i):
        d = sets[i % ring]
        fused.step(d["q"], d["dq"], d["s"], d["alpha"], params, n_ctrl_joints=N_JOINTS, s_out=d["s_out"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(values):
        if world == 1:
            return list(values)
        t = torch.tensor(list(values), device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def timed_eager(fn):
        for i in range(args.warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            fn(args.warmup + i)
        e1.record()
        barrier()
        return max_over_ranks([e0.elapsed_time(e1)])[0]

    def timed_graphs(fn, finish=None, reset=None):
        """K steps per CUDA graph (n_graphs graphs over disjoint ring slots), R >= 5 timed replays cycling through
        them; every replay is exactly K steps between barrier + synchronize.  Returns the per-replay ms (max over
        ranks)."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(3):
                fn(i)
            if finish:
                finish()
        torch.cuda.current_stream().wait_stream(side)
        graphs = []
        for g in range(n_graphs):
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            if reset:
                reset()                                  # a captured step must not depend on events from outside
            with torch.cuda.graph(graph):
                for i in range(K):
                    fn(g * K + i)
                if finish:
                    finish()
            graphs.append(graph)
        if reset:
            reset()
        for graph in graphs:                             # warm-up replays
            graph.replay()
        replays = max(args.replays, n_graphs, 5)
        ms = []
        for r in range(replays):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graphs[r % n_graphs].replay()
            e1.record()
            barrier()
            ms.append(e0.elapsed_time(e1))
        return max_over_ranks(ms)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    # ---- the step as a user runs it: kernel (+ gather at N > 1)
    default_step = step_fused if fused is not None else step_nccl
    default_finish = fused.finish if fused is not None else None
    gather_mode = "single GPU" if world == 1 else (fused_note if fused is not None else
                                                   "one NCCL all-gather of the projected action per step")
    launches0 = _lib.launch_count()
    eager_ms = timed_eager(default_step)
    # kernels of this library per step (2 in the default mode of the step families: step kernel + fix-up), counted by
    # the library itself over the eager leg's warm-up + K steps
    kernels_per_step = max(1, round((_lib.launch_count() - launches0) / float(args.warmup + K)))
    launches_eager = kernels_per_step * K
    if default_finish:
        default_finish()
    if args.no_graph:
        replay_ms, launch_mode = [eager_ms], "eager launches from Python (ctypes)"
    else:
        replay_ms = timed_graphs(default_step, default_finish, fused.reset if fused is not None else None)
        launch_mode = "%d CUDA graph(s) of %d steps, %d timed replays, median" % (n_graphs, K, len(replay_ms))
        if world == 1 and wl != "point_reach" and os.environ.get("ATACOM_PDL", "1") != "0":
            launch_mode += "; programmatic dependent launch"
    total_ms, min_ms = statistics.median(replay_ms), min(replay_ms)
    launches = kernels_per_step * K * len(replay_ms) if not args.no_graph else launches_eager

    # ---- comparison paths at N > 1 (reported in config, never substituted for the default path)
    nccl_ms = kernel_ms_list = other_ms = None
    if world > 1:
        kernel_ms_list = [eager_ms] if args.no_graph else timed_graphs(kernel_only)
        if fused is not None:
            nccl_ms = statistics.median([timed_eager(step_nccl)] if args.no_graph else timed_graphs(step_nccl))
        if fused_other is not None and not args.no_graph:
            def step_other(i):
                d = sets[i % ring]
                fused_other.step(d["q"], d["dq"], d["s"], d["alpha"], params, n_ctrl_joints=N_JOINTS, s_out=d["s_out"])
            other_ms = statistics.median(timed_graphs(step_other, fused_other.finish, fused_other.reset))
    else:
        kernel_ms_list = replay_ms
    kernel_ms = statistics.median(kernel_ms_list)
    canonical_ms = None
    if wl in ("iiwa", "planar") and not args.no_graph:
        canonical_ms = statistics.median(timed_graphs(kernel_canonical)) / K
    # share of the batch the step kernel left to the LAPACK-basis fix-up launch
    redone = None
    if wl in ("iiwa", "planar"):
        d0 = sets[0]
        stt = torch.zeros(B, dtype=torch.uint8, device=dev)
        projection.step(fam, d0["q"], d0["dq"], d0["s"], d0["alpha"], params, n_ctrl_joints=N_JOINTS, ddq=d0["out"],
                        s_out=d0["s_out"], status=stt)
        redone = float(((stt & _lib.ST_LAPACK_PATH) != 0).float().mean())

    # ---- the fused gather delivers what the NCCL all-gather delivers (outside the timed region, every rank)
    gather_verified = None
    if world > 1 and fused is not None:
        d = sets[0]
        buf, _ = fused.step(d["q"], d["dq"], d["s"], d["alpha"], params, n_ctrl_joints=N_JOINTS, s_out=d["s_out"])
        fused.finish()
        torch.cuda.synchronize()
        got = buf.clone()
        step_nccl(0)
        torch.cuda.synchronize()
        # (gather to the learner: the root's buffer is the one that has to be complete)
        mine = fused.root is None or fused.root == rank
        okf = torch.tensor([1.0 if (not mine or torch.equal(got, gathered)) else 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(okf, op=dist.ReduceOp.MIN)
        gather_verified = bool(okf.item() == 1.0)

    # ---- end to end through the host-buffer entry point: pinned host arrays in, pinned host arrays out, every
    # byte of the step's inputs and outputs crosses PCIe inside the timed region.  Both data paths of the call are
    # timed and reported; `e2e.value` is the one HostContext picks by default for pinned buffers (mode "auto").
    # nvidia-smi polling every 100 ms slows driver calls and PCIe traffic measurably, so the sampler runs at 1 s here.
    if rank == 0:
        sampler.pause()
        sampler.start(1000)
    keys = ("q", "dq", "p", "dp", "s", "alpha") if wl == "point_reach" else ("q", "dq", "s", "alpha")
    host = {k_: sets[0][k_].cpu().pin_memory() for k_ in keys}
    out_h = torch.empty(B, n_out).pin_memory()
    s_h = torch.empty(B, host["s"].shape[1]).pin_memory()

    def time_e2e(mode):
        ctx = projection.HostContext(B, chunks=args.chunks, mode=mode)

        def e2e_step():
            if wl == "point_reach":
                ctx.point_reach_step(host["q"], host["dq"], host["p"], host["dp"], host["s"], host["alpha"], out_h, s_h,
                                     params)
            else:
                ctx.step(fam, host["q"], host["dq"], host["s"], host["alpha"], out_h, s_h, params,
                         n_ctrl_joints=N_JOINTS)
            return float(out_h[0, 0])                                # read the result on the host

        for _ in range(max(3, args.warmup)):
            e2e_step()
        reps = []
        for _ in range(5):
            barrier()
            t0 = time.perf_counter()
            for _ in range(K):
                e2e_step()
            torch.cuda.synchronize()
            reps.append(1e3 * (time.perf_counter() - t0))
        ctx.close()
        return statistics.median(max_over_ranks(reps))

    e2e_staged_ms = time_e2e("staged")
    e2e_zero_copy_ms = None if wl == "point_reach" else time_e2e("zero_copy")
    auto_path = "staged" if wl == "point_reach" else projection.HostContext.auto_path(fam, params, N_JOINTS)
    e2e_auto_ms = e2e_staged_ms if auto_path == "staged" else e2e_zero_copy_ms
    e2e_api = ("atacom_%s_step_host, pinned host buffers, mode auto = " % wl) + (
        "staged copies (CUDA graph of H2D, kernels, D2H per chunk)" if auto_path == "staged" else
        "zero-copy (one launch on the caller's buffers, loads/stores cross PCIe via cp.async.bulk; with the LAPACK "
        "basis both passes run inside that launch)")

    clocks = sampler.stop() if rank == 0 else None
    timeouts = _lib.spin_timeouts()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        k_ms = kernel_ms / K
        achieved = bytes_env * B / (k_ms * 1e-3) / 1e9
        # DRAM traffic of the dominant kernel: from the committed ncu capture of this workload (written by
        # profiles/summarize_ncu.py from the .ncu-rep), null when there is none for it
        traffic, traffic_src = None, None
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
            ent = t.get(wl)
            if ent and int(ent.get("batch", 0)) == B:
                traffic, traffic_src = ent["bytes_per_launch"], ent.get("source")
        except Exception:
            pass
        total_envs = world * B * K
        line = dict(
            metric=METRIC, value=total_envs / (total_ms * 1e-3), unit=UNIT, n_gpus=world,
            steps=K, warmup=args.warmup, ms_per_step=total_ms / K, higher_is_better=True,
            scaling=args.scaling, vs_baseline=None,
            dtype="f64 (kinematics and projection; fp32 I/O)", data="synthetic",
            config=dict(workload=workload_label(wl, B, world, args.scaling), batch_per_gpu=B, global_batch=world * B,
                        parallelism="env-shard x%d; %s" % (world, gather_mode) if world > 1 else "single GPU",
                        gather_verified=gather_verified,
                        nccl_all_gather_ms_per_step=(nccl_ms / K) if nccl_ms else None,
                        other_fused_schedule_ms_per_step=(other_ms / K) if other_ms else None,
                        other_fused_schedule=fused_other.describe() if (fused_other is not None and other_ms) else None,
                        kernels_only_ms_per_step=(kernel_ms / K) if world > 1 else None,
                        cold_inputs="ring of %d distinct batches (%.0f MB > 2x L2) rotated every step"
                                    % (ring, ring * B * bytes_env / 1e6),
                        bytes_per_env_step=bytes_env, launch=launch_mode, replays=len(replay_ms),
                        min_ms_per_step=min_ms / K, value_at_min=total_envs / (min_ms * 1e-3),
                        eager_ms_per_step=eager_ms / K, spin_timeouts=timeouts,
                        null_basis="LAPACK gesdd's own (reference-exact on both strata): step kernel + compacted fix-up "
                                   "launch for the environments inside rref's tolerance band" if redone is not None
                                   else "basis-free (k = 1 / tolerance 1e-15)",
                        share_redone_with_lapack_basis=redone, canonical_basis_kernel_ms_per_step=canonical_ms),
            e2e=dict(value=world * B * K / (e2e_auto_ms * 1e-3), unit=UNIT, h2d_bytes_per_step=B * bytes_in,
                     d2h_bytes_per_step=B * bytes_out, api=e2e_api,
                     staged_value=world * B * K / (e2e_staged_ms * 1e-3), staged_chunks=args.chunks,
                     zero_copy_value=(world * B * K / (e2e_zero_copy_ms * 1e-3)) if e2e_zero_copy_ms else None,
                     timing="median of 5 repeats of %d calls, wall clock around the calls, max over ranks" % K),
            gpu_launches=int(launches),
            roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                          traffic=traffic, traffic_source=traffic_src, peak_source=peak_src,
                          kernel=W["kernel"] + (" + atacom_fix_kernel" if redone is not None else ""),
                          kernel_us=k_ms * 1e3, frac_of_8TBs_nominal=achieved / 8000.0,
                          note="nominal bound: the kernel is FP64-issue and latency bound, not bandwidth bound "
                               "(DESIGN.md section 6)"),
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
    ap.add_argument("--workload", default="iiwa", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--batch", type=int, default=0, help="environments per GPU (weak) / in total (strong); 0 = the workload's")
    ap.add_argument("--chunks", type=int, default=2)
    ap.add_argument("--replays", type=int, default=7)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--gather-to", default="root", choices=["root", "all"],
                    help="N > 1: the projected actions are gathered on rank 0 (the learner; north_star: one gather per "
                         "roll-out step) or on every rank (all-gather); the other one is timed beside it")
    ap.add_argument("--gather", default="fused-deferred", choices=["fused", "fused-deferred", "nccl"],
                    help="N > 1: fused peer-store epilogue with the barrier of step t deferred behind kernel t+1 (default), "
                         "with the barrier on the kernel's stream, or a plain NCCL all-gather")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
