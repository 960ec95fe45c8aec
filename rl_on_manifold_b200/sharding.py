"""Multi-GPU layout of the environment batch: independent shards, one gather per roll-out step.

The projection has no cross-environment term (atacom.py:123-139), so rank r of W owns the contiguous
block [lo_r, hi_r) of the environment index and its slices of q, dq, s, alpha, ddq; the slack state `s`
never leaves its rank.  The only exchange is one all-gather of the projected accelerations
ddq[B/W, n] per step, issued on the stream the kernel ran on."""
import ctypes

import torch
import torch.distributed as dist


def shard_bounds(global_B, rank, world):
    """Contiguous block of rank `rank`: sizes differ by at most one when W does not divide B."""
    base, rem = divmod(global_B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_constraint_logs(total, count, c_max, c_dq_max, group=None):
    """Cross-rank merge of one epoch's constraint log (atacom.py:207-216 over the GLOBAL batch): SUM on (sum of
    c_i, number of samples), MAX on (max c_i, max c_dq_i) — two small all-reduces per epoch, nothing per step.
    Returns (total, count, c_max, c_dq_max) of the whole job; a no-op without an initialised process group."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return total, count, c_max, c_dq_max
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    sums = torch.tensor([float(total), float(count)], dtype=torch.float64, device=dev)
    maxs = torch.tensor([float(c_max), float(c_dq_max)], dtype=torch.float64, device=dev)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX, group=group)
    s, m = sums.tolist(), maxs.tolist()
    return s[0], int(round(s[1])), m[0], m[1]


def reduce_stats(stats, group=None):
    """The same for the kernels' device accumulators `stats` [4] = { sum, max c, max c_dq, count }
    (projection.new_stats; atacom_*_constraint_stats, fused roll-outs): reduced in place across the ranks."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return stats
    t = stats if (stats.is_cuda == (dist.get_backend(group) == "nccl")) else stats.cpu()
    sums, maxs = t[[0, 3]].clone(), t[[1, 2]].clone()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX, group=group)
    stats[0], stats[3], stats[1], stats[2] = sums[0], sums[1], maxs[0], maxs[1]
    return stats


class EnvShard:
    def __init__(self, global_B, rank=None, world=None, group=None):
        self.group = group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.global_B = global_B
        self.lo, self.hi = shard_bounds(global_B, self.rank, self.world)
        self.max_local = -(-global_B // self.world)

    @property
    def local_B(self):
        return self.hi - self.lo

    def take(self, t):
        """This rank's rows of a [global_B, ...] tensor."""
        return t[self.lo:self.hi].contiguous()

    def gather(self, local, out=None):
        """All-gather [local_B, d] rows into [global_B, d] (every rank gets the full array)."""
        if self.world == 1:
            return local if out is None else out.copy_(local)
        d = local.shape[1:]
        if out is None:
            out = torch.empty((self.global_B,) + tuple(d), dtype=local.dtype, device=local.device)
        if self.global_B % self.world == 0:
            dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
            return out
        pad = torch.zeros((self.max_local,) + tuple(d), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(parts, pad, group=self.group)
        for r, part in enumerate(parts):
            lo, hi = shard_bounds(self.global_B, r, self.world)
            out[lo:hi] = part[:hi - lo]
        return out


class SymmetricGather:
    """Gather targets for the fused epilogue of `atacom_iiwa_step_gather`.

    Each rank owns `buffers` [world * local_B, n] fp32 tensors in NVLink symmetric memory
    (torch.distributed._symmetric_memory); after the rendezvous every rank holds peer-mapped pointers to all
    of them.  The step kernel stores its shard's rows straight into every rank's buffer (P2P stores over
    NVLink / NVSwitch), so the per-step all-gather costs no extra launch and no extra pass over HBM; one
    device-side barrier per step then publishes the step.  Buffers rotate between steps so that a rank still
    reading step t cannot be overwritten by a peer already writing a later step.

    Two schedules:
      * synchronous (`deferred=False`, 2 buffers): kernel_t, barrier_t on the kernel's stream — the gathered
        rows of step t are complete before step t + 1 starts; the barrier launch (~5 us) and the NVLink transfer
        of step t sit between two kernels;
      * deferred (`deferred=True`, 3 buffers, default of bench.py): barrier_t runs on a side stream behind
        kernel_t, while kernel_{t+1} already computes on the main stream — the barrier and the transfer of step
        t are hidden behind the compute of step t + 1.  The consumer of step t (a central learner's replay
        buffer: nothing in the projection of step t + 1 depends on other ranks' rows) waits on `ready(t)`;
        kernel_{t+3}, which reuses buffer t mod 3, waits for barrier_{t+1}: by then every rank has passed its
        own barrier_{t+1} launch, which is stream-ordered behind whatever consumed step t on the side stream.
    """

    def __init__(self, local_B, n, group=None, buffers=None, in_kernel_barrier=False, deferred=False, root=None):
        """root: None — every rank receives every rank's rows (an all-gather); a rank number — only that rank does (a
        gather to the learner, what SURVEY.md §8e asks for: "one gather over NVLink per roll-out step").  Every rank
        then stores its rows into the root's buffer alone: 1/world of the NVLink egress of the all-gather, and the
        buffer `step` returns is complete on the root only."""
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise ValueError("fused gather supports one NVSwitch domain (<= 8 ranks)")
        if root is not None and (in_kernel_barrier or not 0 <= root < self.world):
            raise ValueError("root must be a rank of the group (and the in-kernel barrier gathers to every rank)")
        self.local_B, self.n = local_B, n
        self.deferred = bool(deferred) and not in_kernel_barrier
        buffers = buffers if buffers is not None else (3 if self.deferred else 2)
        if self.deferred and buffers < 3:
            raise ValueError("the deferred schedule needs at least 3 buffers")
        name = group.group_name
        dev = torch.device("cuda", torch.cuda.current_device())
        self.bufs, self.handles, self.ptrs = [], [], []
        for _ in range(buffers):
            t = symm_mem.empty(self.world * local_B, n, dtype=torch.float32, device=dev)
            hdl = symm_mem.rendezvous(t, name)
            self.bufs.append(t)
            self.handles.append(hdl)
            targets = [int(p) for p in hdl.buffer_ptrs] if root is None else [int(hdl.buffer_ptrs[root])]
            self.ptrs.append((ctypes.c_void_p * len(targets))(*targets))
        self.root = root
        self.n_targets = self.world if root is None else 1
        # Optional in-kernel barrier: a flag array [world] per rank in symmetric memory and two local counters; the
        # last block of every launch publishes / awaits the step, so no separate barrier launch is needed.
        # Measured at N = 2 it is slower than the separate barrier kernel (29.5 vs 25.1 us per step: the system-
        # scope fences sit on the critical path of every block), hence off by default.
        self.in_kernel_barrier = in_kernel_barrier
        if in_kernel_barrier:
            self._flags = symm_mem.empty(64, dtype=torch.int32, device=dev)
            self._flags.zero_()
            self._flag_hdl = symm_mem.rendezvous(self._flags, name)
            self._flag_ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in self._flag_hdl.buffer_ptrs])
            self._local_sync = torch.zeros(2, dtype=torch.int32, device=dev)
            torch.cuda.synchronize()
            self._flag_hdl.barrier()          # every rank's flags are zero before the first step
        self._i = 0
        self._side = torch.cuda.Stream(device=dev) if self.deferred else None
        self._bar_done = [None] * buffers     # event: the barrier of the last step that used buffer b has completed
        self._last = None                     # (buffer index, event) of the newest step

    def describe(self):
        what = self._describe()
        if self.root is not None:
            what = what.replace("fused peer-store epilogue", "fused store epilogue into rank %d's buffer (gather to the "
                                "learner)" % self.root)
        return what

    def _describe(self):
        if self.in_kernel_barrier:
            return "fused peer-store epilogue over NVLink symmetric memory, cross-rank barrier inside the kernel"
        if self.deferred:
            return ("fused peer-store epilogue over NVLink symmetric memory; the barrier of step t runs on a side "
                    "stream behind kernel t+1 (%d rotating buffers; gathered rows of step t are ready one step later)"
                    % len(self.bufs))
        return "fused peer-store epilogue over NVLink symmetric memory + one device barrier per step"

    def step(self, q, dq, s, alpha, params, *, n_ctrl_joints=6, s_out=None, status=None):
        """Project this rank's shard and gather: returns (gathered ddq [world * local_B, n], s_out).  With the
        deferred schedule the returned buffer is complete once `ready()` (or `finish()`) has been waited on."""
        from . import projection
        nb = len(self.bufs)
        b = self._i % nb
        t = self._i
        self._i += 1
        main = torch.cuda.current_stream()
        with torch.cuda.nvtx.range("atacom_step_gather"):
            if self.in_kernel_barrier:
                s_out = projection.iiwa_step_gather(q, dq, s, alpha, params, self.ptrs[b], self.n_targets,
                                                    self.rank * self.local_B, n_ctrl_joints=n_ctrl_joints, s_out=s_out,
                                                    status=status, flag_ptrs=self._flag_ptrs,
                                                    local_sync=self._local_sync, rank=self.rank)
                return self.bufs[b], s_out
            if self.deferred:
                # buffer b was last written at step t - nb; every rank is done with it once barrier_{t-nb+1} ... has
                # completed here — wait for the newest barrier older than two steps (it completed long ago)
                prev = self._bar_done[(t - 2) % nb] if t >= 2 else None
                if prev is not None:
                    main.wait_event(prev)
            s_out = projection.iiwa_step_gather(q, dq, s, alpha, params, self.ptrs[b], self.n_targets,
                                                self.rank * self.local_B, n_ctrl_joints=n_ctrl_joints, s_out=s_out,
                                                status=status)
            if not self.deferred:
                self.handles[b].barrier()
                return self.bufs[b], s_out
            done_k = torch.cuda.Event()
            done_k.record(main)
            self._side.wait_event(done_k)
            with torch.cuda.stream(self._side):
                self.handles[b].barrier()
                ev = torch.cuda.Event()
                ev.record(self._side)
            self._bar_done[b] = ev
            self._last = (b, ev)
        return self.bufs[b], s_out

    def ready(self, stream=None):
        """Make `stream` (default: the current one) wait until the newest step's gathered rows are complete."""
        if self.deferred and self._last is not None:
            (stream or torch.cuda.current_stream()).wait_event(self._last[1])

    def finish(self):
        """Join the side stream: after this, on the current stream, every issued step is gathered everywhere."""
        if self.deferred:
            torch.cuda.current_stream().wait_stream(self._side)

    def reset(self):
        """Forget the bookkeeping of earlier steps (call after `finish()`, e.g. before capturing a CUDA graph: a
        captured step must not wait on an event recorded outside the capture)."""
        self._i = 0
        self._bar_done = [None] * len(self.bufs)
        self._last = None
