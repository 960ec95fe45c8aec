"""Two-GPU check of the fused gather epilogue (atacom_iiwa_step_gather + NVLink symmetric memory):
each rank projects its shard and stores it into every rank's gather buffer; the gathered array must be
bit-identical to the unsharded single-GPU result and to the NCCL all-gather path.
Skipped on a single-GPU box (run with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, out):
    import torch.distributed as dist
    from rl_on_manifold_b200 import _lib, projection, synthetic
    from rl_on_manifold_b200.sharding import EnvShard, SymmetricGather
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    params = _lib.default_params("iiwa", 6)
    gB = world * B
    q, dq, alpha = (t.to(dev) for t in synthetic.state_batch("iiwa", gB, 99, 6, params))
    s = synthetic.slack_mix(projection.slack_init("iiwa", q, dq, params), 99)
    full_ddq, full_s = projection.step("iiwa", q, dq, s, alpha, params)          # unsharded reference
    shard = EnvShard(gB)
    sq, sdq, ss, sal = (shard.take(t) for t in (q, dq, s, alpha))
    ddq_l, s_l = projection.step("iiwa", sq, sdq, ss, sal, params)
    nccl = shard.gather(ddq_l)
    ok = True
    # cross-rank barrier inside the kernel / as a separate launch on the kernel's stream / deferred to a side stream
    for kw in (dict(in_kernel_barrier=True), dict(), dict(deferred=True)):
        fg = SymmetricGather(B, 6, **kw)
        if kw.get("in_kernel_barrier"):
            # the barrier inside the step kernel would publish a step before the fix-up launch of the default
            # (LAPACK-basis) mode has written its rows: this variant exists for the canonical mode only
            pk = params.copy()
            pk.basis_mode = _lib.BASIS_CANONICAL
            want_ddq, want_s = projection.step("iiwa", q, dq, s, alpha, pk)
            for it in range(5):
                gathered, s_f = fg.step(sq, sdq, ss, sal, pk)
                ok = ok and torch.equal(gathered, want_ddq) and torch.equal(s_f, want_s[shard.lo:shard.hi])
            torch.cuda.synchronize()
            continue
        for it in range(7):                                                   # exercises every buffer twice
            gathered, s_f = fg.step(sq, sdq, ss, sal, params)
            # no host synchronisation between launch and check: waiting on the step's own completion (the stream
            # itself, or ready() for the deferred schedule) must imply that every rank's rows have arrived
            fg.ready()
            ok = ok and torch.equal(gathered, full_ddq) and torch.equal(gathered, nccl)
            ok = ok and torch.equal(s_f, full_s[shard.lo:shard.hi])
            if rank == it % world:
                torch.cuda._sleep(2_000_000)                                  # skew the ranks
        fg.finish()
        torch.cuda.synchronize()
    # deferred schedule, steps issued back to back without consuming (what bench.py times): buffers rotate, the
    # last three steps' buffers must hold complete rows after finish()
    fg = SymmetricGather(B, 6, deferred=True)
    outs = [fg.step(sq, sdq, ss, sal, params)[0] for _ in range(9)]
    fg.finish()
    torch.cuda.synchronize()
    ok = ok and all(torch.equal(o, full_ddq) for o in outs[-3:])
    # gather to the learner (rank `root`) instead of to everyone: the root's buffer is complete, nobody else's is written
    for root in (0, world - 1):
        fg = SymmetricGather(B, 6, deferred=True, root=root)
        for it in range(4):
            gathered, s_f = fg.step(sq, sdq, ss, sal, params)
            fg.ready()
            if rank == root:
                ok = ok and torch.equal(gathered, full_ddq)
            ok = ok and torch.equal(s_f, full_s[shard.lo:shard.hi])
        fg.finish()
        torch.cuda.synchronize()
    ok = ok and _lib.spin_timeouts() == 0
    torch.save(dict(ok=bool(ok)), os.path.join(out, "r%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_gather_equals_unsharded(tmp_path, world):
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (run with `gpurun --gpus %d`)" % (world, world))
    mp.spawn(_worker, args=(world, _free_port(), 4096, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert torch.load(tmp_path / ("r%d.pt" % r))["ok"]


def test_step_on_a_second_device_in_one_process():
    """The shared-memory opt-in of the step kernels belongs to the (kernel, device) pair: a process that ran on
    cuda:0 must be able to run on cuda:1 afterwards (and back)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from rl_on_manifold_b200 import _lib, projection, synthetic
    p = _lib.default_params("iiwa", 6)
    outs = []
    for d in (0, 1, 0):
        dev = torch.device("cuda", d)
        q, dq, s, alpha = synthetic.device_batch("iiwa", 3000, 5, dev, 6, p)
        ddq, s_out = projection.step("iiwa", q, dq, s, alpha, p)
        ddq2, _ = projection.iiwa_substeps(q, dq, s, alpha, p, 2)
        torch.cuda.synchronize(dev)
        outs.append((ddq.cpu(), s_out.cpu(), ddq2.cpu()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
