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
    for in_kernel in (True, False):            # cross-rank barrier inside the kernel / as a separate launch
        fg = SymmetricGather(B, 6, in_kernel_barrier=in_kernel)
        for it in range(5):                                                   # exercises both buffers
            gathered, s_f = fg.step(sq, sdq, ss, sal, params)
            # no host synchronisation between launch and check: kernel completion alone must imply that every
            # rank's rows have arrived (the comparison kernels run on the same stream)
            ok = ok and torch.equal(gathered, full_ddq) and torch.equal(gathered, nccl)
            ok = ok and torch.equal(s_f, full_s[shard.lo:shard.hi])
            if rank == it % world:
                torch.cuda._sleep(2_000_000)                                  # skew the ranks
        torch.cuda.synchronize()
    torch.save(dict(ok=bool(ok)), os.path.join(out, "r%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_gather_equals_unsharded(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), 4096, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert torch.load(tmp_path / ("r%d.pt" % r))["ok"]
