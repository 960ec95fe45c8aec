"""N > 1 host logic on CPU: world_size-2 gloo processes shard the environment batch, run the device
algebra on their shard (host build of the device headers) and all-gather ddq; the result must be
bit-identical to the unsharded run (no cross-environment arithmetic anywhere)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rl_on_manifold_b200 import _lib
from rl_on_manifold_b200.sharding import EnvShard, shard_bounds
from tests import helpers


def test_shard_bounds_cover_range():
    for B in (0, 1, 7, 64, 65536, 65537):
        for W in (1, 2, 3, 4, 8):
            spans = [shard_bounds(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q, dq, s, alpha = helpers.synthetic_cpu("circle", B, seed=5)
    shard = EnvShard(B)
    assert (shard.rank, shard.world) == (rank, world)
    h = helpers.load_harness()
    p = _lib.default_params("circle")
    sl = slice(shard.lo, shard.hi)
    ddq, s_out, _, _ = helpers.harness_step(h, "circle", p.flat(), q[sl], dq[sl], s[sl], alpha[sl], np.float32)
    full = shard.gather(torch.from_numpy(ddq))
    np.save(os.path.join(out_dir, "ddq_%d.npy" % rank), full.numpy())
    np.save(os.path.join(out_dir, "s_%d.npy" % rank), s_out)
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [64, 77])
def test_two_rank_gloo_shard_equals_unsharded(tmp_path, B):
    helpers.load_harness()                               # build once before forking
    world = 2
    mp.start_processes(_worker, args=(world, _free_port(), B, str(tmp_path)), nprocs=world, start_method="fork")
    q, dq, s, alpha = helpers.synthetic_cpu("circle", B, seed=5)
    h = helpers.load_harness()
    ddq, s_out, _, _ = helpers.harness_step(h, "circle", _lib.default_params("circle").flat(), q, dq, s, alpha,
                                            np.float32)
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / ("ddq_%d.npy" % r)), ddq)
    s_cat = np.concatenate([np.load(tmp_path / ("s_%d.npy" % r)) for r in range(world)])
    np.testing.assert_array_equal(s_cat, s_out)
