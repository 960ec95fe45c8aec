"""Run the per-environment oracle over a large batch on all host cores, in a process of its own.

    python -m tests.oracle_mp in.npz out.npz

Test infrastructure.  The GPU parity tests at BASELINE.json's full batch sizes call this through
`helpers.oracle_batch_mp`: a separate interpreter (no CUDA context to fork) whose Pool fans the batch out over
os.cpu_count() workers.  in.npz: family, basis, variant, bias and the input arrays; out.npz: what
helpers.oracle_batch returns.
"""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _chunk(task):
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:                                         # pragma: no cover
        pass
    from tests import helpers
    family, arrays, kw = task
    if family == "point_reach":
        return helpers.oracle_point_reach_batch(*arrays)
    return helpers.oracle_batch(family, *arrays, **kw)


def main(src, dst):
    z = np.load(src, allow_pickle=False)
    family = str(z["family"])
    names = ("q", "dq", "p", "dp", "s", "action") if family == "point_reach" else ("q", "dq", "s", "alpha")
    arrays = [z[k] for k in names]
    kw = {} if family == "point_reach" else dict(basis=str(z["basis"]), variant=str(z["variant"]), bias=str(z["bias"]))
    B = arrays[0].shape[0]
    procs = max(1, min(os.cpu_count() or 1, B // 64 or 1))
    idx = np.array_split(np.arange(B), procs * 4)
    with mp.get_context("fork").Pool(procs) as pool:
        parts = pool.map(_chunk, [(family, [a[i] for a in arrays], kw) for i in idx if len(i)])
    out = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    np.savez(dst, **out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
