"""CPU oracle for the ATACOM tangent-space projection step.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`rl_on_manifold_b200/`) imports this; only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` do, and only as the checker / timed CPU baseline.

What it is: a fresh float64 NumPy restatement of the reference's hot path
(`atacom/atacom.py:123-139` and helpers), one environment at a time, the way
the reference runs it.  Every function cites the reference file:line it
restates.

Pinning status (see DESIGN.md "Oracle"):
  * generic algebra, Circle (A / E) and Collision (C) paths: PINNED — checked
    against the reference's own classes imported from /root/reference under
    import shims (`oracle/ref_loader.py`, in-container only) and against the
    golden vectors in `tests/golden/` generated from them.
  * planar-3R and iiwa constraint callbacks: PARITY UNPINNED — the reference
    evaluates them through pinocchio, which is not installed here and is not
    vendored; the restatement follows the URDF chain and is pinned only by
    finite differences and the FK facts recorded in SURVEY.md §8c.
"""
