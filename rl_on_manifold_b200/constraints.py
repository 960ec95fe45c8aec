"""ViabilityConstraint / ConstraintsSet with the reference's constructors (atacom/constraints.py:11-31,
46-58), for BATCHED callbacks.

The reference's callbacks are per-environment Python callables fun(q)->[m], J(q)->[m,n], b(q,dq)->[m];
calling them once per environment cannot feed 65 536 environments.  Here a callback is either
  * a batched tensor callable  fun(q[B,n]) -> [B,m],  J(q[B,n]) -> [B,m,n],  b(q[B,n], dq[B,n]) -> [B,m]
    (evaluated with torch on the GPU, then projected by the generic kernel), or
  * absent, when the owning ConstraintsSet names a built-in device functor (`family`), in which case the
    fused per-family kernel evaluates the constraints itself in registers.
"""
import numpy as np
import torch


class ViabilityConstraint:
    """f(q) + K_f df(q, dq) = 0  /  g(q) + K_g dg(q, dq) <= 0   (constraints.py:4-9)."""

    def __init__(self, dim_q, dim_out, fun=None, J=None, b=None, K=1.0):
        self.dim_q = dim_q
        self.dim_out = dim_out
        self.fun_origin = fun
        self.J = J
        self.b_state = b
        self.K = np.ones(dim_out) * K if np.isscalar(K) else np.asarray(K, dtype=np.float64)
        assert self.K.shape[0] == dim_out

    @property
    def has_callbacks(self):
        return self.fun_origin is not None and self.J is not None and self.b_state is not None

    def _need_callbacks(self):
        if not self.has_callbacks:
            raise ValueError("this ViabilityConstraint has no fun / J / b callbacks: it belongs to a ConstraintsSet "
                             "with a built-in device functor (family=...), whose rows the fused kernels evaluate")

    # batched versions of constraints.py:33-43
    def fun(self, q, dq, origin_constr=False):
        self._need_callbacks()
        c = self.fun_origin(q)
        if origin_constr:
            return c
        K = torch.as_tensor(self.K, dtype=q.dtype, device=q.device)
        return c + K * torch.einsum("bmn,bn->bm", self.J(q), dq)

    def K_J(self, q):
        self._need_callbacks()
        K = torch.as_tensor(self.K, dtype=q.dtype, device=q.device)
        return K[None, :, None] * self.J(q)

    def b(self, q, dq):
        self._need_callbacks()
        K = torch.as_tensor(self.K, dtype=q.dtype, device=q.device)
        return torch.einsum("bmn,bn->bm", self.J(q), dq) + K * self.b_state(q, dq)


class ConstraintsSet:
    """Gathers several constraints (constraints.py:46-82).  `family` names a built-in device functor
    ("circle", "planar", "iiwa") whose rows the fused kernels evaluate themselves."""

    def __init__(self, dim_q, family=None):
        self.dim_q = dim_q
        self.constraints_list = list()
        self.dim_out = 0
        self.family = family

    @staticmethod
    def wrap(c, dim_q):
        """A bare ViabilityConstraint as a one-element set (the reference takes either, atacom.py:18-19)."""
        if c is None or isinstance(c, ConstraintsSet):
            return c
        cs = ConstraintsSet(dim_q, family=getattr(c, "family", None))
        cs.add_constraint(c)
        return cs

    def add_constraint(self, c: ViabilityConstraint):
        assert c.dim_q == self.dim_q
        self.dim_out += c.dim_out
        self.constraints_list.append(c)

    @property
    def K(self):
        return np.concatenate([c.K for c in self.constraints_list]) if self.constraints_list else np.zeros(0)

    @property
    def has_callbacks(self):
        return all(c.has_callbacks for c in self.constraints_list)

    def raw(self, q, dq):
        """(c [B,m], J [B,m,n], b_state [B,m]) stacked over the constraints: the generic kernel's inputs."""
        cs = [c.fun_origin(q) for c in self.constraints_list]
        Js = [c.J(q) for c in self.constraints_list]
        bs = [c.b_state(q, dq) for c in self.constraints_list]
        return torch.cat(cs, 1), torch.cat(Js, 1), torch.cat(bs, 1)

    def fun(self, q, dq, origin_constr=False):
        return torch.cat([c.fun(q, dq, origin_constr) for c in self.constraints_list], 1)

    def K_J(self, q):
        return torch.cat([c.K_J(q) for c in self.constraints_list], 1)

    def b(self, q, dq):
        return torch.cat([c.b(q, dq) for c in self.constraints_list], 1)
