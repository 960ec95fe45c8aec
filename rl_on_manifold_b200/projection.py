"""Batched ATACOM projection step on CUDA tensors — thin typed layer over the C ABI.

Every function takes contiguous float32 CUDA tensors shaped [B, dim] and launches
one kernel of libatacom_b200.so on the current torch CUDA stream.  CPU tensors are
rejected: there is no CPU fallback (the NumPy oracle lives in `oracle/` and is test
infrastructure).
"""
import ctypes
import math

import torch

from . import _lib

FAMILY_DIMS = {
    # family: (n, F, G)   reference: circle_atacom.py:9-18, atacom_air_hockey.py:28-43, iiwa_hit_atacom.py:23-40
    "circle": (2, 1, 1),
    "planar": (3, 0, 6),
    "iiwa6": (6, 1, 11),
    "iiwa7": (7, 1, 12),
}


def family_dims(family, n_ctrl_joints=None):
    if family == "iiwa":
        family = "iiwa%d" % (n_ctrl_joints or 6)
    return FAMILY_DIMS[family]


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _check(t, name, B, dim, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (no CPU fallback)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    want = (B, dim) if dim is not None else (B,)
    if tuple(t.shape) != want:
        raise ValueError("%s: expected shape %s, got %s" % (name, want, tuple(t.shape)))


def _stream(q):
    return ctypes.c_void_p(torch.cuda.current_stream(q.device).cuda_stream)


def _alpha_dim(params, n, F):
    return n if params.variant == _lib.VARIANT_ERROR_CORRECTION else n - F


def step(family, q, dq, s, alpha, params, *, n_ctrl_joints=6, ddq=None, s_out=None, status=None, w_dbg=None):
    """AtacomEnvWrapper.step_action_function for B environments (atacom.py:123-139).

    Returns (ddq, s_out).  `s_out` may be `s` itself for the reference's in-place update."""
    n, F, G = family_dims(family, n_ctrl_joints)
    B = q.shape[0]
    _check(q, "q", B, n)
    _check(dq, "dq", B, n)
    _check(s, "s", B, G)
    _check(alpha, "alpha", B, _alpha_dim(params, n, F))
    if ddq is None:
        ddq = torch.empty_like(q)
    if s_out is None:
        s_out = torch.empty_like(s)
    _check(ddq, "ddq", B, n)
    _check(s_out, "s_out", B, G)
    if status is not None:
        _check(status, "status", B, None, torch.uint8)
    if w_dbg is not None:
        _check(w_dbg, "w_dbg", B, 2 * (n + G))
    with torch.cuda.device(q.device):
        common = (_ptr(q), _ptr(dq), _ptr(s), _ptr(alpha), _ptr(ddq), _ptr(s_out), _ptr(status), _ptr(w_dbg),
                  B, ctypes.byref(params), _stream(q))
        if family == "circle":
            rc = _lib.lib.atacom_circle_step(*common)
        elif family == "planar":
            rc = _lib.lib.atacom_planar_step(*common)
        elif family.startswith("iiwa"):
            rc = _lib.lib.atacom_iiwa_step(n, *common)
        else:
            raise ValueError(family)
    _lib.check(rc)
    return ddq, s_out


def iiwa_step_gather(q, dq, s, alpha, params, peer_ptrs, world, row_offset, *, n_ctrl_joints=6, ddq=None,
                     s_out=None, status=None, flag_ptrs=None, local_sync=None, rank=0):
    """`step("iiwa", ...)` with the multi-GPU gather fused into the kernel epilogue: every environment's ddq row
    is stored into each rank's [world * B, n] gather buffer (`peer_ptrs`: ctypes array of `world` device pointers,
    peer-mapped; see sharding.SymmetricGather).  Returns s_out."""
    n, F, G = family_dims("iiwa", n_ctrl_joints)
    B = q.shape[0]
    _check(q, "q", B, n)
    _check(dq, "dq", B, n)
    _check(s, "s", B, G)
    _check(alpha, "alpha", B, _alpha_dim(params, n, F))
    if s_out is None:
        s_out = torch.empty_like(s)
    _check(s_out, "s_out", B, G)
    if ddq is not None:
        _check(ddq, "ddq", B, n)
    if status is not None:
        _check(status, "status", B, None, torch.uint8)
    with torch.cuda.device(q.device):
        if flag_ptrs is not None:       # cross-rank barrier inside the kernel as well
            rc = _lib.lib.atacom_iiwa_step_gather_sync(n, _ptr(q), _ptr(dq), _ptr(s), _ptr(alpha), _ptr(ddq),
                                                       _ptr(s_out), _ptr(status), B, ctypes.byref(params),
                                                       _stream(q), peer_ptrs, world, row_offset, flag_ptrs,
                                                       _ptr(local_sync), rank)
        else:
            rc = _lib.lib.atacom_iiwa_step_gather(n, _ptr(q), _ptr(dq), _ptr(s), _ptr(alpha), _ptr(ddq),
                                                  _ptr(s_out), _ptr(status), B, ctypes.byref(params), _stream(q),
                                                  peer_ptrs, world, row_offset)
    _lib.check(rc)
    return s_out


def iiwa_substeps(q, dq, s, alpha, params, K, *, n_ctrl_joints=6, ddq=None, s_out=None, status=None,
                  workspace=None, use_workspace=True):
    """K calls of step_action_function on the same (q, dq, alpha) in one launch — the K simulator sub-steps of one
    agent step of a PyBullet-style environment (env_base.py:161-165).  Returns (ddq [K, B, n], s_out [B, G])."""
    n, F, G = family_dims("iiwa", n_ctrl_joints)
    B = q.shape[0]
    _check(q, "q", B, n)
    _check(dq, "dq", B, n)
    _check(s, "s", B, G)
    _check(alpha, "alpha", B, _alpha_dim(params, n, F))
    if ddq is None:
        ddq = torch.empty(K, B, n, device=q.device, dtype=torch.float32)
    if tuple(ddq.shape) != (K, B, n) or not ddq.is_contiguous() or not ddq.is_cuda or ddq.dtype != torch.float32:
        raise ValueError("ddq: expected a contiguous CUDA float32 tensor of shape [%d, %d, %d]" % (K, B, n))
    if s_out is None:
        s_out = torch.empty_like(s)
    _check(s_out, "s_out", B, G)
    if status is not None:
        _check(status, "status", B, None, torch.uint8)
    if workspace is None and use_workspace and K > 1:
        workspace = torch.empty(_lib.lib.atacom_iiwa_substeps_workspace_doubles(n) * B, device=q.device,
                                dtype=torch.float64)
    with torch.cuda.device(q.device):
        rc = _lib.lib.atacom_iiwa_step_substeps(n, K, _ptr(q), _ptr(dq), _ptr(s), _ptr(alpha), _ptr(ddq), _ptr(s_out),
                                                _ptr(status), _ptr(workspace), B, ctypes.byref(params), _stream(q))
    _lib.check(rc)
    return ddq, s_out


def slack_init(family, q, dq, params, *, n_ctrl_joints=6, s=None, mask=None):
    """AtacomEnvWrapper._compute_slack_variables (atacom.py:145-149)."""
    n, F, G = family_dims(family, n_ctrl_joints)
    B = q.shape[0]
    _check(q, "q", B, n)
    _check(dq, "dq", B, n)
    if s is None:
        s = torch.zeros(B, G, device=q.device, dtype=torch.float32)
    _check(s, "s", B, G)
    if mask is not None:
        _check(mask, "mask", B, None, torch.uint8)
    with torch.cuda.device(q.device):
        common = (_ptr(q), _ptr(dq), _ptr(s), _ptr(mask), B, ctypes.byref(params), _stream(q))
        if family == "circle":
            rc = _lib.lib.atacom_circle_slack_init(*common)
        elif family == "planar":
            rc = _lib.lib.atacom_planar_slack_init(*common)
        elif family.startswith("iiwa"):
            rc = _lib.lib.atacom_iiwa_slack_init(n, *common)
        else:
            raise ValueError(family)
    _lib.check(rc)
    return s


def point_reach_step(q, dq, obs_p, obs_dp, s, action, params, *, w=None, s_out=None, status=None, w_dbg=None):
    """PointReachAtacom.step up to the base-env call (collision_avoidance_atacom.py:29-47)."""
    B = q.shape[0]
    G = s.shape[1]
    _check(q, "q", B, 2)
    _check(dq, "dq", B, 2)
    _check(obs_p, "obs_p", B, 2 * G)
    _check(obs_dp, "obs_dp", B, 2 * G)
    _check(s, "s", B, G)
    _check(action, "action", B, 2)
    if w is None:
        w = torch.empty_like(q)
    if s_out is None:
        s_out = torch.empty_like(s)
    _check(w, "w", B, 2)
    _check(s_out, "s_out", B, G)
    if status is not None:
        _check(status, "status", B, None, torch.uint8)
    if w_dbg is not None:
        _check(w_dbg, "w_dbg", B, 2 * (2 + G))
    with torch.cuda.device(q.device):
        rc = _lib.lib.atacom_point_reach_step(G, _ptr(q), _ptr(dq), _ptr(obs_p), _ptr(obs_dp), _ptr(s),
                                              _ptr(action), _ptr(w), _ptr(s_out), _ptr(status), _ptr(w_dbg), B,
                                              ctypes.byref(params), _stream(q))
    _lib.check(rc)
    return w, s_out


def point_reach_slack_init(q, obs_p, params, *, s=None, mask=None):
    """PointReachAtacom.reset slack (collision_avoidance_atacom.py:25)."""
    B = q.shape[0]
    G = obs_p.shape[1] // 2
    _check(q, "q", B, 2)
    _check(obs_p, "obs_p", B, 2 * G)
    if s is None:
        s = torch.zeros(B, G, device=q.device, dtype=torch.float32)
    _check(s, "s", B, G)
    if mask is not None:
        _check(mask, "mask", B, None, torch.uint8)
    with torch.cuda.device(q.device):
        rc = _lib.lib.atacom_point_reach_slack_init(G, _ptr(q), _ptr(obs_p), _ptr(s), _ptr(mask), B,
                                                    ctypes.byref(params), _stream(q))
    _lib.check(rc)
    return s


def generic_supported(n, F, G):
    return bool(_lib.lib.atacom_generic_supported(n, F, G))


def generic_step(n, F, G, c, J, b, dq, s, alpha, params, *, ddq=None, s_out=None, status=None, w_dbg=None):
    """Projection for a user-defined ConstraintsSet (constraints.py:46-82) whose callbacks were
    evaluated batched: c [B,C] = fun(q), J [B,C,n], b [B,C] = b_state(q,dq); equality rows first."""
    B = dq.shape[0]
    C = F + G
    _check(c, "c", B, C)
    _check(b, "b", B, C)
    if tuple(J.shape) != (B, C, n) or not J.is_contiguous() or not J.is_cuda or J.dtype != torch.float32:
        raise ValueError("J must be a contiguous float32 CUDA tensor [B, C, n]")
    _check(dq, "dq", B, n)
    if G > 0:
        _check(s, "s", B, G)
    _check(alpha, "alpha", B, _alpha_dim(params, n, F))
    if ddq is None:
        ddq = torch.empty_like(dq)
    if s_out is None:
        s_out = torch.empty(B, G, device=dq.device, dtype=torch.float32)
    if status is not None:
        _check(status, "status", B, None, torch.uint8)
    if w_dbg is not None:
        _check(w_dbg, "w_dbg", B, 2 * (n + G))
    with torch.cuda.device(dq.device):
        rc = _lib.lib.atacom_generic_step(n, F, G, _ptr(c), _ptr(J), _ptr(b), _ptr(dq), _ptr(s) if G else None,
                                          _ptr(alpha), _ptr(ddq), _ptr(s_out) if G else None, _ptr(status),
                                          _ptr(w_dbg), B, ctypes.byref(params), _stream(dq))
    _lib.check(rc)
    return ddq, s_out


# ------------------------------------------------------------------ constraint statistics, fused roll-outs
def new_stats(device):
    """Device accumulators of the constraint log: { sum c_i, max c_i, max c_dq_i, count } (float64 [4])."""
    return torch.tensor([0.0, -math.inf, -math.inf, 0.0], dtype=torch.float64, device=device)


def read_stats(stats):
    """(c_avg, c_max, c_dq_max) as get_constraints_logs returns them (atacom.py:207-216)."""
    v = stats.cpu().tolist()
    return v[0] / max(v[3], 1.0), v[1], v[2]


def constraint_stats(family, q, dq, params, *, n_ctrl_joints=6, per_env=None, stats=None):
    """AtacomEnvWrapper._update_constraint_stats (atacom.py:201-205) for B environments: returns per_env [B, 2]
    = (max(|c_f|, c_g), max_j(|dq_j| - vel_max_j)); `stats` (see new_stats) is updated in place if given."""
    n, F, G = family_dims(family, n_ctrl_joints)
    B = q.shape[0]
    _check(q, "q", B, n)
    _check(dq, "dq", B, n)
    if per_env is None:
        per_env = torch.empty(B, 2, device=q.device, dtype=torch.float32)
    _check(per_env, "per_env", B, 2)
    if stats is not None and (stats.dtype != torch.float64 or stats.numel() != 4 or not stats.is_cuda):
        raise ValueError("stats must be a CUDA float64 tensor of 4 elements (projection.new_stats)")
    with torch.cuda.device(q.device):
        args = (_ptr(q), _ptr(dq), _ptr(per_env), _ptr(stats), B, ctypes.byref(params), _stream(q))
        if family == "circle":
            rc = _lib.lib.atacom_circle_constraint_stats(*args)
        elif family == "planar":
            rc = _lib.lib.atacom_planar_constraint_stats(*args)
        elif family.startswith("iiwa"):
            rc = _lib.lib.atacom_iiwa_constraint_stats(n, *args)
        else:
            raise ValueError(family)
    _lib.check(rc)
    return per_env


def circle_rollout(state, s, actions, params, *, rewards=None, stats=None, status=None):
    """T fused agent steps of CircleEnvAtacom / CircleEnvErrorCorrection (atacom.py:106-115,
    circle_base.py:53-67): `state` [B, 4] and `s` [B, 1] are advanced in place, `actions` is [T, B, 1]
    (or [T, B, 2] for the error-correction variant); returns rewards [T, B]."""
    B = state.shape[0]
    na = 2 if params.variant == _lib.VARIANT_ERROR_CORRECTION else 1
    _check(state, "state", B, 4)
    _check(s, "s", B, 1)
    if actions.dim() != 3 or actions.shape[1:] != (B, na) or not actions.is_contiguous() or not actions.is_cuda \
            or actions.dtype != torch.float32:
        raise ValueError("actions: expected a contiguous CUDA float32 tensor of shape [T, %d, %d]" % (B, na))
    T = actions.shape[0]
    if rewards is None:
        rewards = torch.empty(T, B, device=state.device, dtype=torch.float32)
    if status is not None:
        _check(status, "status", B, None, torch.uint8)
    with torch.cuda.device(state.device):
        rc = _lib.lib.atacom_circle_rollout(_ptr(state), _ptr(s), _ptr(actions), _ptr(rewards), _ptr(stats),
                                            _ptr(status), B, T, ctypes.byref(params), _stream(state))
    _lib.check(rc)
    return rewards


def point_reach_rollout(state, s, actions, params, *, obstacle_draws=None, obstacle_centers=None, time0=0.0,
                        rewards=None, stats=None, status=None):
    """T fused steps of PointReachAtacom (collision_avoidance_atacom.py:29-48 + collision_avoidance_base.py:41-76):
    `state` [B, 4 + 4 G] and `s` [B, G] are advanced in place; `actions` [T, B, 2]; `obstacle_draws` [T, B, 2 G]
    are the U(-1, 1) draws of the obstacles' random walk.  Returns rewards [T, B]."""
    B, G = s.shape
    _check(state, "state", B, 4 + 4 * G)
    _check(s, "s", B, G)
    T = actions.shape[0]
    for name, t, shape in (("actions", actions, (T, B, 2)), ("obstacle_draws", obstacle_draws, (T, B, 2 * G)),
                           ("obstacle_centers", obstacle_centers, (B, 2 * G))):
        if t is not None and (tuple(t.shape) != shape or not t.is_contiguous() or not t.is_cuda
                              or t.dtype != torch.float32):
            raise ValueError("%s: expected a contiguous CUDA float32 tensor of shape %s" % (name, shape))
    if rewards is None:
        rewards = torch.empty(T, B, device=state.device, dtype=torch.float32)
    with torch.cuda.device(state.device):
        rc = _lib.lib.atacom_point_reach_rollout(G, _ptr(state), _ptr(s), _ptr(actions), _ptr(obstacle_draws),
                                                 _ptr(obstacle_centers), float(time0), _ptr(rewards), _ptr(stats),
                                                 _ptr(status), B, T, ctypes.byref(params), _stream(state))
    _lib.check(rc)
    return rewards


class HostContext:
    """Host-buffer entry points (`atacom_*_step_host`): NumPy / pinned-tensor in, NumPy out,
    copies inside the call — what a caller of the NumPy reference binds."""

    MODES = {"auto": _lib.HOST_AUTO, "staged": _lib.HOST_STAGED, "zero_copy": _lib.HOST_ZERO_COPY,
             "hybrid": _lib.HOST_HYBRID}

    def __init__(self, max_B, chunks=2, mode="auto"):
        """mode: 'staged' (device staging buffers + copy engines), 'hybrid' (inputs through the copy engines,
        outputs stored by the kernels straight into the caller's page-locked buffers), 'zero_copy' (one kernel
        works on the caller's page-locked buffers over PCIe) or 'auto' (zero-copy when the buffers are page-locked,
        else staged).
        The pipeline of a call is replayed as a CUDA graph while the same buffers are passed."""
        self._ctx = ctypes.c_void_p()
        _lib.check(_lib.lib.atacom_host_ctx_create(ctypes.byref(self._ctx), max_B, chunks))
        _lib.check(_lib.lib.atacom_host_ctx_set_mode(self._ctx, self.MODES[mode]))
        self.max_B = max_B

    @staticmethod
    def auto_path(family, params, n_ctrl_joints=6):
        """The data path mode 'auto' takes for page-locked buffers (host_call in csrc/atacom_kernels.cu): zero-copy —
        one launch on the caller's buffers; with the reference-exact null basis both passes of the step run inside
        that launch (atacom_zc_fused_kernel), so every row still crosses PCIe once each way."""
        return "zero_copy"

    @staticmethod
    def _hptr(t):
        """Host pointer of a (pinned or pageable) CPU tensor or a NumPy array."""
        if t is None:
            return None
        if hasattr(t, "data_ptr"):
            if t.is_cuda or not t.is_contiguous():
                raise ValueError("host entry points take contiguous CPU tensors / NumPy arrays")
            return ctypes.c_void_p(t.data_ptr())
        if not t.flags["C_CONTIGUOUS"]:
            raise ValueError("host entry points take C-contiguous NumPy arrays")
        return ctypes.c_void_p(t.ctypes.data)

    def iiwa_step(self, n, q, dq, s, alpha, ddq, s_out, params, status=None):
        p = self._hptr
        _lib.check(_lib.lib.atacom_iiwa_step_host(self._ctx, n, p(q), p(dq), p(s), p(alpha), p(ddq), p(s_out),
                                                  p(status), q.shape[0], ctypes.byref(params)))
        return ddq, s_out

    def step(self, family, q, dq, s, alpha, ddq, s_out, params, status=None, n_ctrl_joints=6):
        """`projection.step` on host arrays: family in {"circle", "planar", "iiwa"}."""
        p = self._hptr
        args = (p(q), p(dq), p(s), p(alpha), p(ddq), p(s_out), p(status), q.shape[0], ctypes.byref(params))
        if family == "circle":
            rc = _lib.lib.atacom_circle_step_host(self._ctx, *args)
        elif family == "planar":
            rc = _lib.lib.atacom_planar_step_host(self._ctx, *args)
        elif family.startswith("iiwa"):
            rc = _lib.lib.atacom_iiwa_step_host(self._ctx, n_ctrl_joints, *args)
        else:
            raise ValueError(family)
        _lib.check(rc)
        return ddq, s_out

    def point_reach_step(self, q, dq, obs_p, obs_dp, s, action, w, s_out, params, status=None):
        """`projection.point_reach_step` on host arrays (PointReachAtacom.step, collision_avoidance_atacom.py:29-47)."""
        p = self._hptr
        _lib.check(_lib.lib.atacom_point_reach_step_host(self._ctx, s.shape[1], p(q), p(dq), p(obs_p), p(obs_dp), p(s),
                                                         p(action), p(w), p(s_out), p(status), q.shape[0],
                                                         ctypes.byref(params)))
        return w, s_out

    def generic_step(self, n, F, G, c, J, b, dq, s, alpha, ddq, s_out, params, status=None):
        """`projection.generic_step` on host arrays (a ConstraintsSet evaluated batched in NumPy)."""
        p = self._hptr
        _lib.check(_lib.lib.atacom_generic_step_host(self._ctx, n, F, G, p(c), p(J), p(b), p(dq), p(s), p(alpha),
                                                     p(ddq), p(s_out), p(status), dq.shape[0], ctypes.byref(params)))
        return ddq, s_out

    def close(self):
        if self._ctx:
            _lib.lib.atacom_host_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
