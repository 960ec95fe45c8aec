"""ErrorCorrectionEnvWrapper — the reference's baseline "E" (atacom/error_correction_wrapper.py): no
null-space projection and no drift term, only  ddq = alpha - Jc^+ (K_c c).  Same constructor; default
Kc = 100 (error_correction_wrapper.py:12); the action has dim_q entries and is scaled by acc_max
element-wise (:102-104); q, dq are refreshed from the simulator state inside the hook (:119-120)."""
from . import _lib
from .atacom import AtacomEnvWrapper


class ErrorCorrectionEnvWrapper(AtacomEnvWrapper):
    variant = _lib.VARIANT_ERROR_CORRECTION

    def __init__(self, base_env, dim_q, vel_max, acc_max, f=None, g=None, Kc=100., Kq=10., time_step=0.01,
                 **kw):
        super().__init__(base_env, dim_q, vel_max, acc_max, f=f, g=g, Kc=Kc, Kq=Kq, time_step=time_step, **kw)

    def _action_dim(self):
        return self.dims['q']                           # error_correction_wrapper.py:49

    def _action_scale(self):
        return self.acc_max                             # error_correction_wrapper.py:104

    def step(self, action):
        # error_correction_wrapper.py:102-107: no q/dq refresh and no constraint statistics in step()
        import numpy as np
        import torch
        self._numpy_io = isinstance(action, np.ndarray)
        action = self._as_batch(action)
        low = torch.as_tensor(self.info.action_space.low, dtype=torch.float32, device=self.device)
        high = torch.as_tensor(self.info.action_space.high, dtype=torch.float32, device=self.device)
        alpha = torch.minimum(torch.maximum(action, low), high)
        alpha = alpha * torch.as_tensor(self.acc_max, dtype=torch.float32, device=self.device)
        self.state, reward, absorb, info = self.env.step(alpha)
        self.state = self._as_batch(self.state)
        return self._ret(self.state.clone()), self._ret(reward), self._ret(absorb), info

    def _refresh_before_projection(self):
        self.q = self._get_q(self.state).contiguous()   # error_correction_wrapper.py:119-120
        self.dq = self._get_dq(self.state).contiguous()

    def get_constraints_logs(self):
        return self.env.get_constraints_logs()          # error_correction_wrapper.py:196-197
