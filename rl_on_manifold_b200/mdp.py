"""Minimal MDP plumbing types.

The reference takes `Box`, `MDPInfo` and `Environment` from MushroomRL (atacom/atacom.py:1,50-51;
circle_base.py:6-8).  When mushroom_rl is importable its classes are used unchanged; otherwise these
duck-typed stand-ins provide the attributes the wrappers and the examples touch.
"""
import copy

import numpy as np

try:                                                     # pragma: no cover - not installed in this image
    from mushroom_rl.utils.spaces import Box             # type: ignore
    from mushroom_rl.core import MDPInfo                 # type: ignore
    HAVE_MUSHROOM_RL = True
except Exception:
    HAVE_MUSHROOM_RL = False

    class Box:
        def __init__(self, low, high, shape=None):
            self._low = np.asarray(low, dtype=np.float64)
            self._high = np.asarray(high, dtype=np.float64)

        @property
        def low(self):
            return self._low

        @property
        def high(self):
            return self._high

        @property
        def shape(self):
            return self._low.shape

    class MDPInfo:
        def __init__(self, observation_space, action_space, gamma, horizon):
            self.observation_space = observation_space
            self.action_space = action_space
            self.gamma = gamma
            self.horizon = horizon

        def copy(self):
            return copy.deepcopy(self)
