"""Import the *unmodified* reference from /root/reference under import shims.

In-container only (the GPU box has no /root/reference): used by
`tests/golden/make_golden.py` to generate fixtures and by the `not gpu` tests
that validate the oracle against the reference (skipped when absent).

The reference needs mushroom_rl / matplotlib, which are not installed.  The
shims below provide just the names the hot-path files touch:
  * `mushroom_rl.utils.spaces` must export `Box` AND `np` — `atacom/atacom.py:1`
    gets NumPy through `from mushroom_rl.utils.spaces import *`.
  * `mushroom_rl.core` -> Environment, MDPInfo, Core, Agent
  * `mushroom_rl.utils.viewer.Viewer`, `mushroom_rl.utils.dataset.compute_J`
  * empty matplotlib / pyplot / ticker (`atacom/utils/__init__.py` pulls
    plot_utils).
"""
import copy
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("ATACOM_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "atacom"))


class _Box:
    def __init__(self, low, high, shape=None):
        self.low = np.asarray(low, dtype=float)
        self.high = np.asarray(high, dtype=float)

    @property
    def shape(self):
        return self.low.shape


class _MDPInfo:
    def __init__(self, observation_space, action_space, gamma, horizon):
        self.observation_space = observation_space
        self.action_space = action_space
        self.gamma = gamma
        self.horizon = horizon

    def copy(self):
        return copy.deepcopy(self)


class _Environment:
    def __init__(self, mdp_info):
        self._mdp_info = mdp_info

    @property
    def info(self):
        return self._mdp_info

    def seed(self, seed):
        np.random.seed(seed)

    def stop(self):
        pass


class _Viewer:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return lambda *a, **k: None


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install_shims():
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "mushroom_rl" not in sys.modules:
        _module("mushroom_rl")
        _module("mushroom_rl.utils", spaces=None)
        sp = _module("mushroom_rl.utils.spaces", Box=_Box, np=np)
        sys.modules["mushroom_rl.utils"].spaces = sp
        _module("mushroom_rl.core", Environment=_Environment, MDPInfo=_MDPInfo,
                Core=object, Agent=object)
        _module("mushroom_rl.utils.viewer", Viewer=_Viewer)
        _module("mushroom_rl.utils.dataset", compute_J=lambda *a, **k: None,
                parse_dataset=lambda *a, **k: None)
    if "matplotlib" not in sys.modules:
        class _Anything(types.ModuleType):
            def __getattr__(self, name):
                if name.startswith("__"):
                    raise AttributeError(name)
                return _Anything(name)

            def __call__(self, *a, **k):
                return _Anything("call")

        for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker",
                     "matplotlib.lines", "matplotlib.patches"):
            sys.modules[name] = _Anything(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def load():
    """Return a namespace with the reference's hot-path classes/functions."""
    install_shims()
    from atacom.atacom import AtacomEnvWrapper
    from atacom.constraints import ViabilityConstraint, ConstraintsSet
    from atacom.error_correction_wrapper import ErrorCorrectionEnvWrapper
    from atacom.utils.null_space_coordinate import pinv_null, rref, rref_sympy
    from atacom.environments.circular_motion.circle_atacom import CircleEnvAtacom
    from atacom.environments.circular_motion.circle_error_correction import CircleEnvErrorCorrection
    from atacom.environments.collision_avoidance.collision_avoidance_atacom import PointReachAtacom
    return types.SimpleNamespace(
        AtacomEnvWrapper=AtacomEnvWrapper, ViabilityConstraint=ViabilityConstraint,
        ConstraintsSet=ConstraintsSet, ErrorCorrectionEnvWrapper=ErrorCorrectionEnvWrapper,
        pinv_null=pinv_null, rref=rref, rref_sympy=rref_sympy,
        CircleEnvAtacom=CircleEnvAtacom, CircleEnvErrorCorrection=CircleEnvErrorCorrection,
        PointReachAtacom=PointReachAtacom, Box=_Box, MDPInfo=_MDPInfo, Environment=_Environment)
