"""Pin of the iiwa constraint callbacks' constants (pinocchio itself is absent): the chain the oracle
(oracle/envs.py) and the kernels (atacom_iiwa_default_params; csrc/atacom_envs.cuh) are written from equals what the
reference's own urdf/iiwa_1.urdf:69-301, env_base.py:50,147-159 and iiwa_hit_atacom.py:23-46,106-107 say — parsed
live when /root/reference is there, and from the committed extraction (tests/golden/make_urdf_constants.py)
everywhere."""
import json
import os

import numpy as np
import pytest

from oracle import envs as oenv
from oracle import ref_loader
from rl_on_manifold_b200 import _lib
from tests import helpers

GOLDEN = os.path.join(helpers.ROOT, "tests", "golden", "iiwa_urdf_constants.json")


def _sources():
    out = [("committed extraction", json.load(open(GOLDEN)))]
    if ref_loader.reference_available():
        import importlib.util
        spec = importlib.util.spec_from_file_location("make_urdf_constants",
                                                      os.path.join(helpers.ROOT, "tests", "golden", "make_urdf_constants.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        out.append(("live reference files", mod.extract(ref_loader.REFERENCE_ROOT)))
    return out


@pytest.mark.parametrize("name,U", _sources(), ids=lambda v: v if isinstance(v, str) else "")
def test_iiwa_chain_constants_equal_the_reference_files(name, U):
    assert len(U["joints"]) == 7
    for j, (xyz, rpy) in zip(U["joints"], oenv.IIWA_ORIGINS):
        np.testing.assert_allclose(j["xyz"], xyz, atol=0, rtol=0)
        np.testing.assert_allclose(j["rpy"], rpy, atol=1e-15)
        assert j["axis"] == [0.0, 0.0, 1.0]                          # every joint revolute about its local z
        assert j["lower"] == -j["upper"]
    np.testing.assert_allclose([j["upper"] for j in U["joints"]], oenv.IIWA_Q_MAX, rtol=0, atol=0)
    np.testing.assert_allclose([j["velocity"] for j in U["joints"]], oenv.IIWA_VEL_MAX, rtol=0, atol=0)
    np.testing.assert_allclose(U["tip_offset"], oenv.IIWA_TIP, rtol=0, atol=0)
    # the tip frame the reference adds to joint 7 is the striker tip of the URDF: joint_ee + striker_joint_1
    if U["fixed_joints"]:
        assert abs(sum(v[2] for v in U["fixed_joints"].values()) - U["tip_offset"][2]) < 1e-12
    assert U["base_translate"][0] == oenv.IIWA_BASE_X
    assert (U["table_length"], U["table_width"], U["mallet_radius"]) == (oenv.TABLE_LENGTH, oenv.TABLE_WIDTH,
                                                                         oenv.MALLET_RADIUS)
    assert U["universal_height"] == oenv.UNIVERSAL_HEIGHT
    assert (U["z_link4_min"], U["z_link7_min"]) == (oenv.Z_LINK4_MIN, oenv.Z_LINK7_MIN)
    assert (U["frame_idx_4"], U["frame_idx_7"]) == (12, 18)
    # constructor gains of AirHockeyIiwaAtacom (iiwa_hit_atacom.py:23-40) as the oracle's spec states them
    spec = oenv.iiwa_spec(6)
    np.testing.assert_allclose(spec.K_f, [U["K_f"]])
    np.testing.assert_allclose(spec.K_g, [U["K_g_cart"]] * 5 + [U["K_g_joint"]] * 6)
    np.testing.assert_allclose(spec.K_c, U["Kc_default"])
    np.testing.assert_allclose(spec.acc_max, U["acc_max"])
    np.testing.assert_allclose(spec.K_q, U["Kq_factor"] * spec.acc_max / spec.vel_max)


@pytest.mark.parametrize("name,U", _sources(), ids=lambda v: v if isinstance(v, str) else "")
@pytest.mark.parametrize("n", [6, 7])
def test_c_abi_default_params_equal_the_reference_files(name, U, n):
    p = _lib.default_params("iiwa", n)
    env = list(p.env)
    assert env[0] == U["base_translate"][0]
    assert abs(env[1] - (U["table_length"] / 2 - U["mallet_radius"])) < 1e-15
    assert abs(env[2] - (U["table_width"] / 2 - U["mallet_radius"])) < 1e-15
    assert (env[3], env[4], env[5]) == (U["universal_height"], U["z_link4_min"], U["z_link7_min"])
    np.testing.assert_allclose(env[6:13], [j["upper"] for j in U["joints"]], rtol=0, atol=0)
    f32 = lambda v: np.asarray(v, dtype=np.float32)
    np.testing.assert_array_equal(f32(list(p.vel_max[:n])), f32([j["velocity"] for j in U["joints"]][:n]))
    np.testing.assert_array_equal(f32(list(p.acc_max[:n])), f32([U["acc_max"]] * n))
    np.testing.assert_array_equal(f32(list(p.K_f[:1])), f32([U["K_f"]]))
    np.testing.assert_array_equal(f32(list(p.K_g[:5 + n])), f32([U["K_g_cart"]] * 5 + [U["K_g_joint"]] * n))
    np.testing.assert_array_equal(f32(list(p.K_c[:6 + n])), f32([U["Kc_default"]] * (6 + n)))
    np.testing.assert_allclose(list(p.K_q[:n]), U["Kq_factor"] * U["acc_max"] / np.array([j["velocity"] for j in U["joints"]][:n]),
                               rtol=1e-6)
    assert abs(p.dt - 1.0 / 240.0) < 1e-9 and p.bias_mode == _lib.BIAS_OMEGA_X_V and p.basis_mode == _lib.BASIS_LAPACK


def test_device_fk_is_the_urdf_chain():
    """The kernels' forward kinematics (host build of csrc/atacom_envs.cuh: IiwaEnv) against a chain built from the
    PARSED URDF origins — through the slack-init rule, which exposes every constraint value c(q) + K J dq."""
    U = json.load(open(GOLDEN))
    origins = [(tuple(j["xyz"]), tuple(j["rpy"])) for j in U["joints"]]
    lib = helpers.load_harness()
    rng = np.random.default_rng(0)
    B = 64
    q = (rng.uniform(-0.9, 0.9, (B, 7)) * oenv.IIWA_Q_MAX).astype(np.float32)
    dq = (rng.uniform(-0.5, 0.5, (B, 7)) * oenv.IIWA_VEL_MAX).astype(np.float32)
    p = _lib.default_params("iiwa", 7)
    z = np.zeros((B, 12), np.float32)
    _, s_dev, _, _ = helpers.harness_step(lib, "iiwa7", p.flat(), q, dq, z, np.zeros((B, 6), np.float32), np.float64,
                                          init_only=True)
    saved = oenv.IIWA_ORIGINS
    try:
        oenv.IIWA_ORIGINS = origins
        spec = oenv.iiwa_spec(7)
        from oracle import atacom_oracle as ao
        want = np.stack([ao.slack_init(spec, oenv.iiwa_eval(q[i].astype(float), dq[i].astype(float)), dq[i].astype(float))
                         for i in range(B)])
    finally:
        oenv.IIWA_ORIGINS = saved
    np.testing.assert_allclose(s_dev, want, atol=2e-7)      # (the C struct carries the gains in fp32)
