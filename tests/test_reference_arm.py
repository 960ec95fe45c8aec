"""The reference arm of bench.py runs the REFERENCE: baseline/reference_driver.py imports the unmodified package
(staged copy under baseline/_ref, else /root/reference) and drives AtacomEnvWrapper.step_action_function.  Checked
here against the oracle on the same seeded IiwaAirHockey-7H inputs, and that the arm's process does not map the
product's native library."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import helpers

ROOT = helpers.ROOT
sys.path.insert(0, ROOT)
from baseline import reference_driver as rd  # noqa: E402

pytestmark = pytest.mark.skipif(not rd.reference_present(), reason="reference tree not present")


def test_reference_driver_matches_oracle_on_the_basis_free_stratum():
    B = 384
    pool = rd.ReferencePool(B, seed=1234, cores=4)
    try:
        wall, inner = pool.step()
        assert 0 < inner <= wall
        ddq, s_new = pool.outputs()
    finally:
        pool.close()
    q, dq, s0, alpha = helpers.synthetic_cpu("iiwa6", B, 1234)
    svd = helpers.oracle_batch("iiwa6", q, dq, s0, alpha, basis="svd")
    can = helpers.oracle_batch("iiwa6", q, dq, s0, alpha, basis="canonical")
    # where the tolerance branch fires (or nearly does) the reference's output depends on LAPACK's null basis
    free = ~svd["fired"] & ~can["fired"] & ~svd["rank_def"] & (svd["margin"] > 0.02) & (can["margin"] > 0.02)
    assert free.mean() > 0.6
    assert np.abs(ddq - svd["ddq"])[free].max() < 1e-8
    assert np.abs(s_new - svd["s_new"])[free].max() < 1e-8
    # the driver's inputs are the benchmark's inputs: slacks from the reference's own reset rule + the mix
    assert np.isfinite(ddq).all()


def test_reference_arm_process_does_not_load_the_product_library():
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from baseline import reference_driver as rd\n"
            "p = rd.ReferencePool(64, cores=2); p.step(); p.close()\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libatacom_b200' not in maps and 'rl_on_manifold_b200' not in sys.modules, 'product code loaded'\n"
            "print('clean')\n" % ROOT)
    out = subprocess.run([sys.executable, "-W", "ignore", "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "clean" in out.stdout, out.stderr[-2000:]


def test_bench_reference_arm_prints_the_contract_line():
    import json
    env = dict(os.environ, ATACOM_REF_ENVS_PER_STEP="256")
    out = subprocess.run([sys.executable, "-W", "ignore", os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "atacom_projection_env_steps_per_sec"
    assert line["cpu_baseline"]["kind"] == "reference" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    assert "workload" in line["config"]
