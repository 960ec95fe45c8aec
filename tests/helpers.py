"""Shared test helpers: host harness loader, oracle drivers, error metrics."""
import ctypes
import os
import subprocess

import numpy as np

from oracle import atacom_oracle as ao
from oracle import envs as oenv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS_DIR = os.path.join(ROOT, "tests", "host_harness")
CSRC = os.path.join(ROOT, "rl_on_manifold_b200", "csrc")

# the null basis the product's default mode reproduces (see oracle/nullspace.py)
REFERENCE_BASIS = "svd"          # SciPy's own SVD null basis: what atacom.py:127 hands to rref

# the two null-basis modes of the kernels and the oracle basis each one reproduces
ORACLE_BASIS = {"lapack": "svd", "canonical": "canonical"}


def with_basis(params, mode):
    """Copy of `params` in the given basis mode ("lapack": the reference's own null basis, the default;
    "canonical": the basis-free fast path alone)."""
    p = params.copy()
    p.basis_mode = {"lapack": 0, "canonical": 1}[mode]
    return p


ENV_ID = {"circle": 0, "planar": 1, "iiwa6": 2, "iiwa7": 3}
DIMS = {"circle": (2, 1, 1), "planar": (3, 0, 6), "iiwa6": (6, 1, 11), "iiwa7": (7, 1, 12)}


def load_harness():
    """Compile the product's device headers for the host (g++) and load them.  Test infrastructure."""
    so = os.path.join(HARNESS_DIR, "_harness.so")
    srcs = [os.path.join(HARNESS_DIR, "harness.cpp")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)
                                                         if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, srcs[0]])
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def harness_step(lib, family, params_flat, q, dq, s, alpha, dtype=np.float32, init_only=False):
    """Run env functor + full step of the device code on the host.  Returns ddq, s_out, w_dbg, status."""
    n, F, G = DIMS[family]
    B = q.shape[0]
    fn = lib.harness_step_f32 if dtype == np.float32 else lib.harness_step_f64
    arrs = [np.ascontiguousarray(params_flat, dtype=np.float64)] + \
        [np.ascontiguousarray(a, dtype=dtype) for a in (q, dq, s, alpha)]
    ddq = np.zeros((B, n), dtype)
    s_out = np.zeros((B, G), dtype)
    dbg = np.zeros((B, 2 * (n + G)), dtype)
    st = np.zeros(B, np.uint8)
    rc = fn(ENV_ID[family], ctypes.c_int64(B), *[_p(a) for a in arrs], _p(ddq), _p(s_out), _p(dbg), _p(st),
            int(init_only))
    assert rc == 0
    return ddq, s_out, dbg, st


def harness_dense(lib, n, F, G, Af, Ag, s, r, alpha, tol, dtype=np.float32):
    B = s.shape[0] if G else alpha.shape[0]
    N = n + G
    fn = lib.harness_dense_f32 if dtype == np.float32 else lib.harness_dense_f64
    arrs = [np.ascontiguousarray(a, dtype=dtype) for a in (Af, Ag, s, r, alpha)]
    wmn, wn, st = np.zeros((B, N), dtype), np.zeros((B, N), dtype), np.zeros(B, np.uint8)
    ct = ctypes.c_float if dtype == np.float32 else ctypes.c_double
    rc = fn(n, F, G, ctypes.c_int64(B), *[_p(a) for a in arrs], ct(tol), _p(wmn), _p(wn), _p(st))
    assert rc == 0, "shape not compiled into the harness"
    return wmn, wn, st


def harness_lapack(lib, n, F, G, Af, Ag, s, r, alpha, tol):
    """The LAPACK-basis routine (atacom_lapack.cuh) built for the host in double."""
    B = s.shape[0] if G else alpha.shape[0]
    N = n + G
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (Af, Ag, s, r, alpha)]
    wmn, wn, st = np.zeros((B, N)), np.zeros((B, N)), np.zeros(B, np.uint8)
    rc = lib.harness_lapack_f64(n, F, G, ctypes.c_int64(B), *[_p(a) for a in arrs], ctypes.c_double(tol), _p(wmn), _p(wn),
                                _p(st))
    assert rc == 0, "shape not compiled into the harness"
    return wmn, wn, st


def harness_structured(lib, n, F, G, Af, Ag, s, r, alpha, tol, tau, dtype=np.float32):
    B = s.shape[0] if G else alpha.shape[0]
    N = n + G
    fn = lib.harness_structured_f32 if dtype == np.float32 else lib.harness_structured_f64
    arrs = [np.ascontiguousarray(a, dtype=dtype) for a in (Af, Ag, s, r, alpha)]
    wmn, wn, st = np.zeros((B, N), dtype), np.zeros((B, N), dtype), np.zeros(B, np.uint8)
    ct = ctypes.c_float if dtype == np.float32 else ctypes.c_double
    rc = fn(n, F, G, ctypes.c_int64(B), *[_p(a) for a in arrs], ct(tol), ct(tau), _p(wmn), _p(wn), _p(st))
    assert rc == 0, "shape not compiled into the harness"
    return wmn, wn, st


def harness_point(lib, G, params_flat, q, dq, p, dp, s, act, dtype=np.float32, init_only=False):
    B = q.shape[0]
    fn = lib.harness_point_f32 if dtype == np.float32 else lib.harness_point_f64
    arrs = [np.ascontiguousarray(params_flat, dtype=np.float64)] + \
        [np.ascontiguousarray(a, dtype=dtype) for a in (q, dq, p, dp, s, act)]
    w, s_out, dbg, st = np.zeros((B, 2), dtype), np.zeros((B, G), dtype), np.zeros((B, 2 * (2 + G)), dtype), \
        np.zeros(B, np.uint8)
    rc = fn(G, ctypes.c_int64(B), *[_p(a) for a in arrs], _p(w), _p(s_out), _p(dbg), _p(st), int(init_only))
    assert rc == 0
    return w, s_out, dbg, st


# ------------------------------------------------------------------ oracle drivers

def oracle_spec(family):
    if family == "circle":
        return oenv.circle_spec()
    if family == "planar":
        return oenv.planar_spec()
    if family == "iiwa6":
        return oenv.iiwa_spec(6)
    if family == "iiwa7":
        return oenv.iiwa_spec(7)
    raise ValueError(family)


def oracle_eval(family, q, dq, bias="omega_x_v"):
    if family == "circle":
        return oenv.circle_eval(q, dq)
    if family == "planar":
        return oenv.planar_eval(q, dq, bias)
    return oenv.iiwa_eval(q, dq, bias)


def oracle_batch(family, q, dq, s, alpha, basis="svd", variant="atacom", bias="omega_x_v", spec=None):
    """Run the per-env oracle over a batch (float64 views of the fp32 inputs).  Returns dict of arrays
    plus per-env flags: fired (tolerance branch), rank_def, margin (distance of the closest pivot
    candidate to the tolerance, relative)."""
    spec = spec or oracle_spec(family)
    B = q.shape[0]
    out = dict(ddq=np.zeros((B, spec.n)), s_new=np.zeros((B, spec.G)), w=np.zeros((B, spec.N)),
               w_mn=np.zeros((B, spec.N)), w_null=np.zeros((B, spec.N)), fired=np.zeros(B, bool),
               rank_def=np.zeros(B, bool), margin=np.full(B, np.inf), cond=np.zeros(B))
    q, dq, s, alpha = (np.asarray(a, dtype=np.float64) for a in (q, dq, s, alpha))
    for i in range(B):
        ev = oracle_eval(family, q[i], dq[i], bias)
        o = ao.atacom_step(spec, ev, dq[i], s[i], alpha[i], basis=basis, variant=variant)
        out["ddq"][i], out["s_new"][i], out["w"][i] = o["ddq"], o["s_new"], o["w"]
        out["w_mn"][i], out["w_null"][i] = o["act_a"] + o["act_err"], o["act_b"]
        tr = o["trace"]
        if tr:
            out["fired"][i] = any(p > 1e-9 for (_, _, p) in tr["dropped"])
            cand = [p for (_, _, p) in tr["dropped"]] + [p for (_, _, p) in tr["pivots"]]
            out["margin"][i] = min(abs(p - tr["tol"]) / tr["tol"] for p in cand) if cand else np.inf
        out["rank_def"][i] = o["rank"] < spec.C
        sv = np.linalg.svd(o["Jc"], compute_uv=False)
        out["cond"][i] = sv[0] / max(sv[-1], 1e-300)
    return out


def oracle_point_reach_batch(q, dq, p, dp, s, action):
    """Env C oracle over a batch (collision_avoidance_atacom.py:29-47).  `ambiguous`: some pivot candidate of the rref
    lies between the reference's tolerance (~1e-15) and the kernels' (2.4e-6 in fp32 terms): decided differently,
    outside the parity domain."""
    B, G = s.shape
    out = dict(w=np.zeros((B, 2 + G)), s_new=np.zeros((B, G)), rank_def=np.zeros(B, bool), ambiguous=np.zeros(B, bool))
    q, dq, p, dp, s, action = (np.asarray(a, dtype=np.float64) for a in (q, dq, p, dp, s, action))
    for i in range(B):
        try:
            o = ao.point_reach_step(q[i], dq[i], p[i].reshape(G, 2), dp[i].reshape(G, 2), s[i], action[i])
        except ValueError:
            out["rank_def"][i] = True
            continue
        out["w"][i], out["s_new"][i] = o["w"], o["s_new"]
        out["rank_def"][i] = o["rank"] < G
        cand = [pv for (_, _, pv) in o["trace"]["pivots"] + o["trace"]["dropped"]]
        out["ambiguous"][i] = any(1e-12 < pv < 1e-4 for pv in cand)
    return out


def oracle_batch_mp(family, *arrays, basis="svd", variant="atacom", bias="omega_x_v", tmpdir=None):
    """`oracle_batch` (or `oracle_point_reach_batch` for family "point_reach") fanned out over all host cores in a
    separate interpreter (tests/oracle_mp.py): what makes parity at BASELINE.json's full batch sizes affordable."""
    import sys
    import tempfile
    d = tmpdir or tempfile.mkdtemp(prefix="oracle_mp_")
    src, dst = os.path.join(str(d), "in_%s.npz" % family), os.path.join(str(d), "out_%s.npz" % family)
    names = ("q", "dq", "p", "dp", "s", "action") if family == "point_reach" else ("q", "dq", "s", "alpha")
    np.savez(src, family=family, basis=basis, variant=variant, bias=bias,
             **{k: np.ascontiguousarray(a) for k, a in zip(names, arrays)})
    subprocess.check_call([sys.executable, "-W", "ignore", "-m", "tests.oracle_mp", src, dst], cwd=ROOT)
    z = np.load(dst)
    return {k: z[k] for k in z.files}


def rel_err(x, ref, scale=None):
    """Per-env max-norm error relative to max(1, max|ref|) — or to max(1, max|scale|) when the
    compared quantity was clipped from a larger one (ddq is w[:n] clipped to +-acc_max)."""
    x, ref = np.asarray(x, np.float64), np.asarray(ref, np.float64)
    if x.ndim == 1:
        x, ref = x[:, None], ref[:, None]
    if x.shape[1] == 0:
        return np.zeros(x.shape[0])
    sc = np.abs(ref if scale is None else np.asarray(scale, np.float64)).max(1)
    return np.abs(x - ref).max(1) / np.maximum(1.0, sc)


def synthetic_cpu(family, B, seed):
    """Synthetic batch as NumPy fp32 arrays with slacks from the ORACLE's reset rule + the mix."""
    import torch
    from rl_on_manifold_b200 import synthetic
    fam, nj = ("iiwa", int(family[-1])) if family.startswith("iiwa") else (family, 6)
    q, dq, alpha = synthetic.state_batch(fam, B, seed, nj)
    spec = oracle_spec(family)
    s = np.stack([ao.slack_init(spec, oracle_eval(family, q[i].double().numpy(), dq[i].double().numpy()),
                                dq[i].double().numpy()) for i in range(B)]).astype(np.float32)
    s = synthetic.slack_mix(torch.from_numpy(s), seed).numpy()
    return q.numpy(), dq.numpy(), s, alpha.numpy()


def exact_params_flat(family, params):
    """`params.flat()` with every constant replaced by the oracle's float64 value, for the float64 host
    build (the C struct stores fp32, e.g. 0.1f != 0.1)."""
    spec = oracle_spec(family)
    f = params.flat()
    off = 0
    for arr, size in ((spec.K_f, 4), (spec.K_g, 16), (spec.K_c, 20), (spec.K_q, 8), (spec.vel_max, 8),
                      (spec.acc_max, 8)):
        f[off:off + len(arr)] = list(arr)
        off += size
    f[off] = spec.dt
    f[off + 1] = spec.tol
    env0 = off + 6
    if family.startswith("iiwa"):
        env = [oenv.IIWA_BASE_X, oenv.TABLE_LENGTH / 2 - oenv.MALLET_RADIUS, oenv.TABLE_WIDTH / 2 - oenv.MALLET_RADIUS,
               oenv.UNIVERSAL_HEIGHT, oenv.Z_LINK4_MIN, oenv.Z_LINK7_MIN] + list(oenv.IIWA_Q_MAX)
    elif family == "planar":
        d = oenv.PLANAR_DEFAULTS
        env = list(d["links"]) + [d["base_x"], d["base_y"]] + list(d["q_max"]) + \
            [d["table_length"] / 2 - d["mallet_radius"], d["table_width"] / 2 - d["mallet_radius"]]
    else:
        env = []
    f[env0:env0 + len(env)] = env
    return f
