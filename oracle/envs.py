"""Oracle: per-env constraint callbacks (float64 NumPy) and their wrapper specs.

TEST INFRASTRUCTURE (see oracle/__init__.py).

  circle_*   restates atacom/environments/circular_motion/circle_atacom.py:9-18,47-69   [PINNED]
  planar_*   restates atacom/environments/planar_air_hockey/atacom_air_hockey.py:28-43,78-107
             with the 3R kinematics written out (the reference calls pinocchio on a URDF that
             ships inside mushroom-rl, un-vendored)                                    [UNPINNED]
  iiwa_*     restates atacom/environments/iiwa_air_hockey/iiwa_hit_atacom.py:23-40,65-139 with
             the serial chain of urdf/iiwa_1.urdf:69-301 and the tip frame of
             env_base.py:147-151 written out (pinocchio absent)                        [UNPINNED]

`bias` selects the b(q,dq) term of the Cartesian rows:
  'jdot_qdot' — the mathematically intended dJ/dt * dq (constraints.py:19-20 docstring);
  'omega_x_v' — what the reference most likely executes: pinocchio's classical frame
                acceleration after a first-order forwardKinematics, i.e. omega x v
                (iiwa_hit_atacom.py:87-89,122-128).
"""
import numpy as np

from .atacom_oracle import ConstraintEval, Spec

# --------------------------------------------------------------------------- circle (env A / E)


def circle_spec(Kc=100.0, dt=0.01):
    """CircleEnvAtacom constructor (circle_atacom.py:8-18)."""
    return Spec(n=2, F=1, G=1, K_f=0.1, K_g=2.0, K_c=Kc, K_q=20.0, vel_max=1.0, acc_max=10.0, dt=dt)


def circle_eval(q, dq):
    """circle_atacom.py:47-69."""
    return ConstraintEval(
        c_f=np.array([q[0] ** 2 + q[1] ** 2 - 1.0]),
        J_f=np.array([[2.0 * q[0], 2.0 * q[1]]]),
        b_f=np.array([2.0 * dq[0] ** 2 + 2.0 * dq[1] ** 2]),
        c_g=np.array([-q[1] - 0.5]),
        J_g=np.array([[0.0, -1.0]]),
        b_g=np.array([0.0]))


def circle_base_step(state, u, dt=0.01):
    """CircularMotion.step after the hook (circle_base.py:59-65): clip, x10, integrate, reward."""
    a = np.clip(u, -1.0, 1.0) * 10.0
    st = np.array(state, dtype=np.float64)
    st[:2] += st[2:4] * dt + a * dt ** 2 / 2
    st[2:4] += a * dt
    reward = np.exp(-np.linalg.norm(np.array([1.0, 0.0]) - st[:2]))
    return st, reward


# --------------------------------------------------------------------------- serial-chain kinematics


def _rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


HALF_PI = np.pi / 2
# urdf/iiwa_1.urdf:72,110,147,184,221,258,295  (xyz, rpy) of joints 1..7; all revolute about local z
IIWA_ORIGINS = [
    ((0.0, 0.0, 0.1575), (0.0, 0.0, 0.0)),
    ((0.0, 0.0, 0.2025), (HALF_PI, 0.0, np.pi)),
    ((0.0, 0.2045, 0.0), (HALF_PI, 0.0, np.pi)),
    ((0.0, 0.0, 0.2155), (HALF_PI, 0.0, 0.0)),
    ((0.0, 0.1845, 0.0), (-HALF_PI, np.pi, 0.0)),
    ((0.0, 0.0, 0.2155), (HALF_PI, 0.0, 0.0)),
    ((0.0, 0.081, 0.0), (-HALF_PI, np.pi, 0.0)),
]
IIWA_TIP = np.array([0.0, 0.0, 0.585])          # env_base.py:148-149 (0.07 + 0.515, urdf:331,370)
IIWA_Q_MAX = np.array([2.9670597283903604, 2.0943951023931953, 2.9670597283903604, 2.0943951023931953,
                       2.9670597283903604, 2.0943951023931953, 3.0543261909900763])   # urdf:74..297
IIWA_VEL_MAX = np.array([1.4835298641951802, 1.4835298641951802, 1.7453292519943295, 1.3089969389957472,
                         2.2689280275926285, 2.356194490192345, 2.356194490192345])   # urdf:74..297
IIWA_BASE_X = -1.51                              # env_base.py:50 (world x of the robot base)
TABLE_LENGTH, TABLE_WIDTH, MALLET_RADIUS = 1.96, 1.02, 0.05      # env_base.py:156-158
UNIVERSAL_HEIGHT = 0.1505                        # env_base.py:159
Z_LINK4_MIN, Z_LINK7_MIN = 0.36, 0.25            # iiwa_hit_atacom.py:106-107


def chain_fk(origins, q):
    """Joint frames of a revolute-z serial chain: list of (R_i, o_i) in the base frame."""
    R = np.eye(3)
    o = np.zeros(3)
    frames = []
    for (xyz, rpy), qi in zip(origins, q):
        o = o + R @ np.array(xyz)
        c, s = np.cos(qi), np.sin(qi)
        R = R @ _rpy(*rpy) @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        frames.append((R, o))
    return frames


def point_kinematics(frames, dq, joint, offset, n_cols):
    """Position, world-aligned linear Jacobian (3 x n_cols), dJ/dt*dq and omega x v of a point
    fixed in the frame of joint `joint` (1-based) at `offset` in that frame."""
    nj = len(frames)
    dqf = np.zeros(nj)
    dqf[:len(dq)] = dq
    Rj, oj = frames[joint - 1]
    p = oj + Rj @ np.asarray(offset)
    J = np.zeros((3, n_cols))
    for i in range(min(joint, n_cols)):
        Ri, oi = frames[i]
        J[:, i] = np.cross(Ri[:, 2], p - oi)
    # zero-joint-acceleration recursion for the point's acceleration
    w = np.zeros(3)
    al = np.zeros(3)
    a_o = np.zeros(3)
    o_prev = np.zeros(3)
    for i in range(joint):
        Ri, oi = frames[i]
        r = oi - o_prev
        a_o = a_o + np.cross(al, r) + np.cross(w, np.cross(w, r))
        z = Ri[:, 2]
        al = al + np.cross(w, z * dqf[i])
        w = w + z * dqf[i]
        o_prev = oi
    r = p - o_prev
    acc = a_o + np.cross(al, r) + np.cross(w, np.cross(w, r))
    v = J @ dqf[:n_cols]
    return p, J, acc, np.cross(w, v)


def iiwa_spec(n=6, Kc=240.0, dt=1.0 / 240.0):
    """AirHockeyIiwaAtacom constructor (iiwa_hit_atacom.py:23-40); n=7 is the
    non-isolated-joint-7 extension (env_single.py:17-20)."""
    acc_max = np.ones(n) * 10.0
    vel_max = IIWA_VEL_MAX[:n].copy()
    return Spec(n=n, F=1, G=5 + n, K_f=0.1, K_g=np.concatenate([np.ones(5) * 0.5, np.ones(n)]),
                K_c=Kc, K_q=4.0 * acc_max / vel_max, vel_max=vel_max, acc_max=acc_max, dt=dt)


def iiwa_eval(q, dq, bias="omega_x_v"):
    """iiwa_hit_atacom.py:70-139.  q, dq have n = 6 or 7 entries; joints beyond n sit at 0
    (`_get_pino_value` pads with zeros, :65-68)."""
    n = len(q)
    qf = np.zeros(7)
    qf[:n] = q
    fr = chain_fk(IIWA_ORIGINS, qf)
    sel = 2 if bias == "jdot_qdot" else 3
    tip = point_kinematics(fr, dq, 7, IIWA_TIP, n)
    l4 = point_kinematics(fr, dq, 4, (0, 0, 0), n)
    l7 = point_kinematics(fr, dq, 7, (0, 0, 0), n)
    p, J, b = tip[0], tip[1], tip[sel]
    if n == 7:
        # the tip lies on the axis of joint 7 (env_base.py:147-151): its Jacobian column is exactly zero, where
        # z_7 x (tip - o_7) leaves rounding noise of either sign
        J[:, 6] = 0.0
    xw = p[0] + IIWA_BASE_X
    half_l = TABLE_LENGTH / 2 - MALLET_RADIUS
    half_w = TABLE_WIDTH / 2 - MALLET_RADIUS
    qmax = IIWA_Q_MAX[:n]
    c_g = np.concatenate([[-xw - half_l, -p[1] - half_w, p[1] - half_w,
                           Z_LINK4_MIN - l4[0][2], Z_LINK7_MIN - l7[0][2]], q ** 2 - qmax ** 2])
    J_g = np.vstack([-J[0], -J[1], J[1], -l4[1][2], -l7[1][2], 2.0 * np.diag(q)])
    b_g = np.concatenate([[-b[0], -b[1], b[1], -l4[sel][2], -l7[sel][2]], 2.0 * dq ** 2])
    return ConstraintEval(c_f=np.array([p[2] - UNIVERSAL_HEIGHT]), J_f=J[2:3].copy(), b_f=np.array([b[2]]),
                          c_g=c_g, J_g=J_g, b_g=b_g)


# --------------------------------------------------------------------------- planar 3R (env H)

# The planar URDF lives inside mushroom-rl (atacom_air_hockey.py:45) and is NOT available here:
# every number below is a parameter; the defaults are recalled, unverified.
PLANAR_DEFAULTS = dict(links=(0.55, 0.44, 0.44), base_x=-1.51, base_y=0.0,
                       q_max=(2.9670597283903604, 2.0943951023931953, 2.0943951023931953),
                       vel_max=(1.4835298641951802, 1.4835298641951802, 1.7453292519943295),
                       table_length=TABLE_LENGTH, table_width=TABLE_WIDTH, mallet_radius=MALLET_RADIUS)


def planar_spec(Kc=240.0, dt=1.0 / 240.0, params=None):
    """AirHockeyPlanarAtacom constructor (atacom_air_hockey.py:28-43)."""
    pr = dict(PLANAR_DEFAULTS, **(params or {}))
    acc_max = np.ones(3) * 10.0
    vel_max = np.asarray(pr["vel_max"], dtype=np.float64)
    return Spec(n=3, F=0, G=6, K_f=np.zeros(0), K_g=np.array([0.5, 0.5, 0.5, 1.0, 1.0, 1.0]), K_c=Kc,
                K_q=2.0 * acc_max / vel_max, vel_max=vel_max, acc_max=acc_max, dt=dt)


def planar_eval(q, dq, bias="omega_x_v", params=None):
    """atacom_air_hockey.py:78-107 for a planar 3R arm."""
    pr = dict(PLANAR_DEFAULTS, **(params or {}))
    l = np.asarray(pr["links"])
    th = np.cumsum(q)
    om = np.cumsum(dq)
    cx, sy = np.cos(th), np.sin(th)
    x = pr["base_x"] + (l * cx).sum()
    y = pr["base_y"] + (l * sy).sum()
    J = np.zeros((2, 3))
    for j in range(3):
        J[0, j] = -(l[j:] * sy[j:]).sum()
        J[1, j] = (l[j:] * cx[j:]).sum()
    if bias == "jdot_qdot":
        bx = -(l * cx * om ** 2).sum()
        by = -(l * sy * om ** 2).sum()
    else:
        v = J @ dq
        bx, by = -om[-1] * v[1], om[-1] * v[0]
    half_l = pr["table_length"] / 2 - pr["mallet_radius"]
    half_w = pr["table_width"] / 2 - pr["mallet_radius"]
    qmax = np.asarray(pr["q_max"])
    c_g = np.concatenate([[-x - half_l, -y - half_w, y - half_w], q ** 2 - qmax ** 2])
    J_g = np.vstack([-J[0], -J[1], J[1], 2.0 * np.diag(q)])
    b_g = np.concatenate([[-bx, -by, by], 2.0 * dq ** 2])
    e0 = np.zeros(0)
    return ConstraintEval(c_f=e0, J_f=np.zeros((0, 3)), b_f=e0, c_g=c_g, J_g=J_g, b_g=b_g)
