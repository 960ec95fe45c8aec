"""Extract the kinematic constants of the iiwa chain from the reference's own files (run in the build container):

    python tests/golden/make_urdf_constants.py

  urdf/iiwa_1.urdf          joints iiwa_1/joint_1..7: origin xyz / rpy, axis, position and velocity limits
  env_base.py               tip offset, robot base translation, table, mallet, universal height
  iiwa_hit_atacom.py        link-4 / link-7 height limits, constraint gains, frame indices

and write them to tests/golden/iiwa_urdf_constants.json, which tests/test_urdf_pin.py compares with oracle/envs.py
and with atacom_iiwa_default_params on any box (the live files are compared too when /root/reference is there).
"""
import json
import os
import re
import sys
import xml.etree.ElementTree as ET

HERE = os.path.dirname(os.path.abspath(__file__))


def extract(ref_root="/root/reference"):
    base = os.path.join(ref_root, "atacom", "environments", "iiwa_air_hockey")
    root = ET.parse(os.path.join(base, "urdf", "iiwa_1.urdf")).getroot()
    joints = []
    for i in range(1, 8):
        j = next(e for e in root.iter("joint") if e.get("name") == "iiwa_1/joint_%d" % i)
        assert j.get("type") == "revolute"
        o, lim = j.find("origin"), j.find("limit")
        joints.append(dict(xyz=[float(v) for v in o.get("xyz").split()], rpy=[float(v) for v in o.get("rpy").split()],
                           axis=[float(v) for v in j.find("axis").get("xyz").split()],
                           lower=float(lim.get("lower")), upper=float(lim.get("upper")),
                           velocity=float(lim.get("velocity"))))
    fixed = {}
    for name in ("iiwa_1/joint_ee", "iiwa_1/striker_joint_1"):
        j = next((e for e in root.iter("joint") if e.get("name") == name), None)
        if j is not None:
            fixed[name] = [float(v) for v in j.find("origin").get("xyz").split()]
    env_base = open(os.path.join(base, "env_base.py")).read()
    hit = open(os.path.join(base, "iiwa_hit_atacom.py")).read()
    num = r"(-?\d+\.?\d*(?:e-?\d+)?)"
    out = dict(
        joints=joints, fixed_joints=fixed,
        tip_offset=[float(v) for v in re.search(r"pino\.SE3\(np\.eye\(3\), np\.array\(\[%s, %s, %s\]\)\)" % (num, num, num),
                                                env_base).groups()],
        base_translate=[float(v) for v in re.search(r"translate = \[%s, %s, %s\]" % (num, num, num), env_base).groups()],
        table_length=float(re.search(r'"length": %s' % num, env_base).group(1)),
        table_width=float(re.search(r'"width": %s' % num, env_base).group(1)),
        mallet_radius=float(re.search(r"\['mallet'\] = \{\"radius\": %s" % num, env_base).group(1)),
        universal_height=float(re.search(r"\['universal_height'\] = %s" % num, env_base).group(1)),
        z_link4_min=float(re.search(r"g_4 = -ee_pos_4\[2\] \+ %s" % num, hit).group(1)),
        z_link7_min=float(re.search(r"g_5 = -ee_pos_7\[2\] \+ %s" % num, hit).group(1)),
        frame_idx_4=int(re.search(r"self\.frame_idx_4 = (\d+)", hit).group(1)),
        frame_idx_7=int(re.search(r"self\.frame_idx_7 = (\d+)", hit).group(1)),
        K_f=float(re.search(r"b=self\.ee_pos_b_f, K=%s" % num, hit).group(1)),
        K_g_cart=float(re.search(r"b=self\.ee_pos_b_g, K=%s" % num, hit).group(1)),
        K_g_joint=float(re.search(r"b=self\.joint_pos_b_g, K=%s" % num, hit).group(1)),
        acc_max=float(re.search(r"acc_max = np\.ones\(base_env\.n_ctrl_joints\) \* %s" % num, hit).group(1)),
        Kq_factor=float(re.search(r"Kq=%s \* acc_max / vel_max" % num, hit).group(1)),
        Kc_default=float(re.search(r"Kc=%s, random_init" % num, hit).group(1)),
    )
    return out


if __name__ == "__main__":
    data = extract(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    path = os.path.join(HERE, "iiwa_urdf_constants.json")
    json.dump(data, open(path, "w"), indent=1)
    print("wrote", path)
