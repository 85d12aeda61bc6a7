#!/usr/bin/env python3
"""Generate tests/golden/rbd_mujoco.json: rigid-body known answers for the Hunter model computed with the reference's
vendored MuJoCo 3.0.1 binary (/root/reference/mujoco/lib/libmujoco.so.3.0.1) on the reference's MJCF
(mujoco/model/hunter/hunter.xml), converted to the reference's generalised coordinates
q = [p, yaw, pitch, roll, q_j], v = q_dot.

Runs only where /root/reference exists. The MJCF is adjusted in memory (written to /tmp):
  * armature / damping / frictionloss zeroed (hunter.xml:6,59) so that M is the rigid-body inertia matrix,
  * the commented-out 0.01 kg imu_link re-added (hunter.xml:51-54) so the total mass matches the URDF,
  * visual mesh geoms and assets stripped (no STL loading needed).
The MJCF stores inertias as rounded quaternion + diagonal values, so agreement with the URDF-derived oracle is
limited to ~1e-4 relative; tests use 2e-3.
"""
import json, os, re, subprocess, sys
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))


def build_probe():
    out = os.path.join(ROOT, "oracle", "_ref", "mj_probe")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["gcc", "-O1", "-o", out, os.path.join(HERE, "mj_probe.c"), "-I", REF + "/mujoco/include",
                           "-L", REF + "/mujoco/lib", "-l:libmujoco.so.3.0.1", "-Wl,-rpath," + REF + "/mujoco/lib", "-lm"])
    return out


def patched_xml():
    s = open(REF + "/mujoco/model/hunter/hunter.xml").read()
    s = re.sub(r'armature="[^"]*"', 'armature="0"', s)
    s = re.sub(r'damping="[^"]*"', 'damping="0"', s)
    s = re.sub(r'frictionloss="[^"]*"', 'frictionloss="0"', s)
    s = re.sub(r"<asset>.*?</asset>", "", s, flags=re.S)
    s = re.sub(r'<geom class="visual"[^>]*/>', "", s)
    s = s.replace('<!-- <body name="imu_link">', '<body name="imu_link">').replace("</body> -->", "</body>")
    s = re.sub(r'<geom size="0.0075 0.0075 0.002"[^>]*/>', "", s)
    p = "/tmp/hunter_patched.xml"
    open(p, "w").write(s)
    return p


def Rzyx(e):
    z, y, x = e
    cz, sz, cy, sy, cx, sx = np.cos(z), np.sin(z), np.cos(y), np.sin(y), np.cos(x), np.sin(x)
    return np.array([[cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx],
                     [sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx],
                     [-sy, cy * sx, cy * cx]])


def Tmap(e):
    z, y, _ = e
    cz, sz, cy, sy = np.cos(z), np.sin(z), np.cos(y), np.sin(y)
    return np.array([[0, -sz, cz * cy], [0, cz, sz * cy], [1, 0, -sy]])


def quat_from_R(R):
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    x = (R[2, 1] - R[1, 2]) / (4 * w); y = (R[0, 2] - R[2, 0]) / (4 * w); z = (R[1, 0] - R[0, 1]) / (4 * w)
    return np.array([w, x, y, z])


def main():
    probe = build_probe()
    xml = patched_xml()
    rng = np.random.default_rng(20240901)
    lo = np.array([-0.2, -0.5, -0.8, 0, -1.1, -0.5, -1, -1.2, 0, -1.1])
    hi = np.array([0.5, 1, 1.2, 1.5, 1.1, 0.2, 0.5, 0.8, 1.5, 1.1])
    cases = []
    lines = []
    for i in range(12):
        q = np.zeros(16); v = np.zeros(16)
        q[0:3] = rng.uniform(-0.3, 0.3, 3) + [0, 0, 0.63]
        q[3:6] = rng.uniform([-np.pi, -0.5, -0.5], [np.pi, 0.5, 0.5])
        q[6:] = rng.uniform(lo, hi)
        v[:] = rng.uniform(-1, 1, 16)
        if i == 0:
            q = np.array([0, 0, 0.63, 0, 0, 0, 0.1, 0, 0.4, 0.93, 0.53, -0.1, 0, -0.4, 0.93, -0.53]); v[:] = 0
        R = Rzyx(q[3:6]); T = Tmap(q[3:6])
        qpos = np.concatenate([q[0:3], quat_from_R(R), q[6:]])
        qvel = np.concatenate([v[0:3], R.T @ T @ v[3:6], v[6:]])
        cases.append((q, v))
        lines.append(" ".join("%.17g" % a for a in np.concatenate([qpos, qvel])))
    out = subprocess.run([probe, xml], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.splitlines()
    nq, nv, mass = out[0].split(); mass = float(mass)
    res = []
    for (q, v), line in zip(cases, out[1:]):
        a = np.array([float(t) for t in line.split()])
        k = 0
        M = a[k:k + 256].reshape(16, 16); k += 256
        bias = a[k:k + 16]; k += 16
        pos = np.zeros((4, 3)); J = np.zeros((12, 16))
        for s in range(4):
            pos[s] = a[k:k + 3]; k += 3
            J[3 * s:3 * s + 3] = a[k:k + 48].reshape(3, 16); k += 48
        com = a[k:k + 3]; k += 3
        linvel = a[k:k + 3]; k += 3
        angmom = a[k:k + 3]; k += 3
        R = Rzyx(q[3:6]); T = Tmap(q[3:6])
        G = np.eye(16); G[3:6, 3:6] = R.T @ T
        # d/dt(omega_world) at zero euler second derivatives, by central differences of T(e + rates t) rates
        eps = 1e-6
        wdot = (Tmap(q[3:6] + eps * v[3:6]) @ v[3:6] - Tmap(q[3:6] - eps * v[3:6]) @ v[3:6]) / (2 * eps)
        Gdot_v = np.zeros(16); Gdot_v[3:6] = R.T @ wdot
        res.append(dict(q=q.tolist(), v=v.tolist(), M=(G.T @ M @ G).tolist(), nle=(G.T @ (M @ Gdot_v + bias)).tolist(),
                        J=(J @ G).tolist(), cpos=pos.reshape(-1).tolist(), com=com.tolist(),
                        h=np.concatenate([mass * linvel, angmom]).tolist()))
    json.dump(dict(source="MuJoCo 3.0.1 (reference vendored binary) on patched mujoco/model/hunter/hunter.xml", mass=mass, cases=res),
              open(os.path.join(HERE, "rbd_mujoco.json"), "w"))
    print("wrote rbd_mujoco.json; mujoco total mass", mass)


if __name__ == "__main__":
    main()
