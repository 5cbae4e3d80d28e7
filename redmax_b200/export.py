"""Trajectory export in the shape of the reference C++ application's Blender JSON (c++/PCG/src/RigidBodyMain.cpp:747-905,
`exportBrender`): a header naming one object per body (`obj`, `name`, `group`) and one record per frame holding, per body, the
`scale`, `location` and `quat` (w, x, y, z) of its world frame.  Body frames come from the GPU (`rmx_body_frames`); nothing here
computes kinematics."""
from __future__ import annotations

import json

import numpy as np


def rotation_to_quaternion(R):
    """Unit quaternions (w, x, y, z) of rotation matrices R [..., 3, 3] (Shepperd's branch on the largest diagonal term)."""
    R = np.asarray(R, dtype=float)
    out = np.empty(R.shape[:-2] + (4,))
    flatR = R.reshape(-1, 3, 3)
    flat = out.reshape(-1, 4)
    for i, r in enumerate(flatR):
        t = np.trace(r)
        if t > 0:
            s = np.sqrt(t + 1.0) * 2
            q = [0.25 * s, (r[2, 1] - r[1, 2]) / s, (r[0, 2] - r[2, 0]) / s, (r[1, 0] - r[0, 1]) / s]
        else:
            k = int(np.argmax(np.diag(r)))
            a, b = (k + 1) % 3, (k + 2) % 3
            s = np.sqrt(1.0 + r[k, k] - r[a, a] - r[b, b]) * 2
            q = [0.0, 0.0, 0.0, 0.0]
            q[0] = (r[b, a] - r[a, b]) / s
            q[1 + k] = 0.25 * s
            q[1 + a] = (r[a, k] + r[k, a]) / s
            q[1 + b] = (r[b, k] + r[k, b]) / s
        flat[i] = q
    return out


def brender_scene(scene, q_traj, names=None, obj_paths=('cube.obj',), group='bodies', every=1, chart=None):
    """Dict in the reference exporter's layout for one rollout q_traj [nsteps, nr]: {'header': {'objs', 'states'},
    'body': [{'frame': k, name: {'scale', 'location', 'quat'}, ...}, ...]}.  `scale` is the cuboid's side lengths (the exporter
    scales a unit cube), `location` / `quat` the body frame.  chart [nsteps, nspherical]: scene.chart_history(out, b) for a
    rollout that re-parameterised its spherical joints."""
    q_traj = np.asarray(q_traj, dtype=float)
    nb = len(scene.bodies)
    names = names or [(b.name or 'body%d' % i) for i, b in enumerate(scene.bodies)]
    frames = np.arange(0, q_traj.shape[0], every)
    E = scene.body_frames(q_traj[frames], chart=None if chart is None else np.asarray(chart)[frames])  # [nframes, nb, 4, 4]
    quat = rotation_to_quaternion(E[:, :, :3, :3])
    header = {'objs': list(obj_paths), 'states': [{'obj': 0, 'name': names[i], 'group': group} for i in range(nb)]}
    body = []
    for fi, k in enumerate(frames):
        rec = {'frame': int(k)}
        for i in range(nb):
            rec[names[i]] = {'scale': [float(x) for x in scene.bodies[i].sides],
                             'location': [float(x) for x in E[fi, i, :3, 3]],
                             'quat': [float(x) for x in quat[fi, i]]}
        body.append(rec)
    return {'header': header, 'body': body}


def export_brender(scene, q_traj, path, **kw):
    with open(path, 'w') as f:
        json.dump(brender_scene(scene, q_traj, **kw), f)
    return path
