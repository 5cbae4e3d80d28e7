"""Host side of the Euler-chart handling of +redmax/JointSpherical.m: what the reference does between time steps in
`jroot.reparam()` (driverRedMaxBDF2.m:112 -> JointSpherical.reparam_, JointSpherical.m:63-103).  Coordinates only -- a few
3x3 products per chart switch, no dynamics: the rollouts themselves run in the CUDA library under the chart in force
(rmx_scene_desc.chart) and are resumed after a switch with rmx_rollout_resume.

Charts in the reference's numbering 1..12 (JointSpherical.m:5-16): R = R_a(q1) R_b(q2) R_c(q3), body angular velocity
omega = T(q) qdot with T = [(R_b R_c)' e_a, R_c' e_b, e_c].
"""
from __future__ import annotations

import math

import numpy as np

CHARTS = {1: (0, 1, 0), 2: (0, 2, 0), 3: (1, 2, 1), 4: (1, 0, 1), 5: (2, 0, 2), 6: (2, 1, 2),
          7: (0, 1, 2), 8: (0, 2, 1), 9: (1, 2, 0), 10: (1, 0, 2), 11: (2, 0, 1), 12: (2, 1, 0)}
CHART_NAMES = {k: ''.join('XYZ'[i] for i in v) for k, v in CHARTS.items()}  # JointSpherical.getChartName
CHART_XYZ = 7


def _rot(axis, angle):
    c, s = math.cos(angle), math.sin(angle)
    i, j = (axis + 1) % 3, (axis + 2) % 3
    R = np.eye(3)
    R[i, i] = c
    R[i, j] = -s
    R[j, i] = s
    R[j, j] = c
    return R


def chart_R_T(chart, q):
    """R(q) and T(q) of JointSpherical.getEuler (outputs 1 and 5) for one chart."""
    a, b, c = CHARTS[chart]
    Ra, Rb, Rc = _rot(a, q[0]), _rot(b, q[1]), _rot(c, q[2])
    E = np.eye(3)
    T = np.column_stack([Rc.T @ (Rb.T @ E[a]), Rc.T @ E[b], E[c]])
    return Ra @ Rb @ Rc, T


def chart_det(chart, q2):
    """|det T| of a chart as a function of its middle angle: sin for the proper Euler charts 1..6, cos for 7..12.
    Vectorised over q2."""
    return np.abs(np.sin(q2)) if chart <= 6 else np.abs(np.cos(q2))


def chart_inv(chart, R):
    """JointSpherical.getEulerInv (JointSpherical.m:186, :1809-1964); NaNs at gimbal lock."""
    a, b, c3 = CHARTS[chart]
    eps = 1.0 if (b - a) % 3 == 1 else -1.0
    if a == c3:
        c = 3 - a - b
        if not (-1.0 < R[a, a] < 1.0):
            return np.full(3, np.nan)
        return np.array([math.atan2(R[b, a], -eps * R[c, a]), math.acos(R[a, a]), math.atan2(R[a, b], eps * R[a, c])])
    c = c3
    if not (-1.0 < R[a, c] < 1.0):
        return np.full(3, np.nan)
    return np.array([math.atan2(-eps * R[b, c], R[c, c]), math.asin(eps * R[a, c]), math.atan2(-eps * R[a, b], R[a, a])])


def reparam(chart, q, qdot, chart1, q1, qdot1):
    """JointSpherical.reparam_ (JointSpherical.m:63-103) for one joint whose chart has |det T| <= 0.5 at (q, qdot): picks the
    chart with the largest min(|det T(q)|, |det T(q1)|) (first maximum, as MATLAB's max) and re-expresses the current state
    and the BDF2 history in it.  Returns (new chart, q, qdot, q1, qdot1)."""
    R, Told = chart_R_T(chart, q)
    R1, Told1 = chart_R_T(chart1, q1)
    dets = np.zeros((2, 12))
    for k in range(1, 13):
        for row, Rk in ((0, R), (1, R1)):
            qk = chart_inv(k, Rk)
            # closed form, as the reference's detS: the pairs XYX / XZX, YZY / YXY, ZXZ / ZYZ share their middle angle and tie
            # exactly; argmax takes the first, as MATLAB's max
            dets[row, k - 1] = chart_det(k, qk[1]) if np.isfinite(qk).all() else 0.0
    new = int(np.argmax(np.min(dets, axis=0))) + 1
    qn = chart_inv(new, R)
    qdn = np.linalg.solve(chart_R_T(new, qn)[1], Told @ qdot)
    q1n = chart_inv(new, R1)
    qd1n = np.linalg.solve(chart_R_T(new, q1n)[1], Told1 @ qdot1)
    return new, qn, qdn, q1n, qd1n
