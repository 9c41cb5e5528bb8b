"""Measured effect of deviation D4 (docs/STEP_SPEC.md section 9): the reference's robot is a btCompoundShape of the
chassis cylinder and four wheel discs (src/sslsim.cpp:81-98); the spec uses the chassis cylinder alone for robot-robot,
ball-robot and robot-wall contacts.  This script steps the config-2 workload on the CPU oracle (test infrastructure)
and counts, per world-step state, the contact rows the spec generates and the rows a wheel disc would add or move:

  wheel disc k of a robot at yaw t: centre 0.095 m from the axis at angle t - a_k (a = 60, 135, -135, -60 deg, rotation
  about (0,0,-1), reference :27-28, :93-97), radial extent [0.09, 0.10], tangential half extent 0.025, z 0..0.05 when
  the robot stands on the ground.  Horizontal tests only (robots and the ball are on the ground in this workload):
  the disc is treated as the 10 mm x 50 mm rectangle it projects to.

    ball-robot:   a wheel row would exist if the ball centre is within BALL_R + margin of a wheel rectangle
    robot-robot:  ... if robot j's chassis circle (r 0.09) is within margin of a wheel rectangle of robot i (either way;
                  wheel-wheel contacts, rarer still, are not counted)
    robot-wall:   ... if the farthest wheel corner towards the wall is within margin of the wall

  `extra`  = the wheel row exists in a state where the spec has no row for that pair  (a contact Bullet would have, we not)
  `moved`  = both exist and the wheel is the closer feature: the binding contact point sits up to 10 mm further out

usage: python tools/wheel_shape_effect.py [worlds] [steps]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 64
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
R = 12
BALL_R, ROBOT_R, MARGIN = 0.0215, 0.09, 0.02
ANG = np.deg2rad([60.0, 135.0, -135.0, -60.0])
F = ob.FIELD_2015
LIMX = F[1] / 2 + F[3] + F[4]
LIMY = F[2] / 2 + F[3] + F[4]


def rect_dist(px, py, cx, cy, yaw):
    """distance from points (px,py) to the 4 wheel rectangles of robots at (cx,cy,yaw): min over wheels. Shapes broadcast."""
    dx, dy = px - cx, py - cy
    best = None
    for a in ANG:
        t = yaw - a
        u = dx * np.cos(t) + dy * np.sin(t)       # radial coordinate in the wheel frame
        v = -dx * np.sin(t) + dy * np.cos(t)      # tangential
        ex = np.maximum(np.maximum(0.09 - u, u - 0.10), 0.0)
        ey = np.maximum(np.abs(v) - 0.025, 0.0)
        d = np.hypot(ex, ey)
        best = d if best is None else np.minimum(best, d)
    return best


def wall_support(yaw, nx, ny):
    """largest extent of the wheel rectangles of a robot along the horizontal direction (nx,ny)"""
    best = None
    for a in ANG:
        t = yaw - a
        for u in (0.09, 0.10):
            for v in (-0.025, 0.025):
                x = u * np.cos(t) - v * np.sin(t)
                y = u * np.sin(t) + v * np.cos(t)
                s = x * nx + y * ny
                best = s if best is None else np.maximum(best, s)
    return best


w = ob.Worlds(W, R, seed=0x5351)
cnt = dict(br_spec=0, br_extra=0, br_moved=0, rr_spec=0, rr_extra=0, rr_moved=0, wall_spec=0, wall_extra=0, wall_moved=0, states=0)
iu = np.triu_indices(R, 1)
period = 0
for step in range(STEPS):
    if step % 30 == 0:
        w.gen_commands(period, 0)
        period += 1
    w.step(1, threads=os.cpu_count() or 1)
    if step < 120:
        continue   # the drop from 0.5 m
    p = w.pos.astype(np.float64)
    x, y, yaw = p[:, :R, 0], p[:, :R, 1], p[:, :R, 3]
    bx, by = p[:, R, 0][:, None], p[:, R, 1][:, None]
    sp = np.linalg.norm(w.vel[:, R, :3].astype(np.float64), axis=1)[:, None]
    bm = MARGIN + sp / 120.0
    # ball-robot
    dc = np.hypot(bx - x, by - y) - ROBOT_R - BALL_R
    dw = rect_dist(bx, by, x, y, yaw) - BALL_R
    spec, wheel = dc <= bm, dw <= bm
    cnt["br_spec"] += int(spec.sum()); cnt["br_extra"] += int((wheel & ~spec).sum()); cnt["br_moved"] += int((spec & (dw < dc)).sum())
    # robot-robot (i<j), wheel of i against chassis of j and the other way round
    xi, yi, ti = x[:, iu[0]], y[:, iu[0]], yaw[:, iu[0]]
    xj, yj, tj = x[:, iu[1]], y[:, iu[1]], yaw[:, iu[1]]
    dc = np.hypot(xi - xj, yi - yj) - 2 * ROBOT_R
    dw = np.minimum(rect_dist(xj, yj, xi, yi, ti), rect_dist(xi, yi, xj, yj, tj)) - ROBOT_R
    spec, wheel = dc <= MARGIN, dw <= MARGIN
    cnt["rr_spec"] += int(spec.sum()); cnt["rr_extra"] += int((wheel & ~spec).sum()); cnt["rr_moved"] += int((spec & (dw < dc)).sum())
    # robot-wall: the four walls
    for nx, ny, dist in ((1, 0, LIMX - x), (-1, 0, x + LIMX), (0, 1, LIMY - y), (0, -1, y + LIMY)):
        sup = wall_support(yaw, nx, ny)
        spec, wheel = dist - ROBOT_R <= MARGIN, dist - sup <= MARGIN
        cnt["wall_spec"] += int(spec.sum()); cnt["wall_extra"] += int((wheel & ~spec).sum()); cnt["wall_moved"] += int((spec & (sup > ROBOT_R)).sum())
    cnt["states"] += W

print("config-2 workload on the oracle: %d worlds x %d world-steps (first 120 skipped), %d world states examined" % (W, STEPS, cnt["states"]))
for k, name in (("br", "ball-robot"), ("rr", "robot-robot"), ("wall", "robot-wall")):
    s, e, m = cnt[k + "_spec"], cnt[k + "_extra"], cnt[k + "_moved"]
    print("%-12s rows in the spec %9d (%.4f per world state)   wheel rows the spec lacks %8d (+%.1f %%)   spec rows whose closest feature is a wheel %8d (%.1f %%)"
          % (name, s, s / cnt["states"], e, 100.0 * e / max(s, 1), m, 100.0 * m / max(s, 1)))
tot_s = cnt["br_spec"] + cnt["rr_spec"] + cnt["wall_spec"] + 13 * cnt["states"]
tot_e = cnt["br_extra"] + cnt["rr_extra"] + cnt["wall_extra"]
print("all rows incl. 13 ground rows per state: %d; missing wheel rows %d = %.3f %% of the contact count" % (tot_s, tot_e, 100.0 * tot_e / tot_s))
