"""Generates tests/golden/*.npz.

The reference has no tests, fixtures or golden vectors and its step (Bullet) cannot be built or imported
here, so these vectors are produced by the repo's own CPU oracle (oracle/ssl_oracle.c) after
it passed the analytic known-answer tests (tests/test_oracle_known_answers.py).  They freeze the
specification: a later change to the oracle or the kernel that alters results fails against them, on the
CPU (oracle vs golden) and on the GPU (CUDA vs golden).  PARITY AGAINST BULLET REMAINS UNPINNED.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_binding as ob  # noqa: E402

CASES = {
    # name: (worlds, robots, field, init_mode, flags, command kind (None = passive), periods, steps per period)
    "config1_reference_mode": (1, 12, "FIELD_2015", 0, 0, None, 4, 150),
    "config2_wheels": (8, 12, "FIELD_2015", 0, 0, 0, 10, 30),
    "config4_kick_dribble": (8, 12, "FIELD_2015", 0, 0, 1, 200, 1),
    "config4_kick_dribble_spin": (8, 12, "FIELD_2015", 0, ob.F_BALL_SPIN, 1, 200, 1),
    "config3_crowded_11v11": (2, 22, "FIELD_DIV_A", 1, 0, 1, 5, 30),
}


def prepare(name, w):
    """Deterministic state edits applied before stepping (both by the generator and by the tests)."""
    if name.startswith("config4"):
        R = w.R
        w.pos[:, :, 2] = 0.025
        w.pos[:, R, 2] = 0.0215
        w.pos[:, 0, 3] = 0.0   # robot 0 faces +x with the ball right in front of its kicker (no trig: portable)
        w.pos[:, R, 0] = w.pos[:, 0, 0] + np.float32(0.12)
        w.pos[:, R, 1] = w.pos[:, 0, 1]


def run(name, stepper=None):
    W, R, fld, mode, flags, kind, periods, spp = CASES[name]
    w = ob.Worlds(W, R, fld=getattr(ob, fld), params=ob.default_params(flags=flags), mode=mode)
    prepare(name, w)
    snaps = []
    for p in range(periods):
        if kind is not None:
            w.gen_commands(p, kind)
        w.step(spp)
        if p in (0, periods // 2, periods - 1):
            snaps.append((p, w.pos.copy(), w.vel.copy(), w.aux.copy()))
    return snaps


if __name__ == "__main__":
    ob.build()
    for name in CASES:
        snaps = run(name)
        out = {}
        for p, pos, vel, aux in snaps:
            out["pos_%d" % p], out["vel_%d" % p], out["aux_%d" % p] = pos, vel, aux
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if k.startswith("pos")})
