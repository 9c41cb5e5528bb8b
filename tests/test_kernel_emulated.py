"""The step kernel's own source (ssl-sim_b200/csrc/step_kernel.cu), compiled for the host and run with its 32
lanes as coroutines (tests/warp_emu/, TEST HARNESS ONLY), against the CPU oracle -- bit for bit.

This is what lets kernel changes be checked where there is no GPU: it exercises the shuffle broadphase, the level
schedule, the register rows and the out-of-line general path exactly as written, and aborts if the lanes of a warp
do not meet at the same collective.  It proves nothing about nvcc's code generation or the hardware; the `-m gpu`
parity tests do that, through the C ABI.  The product never loads this library and has no CPU fallback.
"""
import numpy as np
import pytest

import oracle_binding as ob
import kernel_emu_binding as ke


def _same(o, e, what=""):
    np.testing.assert_array_equal(o.pos, e.pos, err_msg="pos " + what)
    np.testing.assert_array_equal(o.vel, e.vel, err_msg="vel " + what)
    np.testing.assert_array_equal(o.aux, e.aux, err_msg="aux " + what)


def _pair(W, R, **kw):
    return ob.Worlds(W, R, **kw), ke.EmuWorlds(W, R, **kw)


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("lanes", [0, 32])
def test_emulated_kernel_6v6_commands(kind, lanes):
    o, e = _pair(24, 12)
    for p in range(6):
        o.gen_commands(p, kind); e.gen_commands(p, kind)
        o.step(30); e.step(30, lanes=lanes)
        _same(o, e, "period %d" % p)
    assert int(o.aux[:, 7].sum()) > 0


def test_emulated_kernel_reference_mode_drop_and_settle():
    o, e = _pair(6, 12)
    for k in range(4):
        o.step(75); e.step(75)
        _same(o, e, "block %d" % k)


@pytest.mark.parametrize("flags", [0, ob.F_BALL_SPIN])
def test_emulated_kernel_crowded_11v11(flags):
    kw = dict(fld=ob.FIELD_DIV_A, mode=1, params=ob.default_params(flags=flags))
    o, e = _pair(6, 22, **kw)
    for p in range(5):
        o.gen_commands(p, 1); e.gen_commands(p, 1)
        o.step(30); e.step(30)
        keep = ((o.aux[:, 0] >> 16) & 1) == 0    # overflowed worlds: state unspecified (STEP_SPEC 4.3)
        np.testing.assert_array_equal(o.aux[:, 0] >> 16, e.aux[:, 0] >> 16)
        np.testing.assert_array_equal(o.pos[keep], e.pos[keep])
        np.testing.assert_array_equal(o.vel[keep], e.vel[keep])
        np.testing.assert_array_equal(o.aux[keep], e.aux[keep])


def test_emulated_kernel_command_sequence_equals_per_period():
    W, R, P = 8, 12, 4
    o, e = _pair(W, R)
    seq = np.zeros((P, W, R), ob.CMD_DTYPE)
    for p in range(P):
        o.gen_commands(p, 1)
        seq[p] = o.cmd
        o.step(20)
    e.step(P * 20, cmd_seq=seq, steps_per_period=20)
    _same(o, e)


def test_emulated_kernel_kick_and_goal_area():
    """ball in front of a kicker, robots near the goal solids: lean goal-box rows and the general path"""
    o, e = _pair(8, 12, params=ob.default_params(flags=ob.F_BALL_SPIN))
    for w in (o, e):
        R = w.R
        w.pos[:, :, 2] = 0.025
        w.pos[:, R, 2] = 0.0215
        w.pos[:, 0, 3] = 0.0
        w.pos[:, 0, 0] = 4.0 + 0.05 * np.arange(w.W, dtype=np.float32)
        w.pos[:, 0, 1] = 0.3
        w.pos[:, R, 0] = w.pos[:, 0, 0] + np.float32(0.12)
        w.pos[:, R, 1] = w.pos[:, 0, 1]
        w.pos[:, 1, 0] = 4.55; w.pos[:, 1, 1] = 0.45   # next to a goal side wall
    for p in range(120):
        o.gen_commands(p, 1); e.gen_commands(p, 1)
        o.step(1); e.step(1)
    _same(o, e)
    assert int((o.aux[:, 5] >> 16).sum()) > 0   # kicks happened


@pytest.mark.parametrize("flags", [0, ob.F_BALL_SPIN])
def test_emulated_kernel_ball_sleep_and_wake(flags):
    """balls fall asleep 2 s after they stopped (STEP_SPEC 4.0 / 4.8) and roaming robots wake them again"""
    kw = dict(params=ob.default_params(flags=flags))
    o, e = _pair(32, 12, **kw)
    states = set()
    woken = 0
    prev = None
    for p in range(30):
        o.gen_commands(p, 0); e.gen_commands(p, 0)
        o.step(30); e.step(30)
        _same(o, e, "period %d" % p)
        act = (o.aux[:, 0] >> 18) & 3
        states |= set(int(a) for a in act)
        if prev is not None:
            woken += int(((prev == 2) & (act != 2)).sum())
        prev = act
    assert 2 in states            # some ball slept
    assert woken > 0              # and some sleeping ball was woken by a robot


def test_emulated_kernel_restitution_threshold_parameter():
    kw = dict(params=ob.default_params(restitution_threshold=0.2))
    o, e = _pair(8, 12, **kw)
    for p in range(4):
        o.gen_commands(p, 1); e.gen_commands(p, 1)
        o.step(40); e.step(40)
    _same(o, e)
