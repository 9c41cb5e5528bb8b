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
    for w in range(o.W):   # the warm tables (STEP_SPEC 4.5a) hold the same impulses under the same keys
        if not (int(o.aux[w, 0]) >> 16) & 1:
            assert ob.warm_entries(o.warm[w], o.R) == ob.warm_entries(e.warm[w], e.R), "warm table of world %d %s" % (w, what)


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


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("SSLSIM_EMU_FUZZ_SEEDS", "10"))))   # more: SSLSIM_EMU_FUZZ_SEEDS=300
def test_emulated_kernel_random_fields_tunables_and_commands(seed):
    """The GPU property test (test_gpu_parity.py::test_random_fields_tunables_and_commands) on the emulated kernel:
    random field geometry (narrow and shallow goals included: the goal-solid pre-tests), tunables, robot counts, lane
    mappings, substep schemes, initial velocities and commands -- bit-identical to the oracle."""
    rng = np.random.default_rng(5000 + seed)
    f32 = lambda a, b: float(np.float32(rng.uniform(a, b)))
    R = int(rng.choice([1, 3, 6, 12, 15, 16, 22, 31]))
    W = int(rng.integers(5, 24))
    fld = (0.01, f32(2.0, 12.0), f32(1.5, 9.0), f32(0.05, 0.4), f32(0.0, 0.5), f32(0.1, 1.8), f32(0.05, 0.3),
           f32(0.01, 0.05), 0.5, 1.0, 0.5, 0.2, 1.0, 0.4, f32(0.1, 0.3))
    if fld[5] + 2 * fld[7] + 0.3 > fld[2]:             # the goal has to fit the field width
        fld = fld[:5] + (f32(0.1, 0.6),) + fld[6:]
    kw = dict(flags=int(rng.integers(0, 2)), solver_iterations=int(rng.choice([1, 2, 5, 10, 13])),
              kick_reach=f32(0.0, 0.05), kick_halfwidth=f32(0.01, 0.08), kick_max_height=f32(0.0, 0.1),
              wheel_kp=f32(1.0, 60.0), wheel_fmax=f32(0.5, 10.0), robot_yaw_inertia=f32(0.002, 0.05),
              dribble_kd=f32(0.0, 120.0), dribble_cd=f32(0.0, 20.0), dribble_amax=f32(0.0, 30.0),
              dribble_reach=f32(0.0, 0.06), mu_roll=f32(0.0, 0.2),
              restitution_threshold=float(rng.choice([0.0, 0.2])),
              warmstart_factor=float(rng.choice([0.85, 0.85, 0.0, 0.5, 1.0])))
    o = ob.Worlds(W, R, fld=fld, params=ob.default_params(**kw), seed=77 + seed)
    e = ke.EmuWorlds(W, R, fld=fld, params=ob.default_params(**kw), seed=77 + seed)
    lanes = 32 if (R + 1 <= 16 and rng.integers(0, 2)) else 0
    pos, vel = o.pos.copy(), o.vel.copy()
    pos[:, :, 2] = np.where(rng.random((W, R + 1)) < 0.7, 0.025, pos[:, :, 2]).astype(np.float32)   # most on the ground
    pos[:, R, 2] = np.where(rng.random(W) < 0.7, 0.0215, 0.3).astype(np.float32)
    if rng.integers(0, 2):      # half of the cases: everybody crowds one goal
        L2 = fld[1] / 2
        pos[:, :, 0] = (L2 + rng.uniform(-0.5, 0.35, (W, R + 1))).astype(np.float32) * (1 if rng.integers(0, 2) else -1)
        pos[:, :, 1] = rng.uniform(-fld[5], fld[5], (W, R + 1)).astype(np.float32)
    vel[:, :, :2] = rng.normal(0, 1.5, (W, R + 1, 2)).astype(np.float32)
    vel[:, R, :3] = rng.normal(0, 3.0, (W, 3)).astype(np.float32)
    vel[:, :R, 3] = rng.normal(0, 3.0, (W, R)).astype(np.float32)
    if kw["flags"] == 0:
        pos[:, R, 3] = 0.0; vel[:, R, 3] = 0.0        # no ball spin state in the sliding model
    for w in (o, e):
        w.pos[...] = pos; w.vel[...] = vel
    sub, h = [(2, 1.0 / 120), (1, 1.0 / 60), (3, 1.0 / 200), (4, 1.0 / 240)][int(rng.integers(0, 4))]
    for period in range(4):
        cmd = np.zeros((W, R), ob.CMD_DTYPE)
        mode = rng.integers(0, 3, (W, R))              # passive / wheel speeds / velocity commands
        cmd["flags"] = np.where(mode == 0, 0, np.where(mode == 1, ob.CMD_DRIVE | ob.CMD_WHEELS, ob.CMD_DRIVE))
        cmd["flags"] |= np.where(rng.random((W, R)) < 0.4, ob.CMD_SPINNER, 0).astype(np.uint32)
        cmd["wheel"] = np.where((mode == 1)[..., None], rng.uniform(-60, 60, (W, R, 4)), rng.uniform(-2, 2, (W, R, 4))).astype(np.float32)
        cmd["kickspeedx"] = np.where(rng.random((W, R)) < 0.3, rng.uniform(0, 7, (W, R)), 0).astype(np.float32)
        cmd["kickspeedz"] = np.where(rng.random((W, R)) < 0.3, rng.uniform(0, 4, (W, R)), 0).astype(np.float32)
        o.cmd = cmd.copy(); e.cmd = cmd.copy()
        n = int(rng.integers(1, 40))
        o.step(n, substeps=sub, h=np.float32(h)); e.step(n, substeps=sub, h=np.float32(h), lanes=lanes)
        keep = ((o.aux[:, 0] >> 16) & 1) == 0    # overflowed worlds: state unspecified (STEP_SPEC 4.3)
        what = "seed %d period %d R=%d sub=%d" % (seed, period, R, sub)
        np.testing.assert_array_equal(o.aux[:, 0] >> 16, e.aux[:, 0] >> 16, err_msg=what)
        np.testing.assert_array_equal(o.pos[keep], e.pos[keep], err_msg="pos " + what)
        np.testing.assert_array_equal(o.vel[keep], e.vel[keep], err_msg="vel " + what)
        for w in np.nonzero(keep)[0]:
            assert ob.warm_entries(o.warm[w], R) == ob.warm_entries(e.warm[w], R), "warm table of world %d, %s" % (w, what)
        np.testing.assert_array_equal(o.aux[keep], e.aux[keep], err_msg="aux " + what)
