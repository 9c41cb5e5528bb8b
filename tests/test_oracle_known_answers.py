"""CPU tests that pin the oracle (oracle/ssl_oracle.c) to analytic known answers derived from the
reference's scene constants (reference src/sslsim.cpp:17-28, 47, 87, 141, 240; src/field.c:47-63).

The reference ships no tests or golden vectors and its step is un-vendored Bullet (SURVEY 8c:
"parity unpinned"), so these closed-form checks are what anchors the oracle.
"""
import ctypes as C

import numpy as np
import pytest

G = np.float32(9.80665)
H = np.float32(1.0 / 60 / 2)
_fmaf = C.CDLL("libm.so.6").fmaf            # an independent fused multiply-add (STEP_SPEC section 0)
_fmaf.restype = C.c_float
_fmaf.argtypes = [C.c_float] * 3


def fma32(a, b, c):
    return np.float32(_fmaf(float(a), float(b), float(c)))


def _one_world(orc, R=0, **kw):
    w = orc.Worlds(1, R, **kw)
    return w


def test_free_fall_semi_implicit_euler(orc):
    """Ball dropped from 0.5 m (reference src/sslsim.cpp:22, :251): v -= g h, z += v h each 1/120 s substep."""
    w = _one_world(orc)
    z, v = np.float32(0.5), np.float32(0.0)
    for n in range(30):
        w.step(1, substeps=1)
        v = np.float32(v + (np.float32(0.0) - np.float32(G * H)))
        z = fma32(v, H, z)       # position integration is one fused multiply-add (STEP_SPEC 0, 4.7)
        assert w.vel[0, 0, 2] == v and w.pos[0, 0, 2] == z, n
    # closed form: z_n = 0.5 - g h^2 n(n+1)/2
    assert abs(float(z) - (0.5 - 9.80665 * float(H) ** 2 * 30 * 31 / 2)) < 1e-5


def test_ball_first_bounce_restitution_half(orc):
    """ball (0.5) x ground (1.0) restitution by Bullet's product rule -> rebound speed = 0.5 x impact speed."""
    w = _one_world(orc)
    prev_v = 0.0
    for n in range(200):
        w.step(1, substeps=1)
        v = float(w.vel[0, 0, 2])
        if v > 0:
            break
        prev_v = v
    assert prev_v < -2.5            # sqrt(2 g 0.48) ~ 3.07 m/s
    assert v == pytest.approx(-0.5 * prev_v, rel=0.03)


def test_robot_first_bounce_restitution(orc):
    """robot (0.2) x ground (1.0) -> 0.2, plus Bullet's ERP term (0.2 x penetration / h) when the robot
    is first seen already penetrating (robots get no predictive contact: only the ball has CCD upstream,
    src/sslsim.cpp:58-59)."""
    w = _one_world(orc, R=1)
    prev_v, prev_z = 0.0, 0.5
    for n in range(200):
        w.step(1, substeps=1)
        v = float(w.vel[0, 0, 2])
        if v > 0:
            break
        prev_v, prev_z = v, float(w.pos[0, 0, 2])
    dist = prev_z - 0.025
    expect = -0.2 * prev_v + (0.2 * -dist / float(H) if dist <= 0 else 0.0)
    assert prev_v < -2.5
    assert v == pytest.approx(expect, rel=1e-4)


def test_rest_heights(orc):
    """ball rests at its radius 0.0215 (src/sslsim.cpp:17), the robot origin at the wheel radius 0.025 (:23)."""
    w = _one_world(orc, R=2)
    w.pos[0, 1, :2] = (2.0, 2.0)
    w.pos[0, 0, :2] = (-2.0, -2.0)
    w.step(600)
    assert abs(float(w.pos[0, 2, 2]) - 0.0215) < 1e-4
    assert abs(float(w.pos[0, 0, 2]) - 0.025) < 1e-4 and abs(float(w.pos[0, 1, 2]) - 0.025) < 1e-4
    assert np.abs(w.vel[0, :, :3]).max() < 1e-3


def test_sliding_friction_and_no_spin(orc):
    """Zero inertia upstream (src/sslsim.cpp:46) => the ball slides with mu g = 0.25 x 9.80665 and never spins."""
    w = _one_world(orc)
    w.pos[0, 0, 2] = 0.0215
    w.vel[0, 0, 0] = 2.0
    w.step(30)   # 0.5 s
    assert float(w.vel[0, 0, 0]) == pytest.approx(2.0 - 0.25 * 9.80665 * 0.5, abs=2e-3)
    assert w.pos[0, 0, 3] == 0 and w.vel[0, 0, 3] == 0
    w.step(60)
    assert abs(float(w.vel[0, 0, 0])) < 1e-3       # stopped: 2 / 2.4517 = 0.816 s
    assert float(w.pos[0, 0, 0]) == pytest.approx(2.0 ** 2 / (2 * 0.25 * 9.80665), abs=0.02)


def test_ball_spin_model_rolls(orc):
    """Builder-defined rolling model: sliding turns into rolling at 5/7 v0, then decays with mu_roll g."""
    p = orc.default_params(flags=orc.F_BALL_SPIN)
    w = orc.Worlds(1, 0, params=p)
    w.pos[0, 0, 2] = 0.0215
    w.vel[0, 0, 0] = 2.0
    w.step(40)
    vx, wy = float(w.vel[0, 0, 0]), float(w.vel[0, 0, 3])
    assert vx == pytest.approx(wy * 0.0215, rel=2e-2)       # rolling without slipping: v = omega r
    assert 1.2 < vx < 2.0 * 5 / 7 + 0.01


def test_walls_and_field_limits(orc):
    """FIELD_2015 walls at +-5.2 / +-3.7 (src/field.h:36-42, src/field.c:47-63)."""
    w = _one_world(orc, R=2)
    w.pos[0, 0] = (4.5, 2.0, 0.025, 0.0); w.vel[0, 0, 0] = 5.0
    w.pos[0, 1] = (0.0, -3.0, 0.025, 0.0); w.vel[0, 1, 1] = -5.0
    w.pos[0, 2] = (0.0, 3.0, 0.0215, 0.0); w.vel[0, 2, 1] = 6.0
    mx = my = by = 0.0
    bounced = False
    for _ in range(120):
        w.step(1)
        bounced |= float(w.vel[0, 2, 1]) < -1.0
        mx = max(mx, float(w.pos[0, 0, 0])); my = min(my, float(w.pos[0, 1, 1])); by = max(by, float(w.pos[0, 2, 1]))
    assert 5.2 - 0.09 - 0.02 < mx < 5.2 - 0.09 + 5e-3
    assert -(3.7 - 0.09) - 5e-3 < my < -(3.7 - 0.09) + 0.02
    assert 3.7 - 0.0215 - 0.08 < by < 3.7 - 0.0215 + 5e-3   # reflects within one substep of travel (6 m/s x 1/120 s) of the wall
    assert bounced     # came back off the wall (e = 0.5)


def test_goal_and_out_events(orc):
    w = orc.Worlds(2, 0)
    w.pos[:, 0, 2] = 0.0215
    w.pos[0, 0, :2] = (4.0, 0.1); w.vel[0, 0, 0] = 4.0      # into the +x goal
    w.pos[1, 0, :2] = (-4.0, 1.5); w.vel[1, 0, 0] = -4.0     # past the -x goal line beside the goal: out
    seen = [0, 0]
    for _ in range(120):
        w.step(1)
        for k in range(2):
            seen[k] |= int(w.aux[k, 0]) & (0xF << 12)
    assert seen[0] & (1 << 12) and not seen[0] & (1 << 14)
    assert seen[1] & (1 << 14) and not seen[1] & (3 << 12)
    assert int(w.aux[0, 4]) == 1 and int(w.aux[1, 5]) & 0xFFFF == 1
    # the goal's rear wall (e = 0: goal bodies keep Bullet's default restitution) stops the ball inside the goal
    assert 4.5 < float(w.pos[0, 0, 0]) < 4.5 + 0.18 and abs(float(w.vel[0, 0, 0])) < 0.05


def test_kick_and_touch_bookkeeping(orc):
    w = orc.Worlds(1, 2)
    w.pos[0, 0] = (0.0, 0.0, 0.025, 0.0)
    w.pos[0, 1] = (2.0, 0.0, 0.025, np.pi)
    w.pos[0, 2] = (0.115, 0.0, 0.0215, 0.0)
    w.cmd = np.zeros((1, 2), orc.CMD_DTYPE)
    w.cmd[0, 0]["kickspeedx"] = 3.0
    w.step(1)
    assert int(w.aux[0, 0]) & (1 << 15) and int(w.aux[0, 0]) & 0xFF == 0 and int(w.aux[0, 5]) >> 16 == 1
    peak = np.frombuffer(w.aux[0, 1:2].tobytes(), np.float32)[0]
    assert peak == pytest.approx(9.0, rel=1e-3)
    w.cmd[0, 0]["kickspeedx"] = 0.0
    touched = 0
    for _ in range(80):
        w.step(1)
        touched |= int(w.aux[0, 2])
    assert touched & 2 and int(w.aux[0, 0]) & 0xFF == 1     # the ball reached robot 1 and touched it
    assert int(w.aux[0, 6]) >= 1


def test_wheel_drive_moves_robot(orc):
    """velocity-mode command (veltangent) accelerates the robot along its heading; yaw rate follows velangular."""
    w = orc.Worlds(1, 1)
    w.pos[0, 0] = (0.0, 0.0, 0.025, 0.5)
    w.cmd = np.zeros((1, 1), orc.CMD_DTYPE)
    w.cmd[0, 0]["wheel"] = (1.0, 0.0, 2.0, 0.0)
    w.cmd[0, 0]["flags"] = orc.CMD_DRIVE
    w.step(120)
    vx, vy, wz = (float(w.vel[0, 0, k]) for k in (0, 1, 3))
    assert wz == pytest.approx(2.0, abs=0.2)    # P-controlled wheels: steady-state lag while turning
    assert np.hypot(vx, vy) == pytest.approx(1.0, abs=0.2)


def test_sincos_accuracy(orc):
    L = orc.lib()
    s, c = C.c_float(), C.c_float()
    worst = 0.0
    for x in np.linspace(-7.0, 7.0, 2001, dtype=np.float32):
        L.orc_sincos(C.c_float(float(x)), C.byref(s), C.byref(c))
        worst = max(worst, abs(s.value - np.sin(np.float64(x))), abs(c.value - np.cos(np.float64(x))))
    assert worst < 5e-7


def test_determinism_and_world_independence(orc):
    a = orc.Worlds(8, 12)
    b = orc.Worlds(4, 12, world0=4)
    for p in range(3):
        a.gen_commands(p, 1); b.gen_commands(p, 1)
        a.step(30, threads=3); b.step(30)
    np.testing.assert_array_equal(a.pos[4:], b.pos)
    np.testing.assert_array_equal(a.aux[4:], b.aux)


def test_energy_does_not_grow_without_commands(orc):
    w = orc.Worlds(16, 12)
    def energy():
        m = np.array([2.0] * 12 + [0.046])
        return (0.5 * m * (w.vel[..., :3].astype(np.float64) ** 2).sum(-1) + m * 9.80665 * w.pos[..., 2]).sum(-1)
    e0 = energy()
    for _ in range(20):
        w.step(10)
        e1 = energy()
        assert (e1 <= e0 + 1e-3).all()
        e0 = e1


def _act(w, k=0):
    st = int(w.aux[k, 0])
    return (st >> 18) & 3, st >> 20


def test_ball_goes_to_sleep_after_two_seconds(orc):
    """The reference ball is not DISABLE_DEACTIVATION (src/sslsim.cpp:51-61; robots are, :113): Bullet puts it to
    sleep once it has stayed below 0.8 m/s for more than 2 s (binary32 accumulation of h = 1/120: 241 substeps) and
    no robot shares its island.  Asleep = velocity exactly zero, no gravity, no contact row, no integration."""
    assert orc.lib().orc_sleep_substeps(H) == 241
    w = orc.Worlds(1, 1)
    w.pos[0, 0] = (3.0, 2.0, 0.025, 0.0)        # robot far away, at rest
    w.pos[0, 1] = (0.0, 0.0, 0.0215, 0.0)       # ball at rest on the ground
    w.step(120)                                  # 240 slow substeps: not yet
    assert _act(w) == (0, 240)
    w.step(1, substeps=1)                        # 241st: m_deactivationTime > 2 -> wants deactivation
    assert _act(w) == (1, 241)
    c0 = int(w.aux[0, 7])
    w.step(1, substeps=1)                        # next island build: alone -> asleep, velocities zeroed
    assert _act(w)[0] == 2
    assert (w.vel[0, 1, :3] == 0).all()
    assert int(w.aux[0, 7]) - c0 == 1            # only the robot's ground row was solved
    z = w.pos[0, 1, 2].copy()
    w.step(60)
    assert _act(w)[0] == 2 and w.pos[0, 1, 2] == z and (w.vel[0, 1, :3] == 0).all()


def test_moving_ball_does_not_sleep_and_robot_wakes_it(orc):
    w = orc.Worlds(1, 1)
    w.pos[0, 0] = (3.0, 2.0, 0.025, 0.0)
    w.pos[0, 1] = (0.0, 0.0, 0.0215, 0.0)
    w.vel[0, 1, 0] = 3.0                         # sliding: above 0.8 m/s for ~0.9 s
    w.step(30)
    assert _act(w) == (0, 0)
    w.step(600)
    assert _act(w)[0] == 2                       # stopped, then slept
    # a robot that comes within contact reach of the ball (the pair passes the broadphase: axis distance <= 0.09 +
    # 0.0215 + margin 0.02) wakes it: asleep -> wants, time 0; one that stays just outside does not
    bx = float(w.pos[0, 1, 0])
    w.pos[0, 0, :2] = (bx + 0.14, 0.0)
    w.step(1, substeps=1)
    assert _act(w)[0] == 2
    w.pos[0, 0, :2] = (bx + 0.13, 0.0)
    w.step(1, substeps=1)
    assert _act(w) == (1, 1) and float(w.vel[0, 1, 2]) == 0.0   # awake, but gravity was not applied this step
    w.step(1)
    assert _act(w)[0] == 1 and abs(float(w.vel[0, 1, 2])) < 1e-6
    w.pos[0, 0, :2] = (bx + 1.0, 0.0)            # the robot leaves: the drowsy ball sleeps at once
    w.step(1, substeps=1)
    assert _act(w)[0] == 2


def test_restitution_threshold_is_a_bullet_version_switch(orc):
    """Bullet 2.82 (the spec's column): restitution e*(-vn) for any approach speed.  0.2 (Bullet >= 2.87's
    m_restitutionVelocityThreshold) is available as a parameter."""
    out = []
    for thr in (0.0, 0.2):
        w = orc.Worlds(1, 0, params=orc.default_params(restitution_threshold=thr))
        w.pos[0, 0, 2] = 0.0215 + 0.0005
        w.vel[0, 0, 2] = -0.1                    # slow approach: below 0.2 m/s
        w.step(1, substeps=1)
        out.append(float(w.vel[0, 0, 2]))
    assert out[0] == pytest.approx(0.05, rel=1e-3)      # bounces with e = 0.5
    assert out[1] <= 0.0                                # no restitution below the threshold


def test_readout_yaw_follows_bullets_quaternion_readout(orc):
    """robot_get_pos reports sign(axis.z) * getAngle() of the quaternion btMatrix3x3::getRotation extracts
    (reference src/sslsim.cpp:437-443): yaw wrapped to (-pi, pi], except that (-pi, -2pi/3) comes back as
    2pi - |yaw|; the identity reads -0.0."""
    two_pi = 2 * np.pi
    for yaw, want in [(0.5, 0.5), (5.0, 5.0 - two_pi), (4.0, 4.0), (3.5, 3.5), (-3.0, two_pi - 3.0), (-2.0, -2.0),
                      (-6.0, two_pi - 6.0), (2.5, 2.5), (-2.2, two_pi - 2.2)]:
        assert float(orc.readout_yaw(yaw)) == pytest.approx(want, abs=1e-6), yaw
    z = orc.readout_yaw(0.0)
    assert z == 0.0 and np.signbit(z)


def test_declared_deviation_d4_contacts_use_the_chassis_radius_only(orc):
    """STEP_SPEC section 9, D4 (declared deviation, pinned here so that it cannot change unnoticed): the reference's robot
    is a compound of the chassis cylinder (r = 0.09) and four wheel discs reaching to 0.10 (src/sslsim.cpp:81-98); the
    spec's wall, robot-robot and ball-robot contacts use the chassis cylinder alone.  A robot driven into the +x wall
    with a wheel (mounted at 60 degrees from its heading) pointing straight at the wall is turned back with its axis
    0.09 m from the wall, where Bullet's compound shape would stop it at 0.10 m; two robots pushed together get as close as
    0.18 m axis to axis whatever their headings."""
    w = _one_world(orc, R=3)
    yaw = np.float32(np.deg2rad(60.0))                         # wheel 0 (60 degrees about -z) points along +x
    w.pos[0, 0] = (5.05, 0.0, 0.025, yaw); w.vel[0, 0, 0] = 0.6
    w.pos[0, 1] = (0.0, 1.0, 0.025, 0.0); w.vel[0, 1, 0] = 1.0  # robot 1 runs into robot 2
    w.pos[0, 2] = (0.3, 1.0, 0.025, yaw)
    w.pos[0, 3] = (0.0, -2.0, 0.0215, 0.0)
    min_gap, max_x = 1.0, 0.0
    for _ in range(240):
        w.step(1)
        min_gap = min(min_gap, float(w.pos[0, 2, 0] - w.pos[0, 1, 0]))
        max_x = max(max_x, float(w.pos[0, 0, 0]))
    wall = 9.0 / 2 + 0.25 + 0.45                                # FIELD_2015: field_length / 2 + boundary + referee = 5.2
    assert wall - 0.09 - 4e-3 < max_x < wall - 0.09 + 1e-3      # turned back at the chassis radius (within a substep's travel), not at the 0.10 wheel reach
    assert 0.18 - 3e-3 < min_gap < 0.18 + 8e-3                  # chassis to chassis (0.20 with both wheels in line)


def test_warm_starting_known_answers(orc):
    """STEP_SPEC 4.5a (Bullet's SOLVER_USE_WARMSTARTING, factor 0.85).  warmstart_factor = 0 is the cold-start step of spec
    v2; a body resting on the ground carries the impulse m g h of its ground row in the warm table; a single resting row
    gives the same motion either way (the first update of an isolated row solves it exactly from any start), while a
    bounce off a wall does not (Bullet's artefact: the warm-started friction impulse stays applied once the normal
    impulse has been clamped back to zero, because friction rows are skipped while their normal impulse is zero)."""
    def run(factor, vy, steps=40):
        w = _one_world(orc, R=1, params=orc.default_params(warmstart_factor=factor))
        w.pos[0, 0] = (0.0, 3.55, 0.025, 0.0); w.vel[0, 0] = (0.4, vy, 0.0, 0.0)   # resting on the ground, 0.06 m from the +y wall
        w.pos[0, 1] = (2.0, 0.0, 0.0215, 0.0)
        w.step(steps)
        return w

    assert orc.default_params().warmstart_factor == np.float32(0.85)
    cold, warm = run(0.0, 0.0), run(0.85, 0.0)
    np.testing.assert_allclose(warm.pos, cold.pos, atol=2e-6)        # sliding on the ground only: same physics
    ent = orc.warm_entries(warm.warm[0], 1)
    assert set(ent) == {("ground", 0), ("ground", 1)}                # the two ground rows persist in the table
    gn = np.array([ent[("ground", 0)][0], ent[("ground", 1)][0]], np.uint32).view(np.float32)
    np.testing.assert_allclose(gn, [2.0 * 9.80665 / 120, 0.046 * 9.80665 / 120], rtol=1e-5)   # m g h
    gf = np.array([ent[("ground", 0)][1]], np.uint32).view(np.float32)
    assert abs(gf[0]) <= 0.25 * gn[0] * (1 + 1e-6)                   # inside the friction cone
    # factor 0 with a table is bit for bit the step without one
    ref = _one_world(orc, R=1, params=orc.default_params(warmstart_factor=0.0))
    ref.pos[...] = 0; ref.pos[0, 0] = (0.0, 3.55, 0.025, 0.0); ref.vel[0, 0] = (0.4, 1.5, 0.0, 0.0); ref.pos[0, 1] = (2.0, 0.0, 0.0215, 0.0)
    import ctypes as C
    L = orc.lib()
    L.orc_step.argtypes = [C.POINTER(orc.Field), C.POINTER(orc.Params), C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_float]
    for _ in range(40):
        L.orc_step(C.byref(ref.field), C.byref(ref.params), 1, orc._ptr(ref.pos[0]), orc._ptr(ref.vel[0]), orc._ptr(ref.aux[0]),
                   None, 1, 2, C.c_float(float(orc.H)))
    pc = run(0.0, 1.5)
    np.testing.assert_array_equal(pc.pos, ref.pos); np.testing.assert_array_equal(pc.vel, ref.vel)
    pw = run(0.85, 1.5)                                              # runs into the wall and bounces
    assert np.abs(pw.pos - pc.pos).max() > 1e-4                      # what deviation D2 of spec v2 amounted to
    # while the robot leans on the wall its wall row is in the table under the wall's axis
    lean = _one_world(orc, R=1)
    lean.pos[0, 0] = (0.0, 3.612, 0.025, 0.0); lean.vel[0, 0] = (0.0, 0.5, 0.0, 0.0)
    lean.pos[0, 1] = (2.0, 0.0, 0.0215, 0.0)
    lean.step(1, substeps=1)
    assert ("wall y", 0) in orc.warm_entries(lean.warm[0], 1)
    lean.clear_warm([0])
    assert not orc.warm_entries(lean.warm[0], 1)


def test_warm_start_needs_a_persisting_contact_point(orc):
    """STEP_SPEC 4.5a, Bullet's refreshContactPoints: a point whose anchors came apart by more than the breaking threshold
    (0.02 m along the normal, or across it within one substep = 2.4 m/s of lateral relative velocity) does not persist, so
    its row starts cold.  A ball bouncing off the ground at speed is 4 cm up one substep later: without the rule the
    bounce's friction impulse would be re-applied there and throw the ball backwards."""
    def bounce(factor):
        w = _one_world(orc, R=0, params=orc.default_params(warmstart_factor=factor))
        w.pos[0, 0] = (0.0, 0.0, 0.10, 0.0); w.vel[0, 0] = (-2.4, -0.5, -5.8, 0.0)
        w.step(6)
        return w
    cold, warm = bounce(0.0), bounce(0.85)
    assert cold.vel[0, 0, 2] > 2.0                                    # it bounced
    np.testing.assert_array_equal(warm.pos, cold.pos)                 # and nothing was warm-started on the way
    np.testing.assert_array_equal(warm.vel, cold.vel)
    # a robot sliding into the +y wall: along the wall slower than 2.4 m/s its points persist and the bounce differs
    # from the cold one (see test_warm_starting_known_answers); faster than that they do not, and nothing differs
    def wall(vx, factor):
        w = _one_world(orc, R=1, params=orc.default_params(warmstart_factor=factor))
        w.pos[0, 0] = (0.0, 3.55, 0.025, 0.0); w.vel[0, 0] = (vx, 1.5, 0.0, 0.0)
        w.pos[0, 1] = (2.0, 0.0, 0.0215, 0.0)
        w.step(4)                                                     # (the fast one is still above 2.4 m/s after 4 steps)
        return w
    assert np.abs(wall(0.4, 0.85).pos - wall(0.4, 0.0).pos).max() > 1e-4
    np.testing.assert_array_equal(wall(3.0, 0.85).pos[0, 0], wall(3.0, 0.0).pos[0, 0])
    np.testing.assert_array_equal(wall(3.0, 0.85).vel[0, 0], wall(3.0, 0.0).vel[0, 0])


def test_warm_start_converges_a_stack_faster(orc):
    """what warm starting is for: a robot pushed against a wall by another robot (three coupled rows) is closer to the
    solver's fixed point after the same 10 iterations when the rows start from last substep's impulses"""
    def residual(factor):
        w = _one_world(orc, R=2, params=orc.default_params(warmstart_factor=factor))
        w.pos[0, 0] = (5.2 - 0.09, 0.0, 0.025, 0.0)                  # touching the +x wall (FIELD_2015: 5.2)
        w.pos[0, 1] = (5.2 - 0.09 - 0.18, 0.0, 0.025, 0.0)           # touching robot 0
        w.pos[0, 2] = (0.0, 2.0, 0.0215, 0.0)
        w.cmd = np.zeros((1, 2), orc.CMD_DTYPE)
        w.cmd[0, 1]["wheel"] = (2.0, 0.0, 0.0, 0.0); w.cmd[0, 1]["flags"] = orc.CMD_DRIVE   # drives forward (+x) into robot 0
        w.step(60)
        return float(np.abs(w.vel[0, :2, 0]).max())
    assert residual(0.85) < residual(0.0)                            # the pushed pair is nearer to rest
