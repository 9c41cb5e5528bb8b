"""ctypes binding of the CPU oracle (oracle/libssl_oracle.so).

Test infrastructure only: imported by tests/, bench.py's cpu_baseline / --impl reference
leg and __graft_entry__.smoke().  Never imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None

FIELD_2014_SINGLE = (0.010, 6.050, 4.050, 0.250, 0.425, 0.700, 0.180, 0.020, 0.500, 0.800, 0.350, 0.200, 0.750, 0.400, 0.165)
FIELD_2014_DOUBLE = (0.010, 8.090, 6.050, 0.250, 0.425, 1.000, 0.180, 0.020, 0.500, 1.000, 0.500, 0.200, 1.000, 0.400, 0.165)
FIELD_2015 = (0.010, 9.000, 6.000, 0.250, 0.450, 1.000, 0.180, 0.020, 0.500, 1.000, 0.500, 0.200, 1.000, 0.400, 0.165)
FIELD_DIV_A = (0.010, 12.0, 9.0, 0.30, 0.40, 1.8, 0.18, 0.02, 0.5, 1.0, 0.5, 0.2, 1.0, 0.4, 0.16)

H = np.float32(1.0 / 60 / 2)  # reference src/sslsim.cpp:309


class Field(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "line_width", "field_length", "field_width", "boundary_width", "referee_width", "goal_width",
        "goal_depth", "goal_wall_width", "center_circle_radius", "defense_radius", "defense_stretch",
        "free_kick_from_defense_dist", "penalty_spot_from_field_line_dist", "penalty_line_from_spot_dist",
        "goal_height")]


class Params(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("solver_iterations", C.c_int32),
                ("kick_reach", C.c_float), ("kick_halfwidth", C.c_float), ("kick_max_height", C.c_float),
                ("wheel_kp", C.c_float), ("wheel_fmax", C.c_float), ("robot_yaw_inertia", C.c_float),
                ("dribble_kd", C.c_float), ("dribble_cd", C.c_float), ("dribble_amax", C.c_float),
                ("dribble_reach", C.c_float), ("mu_roll", C.c_float), ("restitution_threshold", C.c_float),
                ("warmstart_factor", C.c_float), ("reserved", C.c_float * 1)]


CMD_DTYPE = np.dtype([("wheel", np.float32, 4), ("kickspeedx", np.float32), ("kickspeedz", np.float32),
                      ("flags", np.uint32), ("pad", np.uint32)])
CMD_DRIVE, CMD_WHEELS, CMD_SPINNER = 1, 2, 4
F_BALL_SPIN = 1


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def _cpu_has_fma():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("flags"):
                    return " fma " in ln + " "
    except OSError:
        pass
    return False


def lib():
    global _LIB
    if _LIB is None:
        name = "libssl_oracle.so" if _cpu_has_fma() else "libssl_oracle_nofma.so"   # same results, see oracle/Makefile
        path = os.path.join(ORACLE_DIR, name)
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        fp = C.c_void_p
        L.orc_params_default.argtypes = [C.POINTER(Params)]
        L.orc_step.argtypes = [C.POINTER(Field), C.POINTER(Params), C.c_int, fp, fp, fp, fp, C.c_int, C.c_int, C.c_float]
        L.orc_init_world.argtypes = [C.POINTER(Field), C.c_int, C.c_uint64, C.c_uint64, C.c_int, fp, fp, fp]
        L.orc_gen_commands.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, fp]
        L.orc_gather.argtypes = [C.c_int, fp, fp, fp]
        L.orc_gather.restype = C.c_int
        L.orc_referee.argtypes = [C.POINTER(Field), C.c_int, fp, fp, fp]
        L.orc_referee.restype = C.c_int
        L.orc_readout_yaw.argtypes = [C.c_float]
        L.orc_readout_yaw.restype = C.c_float
        L.orc_sleep_substeps.argtypes = [C.c_float]
        L.orc_sleep_substeps.restype = C.c_int
        L.orc_u01.argtypes = [C.c_uint64] * 4
        L.orc_u01.restype = C.c_float
        L.orc_sincos.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_step_many.argtypes = [C.POINTER(Field), C.POINTER(Params), C.c_int, C.c_int, fp, fp, fp, fp,
                                    C.c_int, C.c_int, C.c_float, C.c_int, fp]
        L.orc_warm_words.argtypes = [C.c_int]
        L.orc_warm_words.restype = C.c_int
        _LIB = L
    return _LIB


def readout_yaw(yaw):
    """the orientation robot_get_pos / the detection gather report for a state yaw (Bullet's quaternion readout)"""
    return np.float32(lib().orc_readout_yaw(C.c_float(float(yaw))))


def field(vals=FIELD_2015):
    return Field(*vals)


def default_params(**kw):
    p = Params()
    lib().orc_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Worlds:
    """W independent oracle worlds, world-major arrays pos/vel [W, R+1, 4], aux [W, 8], cmd [W, R]."""

    def __init__(self, n_worlds, n_robots, fld=FIELD_2015, params=None, seed=0x5351, mode=0, world0=0, world_ids=None):
        """world_ids: explicit global world ids (sampled worlds of a larger batch); default world0, world0+1, ..."""
        self.W, self.R = n_worlds, n_robots
        self.ids = [world0 + w for w in range(n_worlds)] if world_ids is None else [int(i) for i in world_ids]
        assert len(self.ids) == n_worlds
        self.field = field(fld)
        self.params = params if params is not None else default_params()
        self.pos = np.zeros((n_worlds, n_robots + 1, 4), np.float32)
        self.vel = np.zeros_like(self.pos)
        self.aux = np.zeros((n_worlds, 8), np.uint32)
        self.cmd = None
        self.seed, self.world0 = seed, world0
        L = lib()
        # the warm table (STEP_SPEC 4.5a): part of the world state; an external write of a world's state clears its row
        self.warm = np.zeros((n_worlds, L.orc_warm_words(n_robots)), np.uint32)
        for w in range(n_worlds):
            L.orc_init_world(C.byref(self.field), n_robots, seed, self.ids[w], mode,
                             _ptr(self.pos[w]), _ptr(self.vel[w]), _ptr(self.aux[w]))

    def gen_commands(self, period, kind=0):
        if self.cmd is None:
            self.cmd = np.zeros((self.W, self.R), CMD_DTYPE)
        L = lib()
        for w in range(self.W):
            L.orc_gen_commands(self.R, self.seed, self.ids[w], period, kind, _ptr(self.cmd[w]))

    def step(self, n_steps=1, substeps=2, h=H, threads=1):
        lib().orc_step_many(C.byref(self.field), C.byref(self.params), self.R, self.W, _ptr(self.pos),
                            _ptr(self.vel), _ptr(self.aux), _ptr(self.cmd), n_steps, substeps,
                            C.c_float(float(h)), threads, _ptr(self.warm))

    def clear_warm(self, worlds=None):
        """what every external state write does on the device (write_state, replacements, referee placement, setters)"""
        if worlds is None:
            self.warm[...] = 0
        else:
            self.warm[worlds] = 0

    def referee(self):
        L = lib()
        n = 0
        for w in range(self.W):
            if L.orc_referee(C.byref(self.field), self.R, _ptr(self.pos[w]), _ptr(self.vel[w]), _ptr(self.aux[w])):
                self.warm[w] = 0
                n += 1
        return n

    def gather(self, w, ids):
        out = np.zeros(6 + 7 * self.R, np.float32)
        ids = np.asarray(ids, np.int32)
        n = lib().orc_gather(self.R, _ptr(self.pos[w]), _ptr(ids), _ptr(out))
        return out[:n]


def warm_entries(row, n_robots):
    """one world's warm table as {key: (normal impulse bits, friction impulse bits)}: the slot / entry order, which
    carries no meaning, drops out; entries whose two impulses are zero are the same as no entry"""
    N = n_robots + 1
    cap = 32 if N <= 16 else 64
    row = np.asarray(row, np.uint32)
    out = {}
    for b in range(N):
        sig = int(row[b])
        for k in range(4):
            sid = (sig >> (8 * k)) & 0xff
            if sid:
                out[("static", b, sid - 1)] = (int(row[N + k * N + b]), int(row[5 * N + k * N + b]))
        out[("ground", b)] = (int(row[9 * N + b]), int(row[10 * N + b]))
        out[("wall x", b)] = (int(row[11 * N + b]), int(row[12 * N + b]))
        out[("wall y", b)] = (int(row[13 * N + b]), int(row[14 * N + b]))
    pt = row[15 * N:]
    for c in range(int(pt[0])):
        ab = int(pt[1 + c])
        out[("pair", ab & 0xff, ab >> 8)] = (int(pt[1 + cap + c]), int(pt[1 + 2 * cap + c]))
    z = lambda u: 0 if (u & 0x7fffffff) == 0 else u          # the sign of a zero is not part of the contract (STEP_SPEC 4.8)
    return {k: (z(v[0]), z(v[1])) for k, v in out.items() if (v[0] & 0x7fffffff) or (v[1] & 0x7fffffff)}
