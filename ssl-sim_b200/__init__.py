"""sslsim_b200 -- Python host-side mirror of the C ABI in include/sslsim.h and
include/sslsim_batch.h (libsslsim_b200.so, hand-written sm_100a kernels).

The directory name carries a hyphen (``ssl-sim_b200``), so import it with
``importlib.import_module("ssl-sim_b200")`` after putting the repo root on
``sys.path`` (tests/conftest.py, bench.py and __graft_entry__.py do this).

This module is a thin ctypes binding: every compute call goes to the CUDA
library.  There is no CPU implementation and no fallback: if the shared library
is missing or no CUDA device is usable, the calls raise.
Names, argument meaning and error behaviour follow the reference's C API
(reference src/world.h, src/robot.h, src/ball.h) and its batched extension.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsslsim_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

TEAM_NONE, TEAM_BLUE, TEAM_YELLOW = -1, 0, 1
CMD_DRIVE, CMD_WHEELS, CMD_SPINNER = 1, 2, 4
F_BALL_SPIN = 1
AUX_STATUS, AUX_PEAK2, AUX_TOUCH, AUX_RESTARTS, AUX_GOALS, AUX_OUTS, AUX_TOUCHES, AUX_CONTACTS = range(8)
MAX_ROBOTS = 31

CMD_DTYPE = np.dtype([("wheel", np.float32, 4), ("kickspeedx", np.float32), ("kickspeedz", np.float32),
                      ("flags", np.uint32), ("pad", np.uint32)])
ROBOT_REPLACEMENT_DTYPE = np.dtype([("world", np.int32), ("robot", np.int32), ("x", np.float32),
                                    ("y", np.float32), ("dir", np.float32)])
BALL_REPLACEMENT_DTYPE = np.dtype([("world", np.int32), ("x", np.float32), ("y", np.float32),
                                   ("vx", np.float32), ("vy", np.float32)])


class FieldGeometry(C.Structure):
    """reference src/field.h:18-34"""
    _fields_ = [(n, C.c_float) for n in (
        "line_width", "field_length", "field_width", "boundary_width", "referee_width", "goal_width",
        "goal_depth", "goal_wall_width", "center_circle_radius", "defense_radius", "defense_stretch",
        "free_kick_from_defense_dist", "penalty_spot_from_field_line_dist", "penalty_line_from_spot_dist",
        "goal_height")]

    def as_tuple(self):
        return tuple(getattr(self, n) for n, _ in self._fields_)


class SimParams(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("solver_iterations", C.c_int32),
                ("kick_reach", C.c_float), ("kick_halfwidth", C.c_float), ("kick_max_height", C.c_float),
                ("wheel_kp", C.c_float), ("wheel_fmax", C.c_float), ("robot_yaw_inertia", C.c_float),
                ("dribble_kd", C.c_float), ("dribble_cd", C.c_float), ("dribble_amax", C.c_float),
                ("dribble_reach", C.c_float), ("mu_roll", C.c_float), ("restitution_threshold", C.c_float),
                ("warmstart_factor", C.c_float), ("reserved", C.c_float * 1)]


class Vec2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Pos2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("w", C.c_float)]


class Pos3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float),
                ("wx", C.c_float), ("wy", C.c_float), ("wz", C.c_float)]


class BatchStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("goals_blue", "goals_yellow", "ball_out", "kicks", "touches",
                                          "contacts", "bad_worlds", "checksum")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


# every symbol include/*.h declares: (name, restype, argtypes)
_VP, _I, _F, _U64 = C.c_void_p, C.c_int, C.c_float, C.c_uint64
LEGACY_SYMBOLS = [
    ("new_world", _VP, [C.POINTER(FieldGeometry)]), ("clone_world", _VP, [_VP]), ("delete_world", None, [_VP]),
    ("world_step", None, [_VP]), ("world_step_delta", None, [_VP, _F, _I, _F]),
    ("world_get_field", C.POINTER(FieldGeometry), [_VP]), ("world_ball_count", _I, [_VP]),
    ("world_get_ball", _VP, [_VP, _I]), ("world_get_game_ball", _VP, [_VP]),
    ("world_get_mut_ball", _VP, [_VP, _I]), ("world_get_mut_game_ball", _VP, [_VP]),
    ("world_robot_count", _I, [_VP]), ("world_get_robot", _VP, [_VP, _I]), ("world_get_mut_robot", _VP, [_VP, _I]),
    ("world_add_robot", None, [_VP, _I, _I]), ("world_get_frame_number", C.c_uint, [_VP]),
    ("world_get_timestamp", C.c_double, [_VP]), ("world_bt_dynamics", _VP, [_VP]),
    ("ball_get_vec", Vec3, [_VP]), ("ball_set_vec", None, [_VP, Vec2]), ("ball_get_pos", Pos3, [_VP]),
    ("ball_set_pos", None, [_VP, Pos3]), ("ball_get_vel", Pos3, [_VP]), ("ball_set_vel", None, [_VP, Pos3]),
    ("ball_is_touching_robot", _I, [_VP, _VP]), ("ball_last_touching_robot", _VP, [_VP]),
    ("ball_get_speed2", _F, [_VP]), ("ball_get_peak_speed2_from_last_kick", _F, [_VP]),
    ("ball_bt_rigid_body", _VP, [_VP]),
    ("get_id", _I, [_VP]), ("get_team", _I, [_VP]), ("robot_get_pos", Pos2, [_VP]),
    ("robot_set_pos", None, [_VP, Pos2]), ("robot_get_vel", Pos2, [_VP]), ("robot_set_vel", None, [_VP, Pos2]),
    ("robot_is_touching_robot", _I, [_VP, _VP]), ("robot_get_speed2", _F, [_VP]),
]
BATCH_SYMBOLS = [
    ("sslsim_params_default", None, [C.POINTER(SimParams)]),
    ("new_world_batch", _VP, [C.POINTER(FieldGeometry), _I, _I, _I, _U64]),
    ("new_world_batch_ex", _VP, [C.POINTER(FieldGeometry), _I, _I, _I, _U64, _U64, _I, C.POINTER(SimParams)]),
    ("delete_world_batch", None, [_VP]), ("world_batch_reset", _I, [_VP, _U64, _I]),
    ("world_batch_world_count", _I, [_VP]), ("world_batch_robot_count", _I, [_VP]),
    ("world_batch_device", _I, [_VP]), ("world_batch_frame_number", C.c_uint, [_VP]),
    ("world_batch_timestamp", C.c_double, [_VP]),
    ("world_batch_step", _I, [_VP, _I]), ("world_batch_step_delta", _I, [_VP, _I, _I, _F]),
    ("world_batch_sync", _I, [_VP]),
    ("world_batch_set_chunks", _I, [_VP, _I]), ("world_batch_chunk_count", _I, [_VP]),
    ("world_batch_chunk_range", _I, [_VP, _I, C.POINTER(_I), C.POINTER(_I)]),
    ("world_batch_chunk_step", _I, [_VP, _I, _VP, _I, _VP]), ("world_batch_chunk_wait", _I, [_VP, _I]),
    ("world_batch_chunk_step_ex", _I, [_VP, _I, _VP, _I, _VP, _VP, _VP]),
    ("world_batch_chunk_step_sequence", _I, [_VP, _I, _VP, _I, _I]),
    ("world_batch_read_worlds", _I, [_VP, _VP, _I, _VP, _VP, _VP]),
    ("world_batch_step_sequence", _I, [_VP, _VP, _I, _I]), ("world_batch_step_sequence_host", _I, [_VP, _VP, _I, _I]),
    ("world_batch_set_commands", _I, [_VP, _VP]), ("world_batch_set_commands_device", _I, [_VP, _VP]),
    ("world_batch_clear_commands", _I, [_VP]), ("world_batch_gen_commands", _I, [_VP, _U64, _I]),
    ("world_batch_gen_commands_into", _I, [_VP, _VP, _U64, _I]),
    ("world_batch_replace_robots", _I, [_VP, _VP, _I]), ("world_batch_replace_balls", _I, [_VP, _VP, _I]),
    ("world_batch_read_state", _I, [_VP, _VP, _VP, _VP]), ("world_batch_write_state", _I, [_VP, _VP, _VP, _VP]),
    ("world_batch_read_aux", _I, [_VP, _VP]),
    ("world_batch_read_warm", _I, [_VP, _VP]), ("world_batch_write_warm", _I, [_VP, _VP]), ("world_batch_warm_words", _I, [_VP]),
    ("world_batch_device_state", _I, [_VP, C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_VP)]),
    ("world_batch_gather_detections", _I, [_VP, _VP, _I, _VP, C.c_size_t]),
    ("world_batch_detection_record_size", C.c_size_t, [_VP]),
    ("world_batch_get_world", _VP, [_VP, _I]), ("world_get_batch", _VP, [_VP, C.POINTER(_I)]),
    ("world_batch_reduce_stats", _I, [_VP, C.POINTER(BatchStats)]),
    ("world_batch_snapshot", _VP, [_VP]), ("world_batch_restore", _I, [_VP, _VP]), ("world_snapshot_free", None, [_VP]),
    ("world_batch_fork", _I, [_VP, _I, _VP, _I]), ("world_batch_referee", _I, [_VP]),
    ("world_batch_stream", _VP, [_VP]), ("world_batch_set_stream", _I, [_VP, _VP]),
    ("world_batch_timer_start", _I, [_VP]), ("world_batch_timer_stop", _I, [_VP, C.POINTER(_F)]),
    ("world_batch_enable_kernel_timing", _I, [_VP, _I]),
    ("world_batch_kernel_time", _I, [_VP, C.POINTER(_F), C.POINTER(_I), _I]),
    ("world_batch_set_lanes_per_world", _I, [_VP, _I]),
    ("sslsim_selftest_math", _I, [_I, _U64, _U64, C.POINTER(_U64), C.POINTER(_U64)]),
    ("world_batch_last_error", _I, [_VP]), ("world_batch_error_string", C.c_char_p, [_VP]),
    ("sslsim_b200_version", C.c_char_p, []),
]
WIRE_SYMBOLS = [
    ("serialize_world", _I, [_VP, _VP, _I]), ("serialize_field", _I, [C.POINTER(FieldGeometry), _VP, _I]),
    ("sslsim_encode_detection", _I, [C.c_uint, C.c_double, C.c_double, _VP, _VP, _I, _VP, _I]),
    ("world_batch_serialize_detections", _I, [_VP, _VP, _I, _VP, C.c_size_t, _VP]),
    ("sslsim_decode_grsim_packet", _I, [_VP, C.c_size_t, _I, _VP, _VP, C.POINTER(_I), _VP, C.POINTER(_I),
                                        C.POINTER(C.c_double)]),
]
_LL = C.c_longlong
MULTI_SYMBOLS = [
    ("sslsim_shard_range", None, [_LL, _I, _I, C.POINTER(_LL), C.POINTER(_I)]),
    ("new_world_multi", _VP, [C.POINTER(FieldGeometry), _LL, _I, _VP, _I, _U64, _I, C.POINTER(SimParams)]),
    ("sslsim_multi_unique_id", _I, [_VP]),
    ("new_world_multi_rank", _VP, [C.POINTER(FieldGeometry), _LL, _I, _I, _I, _I, _VP, _U64, _I, C.POINTER(SimParams)]),
    ("delete_world_multi", None, [_VP]), ("world_multi_rank_count", _I, [_VP]), ("world_multi_local_count", _I, [_VP]),
    ("world_multi_shard", _VP, [_VP, _I]), ("world_multi_shard_rank", _I, [_VP, _I]), ("world_multi_world_count", _LL, [_VP]),
    ("world_multi_step", _I, [_VP, _I]), ("world_multi_gen_commands", _I, [_VP, _U64, _I]),
    ("world_multi_set_commands", _I, [_VP, _VP]), ("world_multi_sync", _I, [_VP]),
    ("world_multi_read_worlds", _I, [_VP, _VP, _I, _VP, _VP, _VP]),
    ("world_multi_allgather_stats", _I, [_VP, _VP, C.POINTER(BatchStats)]),
    ("world_multi_timer_start", _I, [_VP]), ("world_multi_timer_stop", _I, [_VP, C.POINTER(_F), _VP]),
    ("world_multi_error_string", C.c_char_p, [_VP]), ("sslsim_multi_nccl_version", C.c_char_p, []),
]
INGEST_SYMBOLS = [
    ("new_grsim_ingest", _VP, [_VP, _VP, _I, _I, C.c_char_p, C.c_char_p]), ("delete_grsim_ingest", None, [_VP]),
    ("grsim_ingest_poll", _I, [_VP, _I]), ("grsim_ingest_apply", _I, [_VP]), ("grsim_ingest_port", _I, [_VP, _I]),
    ("grsim_ingest_packets", C.c_ulonglong, [_VP]), ("grsim_ingest_bad_packets", C.c_ulonglong, [_VP]),
    ("grsim_ingest_last_timestamp", C.c_double, [_VP, _I]),
]
DATA_SYMBOLS = ["FIELD_2014_SINGLE", "FIELD_2014_DOUBLE", "FIELD_2015", "FIELD_DIV_A"]

_LIB = None


def build(verbose=False):
    """Compile libsslsim_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    subprocess.run(["make", "-s" if not verbose else "-k", "-C", CSRC_DIR], check=True)
    return LIB_PATH


def lib():
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("sslsim_b200: %s is missing; build it with __graft_entry__.build() or "
                               "`make -C ssl-sim_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, res, args in LEGACY_SYMBOLS + BATCH_SYMBOLS + WIRE_SYMBOLS + MULTI_SYMBOLS + INGEST_SYMBOLS:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def encode_detection(frame_number, t_capture, t_sent, record, teams):
    rec = np.ascontiguousarray(record, np.float32)
    tm = np.ascontiguousarray(teams, np.int32)
    buf = np.empty(64 + 48 * (tm.size + 1), np.uint8)
    n = lib().sslsim_encode_detection(frame_number, t_capture, t_sent, _ptr(rec), _ptr(tm), tm.size, _ptr(buf), buf.size)
    if n < 0:
        raise SslSimError("sslsim_encode_detection failed")
    return buf[:n].tobytes()


def serialize_field(fld):
    buf = np.empty(256, np.uint8)
    n = lib().serialize_field(C.byref(fld), _ptr(buf), buf.size)
    if n < 0:
        raise SslSimError("serialize_field failed")
    return buf[:n].tobytes()


def decode_grsim_packet(data, n_robots, cmds=None):
    """-> (cmds [n_robots] CMD_DTYPE, robot replacements, ball replacement or None, timestamp, n decoded)."""
    cmds = np.zeros(n_robots, CMD_DTYPE) if cmds is None else np.ascontiguousarray(cmds, CMD_DTYPE)
    rr = np.zeros(max(n_robots, 1), ROBOT_REPLACEMENT_DTYPE)
    br = np.zeros(1, BALL_REPLACEMENT_DTYPE)
    nrr, hb, ts = C.c_int(rr.size), C.c_int(0), C.c_double(0.0)
    raw = np.frombuffer(bytes(data), np.uint8)
    n = lib().sslsim_decode_grsim_packet(_ptr(raw) if raw.size else None, raw.size, n_robots, _ptr(cmds), _ptr(rr),
                                         C.byref(nrr), _ptr(br), C.byref(hb), C.byref(ts))
    if n < 0:
        raise SslSimError("malformed grSim packet")
    return cmds, rr[:nrr.value], (br[0] if hb.value else None), ts.value, n


def field(name="FIELD_2015"):
    """One of the library's FieldGeometry constants (reference src/field.c:11-63, + FIELD_DIV_A)."""
    return FieldGeometry.in_dll(lib(), name)


def default_params(**kw):
    p = SimParams()
    lib().sslsim_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class SslSimError(RuntimeError):
    pass


class WorldBatch:
    """W independent SSL worlds on one CUDA device (struct WorldBatch)."""

    def __init__(self, n_worlds, n_robots=12, fld="FIELD_2015", device=0, seed=0x5351, first_world_id=0,
                 init_mode=0, params=None):
        L = lib()
        self._field = field(fld) if isinstance(fld, str) else fld
        self._params = params
        self._h = L.new_world_batch_ex(C.byref(self._field), n_worlds, n_robots, device, seed, first_world_id,
                                       init_mode, C.byref(params) if params is not None else None)
        if not self._h:
            raise SslSimError("new_world_batch failed (no CUDA device, bad arguments or out of memory); "
                              "sslsim_b200 has no CPU fallback")
        self.W, self.R, self.N = n_worlds, n_robots, n_robots + 1
        self.seed, self.first_world_id = seed, first_world_id

    @classmethod
    def from_handle(cls, handle, owner, seed=0x5351, first_world_id=0):
        """Wrap a `struct WorldBatch *` owned by something else (a WorldMulti shard): close() does not delete it."""
        L = lib()
        self = cls.__new__(cls)
        self._h, self._owner, self._borrowed = handle, owner, True
        self.W, self.R = L.world_batch_world_count(handle), L.world_batch_robot_count(handle)
        self.N = self.R + 1
        self.seed, self.first_world_id = seed, first_world_id
        return self

    def _ck(self, rc):
        if rc < 0:
            raise SslSimError("sslsim_b200 error %d: %s" % (rc, lib().world_batch_error_string(self._h).decode()))
        return rc

    def close(self):
        if getattr(self, "_h", None):
            if not getattr(self, "_borrowed", False):
                lib().delete_world_batch(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- step ----
    def step(self, n_steps=1):
        self._ck(lib().world_batch_step(self._h, n_steps))

    def step_delta(self, n_steps, substeps, h):
        self._ck(lib().world_batch_step_delta(self._h, n_steps, substeps, C.c_float(float(h))))

    def step_sequence(self, cmds, steps_per_period):
        """cmds: host array [P, W, R] of CMD_DTYPE; one launch runs P periods of steps_per_period world-steps."""
        cmds = np.ascontiguousarray(cmds, dtype=CMD_DTYPE)
        assert cmds.ndim == 3 and cmds.shape[1] * cmds.shape[2] == self.W * self.R
        self._ck(lib().world_batch_step_sequence_host(self._h, _ptr(cmds), cmds.shape[0], steps_per_period))

    def step_sequence_device(self, dev_ptr, n_periods, steps_per_period):
        self._ck(lib().world_batch_step_sequence(self._h, C.c_void_p(dev_ptr), n_periods, steps_per_period))

    def sync(self):
        self._ck(lib().world_batch_sync(self._h))

    # ---- chunked pipeline (closed loop without a batch-wide barrier) ----
    def set_chunks(self, n_chunks):
        self._ck(lib().world_batch_set_chunks(self._h, n_chunks))

    def chunk_range(self, chunk):
        first, count = C.c_int(0), C.c_int(0)
        self._ck(lib().world_batch_chunk_range(self._h, chunk, C.byref(first), C.byref(count)))
        return first.value, count.value

    def chunk_step(self, chunk, host_cmd_ptr, n_steps, host_aux_ptr):
        """queue H2D(commands of the chunk) + n_steps + D2H(aux of the chunk) on the chunk's stream; returns at once"""
        self._ck(lib().world_batch_chunk_step(self._h, chunk, host_cmd_ptr, n_steps, host_aux_ptr))

    def chunk_step_sequence(self, chunk, dev_ptr, n_periods, steps_per_period):
        self._ck(lib().world_batch_chunk_step_sequence(self._h, chunk, C.c_void_p(dev_ptr), n_periods, steps_per_period))

    def read_worlds(self, ids):
        """state of selected worlds: (pos [n, N, 4], vel [n, N, 4], aux [n, 8])"""
        ids = np.ascontiguousarray(ids, np.int32)
        pos = np.empty((ids.size, self.N, 4), np.float32)
        vel = np.empty_like(pos)
        aux = np.empty((ids.size, 8), np.uint32)
        self._ck(lib().world_batch_read_worlds(self._h, _ptr(ids), ids.size, _ptr(pos), _ptr(vel), _ptr(aux)))
        return pos, vel, aux

    def chunk_wait(self, chunk):
        self._ck(lib().world_batch_chunk_wait(self._h, chunk))

    def reset(self, seed=None, init_mode=0):
        self._ck(lib().world_batch_reset(self._h, self.seed if seed is None else seed, init_mode))

    # ---- input ----
    def set_commands(self, cmds):
        """cmds: array [W, R] of CMD_DTYPE (host) or None for passive robots."""
        if cmds is None:
            self._ck(lib().world_batch_clear_commands(self._h))
            return
        cmds = np.ascontiguousarray(cmds, dtype=CMD_DTYPE)
        assert cmds.size == self.W * self.R
        self._ck(lib().world_batch_set_commands(self._h, _ptr(cmds)))

    def set_commands_ptr(self, host_ptr):
        self._ck(lib().world_batch_set_commands(self._h, host_ptr))

    def gen_commands(self, period, kind=0):
        self._ck(lib().world_batch_gen_commands(self._h, period, kind))

    def gen_commands_into(self, dev_ptr, period, kind=0):
        self._ck(lib().world_batch_gen_commands_into(self._h, C.c_void_p(dev_ptr), period, kind))

    def set_commands_device(self, dev_ptr):
        self._ck(lib().world_batch_set_commands_device(self._h, C.c_void_p(dev_ptr)))

    def replace_robots(self, recs):
        recs = np.ascontiguousarray(recs, dtype=ROBOT_REPLACEMENT_DTYPE)
        self._ck(lib().world_batch_replace_robots(self._h, _ptr(recs), recs.size))

    def replace_balls(self, recs):
        recs = np.ascontiguousarray(recs, dtype=BALL_REPLACEMENT_DTYPE)
        self._ck(lib().world_batch_replace_balls(self._h, _ptr(recs), recs.size))

    # ---- output ----
    def read_state(self, want_pos=True, want_vel=True, want_aux=True):
        pos = np.empty((self.W, self.N, 4), np.float32) if want_pos else None
        vel = np.empty((self.W, self.N, 4), np.float32) if want_vel else None
        aux = np.empty((self.W, 8), np.uint32) if want_aux else None
        self._ck(lib().world_batch_read_state(self._h, _ptr(pos), _ptr(vel), _ptr(aux)))
        return pos, vel, aux

    def read_aux_into(self, host_ptr):
        self._ck(lib().world_batch_read_aux(self._h, host_ptr))

    def write_state(self, pos=None, vel=None, aux=None):
        pos = None if pos is None else np.ascontiguousarray(pos, np.float32)
        vel = None if vel is None else np.ascontiguousarray(vel, np.float32)
        aux = None if aux is None else np.ascontiguousarray(aux, np.uint32)
        self._ck(lib().world_batch_write_state(self._h, _ptr(pos), _ptr(vel), _ptr(aux)))

    def read_warm(self):
        """the warm tables (docs/STEP_SPEC.md 4.5a; layout in include/sslsim_batch.h): [W, world_batch_warm_words] u32"""
        warm = np.empty((self.W, lib().world_batch_warm_words(self._h)), np.uint32)
        self._ck(lib().world_batch_read_warm(self._h, _ptr(warm)))
        return warm

    def write_warm(self, warm):
        warm = np.ascontiguousarray(warm, np.uint32)
        assert warm.shape == (self.W, lib().world_batch_warm_words(self._h))
        self._ck(lib().world_batch_write_warm(self._h, _ptr(warm)))

    def gather_detections(self, world_ids):
        ids = np.ascontiguousarray(world_ids, np.int32)
        out = np.empty((ids.size, 6 + 7 * self.R), np.float32)
        n = self._ck(lib().world_batch_gather_detections(self._h, _ptr(ids), ids.size, _ptr(out), out.nbytes))
        assert n == out.nbytes
        return out

    def serialize_detections(self, world_ids):
        """One SSL_WrapperPacket{detection} per observed world (bytes objects)."""
        ids = np.ascontiguousarray(world_ids, np.int32)
        out = np.empty(ids.size * (64 + 48 * (self.R + 1)), np.uint8)
        sizes = np.zeros(ids.size, np.int32)
        n = self._ck(lib().world_batch_serialize_detections(self._h, _ptr(ids), ids.size, _ptr(out), out.nbytes, _ptr(sizes)))
        offs = np.concatenate([[0], np.cumsum(sizes)])
        assert offs[-1] == n
        return [out[offs[i]:offs[i + 1]].tobytes() for i in range(ids.size)]

    def referee(self):
        self._ck(lib().world_batch_referee(self._h))

    def snapshot(self):
        h = lib().world_batch_snapshot(self._h)
        if not h:
            raise SslSimError("world_batch_snapshot failed")
        return h

    def restore(self, snap):
        self._ck(lib().world_batch_restore(self._h, snap))

    def free_snapshot(self, snap):
        lib().world_snapshot_free(snap)

    def fork(self, src, dst):
        dst = np.ascontiguousarray(dst, np.int32)
        self._ck(lib().world_batch_fork(self._h, src, _ptr(dst), dst.size))

    def reduce_stats(self):
        st = BatchStats()
        self._ck(lib().world_batch_reduce_stats(self._h, C.byref(st)))
        return st.as_dict()

    def device_pointers(self):
        p = [C.c_void_p() for _ in range(4)]
        self._ck(lib().world_batch_device_state(self._h, *[C.byref(x) for x in p]))
        return tuple(x.value for x in p)

    def get_world(self, index):
        h = lib().world_batch_get_world(self._h, index)
        if not h:
            raise IndexError(index)
        return World(handle=h, owner=self)

    @property
    def frame_number(self):
        return lib().world_batch_frame_number(self._h)

    @property
    def timestamp(self):
        return lib().world_batch_timestamp(self._h)

    # ---- timing / tuning ----
    def timer_start(self):
        self._ck(lib().world_batch_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(lib().world_batch_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def enable_kernel_timing(self, on=True):
        self._ck(lib().world_batch_enable_kernel_timing(self._h, int(on)))

    def kernel_time(self, reset=True):
        ms, n = C.c_float(), C.c_int()
        self._ck(lib().world_batch_kernel_time(self._h, C.byref(ms), C.byref(n), int(reset)))
        return ms.value, n.value

    def set_lanes_per_world(self, lanes):
        self._ck(lib().world_batch_set_lanes_per_world(self._h, lanes))

    def set_stream(self, cuda_stream):
        self._ck(lib().world_batch_set_stream(self._h, C.c_void_p(cuda_stream)))


class GrSimIngest:
    """include/sslsim_ingest.h: UDP receive -> grSim_Packet decode -> command rows / replacements of the steered worlds."""

    def __init__(self, batch, world_ids, base_port, addr="", iface=""):
        ids = np.ascontiguousarray(world_ids, np.int32)
        self._batch = batch
        self._h = lib().new_grsim_ingest(batch._h, _ptr(ids), ids.size, base_port, addr.encode(), iface.encode())
        if not self._h:
            raise SslSimError("new_grsim_ingest failed (bad arguments or socket error)")
        self.n = ids.size

    def port(self, k):
        return lib().grsim_ingest_port(self._h, k)

    def poll(self, timeout_ms=0):
        n = lib().grsim_ingest_poll(self._h, timeout_ms)
        if n < 0:
            raise SslSimError("grsim_ingest_poll failed")
        return n

    def apply(self):
        return self._batch._ck(lib().grsim_ingest_apply(self._h))

    @property
    def packets(self):
        return lib().grsim_ingest_packets(self._h)

    @property
    def bad_packets(self):
        return lib().grsim_ingest_bad_packets(self._h)

    def last_timestamp(self, k):
        return lib().grsim_ingest_last_timestamp(self._h, k)

    def close(self):
        if getattr(self, "_h", None):
            lib().delete_grsim_ingest(self._h)
            self._h = None


def _share_torch_nccl():
    """A Python-process matter, not the library's: PyTorch bundles its own libnccl.so.2 and fails to import if an older
    one with the same soname is already mapped.  The C library binds NCCL lazily by soname (dlopen), so importing torch
    first -- when it is installed -- makes both use the one NCCL.  A C/C++ host has no such issue."""
    try:
        import torch  # noqa: F401
    except Exception:
        pass


class WorldMulti:
    """include/sslsim_multi.h: a batch sharded over several GPUs by contiguous world range, no per-step collective,
    ncclAllGather of the episode statistics.  devices=[...] = one process driving those GPUs;
    rank/n_ranks/device/nccl_id = one process per GPU (nccl_id: the 128 bytes of WorldMulti.unique_id() from rank 0)."""

    def __init__(self, total_worlds, n_robots=12, fld="FIELD_2015", devices=None, seed=0x5351, init_mode=0, params=None,
                 rank=None, n_ranks=None, device=None, nccl_id=None):
        _share_torch_nccl()
        L = lib()
        self._field = field(fld) if isinstance(fld, str) else fld
        pp = C.byref(params) if params is not None else None
        if rank is None:
            devs = np.ascontiguousarray(devices if devices is not None else [0], np.int32)
            self._h = L.new_world_multi(C.byref(self._field), total_worlds, n_robots, _ptr(devs), devs.size, seed, init_mode, pp)
        else:
            buf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
            self._h = L.new_world_multi_rank(C.byref(self._field), total_worlds, n_robots, rank, n_ranks, device, buf, seed, init_mode, pp)
        if not self._h:
            raise SslSimError("new_world_multi failed (no CUDA device, NCCL missing or bad arguments); no CPU fallback")
        self.total, self.R, self.N = total_worlds, n_robots, n_robots + 1
        self.n_ranks, self.n_local = L.world_multi_rank_count(self._h), L.world_multi_local_count(self._h)

    @staticmethod
    def unique_id():
        _share_torch_nccl()
        buf = (C.c_char * 128)()
        if lib().sslsim_multi_unique_id(buf) < 0:
            raise SslSimError("ncclGetUniqueId failed")
        return bytes(buf)

    def _ck(self, rc):
        if rc < 0:
            raise SslSimError("sslsim_b200 multi error %d: %s" % (rc, lib().world_multi_error_string(self._h).decode()))
        return rc

    def close(self):
        if getattr(self, "_h", None):
            lib().delete_world_multi(self._h)
            self._h = None

    def shard_handle(self, i):
        return lib().world_multi_shard(self._h, i)

    def shard(self, i=0, seed=0x5351):
        """local shard i as a WorldBatch (borrowed: every per-batch call works on it)"""
        L = lib()
        first, count = C.c_longlong(0), C.c_int(0)
        L.sslsim_shard_range(self.total, L.world_multi_shard_rank(self._h, i), self.n_ranks, C.byref(first), C.byref(count))
        return WorldBatch.from_handle(L.world_multi_shard(self._h, i), self, seed=seed, first_world_id=first.value)

    def step(self, n_steps=1):
        self._ck(lib().world_multi_step(self._h, n_steps))

    def gen_commands(self, period, kind=0):
        self._ck(lib().world_multi_gen_commands(self._h, period, kind))

    def set_commands(self, cmds):
        cmds = None if cmds is None else np.ascontiguousarray(cmds, dtype=CMD_DTYPE)
        self._ck(lib().world_multi_set_commands(self._h, _ptr(cmds)))

    def sync(self):
        self._ck(lib().world_multi_sync(self._h))

    def read_worlds(self, ids):
        ids = np.ascontiguousarray(ids, np.int64)
        pos = np.zeros((ids.size, self.N, 4), np.float32)
        vel = np.zeros_like(pos)
        aux = np.zeros((ids.size, 8), np.uint32)
        n = self._ck(lib().world_multi_read_worlds(self._h, _ptr(ids), ids.size, _ptr(pos), _ptr(vel), _ptr(aux)))
        return pos, vel, aux, n

    def allgather_stats(self):
        per = (BatchStats * self.n_ranks)()
        tot = BatchStats()
        self._ck(lib().world_multi_allgather_stats(self._h, per, C.byref(tot)))
        return [p.as_dict() for p in per], tot.as_dict()

    def timer_start(self):
        self._ck(lib().world_multi_timer_start(self._h))

    def timer_stop(self):
        mx = C.c_float()
        per = (C.c_float * self.n_local)()
        self._ck(lib().world_multi_timer_stop(self._h, C.byref(mx), per))
        return mx.value, list(per)


class World:
    """The reference's single-world API (struct World *), device-backed."""

    def __init__(self, fld="FIELD_2015", handle=None, owner=None):
        L = lib()
        self._owner = owner
        if handle is None:
            self._field = field(fld) if isinstance(fld, str) else fld
            handle = L.new_world(C.byref(self._field))
            if not handle:
                raise SslSimError("new_world failed (no CUDA device?); sslsim_b200 has no CPU fallback")
            self._owned = True
        else:
            self._owned = False
        self._h = handle

    def close(self):
        if self._owned and self._h:
            lib().delete_world(self._h)
        self._h = None

    def clone(self):
        h = lib().clone_world(self._h)
        if not h:
            raise SslSimError("clone_world failed")
        w = World(handle=h)
        w._owned = True
        return w

    def full_state(self):
        """(pos [N, 4], vel [N, 4], aux [8], warm [words]) of this world through world_get_batch: what the reference's
        per-world getters do not expose (a robot's height and vertical velocity, the aux words, the warm table)"""
        L = lib()
        idx = _I(0)
        bh = L.world_get_batch(self._h, C.byref(idx))
        W, R = L.world_batch_world_count(bh), L.world_batch_robot_count(bh)
        ids = np.array([idx.value], np.int32)
        pos = np.empty((1, R + 1, 4), np.float32); vel = np.empty_like(pos); aux = np.empty((1, 8), np.uint32)
        if L.world_batch_read_worlds(bh, _ptr(ids), 1, _ptr(pos), _ptr(vel), _ptr(aux)):
            raise SslSimError("world_batch_read_worlds failed")
        warm = np.empty((W, L.world_batch_warm_words(bh)), np.uint32)
        if L.world_batch_read_warm(bh, _ptr(warm)):
            raise SslSimError("world_batch_read_warm failed")
        return pos[0], vel[0], aux[0], warm[idx.value]

    def step(self):
        lib().world_step(self._h)

    def step_delta(self, time_step, max_substeps, fixed_time_step):
        lib().world_step_delta(self._h, time_step, max_substeps, fixed_time_step)

    def add_robot(self, rid, team):
        lib().world_add_robot(self._h, rid, team)

    @property
    def robot_count(self):
        return lib().world_robot_count(self._h)

    @property
    def ball_count(self):
        return lib().world_ball_count(self._h)

    @property
    def frame_number(self):
        return lib().world_get_frame_number(self._h)

    @property
    def timestamp(self):
        return lib().world_get_timestamp(self._h)

    def robot(self, i):
        return lib().world_get_mut_robot(self._h, i)

    def ball(self):
        return lib().world_get_mut_game_ball(self._h)

    def robot_pos(self, i):
        p = lib().robot_get_pos(self.robot(i))
        return (p.x, p.y, p.w)

    def robot_vel(self, i):
        p = lib().robot_get_vel(self.robot(i))
        return (p.x, p.y, p.w)

    def serialize(self):
        buf = np.empty(10240, np.uint8)     # BUFFER_SIZE of reference src/main.c:74
        n = lib().serialize_world(self._h, _ptr(buf), buf.size)
        if n < 0:
            raise SslSimError("serialize_world failed")
        return buf[:n].tobytes()

    def ball_vec(self):
        v = lib().ball_get_vec(self.ball())
        return (v.x, v.y, v.z)


from .sharding import STAT_KEYS, allgather_stats, shard_range  # noqa: E402,F401
