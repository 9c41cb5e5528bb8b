"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/*.h declares, the PODs have the reference's layout, and a plain-C client written like the
reference's CLI (reference src/main.c) compiles with the reference's warning flags and links.
No compute call is made here (no GPU)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def _declared_functions(path):
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)
    src = re.sub(r"\{[^{}]*\}", "{}", src)     # drop struct / inline bodies (no nesting in these headers)
    names = []
    for decl in src.split(";"):
        decl = " ".join(decl.split())
        if "(" not in decl or "static inline" in decl or decl.startswith("typedef"):
            continue
        m = re.search(r"(\w+)\s*\(", decl)
        if m:
            names.append(m.group(1))
    return names


def test_every_declared_symbol_is_exported(sim):
    L = sim.lib()
    declared = (_declared_functions(os.path.join(INC, "sslsim.h")) + _declared_functions(os.path.join(INC, "sslsim_batch.h")) +
                _declared_functions(os.path.join(INC, "serialize.h")) + _declared_functions(os.path.join(INC, "sslsim_multi.h")) +
                _declared_functions(os.path.join(INC, "sslsim_ingest.h")))
    assert len(declared) >= 37 + 30 + 5 + 15
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, missing
    bound = {n for n, _, _ in sim.LEGACY_SYMBOLS + sim.BATCH_SYMBOLS + sim.WIRE_SYMBOLS + sim.MULTI_SYMBOLS + sim.INGEST_SYMBOLS}
    assert set(declared) <= bound, set(declared) - bound      # the Python mirror binds all of them
    for d in sim.DATA_SYMBOLS:
        sim.FieldGeometry.in_dll(L, d)


def test_reference_api_surface_is_complete(sim):
    """The 37 functions + 3 constants of reference src/world.h:21-46, src/ball.h:18-40, src/robot.h:19-33, src/field.h:44-46."""
    ref = """new_world clone_world delete_world world_step world_step_delta world_get_field world_ball_count
    world_get_ball world_get_game_ball world_get_mut_ball world_get_mut_game_ball world_robot_count world_get_robot
    world_get_mut_robot world_add_robot world_get_frame_number world_get_timestamp world_bt_dynamics
    ball_get_vec ball_set_vec ball_get_pos ball_set_pos ball_get_vel ball_set_vel ball_is_touching_robot
    ball_last_touching_robot ball_get_speed2 ball_get_peak_speed2_from_last_kick ball_bt_rigid_body
    get_id get_team robot_get_pos robot_set_pos robot_get_vel robot_set_vel robot_is_touching_robot
    robot_get_speed2""".split()
    assert len(ref) == 37
    L = sim.lib()
    for n in ref + ["FIELD_2014_SINGLE", "FIELD_2014_DOUBLE", "FIELD_2015"]:
        assert hasattr(L, n), n


def test_pod_layouts_and_field_constants(sim):
    assert C.sizeof(sim.FieldGeometry) == 60          # 15 floats, reference src/field.h:18-34
    assert C.sizeof(sim.Vec2) == 8 and C.sizeof(sim.Vec3) == 12 and C.sizeof(sim.Pos2) == 12 and C.sizeof(sim.Pos3) == 24
    assert C.sizeof(sim.SimParams) == 64 and sim.CMD_DTYPE.itemsize == 32
    assert sim.ROBOT_REPLACEMENT_DTYPE.itemsize == 20 and sim.BALL_REPLACEMENT_DTYPE.itemsize == 20
    f = sim.field("FIELD_2015")                       # reference src/field.c:47-63
    assert (f.field_length, f.field_width, f.goal_width) == (9.0, 6.0, 1.0)
    assert f.boundary_width == C.c_float(0.25).value and f.referee_width == C.c_float(0.45).value
    s = sim.field("FIELD_2014_SINGLE")                # src/field.c:11-27
    assert (s.field_length, s.field_width) == (C.c_float(6.05).value, C.c_float(4.05).value)
    d = sim.field("FIELD_2014_DOUBLE")                # src/field.c:29-45
    assert d.field_length == C.c_float(8.09).value and d.defense_radius == 1.0
    p = sim.default_params()
    assert p.solver_iterations == 10 and p.flags == 0


def test_fields_match_oracle_binding(sim, orc):
    for name in ("FIELD_2014_SINGLE", "FIELD_2014_DOUBLE", "FIELD_2015", "FIELD_DIV_A"):
        assert sim.field(name).as_tuple() == tuple(C.c_float(v).value for v in getattr(orc, name))
    po, ps = orc.default_params(), sim.default_params()
    assert bytes(po) == bytes(ps)


def test_c_client_compiles_and_links(sim, tmp_path):
    """C11, the reference's warning flags (reference CMakeLists.txt:11), headers from include/, link against the .so."""
    exe = tmp_path / "c_api_client"
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-pedantic-errors", "-Werror", "-I", INC,
           os.path.join(ROOT, "tests", "c_api_client.c"), "-o", str(exe), sim.LIB_PATH,
           "-Wl,-rpath," + os.path.dirname(sim.LIB_PATH)]
    subprocess.run(cmd, check=True)
    assert exe.exists()
    # the per-topic forwarding headers compile stand-alone as C and as C++
    for h in ("world.h", "robot.h", "ball.h", "team.h", "vector.h", "field.h", "sslsim.h", "sslsim_batch.h", "serialize.h"):
        for comp, std in (("gcc", "-std=c11"), ("g++", "-std=c++11")):
            subprocess.run([comp, std, "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", INC,
                            "-x", "c" if comp == "gcc" else "c++", os.path.join(INC, h)], check=True)


def test_no_device_fails_loudly(sim):
    """Without a CUDA device the product path must fail, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sim.SslSimError):
        sim.WorldBatch(4, 12)
    with pytest.raises(sim.SslSimError):
        sim.World()


def test_null_handles_are_argument_errors(sim):
    """The batched calls return SSLSIM_E_ARG (-1) on a NULL batch instead of touching it (no GPU needed)."""
    L = sim.lib()
    assert L.world_batch_step(None, 1) == -1
    assert L.world_batch_sync(None) == -1
    assert L.world_batch_set_chunks(None, 4) == -1
    assert L.world_batch_chunk_step(None, 0, None, 1, None) == -1
    assert L.world_batch_chunk_wait(None, 0) == -1
    assert L.world_batch_chunk_count(None) == 0
    assert L.world_batch_chunk_range(None, 0, None, None) == -1
    assert L.world_batch_read_aux(None, None) == -1
    assert L.world_batch_world_count(None) == 0 and L.world_batch_device(None) == -1
    assert L.world_batch_last_error(None) == -1
    assert L.world_batch_error_string(None) == b"null batch"


def test_product_does_not_touch_oracle():
    """Nothing under ssl-sim_b200/ or include/ may reference the oracle."""
    for base in ("ssl-sim_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".h", ".cpp", ".c", "Makefile")):
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    assert not re.search(r'#include\s*[<"][^>"]*oracle|import\s+oracle|from\s+oracle|\borc_\w+\s*\(|'
                                         r'libssl_oracle|warp_emu|step_kernel_emu', txt.replace("tests/warp_emu", "")), (dp, fn)
                    # the only run-time loading the product does is NCCL (sslsim_multi.h): every dlopen names libnccl
                    for call in re.findall(r'dlopen\s*\(([^;]*);', txt):
                        assert fn == "multi.cu" and '"libnccl' in call, (dp, fn, call)
