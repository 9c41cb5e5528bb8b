"""The one part of the hot path's scene that the reference can still build here: its field geometry
(reference src/field.c:11-63, src/field.h:18-42).  oracle/Makefile compiles src/field.c
from /root/reference, unmodified and in place, into oracle/_ref/libfield_ref.so; this pins our exported
constants, the FieldGeometry layout and the wall positions derived from them against the real reference,
byte for byte.  (The step itself is Bullet and cannot be built: DESIGN.md, "Parity status".)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libfield_ref.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO) and os.path.exists("/root/reference/src/field.c"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libfield_ref.so not built and no reference tree to build it from")
    L = C.CDLL(REF_SO)
    L.ref_field_limit_x.restype = C.c_float
    L.ref_field_limit_y.restype = C.c_float
    L.ref_field_limit_x.argtypes = [C.c_void_p]
    L.ref_field_limit_y.argtypes = [C.c_void_p]
    return L


def _bytes(lib, name, n=60):
    return bytes((C.c_ubyte * n).in_dll(lib, name))


def test_field_constants_are_the_references_bytes(sim, ref):
    assert ref.ref_field_geometry_size() == C.sizeof(sim.FieldGeometry) == 60
    ours = sim.lib()
    for name in ("FIELD_2014_SINGLE", "FIELD_2014_DOUBLE", "FIELD_2015"):
        assert _bytes(ours, name) == _bytes(ref, name), name


def test_oracle_fields_are_the_references(orc, ref):
    for name in ("FIELD_2014_SINGLE", "FIELD_2014_DOUBLE", "FIELD_2015"):
        want = (C.c_float * 15).in_dll(ref, name)
        assert tuple(C.c_float(v).value for v in getattr(orc, name)) == tuple(want), name


def test_wall_positions_match_field_limit(sim, orc, ref):
    """field_limit_x / field_limit_y (reference src/field.h:36-42) place the four wall planes
    (src/sslsim.cpp:145-168): ours (include/field.h inline, same expression) against the reference's compiled code."""
    for name in ("FIELD_2014_SINGLE", "FIELD_2014_DOUBLE", "FIELD_2015"):
        g = (C.c_float * 15).in_dll(ref, name)
        lx, ly = ref.ref_field_limit_x(C.addressof(g)), ref.ref_field_limit_y(C.addressof(g))
        f = sim.field(name)
        F = np.float32   # the expression of include/sslsim.h:64-69 (= sslsim_derive_field's limit_x / limit_y) in fp32
        assert F(F(F(f.field_length) / F(2)) + F(f.boundary_width)) + F(f.referee_width) == F(lx)
        assert F(F(F(f.field_width) / F(2)) + F(f.boundary_width)) + F(f.referee_width) == F(ly)
    # known answers for the CLI's field (src/main.c:62): walls at +-5.2 and +-3.7
    g = (C.c_float * 15).in_dll(ref, "FIELD_2015")
    assert ref.ref_field_limit_x(C.addressof(g)) == C.c_float(5.2).value
    assert ref.ref_field_limit_y(C.addressof(g)) == C.c_float(3.7).value
