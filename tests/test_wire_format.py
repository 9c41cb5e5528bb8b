"""Wire format of the packet path (reference src/serialize.cpp:17-104, protos/*.proto) against python-protobuf:
the hand-rolled proto2 encoder must produce exactly the bytes protobuf produces for the same values, and the
grSim packet decoder must read what protobuf writes.  Host-only: no GPU needed (the GPU leg is in
test_gpu_parity.py::test_serialize_world_and_batch)."""
import numpy as np
import pytest

import ssl_protos


@pytest.fixture(scope="module")
def pb():
    return ssl_protos.build()


def _reference_packet(pb, frame, t_capture, t_sent, record, teams):
    """What reference src/serialize.cpp:17-76 feeds to protobuf, built with python-protobuf."""
    w = pb["SSL_WrapperPacket"]()
    d = w.detection
    d.frame_number, d.t_capture, d.t_sent, d.camera_id = frame, t_capture, t_sent, 0
    b = d.balls.add()
    b.confidence, b.x, b.y, b.z, b.pixel_x, b.pixel_y = (float(v) for v in record[:6])
    for i, team in enumerate(teams):
        r = record[6 + 7 * i: 13 + 7 * i]
        m = d.robots_blue.add() if team == 0 else d.robots_yellow.add()
        m.confidence, m.robot_id, m.x, m.y, m.orientation, m.pixel_x, m.pixel_y = \
            float(r[0]), int(r[1]), float(r[2]), float(r[3]), float(r[4]), float(r[5]), float(r[6])
    return w


def test_detection_packet_bytes_match_protobuf(sim, pb):
    rng = np.random.default_rng(3)
    for R in (0, 1, 12, 22):
        teams = [i & 1 for i in range(R)]
        rec = rng.uniform(-5, 5, 6 + 7 * R).astype(np.float32)
        rec[0] = 1.0
        for i in range(R):
            rec[6 + 7 * i] = 1.0
            rec[7 + 7 * i] = i // 2
        rec[10:11] = -0.0 if R else rec[10:11]                 # the reference's yaw-0 quirk travels as -0.0
        mine = sim.encode_detection(1234567, 20.5 + R, 1.7e9 + 0.25, rec, teams)
        ref = _reference_packet(pb, 1234567, 20.5 + R, 1.7e9 + 0.25, rec, teams)
        assert mine == ref.SerializeToString(), R
        back = pb["SSL_WrapperPacket"].FromString(mine)
        assert len(back.detection.robots_blue) == (R + 1) // 2 and len(back.detection.robots_yellow) == R // 2
    # size quoted in SURVEY 8a13: ~475 B for 6v6
    assert 440 < len(sim.encode_detection(1, 0.0, 0.0, np.zeros(90, np.float32), [i & 1 for i in range(12)])) < 500


def test_detection_buffer_too_small_returns_minus_one(sim):
    rec = np.zeros(6 + 7 * 12, np.float32)
    tm = np.array([i & 1 for i in range(12)], np.int32)
    buf = np.zeros(100, np.uint8)
    L = sim.lib()
    assert L.sslsim_encode_detection(1, 0.0, 0.0, rec.ctypes.data, tm.ctypes.data, 12, buf.ctypes.data, 100) == -1


def test_geometry_packet_bytes_match_protobuf_with_truncation(sim, pb):
    for name in ("FIELD_2015", "FIELD_2014_SINGLE", "FIELD_2014_DOUBLE", "FIELD_DIV_A"):
        f = sim.field(name)
        w = pb["SSL_WrapperPacket"]()
        g = w.geometry.field
        for fname, _ in sim.FieldGeometry._fields_[:14]:
            setattr(g, fname, int(getattr(f, fname)))          # float -> int32 truncation, as src/serialize.cpp:82-98
        mine = sim.serialize_field(f)
        assert mine == w.SerializeToString(), name
        back = pb["SSL_WrapperPacket"].FromString(mine).geometry.field
        assert back.line_width == 0 and back.field_length == int(f.field_length)


def test_grsim_packet_decode(sim, pb):
    p = pb["grSim_Packet"]()
    p.commands.timestamp, p.commands.isteamyellow = 12.5, True
    c = p.commands.robot_commands.add()
    c.id, c.kickspeedx, c.kickspeedz, c.veltangent, c.velnormal, c.velangular, c.spinner, c.wheelsspeed = 3, 4.5, 1.0, 0.7, -0.2, 1.5, True, False
    c = p.commands.robot_commands.add()
    c.id, c.kickspeedx, c.kickspeedz, c.veltangent, c.velnormal, c.velangular, c.spinner, c.wheelsspeed = 0, 0.0, 0.0, 0.0, 0.0, 0.0, False, True
    c.wheel1, c.wheel2, c.wheel3, c.wheel4 = 10.0, -11.0, 12.0, -13.0
    p.replacement.ball.x, p.replacement.ball.y, p.replacement.ball.vx, p.replacement.ball.vy = 1.0, -2.0, 0.5, 0.25
    r = p.replacement.robots.add()
    r.x, r.y, r.dir, r.id, r.yellowteam = -1.5, 0.5, 3.0, 2, False
    prev = np.zeros(12, sim.CMD_DTYPE)
    prev["kickspeedx"] = 9.0                                     # slots not named by the packet are kept
    cmds, rr, br, ts, n = sim.decode_grsim_packet(p.SerializeToString(), 12, prev)
    assert n == 2 and ts == 12.5
    assert cmds[7]["flags"] == sim.CMD_DRIVE | sim.CMD_SPINNER and tuple(cmds[7]["wheel"][:3]) == (np.float32(0.7), np.float32(-0.2), 1.5)
    assert cmds[7]["kickspeedx"] == 4.5 and cmds[7]["kickspeedz"] == 1.0
    assert cmds[1]["flags"] == sim.CMD_DRIVE | sim.CMD_WHEELS and tuple(cmds[1]["wheel"]) == (10.0, -11.0, 12.0, -13.0)
    assert cmds[0]["kickspeedx"] == 9.0 and cmds[0]["flags"] == 0
    assert len(rr) == 1 and (rr[0]["robot"], rr[0]["x"], rr[0]["y"], rr[0]["dir"]) == (4, -1.5, 0.5, 3.0)
    assert (br["x"], br["y"], br["vx"], br["vy"]) == (1.0, -2.0, 0.5, 0.25)
    # empty and malformed packets
    assert sim.decode_grsim_packet(b"", 12)[4] == 0
    with pytest.raises(sim.SslSimError):
        sim.decode_grsim_packet(p.SerializeToString()[:-3], 12)
    # blue team packet, id beyond the roster is ignored
    q = pb["grSim_Packet"]()
    q.commands.timestamp, q.commands.isteamyellow = 0.0, False
    c = q.commands.robot_commands.add()
    c.id, c.kickspeedx, c.kickspeedz, c.veltangent, c.velnormal, c.velangular, c.spinner, c.wheelsspeed = 40, 0, 0, 1, 0, 0, False, False
    assert sim.decode_grsim_packet(q.SerializeToString(), 12)[4] == 0
