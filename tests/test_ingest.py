"""grSim ingest (include/sslsim_ingest.h): python-protobuf grSim_Packets sent over UDP -- loopback
unicast, loopback multicast, and through the reference's own sender (src/utils/net.c compiled in place) -- reach the
command rows and K2 of the steered worlds; the batch then steps bit-equal to the oracle given the same commands and
replacements."""
import os
import socket
import subprocess

import numpy as np
import pytest

import ssl_protos

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NET_SEND = os.path.join(ROOT, "oracle", "_ref", "net_send")


@pytest.fixture(scope="module")
def pb():
    return ssl_protos.build()


def test_ingest_rejects_bad_arguments(sim):
    L = sim.lib()
    assert not L.new_grsim_ingest(None, None, 0, 0, b"", b"")
    assert L.grsim_ingest_poll(None, 0) == -1 and L.grsim_ingest_apply(None) == -1
    assert L.grsim_ingest_port(None, 0) == -1 and L.grsim_ingest_packets(None) == 0


def _packets(pb):
    p3 = pb["grSim_Packet"]()
    p3.commands.timestamp, p3.commands.isteamyellow = 7.25, True
    for rid, vt, vn, va, spin in ((0, 0.8, 0.0, 0.0, False), (2, -0.5, 0.3, 2.0, True)):
        c = p3.commands.robot_commands.add()
        c.id, c.kickspeedx, c.kickspeedz, c.veltangent, c.velnormal, c.velangular, c.spinner, c.wheelsspeed = rid, 0.0, 0.0, vt, vn, va, spin, False
    p3.replacement.ball.x, p3.replacement.ball.y, p3.replacement.ball.vx, p3.replacement.ball.vy = 1.25, -0.75, 2.0, 0.5
    p40 = pb["grSim_Packet"]()
    p40.commands.timestamp, p40.commands.isteamyellow = 9.5, False
    c = p40.commands.robot_commands.add()
    c.id, c.kickspeedx, c.kickspeedz, c.veltangent, c.velnormal, c.velangular, c.spinner, c.wheelsspeed = 1, 3.0, 0.0, 0.0, 0.0, 0.0, False, True
    c.wheel1, c.wheel2, c.wheel3, c.wheel4 = 20.0, -20.0, -20.0, 20.0
    r = p40.replacement.robots.add()
    r.x, r.y, r.dir, r.id, r.yellowteam = -2.0, 1.0, 1.5, 4, True
    return p3, p40


def _send(kind, addr, port, data, tmp_path):
    if kind == "reference_net_c":
        f = tmp_path / ("pkt_%d.bin" % port)
        f.write_bytes(data)
        subprocess.run([NET_SEND, addr, str(port), str(f)], check=True)
        return
    s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    if kind == "multicast":
        s.setsockopt(socket.IPPROTO_IP, socket.IP_MULTICAST_TTL, 1)
        s.setsockopt(socket.IPPROTO_IP, socket.IP_MULTICAST_LOOP, 1)
        try:
            s.setsockopt(socket.IPPROTO_IP, socket.IP_MULTICAST_IF, socket.inet_aton("127.0.0.1"))
        except OSError:
            pass
    s.sendto(data, (addr, port))
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,addr", [("unicast", "127.0.0.1"), ("multicast", "224.5.23.2"), ("reference_net_c", "224.5.23.2")])
def test_ingest_udp_to_device_commands_and_replacements(sim, orc, pb, tmp_path, kind, addr):
    if kind == "reference_net_c" and not os.path.exists(NET_SEND):
        pytest.skip("oracle/_ref/net_send not built (needs the reference tree at build time)")
    W, R = 64, 12
    b = sim.WorldBatch(W, R)
    o = orc.Worlds(W, R)
    b.step(150); o.step(150)                       # passive robots, dropped and settled (reference mode)
    base = 21000 + (os.getpid() % 2000) * 4 + {"unicast": 0, "multicast": 1, "reference_net_c": 2}[kind] * 9000
    try:
        ing = sim.GrSimIngest(b, [3, 40], base, addr=addr, iface="127.0.0.1" if addr.startswith("224") else "")
    except sim.SslSimError:
        if kind == "unicast":
            raise
        pytest.skip("multicast membership is not available in this sandbox")
    assert ing.port(1) == base + 1
    p3, p40 = _packets(pb)
    _send(kind, addr, base, p3.SerializeToString(), tmp_path)
    _send(kind, addr, base + 1, p40.SerializeToString(), tmp_path)
    _send(kind, addr, base + 1, b"\x0a\x05garbage-not-a-packet", tmp_path)   # malformed: counted, ignored
    got = 0
    for _ in range(50):
        got += ing.poll(100)
        if got >= 3:
            break
    if got == 0 and kind != "unicast":
        ing.close(); b.close()
        pytest.skip("multicast loopback delivers nothing in this sandbox")
    assert got == 3 and ing.packets == 3 and ing.bad_packets == 1
    assert ing.last_timestamp(0) == 7.25 and ing.last_timestamp(1) == 9.5
    assert ing.apply() == 2
    # what the same packets mean, decoded on the host (the decoder is checked against python-protobuf in test_wire_format)
    o.cmd = np.zeros((W, R), orc.CMD_DTYPE)
    c3, _, ball, _, n3 = sim.decode_grsim_packet(p3.SerializeToString(), R)
    c40, rr, _, _, n40 = sim.decode_grsim_packet(p40.SerializeToString(), R)
    assert (n3, n40, len(rr)) == (2, 1, 1) and ball is not None
    o.cmd[3] = c3.view(orc.CMD_DTYPE)
    o.cmd[40] = c40.view(orc.CMD_DTYPE)
    assert o.cmd[3, 1]["flags"] & sim.CMD_DRIVE and o.cmd[3, 5]["flags"] & sim.CMD_SPINNER and o.cmd[40, 2]["flags"] & sim.CMD_WHEELS
    o.pos[3, R] = (ball["x"], ball["y"], 0.5, 0.0)           # ball_set_vec semantics: re-dropped from 0.5 m, woken
    o.vel[3, R] = (ball["vx"], ball["vy"], 0.0, 0.0)
    o.aux[3, 0] &= 0x3FFFF
    slot = int(rr[0]["robot"])
    assert slot == 2 * 4 + 1
    o.pos[40, slot] = (rr[0]["x"], rr[0]["y"], 0.5, rr[0]["dir"])
    o.clear_warm([3, 40])                                    # a replacement empties its world's warm table
    before = b.read_worlds([3, 40])[0].copy()
    b.step(90); o.step(90)
    pos, vel, aux = b.read_state()
    np.testing.assert_array_equal(pos, o.pos)
    np.testing.assert_array_equal(vel, o.vel)
    np.testing.assert_array_equal(aux, o.aux)
    # the steered robots moved, the others did not
    assert np.abs(pos[3, 1, :2] - before[0, 1, :2]).max() > 0.2 and np.abs(pos[40, 2, :2] - before[1, 2, :2]).max() > 0.05
    assert np.abs(pos[5, :R, :2] - o.pos[5, :R, :2]).max() == 0 and float(np.abs(vel[5, :R, :2]).max()) < 1e-3
    ing.close(); b.close()
