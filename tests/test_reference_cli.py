"""The reference's own command-line simulator on the B200 library.

oracle/Makefile (`ref`) compiles the reference's src/main.c and src/utils/net.c UNMODIFIED, where they lie, against
include/ (our sslsim.h / serialize.h) and links them with libsslsim_b200.so instead of the reference's
sslsim.cpp + Bullet + protobuf: oracle/_ref/ssl-sim-cli.  The CPU test checks the binary imports the API from our
library; the GPU test runs it (reference src/main.c:48-104: new_world(&FIELD_2015), 6+6 robots, world_step +
serialize_world every 16 ms, serialize_field every 120 steps, multicast to 224.5.23.2:11002), stops it with
SIGINT like a user would and decodes what it multicast with python-protobuf."""
import os
import signal
import socket
import struct
import subprocess
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "oracle", "_ref", "ssl-sim-cli")


def _have_cli():
    if not os.path.exists(CLI) and os.path.exists("/root/reference/src/main.c"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=False)
    return os.path.exists(CLI)


def test_reference_cli_links_against_our_library(sim):
    if not _have_cli():
        pytest.skip("oracle/_ref/ssl-sim-cli not built and no reference tree to build it from")
    dyn = subprocess.run(["readelf", "-d", CLI], capture_output=True, text=True).stdout
    assert "libsslsim_b200.so" in dyn and "Bullet" not in dyn and "protobuf" not in dyn
    und = subprocess.run(["nm", "-D", "--undefined-only", CLI], capture_output=True, text=True).stdout
    for sym in ("new_world", "world_add_robot", "world_step", "world_get_field", "serialize_world", "serialize_field"):
        assert sym in und, sym
        assert hasattr(sim.lib(), sym)


@pytest.mark.gpu
def test_reference_cli_runs_on_the_gpu_and_multicasts_detection_packets(sim):
    if not _have_cli():
        pytest.skip("oracle/_ref/ssl-sim-cli not built")
    import ssl_protos
    pb = ssl_protos.build()
    rx = None
    try:   # the reference's receiver recipe (src/utils/net.c:72-93): reuse address, bind, join the group
        rx = socket.socket(socket.AF_INET, socket.SOCK_DGRAM, socket.IPPROTO_UDP)
        rx.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        rx.bind(("", 11002))
        rx.setsockopt(socket.IPPROTO_IP, socket.IP_ADD_MEMBERSHIP, struct.pack("4sl", socket.inet_aton("224.5.23.2"), socket.INADDR_ANY))
        rx.settimeout(0.2)
    except OSError:
        rx = None                                  # no multicast in this sandbox: the run itself is still checked
    p = subprocess.Popen([CLI], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    packets, t0 = [], time.time()
    while time.time() - t0 < 4.0 and p.poll() is None:
        if rx is None:
            time.sleep(0.2)
            continue
        try:
            packets.append(rx.recv(65536))
        except socket.timeout:
            pass
        if len(packets) >= 40:
            break
    assert p.poll() is None, "ssl-sim-cli died: %s" % p.stderr.read()
    p.send_signal(signal.SIGINT)                   # src/main.c:39-43: Ctrl+C ends the loop
    out, err = p.communicate(timeout=20)
    assert p.returncode == 0, (p.returncode, err)
    assert "Small Size League Simulator by RoboIME" in out
    assert "closing" in err
    if rx is not None:
        rx.close()
    if not packets:
        pytest.skip("the simulator ran and exited cleanly, but this box does not loop multicast back (no packets to decode)")
    frames, geoms = [], 0
    for data in packets:
        w = pb["SSL_WrapperPacket"].FromString(data)
        if w.HasField("detection"):
            frames.append(w.detection)
        if w.HasField("geometry"):
            geoms += 1
            assert w.geometry.field.field_length == 9 and w.geometry.field.field_width == 6   # src/serialize.cpp:78-104 (metres, int32)
    assert len(frames) >= 10
    nums = [f.frame_number for f in frames]
    assert nums == sorted(nums) and len(set(nums)) == len(nums)
    for f in frames:
        assert len(f.balls) == 1 and len(f.robots_blue) == 6 and len(f.robots_yellow) == 6
        assert sorted(m.robot_id for m in f.robots_blue) == list(range(6))
        assert all(abs(m.x) <= 5.2 and abs(m.y) <= 3.7 for m in list(f.robots_blue) + list(f.robots_yellow))
    # robots are dropped from 0.5 m (src/sslsim.cpp:358-365) and come to rest: the ball ends near the ground
    assert frames[-1].balls[0].z < 0.5
