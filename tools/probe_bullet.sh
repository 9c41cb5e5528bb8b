#!/bin/bash
# Probe the GPU box (or any box) for a real Bullet Physics install (SURVEY 8c "If a real Bullet turns up",
# BASELINE.md 4.4).  Writes a report; exit code 0 always.  If Bullet is found, oracle/Makefile target `bullet`
# can build the reference's src/sslsim.cpp unmodified into oracle/_ref/ as a second oracle.
out=${1:-gpurun_out/bullet_probe.txt}
mkdir -p "$(dirname "$out")"
{
  echo "== bullet probe $(date -u +%FT%TZ) host=$(hostname)"
  echo "-- pkg-config --exists bullet"; if pkg-config --exists bullet 2>/dev/null; then echo "FOUND: $(pkg-config --modversion bullet)"; else echo "not found (rc=$?)"; fi
  echo "-- headers"; find / -xdev \( -name btBulletDynamicsCommon.h -o -name btDiscreteDynamicsWorld.h \) 2>/dev/null | head -5; echo "(end)"
  echo "-- libraries"; find / -xdev \( -name 'libBulletDynamics*' -o -name 'libBulletCollision*' -o -name 'libLinearMath*' \) 2>/dev/null | head -5; echo "(end)"
  echo "-- python"; python -c "import pybullet; print('pybullet', pybullet.__file__)" 2>&1 | tail -1
  echo "-- baseline/_ref"; ls -la baseline/_ref 2>&1 | head -5
  echo "-- reference tree"; ls /root/reference 2>&1 | head -3
  echo "-- cmake bullet module"; find / -xdev -name 'FindBullet.cmake' 2>/dev/null | head -2; echo "(end)"
  echo "-- nproc: $(nproc)"; nvidia-smi -L 2>&1 | head -8
} > "$out" 2>&1
cat "$out"
exit 0
