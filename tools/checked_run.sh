#!/bin/bash
# Memory-safety check without compute-sanitizer (closed on the GPU pool): builds libsslsim_b200.so with
# -DSSLSIM_CHECK (every shared-memory table index asserted, a violation traps), runs the whole GPU test-suite on
# it through gpurun, and restores the normal build.
set -e
cd "$(dirname "$0")/.."
make -s -C ssl-sim_b200/csrc clean all EXTRA=-DSSLSIM_CHECK >/dev/null 2>&1; cp ssl-sim_b200/libsslsim_b200.so tools/_lib_chk.so
make -s -C ssl-sim_b200/csrc clean all >/dev/null 2>&1
cat > tools/_chk_run.sh <<'EOS'
cp ssl-sim_b200/libsslsim_b200.so /tmp/orig.so
cp tools/_lib_chk.so ssl-sim_b200/libsslsim_b200.so
python -m pytest tests -m gpu -q 2>&1 | tail -4
cp /tmp/orig.so ssl-sim_b200/libsslsim_b200.so
EOS
/usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/_chk_run.sh' 2>&1 | grep -v "^\[gpurun\] sending"
rm -f tools/_lib_chk.so tools/_chk_run.sh
