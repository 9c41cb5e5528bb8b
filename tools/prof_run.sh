#!/bin/bash
# Builds a -DSSLSIM_PROFILE library, runs tools/world_profile_fused.py (or the script given as $1) with it on a B200 through gpurun, and
# restores the normal build.
set -e
cd "$(dirname "$0")/.."
make -s -C ssl-sim_b200/csrc clean all EXTRA=-DSSLSIM_PROFILE >/dev/null 2>&1; cp ssl-sim_b200/libsslsim_b200.so tools/_lib_prof.so
make -s -C ssl-sim_b200/csrc clean all >/dev/null 2>&1
cat > tools/_prof_run.sh <<'EOS'
cp ssl-sim_b200/libsslsim_b200.so /tmp/orig.so
cp tools/_lib_prof.so ssl-sim_b200/libsslsim_b200.so
python PROF_SCRIPT
cp /tmp/orig.so ssl-sim_b200/libsslsim_b200.so
EOS
sed -i "s|PROF_SCRIPT|${1:-tools/world_profile_fused.py}|" tools/_prof_run.sh
/usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/_prof_run.sh' 2>&1 | grep -v "^\[gpurun\] sending"
rm -f tools/_lib_prof.so tools/_prof_run.sh
