#!/bin/bash
# ncu --set full capture of the 131,072-world (or $W) fused launch for several builds, one gpurun session.
#   tools/ncu_variants.sh name1:"-DFLAGS" name2:"" ...   -> gpurun_out/ncu_<name>.ncu-rep
set -e
cd "$(dirname "$0")/.."
names=()
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"; names+=("$name")
  if [[ "$flags" == rev=* ]]; then
    rev="${flags#rev=}"; tmp=$(mktemp -d); git archive "$rev" ssl-sim_b200 include | tar -x -C "$tmp"
    make -s -C "$tmp/ssl-sim_b200/csrc" all >/dev/null 2>&1; cp "$tmp/ssl-sim_b200/libsslsim_b200.so" "tools/_lib_$name.so"; rm -rf "$tmp"
  else
    rm -f ssl-sim_b200/csrc/step_kernel.o
    make -s -C ssl-sim_b200/csrc all EXTRA="$flags" >/dev/null 2>&1; cp ssl-sim_b200/libsslsim_b200.so "tools/_lib_$name.so"
  fi
done
rm -f ssl-sim_b200/csrc/step_kernel.o; make -s -C ssl-sim_b200/csrc all >/dev/null 2>&1
cat > tools/_ncu_run.sh <<EOS
cp ssl-sim_b200/libsslsim_b200.so /tmp/orig.so
for v in ${names[@]}; do
  cp tools/_lib_\$v.so ssl-sim_b200/libsslsim_b200.so
  ncu --set full --import-source on --clock-control none -k regex:step_worlds -s 41 -c 1 -f -o gpurun_out/ncu_\$v python tools/var_time.py \$v ${W:-131072} > gpurun_out/ncu_\$v.log 2>&1
  tail -2 gpurun_out/ncu_\$v.log
done
cp /tmp/orig.so ssl-sim_b200/libsslsim_b200.so
EOS
for try in 1 2 3 4 5 6 7 8; do   # the pod sometimes answers "transient" (nothing charged): retry
  /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/_ncu_run.sh' > /tmp/_gpurun_variant.log 2>&1
  grep -q "status=transient" /tmp/_gpurun_variant.log || break
  sleep 120
done
grep -vE "^\[gpurun\] sending" /tmp/_gpurun_variant.log
for n in "${names[@]}"; do rm -f "tools/_lib_$n.so"; done; rm -f tools/_ncu_run.sh
