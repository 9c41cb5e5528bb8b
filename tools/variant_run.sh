#!/bin/bash
# A/B/C... timing of several builds of the step kernel inside ONE gpurun session (boxes differ by a few percent).
#   tools/variant_run.sh name1:"-DFLAG1 -DFLAG2" name2:"" ...     (EXTRA flags per variant; HEAD:<rev> builds git <rev>'s kernel)
# With NCU=1 also collects instructions / duration / active-cycle ratio of the 131,072-world launch per variant.
set -e
cd "$(dirname "$0")/.."
names=()
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  names+=("$name")
  if [[ "$flags" == rev=* ]]; then
    rev="${flags#rev=}"
    tmp=$(mktemp -d); git archive "$rev" ssl-sim_b200 include | tar -x -C "$tmp"
    make -s -C "$tmp/ssl-sim_b200/csrc" all >/dev/null 2>&1; cp "$tmp/ssl-sim_b200/libsslsim_b200.so" "tools/_lib_$name.so"; rm -rf "$tmp"
  else
    rm -f ssl-sim_b200/csrc/step_kernel.o
    make -s -C ssl-sim_b200/csrc all EXTRA="$flags" >/dev/null 2>&1; cp ssl-sim_b200/libsslsim_b200.so "tools/_lib_$name.so"
  fi
done
rm -f ssl-sim_b200/csrc/step_kernel.o; make -s -C ssl-sim_b200/csrc all >/dev/null 2>&1
cat > tools/_var_run.sh <<EOS
cp ssl-sim_b200/libsslsim_b200.so /tmp/orig.so
for rep in 1 2; do for v in ${names[@]}; do
  cp tools/_lib_\$v.so ssl-sim_b200/libsslsim_b200.so
  python tools/var_time.py \$v ${SIZES:-4096,131072}
done; done
if [ -n "$NCU" ]; then for v in ${names[@]}; do
  cp tools/_lib_\$v.so ssl-sim_b200/libsslsim_b200.so
  ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__icc_request_hit_rate.pct --clock-control none -k regex:step_worlds -s 41 -c 1 --csv python tools/var_time.py \$v 131072 2>/dev/null | grep "step_worlds" | awk -F'","' -v v=\$v '{gsub(/"/,"",\$NF); print "ncu", v, \$(NF-2), \$NF}'
done; fi
cp /tmp/orig.so ssl-sim_b200/libsslsim_b200.so
EOS
for try in 1 2 3 4 5 6 7 8; do   # the pod sometimes answers "transient" (nothing charged): retry
  /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/_var_run.sh' > /tmp/_gpurun_variant.log 2>&1
  grep -q "status=transient" /tmp/_gpurun_variant.log || break
  sleep 120
done
grep -vE "^\[gpurun\] sending" /tmp/_gpurun_variant.log
for n in "${names[@]}"; do rm -f "tools/_lib_$n.so"; done; rm -f tools/_var_run.sh
