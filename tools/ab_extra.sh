#!/bin/bash
# A/B timing in one gpurun session: A = working tree as is, B = working tree built with EXTRA="$1" (compile flags).
set -e
cd "$(dirname "$0")/.."
make -s -C ssl-sim_b200/csrc clean all >/dev/null 2>&1; cp ssl-sim_b200/libsslsim_b200.so tools/_lib_A.so
make -s -C ssl-sim_b200/csrc clean all EXTRA="$1" >/dev/null 2>&1; cp ssl-sim_b200/libsslsim_b200.so tools/_lib_B.so
make -s -C ssl-sim_b200/csrc clean all >/dev/null 2>&1
cat > tools/_ab_run.sh <<'EOS'
cp ssl-sim_b200/libsslsim_b200.so /tmp/orig.so
for rep in 1 2; do for v in A B; do
  cp tools/_lib_$v.so ssl-sim_b200/libsslsim_b200.so
  python bench.py --no-cpu-baseline --no-e2e --large-worlds 131072 --steps 100 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', '%.4e' % d['value'], '%.4e' % d['large_batch']['value'], d['episode_stats']['checksum'])"
done; done
cp /tmp/orig.so ssl-sim_b200/libsslsim_b200.so
EOS
/usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/_ab_run.sh' 2>&1 | grep -E "^A |^B |status"
rm -f tools/_lib_A.so tools/_lib_B.so tools/_ab_run.sh
