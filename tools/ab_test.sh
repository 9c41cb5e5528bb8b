#!/bin/bash
# A/B timing of two builds of libsslsim_b200.so in ONE gpurun session (boxes differ by a few percent, so numbers
# from different sessions are not comparable).  A = git HEAD's step_kernel.cu, B = the working tree's.
#   tools/ab_test.sh            builds both, runs them alternately on a B200 through gpurun
set -e
cd "$(dirname "$0")/.."
K=ssl-sim_b200/csrc/step_kernel.cu
make -s -C ssl-sim_b200/csrc all >/dev/null 2>&1; cp ssl-sim_b200/libsslsim_b200.so tools/_lib_B.so
cp $K /tmp/_step_kernel_B.cu
git show HEAD:$K > $K
make -s -C ssl-sim_b200/csrc all >/dev/null 2>&1; cp ssl-sim_b200/libsslsim_b200.so tools/_lib_A.so
cp /tmp/_step_kernel_B.cu $K
make -s -C ssl-sim_b200/csrc all >/dev/null 2>&1
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
