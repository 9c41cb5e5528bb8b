# Round-2 measurement suite: run on a B200 with
#   gpurun --timeout 1700 -- 'bash tools/measure_suite_r2.sh TAG'
# then `python tools/make_profile_summary.py ...` here.  Every ncu pass runs only after the same command exited 0 without ncu.
T=${1:-r2}
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gputests.log 2>&1; tail -3 gpurun_out/${T}_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || { tail -20 gpurun_out/${T}_bench.err; exit 1; }
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2>> gpurun_out/${T}_bench.err
SMALL="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-sub-results --no-parity"
python bench.py $SMALL > gpurun_out/${T}_bench_small.json 2>> gpurun_out/${T}_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py $SMALL > gpurun_out/${T}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_worlds -s 2 -c 1 -f -o gpurun_out/${T}_prof python bench.py $SMALL > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_bench.err
