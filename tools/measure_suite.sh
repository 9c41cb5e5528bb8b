# The measurement suite behind profiles/r1_v14_*: run on a B200 with
#   gpurun --timeout 1500 -- 'bash tools/measure_suite.sh'
# then `python tools/make_profile_summary.py r1_v14 gpurun_out/prof_v14.ncu-rep gpurun_out/launches_v14.csv
# gpurun_out/bench_v14.json 1228800` here.  Every ncu pass runs only after the same command exited 0 without ncu.
set -x
python bench.py > gpurun_out/bench_v14.json 2> gpurun_out/bench_v14.err || exit 1
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_v14_ref.json 2>> gpurun_out/bench_v14.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_v14.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --large-worlds 0 > gpurun_out/ncu_launches_v14.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_worlds -s 4 -c 2 -f -o gpurun_out/prof_v14 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --large-worlds 0 > gpurun_out/ncu_v14.log 2>&1
tail -2 gpurun_out/bench_v14.err
