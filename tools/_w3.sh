python -m pytest tests -m gpu -x -q > gpurun_out/w4_gputests.log 2>&1; tail -3 gpurun_out/w4_gputests.log
for rep in 1 2; do
SSLSIM_WARM=0 python tools/var_time.py cold 4096,131072
SSLSIM_WARM=0.85 python tools/var_time.py warm 4096,131072
done
SSLSIM_WARM=0.85 ncu --set full --clock-control none --import-source on -k regex:step_worlds -s 41 -c 1 -f -o gpurun_out/w4_warm python tools/var_time.py warm 131072 > gpurun_out/w4_ncu.log 2>&1
