python -m pytest tests -m gpu -x -q > gpurun_out/w1_gputests.log 2>&1; tail -15 gpurun_out/w1_gputests.log
for rep in 1 2; do
SSLSIM_WARM=0 python tools/var_time.py cold 4096,131072
SSLSIM_WARM=0.85 python tools/var_time.py warm 4096,131072
done
