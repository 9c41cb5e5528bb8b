SSLSIM_WARM=0.85 python tools/var_time.py warm 131072 > gpurun_out/w2_plain.log 2>&1 || exit 1
SSLSIM_WARM=0.85 ncu --set full --clock-control none --import-source on -k regex:step_worlds -s 41 -c 1 -f -o gpurun_out/w2_warm python tools/var_time.py warm 131072 > gpurun_out/w2_ncu.log 2>&1
tail -2 gpurun_out/w2_ncu.log
