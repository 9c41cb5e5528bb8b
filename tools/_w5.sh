for rep in 1 2; do
SSLSIM_WARM=0 python tools/var_time.py cold 4096,131072
SSLSIM_WARM=0.85 python tools/var_time.py warm 4096,131072
done
for v in 0 0.85; do
SSLSIM_WARM=$v ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__icc_request_hit_rate.pct,launch__occupancy_limit_shared_mem,launch__registers_per_thread --clock-control none -k regex:step_worlds -s 41 -c 1 --csv python tools/var_time.py x 131072 2>/dev/null | grep "step_worlds" | awk -F'","' -v v=$v '{gsub(/"/,"",$NF); print "ncu", v, $(NF-2), $NF}'
done
