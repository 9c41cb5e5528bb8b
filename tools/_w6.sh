for c in 3 4; do for w in 0 0.85; do SSLSIM_CFG=$c SSLSIM_WARM=$w python tools/var_time.py x 131072; done; done
SSLSIM_WARM=0.85 python tools/var_time.py x 4096,131072
