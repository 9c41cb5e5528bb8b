cp ssl-sim_b200/libsslsim_b200.so /tmp/orig.so
cp tools/_lib_prof.so ssl-sim_b200/libsslsim_b200.so
SSLSIM_WARM=0 python tools/world_profile_config3.py
SSLSIM_WARM=0.85 python tools/world_profile_config3.py
cp /tmp/orig.so ssl-sim_b200/libsslsim_b200.so
