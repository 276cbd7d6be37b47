mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2o_gputests.log 2>&1)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_radix|k_ploc|k_wide|k_gather|k_flatten|k_morton' --csv --log-file gpurun_out/r2o_build_launches.csv python tools/build_time.py Instanced:1.0 > gpurun_out/r2o_build_under_ncu.log 2>&1
(PTC_VERBOSE=1 timeout 300 python tools/build_time.py Atrium Instanced:1.0) > gpurun_out/r2o_buildtime.log 2>&1
(timeout 300 python tools/ab_r1.py Atrium:4 Fog:2) > gpurun_out/r2o_ab.log 2>&1
tail -4 gpurun_out/r2o_gputests.log | cut -c1-200; grep -v "^\[render" gpurun_out/r2o_buildtime.log | tail -6 | cut -c1-250; cat gpurun_out/r2o_ab.log | cut -c1-200
