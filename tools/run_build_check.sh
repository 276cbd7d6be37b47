# build kernels: bit-exact structure tests + render parity subset + build times (one B200)
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x -k "${TESTS:-lbvh or wide_bvh or two_level or ray_set or workload_configs or accel_mode}" > gpurun_out/r2b_gputests.log 2>&1)
(PTC_VERBOSE=1 timeout 300 python tools/build_time.py Atrium Instanced:0.25 Instanced:1.0) > gpurun_out/r2b_buildtime.log 2>&1
tail -3 gpurun_out/r2b_gputests.log | cut -c1-200; grep -v "^\[ptc\]" gpurun_out/r2b_buildtime.log | cut -c1-200; grep "^\[ptc\] build" gpurun_out/r2b_buildtime.log | tail -2 | cut -c1-250
