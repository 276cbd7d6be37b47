# build kernels with staged record stores: whole GPU suite + build times (one B200)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_gputests.log 2>&1)
(PTC_VERBOSE=1 timeout 300 python tools/build_time.py Atrium Instanced:0.25 Instanced:1.0) > gpurun_out/r2b_buildtime.log 2>&1
tail -3 gpurun_out/r2b_gputests.log | cut -c1-200; grep -v "^\[ptc\]" gpurun_out/r2b_buildtime.log | cut -c1-200; grep "^\[ptc\] build" gpurun_out/r2b_buildtime.log | tail -3 | cut -c1-250
