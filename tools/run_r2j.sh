mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2j_gputests.log 2>&1)
(timeout 300 python tools/env_sweep.py Fog:1.0:2 - ; timeout 200 python tools/env_sweep.py Atrium:1.0:4 -) > gpurun_out/r2j_fog.log 2>&1
(timeout 400 python tools/partition_probe.py Instanced:1.0 8 64; timeout 200 python tools/partition_probe.py Atrium 8 256) > gpurun_out/r2j_partition.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 4 -c 2 -o gpurun_out/r2j_k_extend -f python tools/profile_run.py 1 Atrium > gpurun_out/r2j_k_extend.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 4 -c 2 -o gpurun_out/r2j_k_shade -f python tools/profile_run.py 1 Atrium > gpurun_out/r2j_k_shade.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 6 -c 2 -o gpurun_out/r2j_fog_k_shade -f python tools/profile_run.py 1 Fog > gpurun_out/r2j_fog_k_shade.log 2>&1
tail -8 gpurun_out/r2j_gputests.log | cut -c1-200; cat gpurun_out/r2j_fog.log gpurun_out/r2j_partition.log | cut -c1-260; ls -la gpurun_out/r2j*.ncu-rep
