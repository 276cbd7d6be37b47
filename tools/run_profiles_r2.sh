mkdir -p gpurun_out
(timeout 300 python tools/env_sweep.py Fog:1.0:2 - ) > gpurun_out/r2i_fog.log 2>&1
(timeout 500 python tools/env_sweep.py Instanced:1.0:2 - PTC_LIB=vviewer_b200/_lib/libptc_cuda_mb10.so PTC_LIB=vviewer_b200/_lib/libptc_cuda_mb12.so; timeout 200 python tools/env_sweep.py Atrium:1.0:4 PTC_LIB=vviewer_b200/_lib/libptc_cuda_mb10.so) > gpurun_out/r2i_occupancy.log 2>&1
(timeout 400 python tools/partition_probe.py Instanced:1.0 8 64; timeout 200 python tools/partition_probe.py Atrium 8 256) > gpurun_out/r2i_partition.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2i_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-frame > gpurun_out/r2i_bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 20 -c 2 -o gpurun_out/r2i_k_extend -f python tools/profile_run.py 1 Atrium > gpurun_out/r2i_k_extend.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 20 -c 2 -o gpurun_out/r2i_k_shade -f python tools/profile_run.py 1 Atrium > gpurun_out/r2i_k_shade.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 40 -c 2 -o gpurun_out/r2i_fog_k_shade -f python tools/profile_run.py 1 Fog > gpurun_out/r2i_fog_k_shade.log 2>&1
cat gpurun_out/r2i_fog.log gpurun_out/r2i_occupancy.log gpurun_out/r2i_partition.log | cut -c1-260
ls -la gpurun_out/*.ncu-rep
