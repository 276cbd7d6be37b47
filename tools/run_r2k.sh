mkdir -p gpurun_out
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_extend|k_shade' --csv --log-file gpurun_out/r2k_traffic.csv python tools/profile_run.py 1 Atrium > gpurun_out/r2k_traffic.log 2>&1
(timeout 600 python tools/workloads.py Cornell Atrium Fog Progressive Instanced:1.0) > gpurun_out/r2k_workloads.log 2>&1
timeout 300 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err
timeout 400 python bench.py --steps 16 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
grep -v "^\[" gpurun_out/r2k_workloads.log | cut -c1-250; tail -c 400 gpurun_out/r2k_bench_ref.json; tail -c 1200 gpurun_out/r2k_bench.json; tail -3 gpurun_out/r2k_traffic.log
