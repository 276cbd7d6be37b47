# final single-GPU evidence of round 2 (one B200); outputs under gpurun_out/r2z_*
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2z_gputests.log 2>&1)
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_extend|k_shade' --csv --log-file gpurun_out/r2z_traffic.csv python tools/profile_run.py 1 Atrium > gpurun_out/r2z_traffic.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 4 -c 2 -o gpurun_out/r2z_k_shade -f python tools/profile_run.py 1 Atrium > gpurun_out/r2z_k_shade.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 4 -c 2 -o gpurun_out/r2z_k_extend -f python tools/profile_run.py 1 Atrium > gpurun_out/r2z_k_extend.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2z_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-frame > gpurun_out/r2z_bench_under_ncu.log 2>&1
(timeout 600 python tools/workloads.py Cornell Atrium Fog Progressive Instanced:1.0) > gpurun_out/r2z_workloads.log 2>&1
(timeout 300 python tools/ab_r1.py Atrium:4 Cornell:8 Fog:2 Progressive:4) > gpurun_out/r2z_ab.log 2>&1
timeout 300 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err
timeout 400 python bench.py --steps 16 --warmup 3 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1
tail -3 gpurun_out/r2z_gputests.log | cut -c1-200; grep -v "^\[" gpurun_out/r2z_workloads.log | cut -c1-230; cat gpurun_out/r2z_ab.log | cut -c1-200; tail -c 1500 gpurun_out/r2z_bench.json; cat gpurun_out/r2z_smoke.log | tail -2
