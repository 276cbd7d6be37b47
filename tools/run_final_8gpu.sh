# final multi-GPU evidence of round 2 (one box, 8 B200); outputs under gpurun_out/r2y_*
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_workloads.py -q -k "multi_device or several_devices or one_process_per_gpu" > gpurun_out/r2y_multigpu_tests.log 2>&1)
(timeout 700 python tools/multigpu_report.py gpurun_out/r2y_multigpu.json > gpurun_out/r2y_multigpu.log 2>&1)
for N in 2 4 8; do timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2y_bench_n$N.json 2> gpurun_out/r2y_bench_n$N.err; done
(PTC_VERBOSE=1 timeout 200 vviewer_b200/_lib/offlinerender --scene Instanced --gpus 8 --split tile --spp 256 --out gpurun_out/r2y_c4 2>&1 | grep -E "^\[render\]|Scene rendered|backend" | tail -12) > gpurun_out/r2y_offlinerender_c4.log 2>&1
tail -3 gpurun_out/r2y_multigpu_tests.log; tail -18 gpurun_out/r2y_multigpu.log; for N in 2 4 8; do tail -c 300 gpurun_out/r2y_bench_n$N.json; echo; done; cat gpurun_out/r2y_offlinerender_c4.log
