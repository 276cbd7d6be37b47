# per-call host work of the multi-GPU path after the geometry-pool cache, the NVLink geometry broadcast and the staged readback
# (one box, 8 B200); outputs under gpurun_out/r2w_*
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_workloads.py -q -k "multi_device or several_devices or one_process_per_gpu or communicator or readback" > gpurun_out/r2w_multigpu_tests.log 2>&1)
(timeout 600 python tools/multigpu_report.py gpurun_out/r2w_multigpu.json --configs=${CONFIGS:-c2,c4} --counts=${COUNTS:-1,8} --verbose > gpurun_out/r2w_multigpu.log 2> gpurun_out/r2w_phases.log)
tail -3 gpurun_out/r2w_multigpu_tests.log; tail -9 gpurun_out/r2w_multigpu.log; grep -A12 "C4 instanced" gpurun_out/r2w_phases.log | cut -c1-120
