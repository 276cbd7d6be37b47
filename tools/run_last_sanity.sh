# sanity of the final tree on one B200: readback / workload tests, smoke, one bench line
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_workloads.py -q -k "readback or workload_configs or offlinerender or ball_on_plane" > gpurun_out/r2s_tests.log 2>&1)
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1
timeout 400 python bench.py --steps 16 --warmup 3 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
tail -2 gpurun_out/r2s_tests.log; tail -1 gpurun_out/r2s_smoke.log; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2s_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','build_ms','s_per_frame','gpu_launches')}, d['e2e']['value'], d['e2e']['seconds'], d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'])
P
