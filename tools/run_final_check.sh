# last validation of the round on a 2-GPU box: whole GPU suite (multi-GPU tests included), smoke, bench at N = 1 and N = 2 (torchrun)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2f_gputests.log 2>&1)
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
timeout 400 python bench.py --steps 16 --warmup 3 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2f_bench_ref_n2.json 2> gpurun_out/r2f_bench_ref_n2.err
tail -3 gpurun_out/r2f_gputests.log | cut -c1-200; tail -2 gpurun_out/r2f_smoke.log; for f in n1 n2 ref_n2; do tail -c 1800 gpurun_out/r2f_bench_$f.json; echo; tail -2 gpurun_out/r2f_bench_$f.err | cut -c1-300; done
