# same-box A/B of library variants: AB_LIBS="label=path,..." bash tools/run_ab_variants.sh Scene:batches ...
mkdir -p gpurun_out
timeout 900 python tools/ab_r1.py "$@" > gpurun_out/ab_variants.log 2>&1
cat gpurun_out/ab_variants.log | cut -c1-220
