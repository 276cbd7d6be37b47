# compute-sanitizer over the round-2 kernels (one B200); output gpurun_out/r2_sanitizer.log
mkdir -p gpurun_out
: > gpurun_out/r2_sanitizer.log
for tool in ${TOOLS:-memcheck racecheck}; do
  echo "===== compute-sanitizer --tool $tool python tools/sanitize_run.py" >> gpurun_out/r2_sanitizer.log
  timeout 270 compute-sanitizer --tool $tool python tools/sanitize_run.py 2>&1 | grep -v "^=========     \|^=========$" | tail -40 >> gpurun_out/r2_sanitizer.log
done
grep -n "=====\|SUMMARY\|done\|rror" gpurun_out/r2_sanitizer.log | cut -c1-200 | head -40
